"""Literal, brute-force Python model of the reference's BED path — a SECOND,
independent restatement used only to cross-check the C oracle on small inputs
(every query scans every block; no index).  Follows /root/reference
source/swiftover/chain.d:272-405, 509-597, 664-743 and bed.d:113-202 line by line.
Test infrastructure only.
"""
WS = b" \t\n\v\f\r"


def to_int(tok: bytes) -> int:
    s = tok.decode("ascii")
    body = s[1:] if s[:1] in "+-" else s
    if not body.isdigit() or not body.isascii():
        raise ValueError(tok)
    v = int(s)
    if not -2**31 <= v < 2**31:
        raise ValueError(tok)
    return v


class PyChainFile:
    def __init__(self, text: bytes):
        lines = text.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()  # byLineCopy: no empty line after a final terminator
        self.trees = {}          # target contig -> list of links (insertion order)
        self.contig_names = []   # chain.d:500
        self.contig_ids = {}
        self.q_sizes = {}        # chain.d:504
        chain_start = 0
        for i, line in enumerate(lines):
            if len(line) > 5 and line[:5] == b"chain":
                if i - 1 > 0:
                    self._add_chain(lines[chain_start:i])
                chain_start = i
        self._add_chain(lines[chain_start:])
        # tree order + duplicate suppression (opCmp == 0 -> existing node kept)
        for k, v in self.trees.items():
            seen, out = set(), []
            for l in sorted(enumerate(v), key=lambda t: (t[1][0], t[1][1], t[1][2], t[0])):
                key = l[1][:3]
                if key in seen:
                    continue
                seen.add(key)
                out.append(l[1])
            self.trees[k] = out

    def _add_chain(self, lines):
        h = lines[0].rstrip(b"\r").split(b" ")
        assert h[0] == b"chain"
        t_name, t_size, t_strand = h[2], to_int(h[3]), h[4][:1]
        t_start = to_int(h[5]) if t_strand == b"+" else t_size - to_int(h[5])
        q_name, q_size, q_strand = h[7], to_int(h[8]), h[9][:1]
        q_start = to_int(h[10]) if q_strand == b"+" else q_size - to_int(h[10])
        to_int(h[12])
        invert = 1 if t_strand == q_strand else -1
        t_from, q_from = t_start, q_start
        for line in lines[1:]:
            if len(line) == 0:
                continue
            d = [to_int(x) for x in line.split()]
            if not d:
                continue
            if q_name not in self.contig_ids:
                self.contig_ids[q_name] = len(self.contig_names)
                self.contig_names.append(q_name)
            link = (t_from, t_from + d[0], q_from, self.contig_ids[q_name], invert)
            self.trees.setdefault(t_name, []).append(link)
            if len(d) == 3:
                t_from += d[0] + d[1]
                q_from += invert * (d[0] + d[2])
            elif len(d) != 1:
                raise ValueError("Unexpected length of alignment data line")
        self.q_sizes[q_name] = q_size

    @staticmethod
    def intersect(link, s, e):
        ts, te, qs, qc, inv = link
        a, b = max(ts, s), min(te, e)
        return (a, b, qs + inv * (a - ts), qc, inv)

    def lift(self, contig: bytes, s: int, e: int):
        return [self.intersect(l, s, e) for l in self.trees.get(contig, []) if l[0] < e and l[1] > s]

    def lift_coord_only(self, contig, c):
        o = self.lift(contig, c, c + 1)
        return (1, o[0][2]) if o else (0, c)


def strand_table(c: int, inv: int) -> int:
    idx = c + inv
    return {0x2A: ord("-"), 0x2C: ord("+"), 0x2E: ord("-")}.get(idx, 33)


def lift_bed(cf: PyChainFile, text: bytes):
    lifted, unmatched = [], []
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    for line in lines:
        if len(line) > 0 and line[:1] == b"#":
            continue
        if len(line) > 4 and line[:5] in (b"track", b"brows"):
            continue
        f = line.split()
        numf = len(f)
        start, end = to_int(f[1]), to_int(f[2])
        links = cf.lift(f[0], start, end)
        has_thick = False
        if numf >= 8:
            ts, te = to_int(f[6]), to_int(f[7])
            if ts != te:
                has_thick = True
                r1, ts = cf.lift_coord_only(f[0], ts)
                if not r1:
                    has_thick = False
                else:
                    r2, te = cf.lift_coord_only(f[0], te)
                    if not r2:
                        has_thick = False
                    else:
                        ts, te = min(ts, te), max(ts, te)
        if not links:
            unmatched.append(b"\t".join(f) + b"\n")
            continue
        f = list(f)
        for l in sorted(links, key=lambda l: l[2]):  # Python's sort is stable
            f[0] = cf.contig_names[l[3]]
            s, e = l[2], l[2] + l[4] * (l[1] - l[0])
            s, e = min(s, e), max(s, e)
            f[1], f[2] = str(s).encode(), str(e).encode()
            if numf >= 6:
                f[5] = bytes([strand_table(f[5][0], l[4])]) + f[5][1:]
            if numf >= 8:
                if not has_thick or ts >= e or te <= s:
                    f[6] = f[7] = f[1]
                else:
                    f[6] = str(max(ts, s)).encode()
                    f[7] = str(min(te, e)).encode()
            lifted.append(b"\t".join(f) + b"\n")
    return b"".join(lifted), b"".join(unmatched)
