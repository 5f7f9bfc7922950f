"""Seeded synthetic chains / BED text for tests (numpy; small sizes)."""
import numpy as np


def random_chain(rng, n_chains=6, n_tcontigs=3, n_qcontigs=4, max_blocks=30, overlap=False, sep=b"\t"):
    """A chain file with target-disjoint chains (like UCSC over.chain) unless overlap=True."""
    out = []
    tnames = [f"chr{i + 1}".encode() for i in range(n_tcontigs)]
    qnames = [f"q{chr(65 + i)}".encode() for i in range(n_qcontigs)]
    tsize, qsize = 2_000_000, 3_000_000
    cursor = {t: int(rng.integers(0, 500)) for t in tnames}
    for cid in range(n_chains):
        t = tnames[int(rng.integers(0, n_tcontigs))]
        q = qnames[int(rng.integers(0, n_qcontigs))]
        qstrand = b"+" if rng.random() < 0.6 else b"-"
        nb = int(rng.integers(1, max_blocks + 1))
        sizes = rng.integers(1, 400, nb)
        dts = rng.integers(0, 60, nb - 1)
        dqs = rng.integers(0, 60, nb - 1)
        span_t = int(sizes.sum() + dts.sum())
        span_q = int(sizes.sum() + dqs.sum())
        if overlap:
            ts = int(rng.integers(0, 20_000))
        else:
            ts = cursor[t] + int(rng.integers(0, 300))
            cursor[t] = ts + span_t
        qs = int(rng.integers(0, qsize - span_q - 1))
        out.append(b"chain %d %s %d + %d %d %s %d %s %d %d %d" % (
            1000 - cid, t, tsize, ts, ts + span_t, q, qsize, qstrand, qs, qs + span_q, cid + 1))
        for i in range(nb):
            if i < nb - 1:
                out.append(sep.join([b"%d" % sizes[i], b"%d" % dts[i], b"%d" % dqs[i]]))
            else:
                out.append(b"%d" % sizes[i])
        out.append(b"")
    return b"\n".join(out) + b"\n", tnames


def random_bed(rng, tnames, n, ncols=3, span=25_000, maxlen=3000, sep=b"\t", unknown_contig_frac=0.0):
    rows = []
    for i in range(n):
        t = tnames[int(rng.integers(0, len(tnames)))]
        if rng.random() < unknown_contig_frac:
            t = b"chrNope"
        s = int(rng.integers(0, span))
        L = int(rng.integers(1, maxlen)) if rng.random() < 0.9 else int(rng.integers(1, 20))
        e = s + L
        f = [t, b"%d" % s, b"%d" % e]
        if ncols >= 4:
            f.append(b"n%d" % i)
        if ncols >= 5:
            f.append(b"%d" % int(rng.integers(0, 1000)))
        if ncols >= 6:
            f.append([b"+", b"-", b".", b"+x"][int(rng.choice(4, p=[0.45, 0.45, 0.07, 0.03]))])
        if ncols >= 8:
            r = rng.random()
            if r < 0.2:
                a = b_ = s
            else:
                a = s + int(rng.integers(0, L))
                b_ = a + int(rng.integers(0, e - a + 1))
            f += [b"%d" % a, b"%d" % b_]
        if ncols >= 9:
            f += [b"0", b"2", b"10,20,", b"0,30,"][: ncols - 8]
        rows.append(sep.join(f))
    return b"\n".join(rows) + (b"\n" if rows else b"")
