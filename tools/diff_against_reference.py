#!/usr/bin/env python
"""Compare this repo's liftBED output with the REAL swiftover's (SURVEY.md 8c policy).

The reference's outputs are not reproducible from its sources alone in one respect: rows of one
input record with equal sort key (raw qStart, bed.d:156) come out in an order that depends on
the splay tree's shape, i.e. on every earlier query, and on Phobos' unstable sort
(bed.d:154-157: "order is undefined").  This repo defines that order (qStart, then block index).
So two comparisons are reported for every (BED fixture, chain) case:

  exact        byte equality of the lifted and the unmatched file
  modulo ties  per input record: same multiset of rows, and every row sits inside the run of
               positions that holds its sort key — i.e. the files differ only by permutations
               inside equal-qStart groups.  Counted: records with ties, tie groups, groups whose
               order differs.  (With >= 6 columns the strand byte of the rows after a permuted
               group may differ too — the flips accumulate in output order, bed.d:173 — such
               rows are compared with column 6 masked and counted separately.)

Reference outputs come from --ref-dir (files <case>.lifted.bed / <case>.unmatched.bed, see
tools/make_reference_outputs.sh) or are produced on the spot by --swiftover BIN (also looked
for on $PATH and in baseline/_ref/).  "Ours" is the CUDA engine (--engine cuda, needs a GPU) or
the CPU oracle (--engine oracle; the two are byte-identical, tests/test_gpu_parity.py).
"""
import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
RES = os.path.join(ROOT, "tests", "golden", "resources")

# (case name, BED fixture, chain fixture): every pair whose BED contigs all exist in the chain
# (the reference's behaviour on an unknown contig is undefined, README.md:59-62)
CASES = [
    ("hg19_col123__hg19ToHg38", "hg19_col123.bed", "hg19ToHg38.over.chain"),
    ("chr1_hg19__hg19ToHg38", "chr1_hg19.bed", "hg19ToHg38.over.chain"),
    ("chr1_hg19__chr1_first", "chr1_hg19.bed", "chr1_first.chain"),
    ("chr1_hg19__minimal", "chr1_hg19.bed", "minimal.chain"),
    ("test__hg19ToHg38", "test.bed", "hg19ToHg38.over.chain"),
    ("test__chr1_first", "test.bed", "chr1_first.chain"),
    ("test__minimal", "test.bed", "minimal.chain"),
    ("nomatch__hg19ToHg38", "nomatch_hg19.bed", "hg19ToHg38.over.chain"),
    ("nomatch__minimal", "nomatch_hg19.bed", "minimal.chain"),
    ("empty__hg19ToHg38", "empty.bed", "hg19ToHg38.over.chain"),
]


def fixture(name):
    return gzip.open(os.path.join(RES, name + ".gz"), "rb").read()


def find_swiftover(explicit):
    for c in (explicit, shutil.which("swiftover"), os.path.join(ROOT, "baseline", "_ref", "swiftover"),
              os.path.join(ROOT, "baseline", "_ref", "bin", "swiftover")):
        if c and os.path.isfile(c) and os.access(c, os.X_OK):
            return c
    return None


def run_reference(binary, bed, chain):
    with tempfile.TemporaryDirectory() as d:
        pc, pi, po, pu = (os.path.join(d, x) for x in ("c.chain", "in.bed", "out.bed", "un.bed"))
        open(pc, "wb").write(chain)
        open(pi, "wb").write(bed)
        subprocess.run([binary, "-t", "bed", "-c", pc, "-i", pi, "-o", po, "-u", pu], check=True, capture_output=True)
        return open(po, "rb").read(), open(pu, "rb").read()


class Ours:
    def __init__(self, engine):
        from oracle_ffi import Oracle, build_oracle
        build_oracle()
        self.oracle = Oracle()
        self.engine = engine

    def lift(self, bed, chain):
        """-> lifted, unmatched, per-record row keys (list of lists of raw qStart in output order)"""
        oc = self.oracle.chain_parse(chain)
        if self.engine == "cuda":
            from swiftover_b200 import ChainFile
            cf = ChainFile(text=chain)
            lifted, unm, _ = cf.lift_bed_text(bed)
            cf.close()
        else:
            lifted, unm, _ = oc.lift_bed(bed)
        keys = []
        for line in bed.splitlines():
            f = line.split()
            if not f or line.startswith((b"#", b"track", b"brows")):
                continue  # bed.d:115-116
            links = oc.lift(f[0].decode(), int(f[1]), int(f[2]))
            keys.append(sorted(l[2] for l in links))  # raw qStart of every trimmed link, ascending (bed.d:156)
        return lifted, unm, keys


def compare_modulo_ties(ours, ref, keys, ncols_strand=True):
    """ours / ref: lifted files.  keys: per record, its rows' sort keys in output order."""
    a, b = ours.split(b"\n"), ref.split(b"\n")
    if a and a[-1] == b"":
        a.pop()
    if b and b[-1] == b"":
        b.pop()
    rep = dict(rows_ours=len(a), rows_ref=len(b), records=len(keys), records_with_ties=0, tie_groups=0,
               tie_groups_permuted=0, rows_strand_only=0, records_different=0)
    if len(a) != len(b) or len(a) != sum(len(k) for k in keys):
        rep["comparable"] = False
        return rep
    rep["comparable"] = True
    mask6 = lambda ln: b"\t".join(c if i != 5 else b"?" for i, c in enumerate(ln.split(b"\t")))
    p = 0
    for ks in keys:
        n = len(ks)
        ra, rb = a[p:p + n], b[p:p + n]
        p += n
        if ra == rb:
            if len(set(ks)) < n:
                rep["records_with_ties"] += 1
                rep["tie_groups"] += sum(1 for k in set(ks) if ks.count(k) > 1)
            continue
        ok = True
        i = 0
        while i < n:
            j = i
            while j < n and ks[j] == ks[i]:
                j += 1
            ga, gb = ra[i:j], rb[i:j]
            if j - i > 1:
                rep["tie_groups"] += 1
            if ga != gb:
                if sorted(ga) == sorted(gb):
                    rep["tie_groups_permuted"] += 1
                elif sorted(map(mask6, ga)) == sorted(map(mask6, gb)):
                    rep["tie_groups_permuted"] += (j - i > 1)
                    rep["rows_strand_only"] += sum(1 for x, y in zip(sorted(ga), sorted(gb)) if x != y)
                else:
                    ok = False
            i = j
        if len(set(ks)) < n:
            rep["records_with_ties"] += 1
        if not ok:
            rep["records_different"] += 1
    return rep


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref-dir", help="directory with <case>.lifted.bed / <case>.unmatched.bed from the reference")
    ap.add_argument("--swiftover", help="reference binary (default: $PATH, baseline/_ref/)")
    ap.add_argument("--engine", default="oracle", choices=["oracle", "cuda"])
    ap.add_argument("--json", action="store_true")
    args = ap.parse_args()
    binary = None if args.ref_dir else find_swiftover(args.swiftover)
    if not args.ref_dir and not binary:
        print("no reference outputs: give --ref-dir (see tools/make_reference_outputs.sh) or a swiftover binary "
              "(--swiftover, $PATH or baseline/_ref/)", file=sys.stderr)
        return 2
    ours = Ours(args.engine)
    report, worst = {}, 0
    for case, bed_name, chain_name in CASES:
        bed, chain = fixture(bed_name), fixture(chain_name)
        if args.ref_dir:
            pl, pu = (os.path.join(args.ref_dir, f"{case}.{k}.bed") for k in ("lifted", "unmatched"))
            if not (os.path.exists(pl) and os.path.exists(pu)):
                report[case] = {"status": "no reference output"}
                continue
            ref_l, ref_u = open(pl, "rb").read(), open(pu, "rb").read()
        else:
            ref_l, ref_u = run_reference(binary, bed, chain)
        our_l, our_u, keys = ours.lift(bed, chain)
        r = {"exact_lifted": our_l == ref_l, "exact_unmatched": our_u == ref_u}
        if not r["exact_lifted"]:
            r["modulo_ties"] = compare_modulo_ties(our_l, ref_l, keys)
            m = r["modulo_ties"]
            r["status"] = ("equal modulo tie order" if m.get("comparable") and m["records_different"] == 0 and r["exact_unmatched"]
                           else "DIFFERENT")
        else:
            r["status"] = "identical" if r["exact_unmatched"] else "DIFFERENT (unmatched file)"
        worst = max(worst, 0 if r["status"] == "identical" else 1 if r["status"].startswith("equal") else 3)
        report[case] = r
    if args.json:
        print(json.dumps(report, indent=1))
    else:
        for case, r in report.items():
            print(f"{case:28s} {r['status']}" + (f"  {r['modulo_ties']}" if "modulo_ties" in r else ""))
    return worst


if __name__ == "__main__":
    sys.exit(main())
