#!/usr/bin/env python
"""GPU fuzz (not a test, not a bench): random chains and batches with random degrees of
sortedness, contig counts and fan-out, binary and text paths against the CPU oracle.
Each case runs in the calling process; a CUDA error or a mismatch stops the run and prints the
seed.  FUZZ_SECONDS (default 120), FUZZ_SEED0 (default 0), FUZZ_BIG=1: 3e5-7e5 records per case,
FUZZ_HG=1: the hg19ToHg38 chain instead of random chains."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth, oracle_ffi
from swiftover_b200 import ChainFile

orc = oracle_ffi.Oracle()
t_end = time.time() + float(os.environ.get("FUZZ_SECONDS", 120))
seed = int(os.environ.get("FUZZ_SEED0", 0))
ncase = 0
while time.time() < t_end:
    rng = np.random.default_rng(50_000 + seed)
    ntc = int(rng.integers(1, 5))
    chain, tn = synth.random_chain(rng, n_chains=int(rng.integers(1, 16)), n_tcontigs=ntc, max_blocks=int(rng.choice([5, 40, 200])),
                                   overlap=bool(rng.random() < 0.3))
    hg = bool(os.environ.get("FUZZ_HG"))  # the reference's hg19ToHg38 chain (53,950 blocks, grouped bottom search)
    if hg:
        import gzip
        chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
    cf, oc = ChainFile(text=chain), orc.chain_parse(chain)
    names = cf.sourceContigs
    big = bool(os.environ.get("FUZZ_BIG"))  # many tiles per CTA (the deferred flush / pipelining paths)
    n = int(rng.choice([300_000, 700_000])) if big else int(rng.choice([1, 7, 1023, 1025, 5000, 20_000, 60_000]))
    tc = rng.integers(-1 if rng.random() < 0.2 else 0, min(len(names), ntc + 1) if hg else len(names), n).astype(np.int32)
    span = int(rng.choice([3_000_000, 60_000_000, 150_000_000])) if hg else int(rng.choice([3_000, 30_000, 120_000]))
    s = rng.integers(0, span, n).astype(np.int32)
    heavy = rng.random(n) < float(rng.choice([0.0, 0.02, 0.3]))
    ln = np.where(heavy, rng.integers(1_000, (2_000_000 if hg else span) + 2, n), rng.integers(1, int(rng.choice([20, 400, 4000])), n))
    order = np.lexsort((s, tc))
    tc, s, ln = tc[order], s[order], ln[order]
    m = int(n * float(rng.choice([0.0, 0.05, 0.12, 0.2, 0.5, 0.9, 1.0])))
    perm = np.concatenate([np.arange(n - m), (n - m) + rng.permutation(m)])
    if rng.random() < 0.3:
        perm = perm[::-1].copy()  # shuffled part first
    tc, s, ln = tc[perm], s[perm], ln[perm]
    e = (s + ln).astype(np.int32)
    strand = rng.choice(np.frombuffer(b"+-.", np.uint8), n)
    ts = (s + rng.integers(0, 200, n)).astype(np.int32)
    te = (ts + rng.integers(0, 200, n)).astype(np.int32)
    what = "?"
    try:
        for what, args in (("bed3", (tc, s, e)), ("bed6", (tc, s, e, strand)), ("bed8", (tc, s, e, strand, ts, te))):
            got, ref = cf.lift_intervals(*args), oc.lift_intervals(*args)
            keys = ["row_offset", "qcid", "start", "end", "src"] + (["strand"] if len(args) > 3 else []) + (["thick_start", "thick_end"] if len(args) > 4 else [])
            assert got["n_rows"] == ref["n_rows"] and got["n_unmatched"] == ref["n_unmatched"], "totals"
            for k in keys:
                assert np.array_equal(got[k], ref[k]), k
        ok = tc >= 0
        cols = int(rng.choice([3, 6, 8, 12]))
        lines = []
        for i in np.nonzero(ok)[0][: (10**9 if big else 30_000)]:
            nm = names[tc[i]] if isinstance(names[tc[i]], bytes) else names[tc[i]].encode()
            f = [nm, b"%d" % s[i], b"%d" % e[i]]
            if cols >= 6: f += [b"n%d" % i, b"0", bytes([strand[i]])]
            if cols >= 8: f += [b"%d" % ts[i], b"%d" % te[i]]
            if cols >= 12: f += [b"0", b"2", b"10,20,", b"0,30,"]
            lines.append(b"\t".join(f))
        bed = b"\n".join(lines) + (b"\n" if lines else b"")
        what = "text%d" % cols
        lifted, unm, st = cf.lift_bed_text(bed)
        o_l, o_u, o_st = oc.lift_bed(bed)
        assert lifted == o_l and unm == o_u, "text bytes"
    except Exception as ex:  # noqa
        print(f"FAIL seed {seed} case {what} n {n} ntc {ntc} shuffled {m}: {type(ex).__name__}: {str(ex)[:300]}", flush=True)
        sys.exit(1)
    cf.close()
    ncase += 1
    seed += 1
print(f"fuzz ok: {ncase} cases, seeds {int(os.environ.get('FUZZ_SEED0', 0))}..{seed - 1}")
