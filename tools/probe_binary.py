#!/usr/bin/env python
"""Box probe (not a bench line): lift_intervals_kernel alone on BASELINE configs[1]
(sorted BED3), the same records shuffled, and shuffled BED8 (strand + thick).
PROBE_N = queries (default 1e8), PROBE_CASES = comma list of labels, PROBE_REPS."""
import gzip, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S

dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = int(float(os.environ.get("PROBE_N", 100_000_000)))
reps = int(os.environ.get("PROBE_REPS", 10))
cases = os.environ.get("PROBE_CASES", "sorted_bed3,shuffled_bed3,shuffled_bed8").split(",")
allc = dict(sorted_bed3=dict(sort=True, len_hi=10_000), shuffled_bed3=dict(sort=False, len_hi=10_000),
            sorted_bed8=dict(sort=True, len_hi=100_000, strand=True, thick=True),
            shuffled_bed8=dict(sort=False, len_hi=100_000, strand=True, thick=True))
out = {}
for label in cases:
    kw = allc[label]
    iv = S.gen_intervals(cf, N, 20240002, device=dev, **kw)
    bufs = S.BinaryLiftBuffers(N, device=dev, rows_factor=1.6, thick="thick_start" in iv)
    st = torch.cuda.current_stream(dev)
    for _ in range(3): S.lift_intervals_dev(cf, iv, bufs, st)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(st); S.lift_intervals_dev(cf, iv, bufs, st); b.record(st)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
    rows, unm = bufs.totals.cpu().tolist()
    bq = 12 + (1 if "strand" in iv else 0) + (8 if "thick_start" in iv else 0)
    bo = 16 + (8 if "thick_start" in iv else 0)
    alg = N * (bq + 4) + rows * bo
    out[label] = dict(ms=round(ms, 4), G_intervals_per_s=round(N / ms / 1e6, 2), rows=rows, unmatched=unm,
                      alg_bytes=alg, GBs=round(alg / ms / 1e6, 1))
    del iv, bufs
    torch.cuda.empty_cache()
print(json.dumps(out))
