#!/usr/bin/env python
"""Box probe (not a bench line): PCIe copy ceilings for the e2e pipeline and the
binary lift kernel's time on BASELINE configs[1] (sorted) / shuffled input."""
import gzip, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S

dev = torch.device("cuda", 0)
out = {}
n = 1 << 31
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best

def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
def chunked(chunk):
    def f():
        for o in range(0, n, chunk):
            with torch.cuda.stream(s1): d_a[o:o+chunk].copy_(h_a[o:o+chunk], non_blocking=True)
            with torch.cuda.stream(s2): h_b[o:o+chunk].copy_(d_b[o:o+chunk], non_blocking=True)
    return f
out["h2d_GBs"] = n / timed(h2d) / 1e9
out["d2h_GBs"] = n / timed(d2h) / 1e9
out["bidir_each_GBs"] = n / timed(both) / 1e9
for c in (8 << 20, 64 << 20, 256 << 20):
    out[f"bidir_chunk{c>>20}MiB_each_GBs"] = n / timed(chunked(c)) / 1e9
del h_a, h_b, d_a, d_b
torch.cuda.empty_cache()

chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = int(os.environ.get("PROBE_N", 100_000_000))
for label, kw in (("sorted_bed3", dict(sort=True, len_hi=10_000)), ("shuffled_bed3", dict(sort=False, len_hi=10_000)),
                  ("shuffled_bed8", dict(sort=False, len_hi=100_000, strand=True, thick=True))):
    iv = S.gen_intervals(cf, N, 20240002, device=dev, **kw)
    bufs = S.BinaryLiftBuffers(N, device=dev, rows_factor=1.6, thick="thick_start" in iv)
    st = torch.cuda.current_stream(dev)
    for _ in range(3): S.lift_intervals_dev(cf, iv, bufs, st)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        a.record(st); S.lift_intervals_dev(cf, iv, bufs, st); b.record(st)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
    rows, unm = bufs.totals.cpu().tolist()
    bq = 12 + (1 if "strand" in iv else 0) + (8 if "thick_start" in iv else 0)
    bo = 16 + (8 if "thick_start" in iv else 0)
    alg = N * (bq + 4) + rows * bo
    out["binary_" + label] = dict(ms=ms, intervals_per_s=N / ms * 1e3, rows=rows, unmatched=unm, alg_bytes=alg,
                                  GBs=alg / ms / 1e6)
    del iv, bufs
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
