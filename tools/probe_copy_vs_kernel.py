#!/usr/bin/env python
"""Box probe: do H2D/D2H memcpys slow down while the text kernel runs (device-resident input)?"""
import gzip, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S, _lib as L
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = 100_000_000
iv = S.gen_intervals(cf, N, 20240002, device=dev, sort=True)
text = S.bed_text(cf, iv, ncols=3)
nb = text.numel()
d_l = torch.empty(int(nb * 1.35) + 65536, dtype=torch.uint8, device=dev)
d_u = torch.empty(nb // 2, dtype=torch.uint8, device=dev)
sizes = torch.zeros(6, dtype=torch.int64, device=dev)
h_a = torch.empty(nb, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(nb, dtype=torch.uint8, device=dev); d_b = torch.empty(nb, dtype=torch.uint8, device=dev)
lib = L.lib()
s1, s2, s3 = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def kern(reps):
    for _ in range(reps):
        rc = lib.swo_lift_bed_text_dev(cf.handle, text.data_ptr(), nb, d_l.data_ptr(), d_l.numel(), d_u.data_ptr(), d_u.numel(), sizes.data_ptr(), s1.cuda_stream)
        assert rc == 0
def copies():
    with torch.cuda.stream(s2): d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s3): h_b.copy_(d_b, non_blocking=True)
def timed(fn):
    for _ in range(2): fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3
out = dict(kernel_x3=timed(lambda: kern(3)), copies_bidir=timed(copies), both=timed(lambda: (kern(3), copies())))
print(json.dumps(out))
