#!/usr/bin/env python
"""Box probe: time the text kernel on consecutive slices of the sorted workload to find slow regions."""
import gzip, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = int(os.environ.get("PROBE_N", 10_000_000))
NS = int(os.environ.get("SLICES", 40))
iv = S.gen_intervals(cf, N, 20240002, device=dev, sort=True)
text = S.bed_text(cf, iv, ncols=3)
nb = text.numel()
st = torch.cuda.current_stream(dev)
res = []
ONLY = os.environ.get('ONLY')
for k in ([int(ONLY)] if ONLY else range(NS)):
    a = nb * k // NS; b = nb * (k + 1) // NS
    a = (a + 15) & ~15
    sl = text[a:b]
    # align the slice start to a line start by blanking the partial first line with a comment marker
    sl = sl.clone()
    nl = int((sl[:200] == 10).nonzero()[0].item())
    sl[:nl] = ord('#') if k else sl[:nl]
    bufs = S.TextLiftBuffers(sl.numel() + 64, device=dev, lifted_factor=2.0)
    pad = torch.zeros(sl.numel() + 64, dtype=torch.uint8, device=dev); pad[:sl.numel()] = sl
    sl = pad[:b - a]
    for _ in range(2): S.lift_text_dev(cf, sl, bufs, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3): S.lift_text_dev(cf, sl, bufs, st)
    e1.record(st); torch.cuda.synchronize()
    sz = bufs.sizes.cpu().tolist()
    first = bytes(sl[nl + 1: nl + 40].cpu().numpy()).split(b"\n")[0].decode()
    res.append((k, round(e0.elapsed_time(e1) / 3, 3), sz[2], sz[3], sz[4], sz[5] != 0, first))
for r in res: print(*r)
