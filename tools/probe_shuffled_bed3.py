#!/usr/bin/env python
"""Box probe: text kernel on cfg2's records in i.i.d. (unsorted) order."""
import gzip, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
cf = ChainFile(text=gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read())
n = int(os.environ.get("PROBE_N", 50_000_000))
out = {}
for label, srt in (("sorted", True), ("shuffled", False)):
    iv = S.gen_intervals(cf, n, 20240002, device=dev, sort=srt)
    text = S.bed_text(cf, iv, ncols=3)
    bufs = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=1.5)
    st = torch.cuda.current_stream(dev)
    for _ in range(3): S.lift_text_dev(cf, text, bufs, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(5): S.lift_text_dev(cf, text, bufs, st)
    b.record(st); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out[label] = dict(ms=ms, G_lines_per_s=n / ms / 1e6)
    del iv, text, bufs
print(json.dumps(out))
