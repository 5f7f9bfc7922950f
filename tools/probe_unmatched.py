#!/usr/bin/env python
"""Box probe: kernel time on an all-unmatched sorted input (centromere gap) vs the normal mix."""
import gzip, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = int(os.environ.get("PROBE_N", 10_000_000))
out = {}
for label in ("mix", "gap"):
    iv = S.gen_intervals(cf, N, 20240002, device=dev, sort=False)
    if label == "gap":
        t = cf.tcid("chr1")
        iv["tcid"] = torch.full_like(iv["tcid"], t)
        ln = (iv["end"] - iv["start"]).clamp(max=5000)
        iv["start"] = (122_000_000 + (torch.arange(N, device=dev, dtype=torch.int64) * 18_000_000 // N)).to(torch.int32)
        iv["end"] = iv["start"] + ln
    else:
        iv = S.gen_intervals(cf, N, 20240002, device=dev, sort=True)
    text = S.bed_text(cf, iv, ncols=3)
    bufs = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=1.35)
    st = torch.cuda.current_stream(dev)
    for _ in range(3): S.lift_text_dev(cf, text, bufs, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(5): S.lift_text_dev(cf, text, bufs, st)
    b.record(st); torch.cuda.synchronize()
    out[label] = dict(ms=a.elapsed_time(b) / 5, sizes=bufs.sizes.cpu().tolist(), bytes=text.numel())
print(json.dumps(out))
