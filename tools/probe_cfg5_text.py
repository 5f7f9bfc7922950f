#!/usr/bin/env python
"""Box probe: text kernel on cfg5-like input (fragmented chain, long intervals, fan-out in the tens)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from swiftover_b200 import ChainFile, synth as S
from test_gpu_fullsize import fragmented_chain
dev = torch.device("cuda", 0)
cf = ChainFile(text=fragmented_chain())
if os.environ.get("TILE"): cf.set_option("tile_bytes", int(os.environ["TILE"]))
n = int(os.environ.get("PROBE_N", 1_000_000))
iv = S.gen_intervals(cf, n, 20240005, len_lo=10_000, len_hi=1_000_000, device=dev)
text = S.bed_text(cf, iv, ncols=3)
bufs = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=80.0)
st = torch.cuda.current_stream(dev)
for _ in range(2): S.lift_text_dev(cf, text, bufs, st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st)
for _ in range(3): S.lift_text_dev(cf, text, bufs, st)
b.record(st); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
sz = bufs.sizes.cpu().tolist()
assert sz[0] <= bufs.lifted.numel()
print(json.dumps(dict(n=n, ms=ms, M_lines_per_s=n / ms / 1e3, rows=sz[3], out_GB=sz[0] / 1e9, out_GBs=sz[0] / ms / 1e6)))
