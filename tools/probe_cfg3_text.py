#!/usr/bin/env python
"""Box probe: text kernel on cfg3-like input (BED8, strand + thick, i.i.d. order)."""
import gzip, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
n = int(os.environ.get("PROBE_N", 20_000_000))
kw = dict(len_hi=int(os.environ.get("LEN_HI", 100_000)), strand=True, thick=os.environ.get("THICK", "1") == "1")
iv = S.gen_intervals(cf, n, 20240003, device=dev, sort=os.environ.get("SORT", "0") == "1", **kw)
text = S.bed_text(cf, iv, ncols=8 if kw["thick"] else 6)
bufs = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=2.0)
st = torch.cuda.current_stream(dev)
for _ in range(3): S.lift_text_dev(cf, text, bufs, st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st)
for _ in range(3): S.lift_text_dev(cf, text, bufs, st)
b.record(st); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
sz = bufs.sizes.cpu().tolist()
print(json.dumps(dict(n=n, ms=ms, G_lines_per_s=n / ms / 1e6, rows=sz[3], unm=sz[4], **{k: str(v) for k, v in kw.items()})))
