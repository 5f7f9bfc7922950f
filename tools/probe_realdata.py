#!/usr/bin/env python
"""Box probe: text kernel on the reference's real fixtures replicated to a few hundred MB
(long intervals, fan-out 2.3 rows per record on hg19_col123.bed; BED12 on chr1_hg19.bed)."""
import gzip, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
rd = lambda n: gzip.open(os.path.join(ROOT, "tests/golden/resources", n + ".gz"), "rb").read()
cf = ChainFile(text=rd("hg19ToHg38.over.chain"))
if os.environ.get("TILE"):
    cf.set_option("tile_bytes", int(os.environ["TILE"]))
out = {}
SETS = (("hg19_col123.bed", 200), ("chr1_hg19.bed", 300))
if os.environ.get("ONLY"):
    SETS = ((os.environ["ONLY"], int(os.environ.get("REPS", 50))),)
for name, reps in SETS:
    data = rd(name) * reps
    text = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    n = data.count(b"\n")
    bufs = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=4.0)
    st = torch.cuda.current_stream(dev)
    for _ in range(2): S.lift_text_dev(cf, text, bufs, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(3): S.lift_text_dev(cf, text, bufs, st)
    b.record(st); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    sz = bufs.sizes.cpu().tolist()
    out[name] = dict(lines=n, in_MB=text.numel() / 1e6, out_MB=(sz[0] + sz[1]) / 1e6, rows=sz[3], ms=ms,
                     M_lines_per_s=n / ms / 1e3, GBs=(text.numel() + sz[0] + sz[1]) / ms / 1e6)
print(json.dumps(out, indent=1))
