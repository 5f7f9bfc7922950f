#!/usr/bin/env python
"""Box probe: end-to-end swo_lift_bed_text time vs pipeline chunk size (not a bench line)."""
import gzip, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = int(os.environ.get("PROBE_N", 100_000_000))
iv = S.gen_intervals(cf, N, 20240002, device=dev, sort=True)
text = S.bed_text(cf, iv, ncols=3)
host = torch.empty(text.numel(), dtype=torch.uint8, pin_memory=True)
host.copy_(text); torch.cuda.synchronize()
del iv, text
torch.cuda.empty_cache()
out = {}
if os.environ.get("D2H"):
    cf.set_option("d2h_streams", int(os.environ["D2H"]))
for mib in [int(x) for x in os.environ.get("CHUNKS", "4,8,16,32,64").split(",")]:
    cf.set_option("chunk_bytes", mib << 20)
    for _ in range(2): cf.lift_bed_text_ptr(host.data_ptr(), host.numel())
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); r = cf.lift_bed_text_ptr(host.data_ptr(), host.numel()); ts.append(time.perf_counter() - t0)
    out[f"{mib}MiB"] = dict(ms_min=min(ts) * 1e3, ms_med=sorted(ts)[len(ts) // 2] * 1e3, G_per_s=N / min(ts) / 1e9)
print(json.dumps(out, indent=1))
