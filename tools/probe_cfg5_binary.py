#!/usr/bin/env python
"""Box probe (not a bench line): BASELINE configs[4] through the binary path — the fragmented chain,
PROBE_N (default 1e7) long intervals in i.i.d. order — for ncu captures of lift_intervals_kernel and
expand_wide_kernel on heavy fan-out (~38 rows per interval)."""
import gzip, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S

dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=S.fragmented_chain(chain))
n = int(float(os.environ.get("PROBE_N", 10_000_000)))
iv = S.gen_intervals(cf, n, 20240005, len_lo=10_000, len_hi=1_000_000, device=dev)
bb = S.BinaryLiftBuffers(n, device=dev, rows_factor=1.0)
bb.cap = 0
S.lift_intervals_dev(cf, iv, bb)  # sizing pass
torch.cuda.synchronize()
fan = int(bb.totals[0].item()) / n
bb = S.BinaryLiftBuffers(n, device=dev, rows_factor=fan * 1.03 + 0.5)
st = torch.cuda.current_stream(dev)
for _ in range(2): S.lift_intervals_dev(cf, iv, bb, st)
torch.cuda.synchronize()
reps = int(os.environ.get("PROBE_REPS", 5))
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, b in ev:
    a.record(st); S.lift_intervals_dev(cf, iv, bb, st); b.record(st)
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
rows = int(bb.totals[0].item())
print(json.dumps(dict(blocks=cf.num_blocks, queries=n, rows=rows, fanout=rows / n, ms=round(ms, 3),
                      G_rows_per_s=round(rows / ms / 1e6, 2), GBs=round((n * 16 + rows * 16) / ms / 1e6, 1))))
