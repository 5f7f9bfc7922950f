#!/usr/bin/env python
"""Box probe: zero-copy-input text kernel running concurrently with a large D2H memcpy."""
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
h_in = torch.empty(nb + 64, dtype=torch.uint8, pin_memory=True); h_in[:nb].copy_(text)
d_l = torch.empty(int(nb * 1.35) + 65536, dtype=torch.uint8, device=dev)
d_u = torch.empty(nb // 2, dtype=torch.uint8, device=dev)
h_out = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
d_src = torch.empty(nb, dtype=torch.uint8, device=dev)
sizes = torch.zeros(6, dtype=torch.int64, device=dev)
lib = L.lib()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def kern():
    rc = lib.swo_lift_bed_text_dev(cf.handle, h_in.data_ptr(), nb, d_l.data_ptr(), d_l.numel(), d_u.data_ptr(), d_u.numel(), sizes.data_ptr(), s1.cuda_stream)
    assert rc == 0
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_src, non_blocking=True)
out = {}
for label, fn in (("kernel_zc_alone", lambda: kern()), ("d2h_alone", lambda: d2h()), ("both", lambda: (kern(), d2h()))):
    for _ in range(2): fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    out[label] = min(ts) * 1e3
print(json.dumps(out))
