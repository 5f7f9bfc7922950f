#!/usr/bin/env python
"""Box probe: the text kernel reading / writing PINNED HOST memory directly (UVA zero-copy)."""
import gzip, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S, _lib as L
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = int(os.environ.get("PROBE_N", 100_000_000))
iv = S.gen_intervals(cf, N, 20240002, device=dev, sort=True)
text = S.bed_text(cf, iv, ncols=3)
nb = text.numel()
h_in = torch.empty(nb + 64, dtype=torch.uint8, pin_memory=True); h_in[:nb].copy_(text)
h_l = torch.empty(int(nb * 1.35) + 65536, dtype=torch.uint8, pin_memory=True)
h_u = torch.empty(nb // 2, dtype=torch.uint8, pin_memory=True)
d_l = torch.empty(int(nb * 1.35) + 65536, dtype=torch.uint8, device=dev)
d_u = torch.empty(nb // 2, dtype=torch.uint8, device=dev)
sizes = torch.zeros(6, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
lib = L.lib()
st = torch.cuda.current_stream(dev).cuda_stream
def run(inp, ol, ou):
    rc = lib.swo_lift_bed_text_dev(cf.handle, inp, nb, ol.data_ptr(), ol.numel(), ou.data_ptr(), ou.numel(), sizes.data_ptr(), st)
    assert rc == 0, lib.swo_last_error(cf.handle)
out = {}
ref = None
for label, inp, ol, ou in (("dev->dev", text.data_ptr(), d_l, d_u), ("host->dev", h_in.data_ptr(), d_l, d_u),
                           ("dev->host", text.data_ptr(), h_l, h_u), ("host->host", h_in.data_ptr(), h_l, h_u)):
    for _ in range(2): run(inp, ol, ou)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); run(inp, ol, ou); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    sz = sizes.cpu().tolist()
    out[label] = dict(ms=min(ts) * 1e3, G_per_s=N / min(ts) / 1e9, sizes=sz[:2])
    chk = (int(ol[:sz[0]].to(torch.int64).sum().item()) if ol.is_cuda else int(ol[:sz[0]].numpy().astype("int64").sum()))
    out[label]["checksum"] = chk
print(json.dumps(out, indent=1))
