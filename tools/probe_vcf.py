#!/usr/bin/env python
"""Box probe (not a bench line): lift_vcf_points_kernel alone on BASELINE configs[3]-shaped records
(PROBE_N, default 1e8; i.i.d. positions, 80 % SNV / 20 % indel REF, synthetic 2-bit genome)."""
import gzip, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S, _lib as L

dev = torch.device("cuda", 0)
cf = ChainFile(text=gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read())
n = int(float(os.environ.get("PROBE_N", 100_000_000)))
names = cf.contigNames
Ls = np.array([cf.qContigSizes[nm] for nm in names], np.int64)
off = np.zeros(len(names), np.int64)
off[1:] = np.cumsum((Ls[:-1] + 31) // 32 * 32)
tot = int(off[-1] + (Ls[-1] + 31) // 32 * 32)
g = torch.Generator(device=dev); g.manual_seed(20240004)
packed = torch.randint(0, 256, (tot // 4,), dtype=torch.uint8, device=dev, generator=g)
cf.load_genome_2bit(packed.cpu().numpy(), off, Ls); del packed
tsz = torch.tensor([cf.source_contig_size(t) for t in range(len(cf.sourceContigs))], dtype=torch.float64, device=dev)
cdf = torch.cumsum(tsz, 0) / tsz.sum()
tc = torch.searchsorted(cdf, S.uniform01(20240004, 0, 0, n, dev)).clamp_(max=len(tsz) - 1)
pos = torch.floor(S.uniform01(20240004, 1, 0, n, dev) * tsz[tc]).to(torch.int32)
kind = S.uniform01(20240004, 2, 0, n, dev)
rl = torch.where(kind < 0.9, 1, 2 + torch.floor(S.uniform01(20240004, 3, 0, n, dev) * 19).to(torch.int64)).to(torch.int32)
tc = tc.to(torch.int32)
ro = torch.cumsum(rl.to(torch.int64), 0) - rl
nref = int((ro[-1] + rl[-1]).item())
ref = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (nref,), device=dev, generator=g)]
oq = torch.empty(n, dtype=torch.int32, device=dev); op, orl = torch.empty_like(oq), torch.empty_like(oq)
fl = torch.empty(n, dtype=torch.uint8, device=dev); oref = torch.empty_like(ref)
st = torch.cuda.current_stream(dev)
def run():
    rc = L.lib().swo_lift_vcf_points_dev(cf.handle, n, tc.data_ptr(), pos.data_ptr(), rl.data_ptr(), ro.data_ptr(), ref.data_ptr(),
                                         oq.data_ptr(), op.data_ptr(), fl.data_ptr(), orl.data_ptr(), oref.data_ptr(), st.cuda_stream)
    assert rc == 0
for _ in range(2): run()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
for a, b in ev:
    a.record(st); run(); b.record(st)
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
print(json.dumps(dict(n=n, ms=round(ms, 3), G_records_per_s=round(n / ms / 1e6, 2), matched=int((fl & 1).sum().item()),
                      GBs=round((n * 65 + 2 * nref) / ms / 1e6, 1))))
