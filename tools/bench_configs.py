#!/usr/bin/env python
"""Kernel-only throughput of the other BASELINE.json configs on one B200 (not bench lines; the
numbers go into DESIGN.md §6).  Inputs are generated on the device and stay in HBM.
  cfg3  BED8 (strand + thick), i.i.d. order, text kernel and binary kernel
  cfg4  VCF records, point lift + REF check against a synthetic 2-bit genome (device-resident
        arrays are not exposed for this kernel: host arrays in/out, so the figure is end to end)
  cfg5  fragmented chain (~540 k blocks), long intervals, binary kernel
"""
import gzip, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
out = {}

def time_dev(fn, reps=5):
    st = torch.cuda.current_stream(dev)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): fn()
    b.record(st); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

cf = ChainFile(text=chain)
# ---- cfg3 ------------------------------------------------------------------------------
n = 50_000_000
iv = S.gen_intervals(cf, n, 20240003, len_hi=100_000, strand=True, thick=True, device=dev)
text = S.bed_text(cf, iv, ncols=8)
bufs = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=2.0)
ms = time_dev(lambda: S.lift_text_dev(cf, text, bufs))
sz = bufs.sizes.cpu().tolist()
out["cfg3_text_bed8_shuffled"] = dict(lines=n, in_bytes=text.numel(), out_bytes=sz[0] + sz[1], rows=sz[3], ms=ms,
                                      G_lines_per_s=n / ms / 1e6, GBs=(text.numel() + sz[0] + sz[1]) / ms / 1e6)
bb = S.BinaryLiftBuffers(n, device=dev, rows_factor=1.6, thick=True)
ms = time_dev(lambda: S.lift_intervals_dev(cf, iv, bb))
rows = int(bb.totals[0].item())
out["cfg3_binary_bed8_shuffled"] = dict(queries=n, rows=rows, ms=ms, G_per_s=n / ms / 1e6,
                                        GBs=(n * (21 + 4) + rows * 24) / ms / 1e6)
del iv, text, bufs, bb
torch.cuda.empty_cache()
# ---- cfg4 ------------------------------------------------------------------------------
rng = np.random.default_rng(20240004)
n = 20_000_000
names = cf.contigNames
Ls = np.array([cf.qContigSizes[nm] for nm in names], np.int64)
off = np.zeros(len(names), np.int64); off[1:] = np.cumsum((Ls[:-1] + 31) // 32 * 32)
tot = int(off[-1] + (Ls[-1] + 31) // 32 * 32)
packed = rng.integers(0, 256, tot // 4, dtype=np.uint8)
cf.load_genome_2bit(packed, off, Ls)
tsz = np.array([cf.source_contig_size(t) for t in range(len(cf.sourceContigs))], np.float64)
tc = rng.choice(len(tsz), n, p=tsz / tsz.sum()).astype(np.int32)
pos = np.floor(rng.random(n) * tsz[tc]).astype(np.int32)
rl = np.where(rng.random(n) < 0.9, 1, rng.integers(2, 21, n)).astype(np.int32)
ro = np.zeros(n, np.uint64); ro[1:] = np.cumsum(rl[:-1].astype(np.uint64))
refb = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rl.sum()))
cf.lift_vcf_points(tc, pos, rl, ro, refb)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); r = cf.lift_vcf_points(tc, pos, rl, ro, refb); ts.append(time.perf_counter() - t0)
out["cfg4_vcf_points_host_arrays"] = dict(records=n, ms=min(ts) * 1e3, M_records_per_s=n / min(ts) / 1e6,
                                          matched=int((r["flags"] & 1).sum()), refchg=int((r["flags"] >> 1).sum()),
                                          note="pageable numpy arrays in/out: dominated by host copies")
# kernel only: the same records resident on the device
from swiftover_b200 import _lib as L
dt = lambda a: torch.from_numpy(a).to(dev)
d_tc, d_pos, d_rl, d_ro, d_ref = dt(tc), dt(pos), dt(rl), dt(ro.astype(np.int64)), dt(refb)
d_oq = torch.empty(n, dtype=torch.int32, device=dev); d_op = torch.empty_like(d_oq); d_orl = torch.empty_like(d_oq)
d_fl = torch.empty(n, dtype=torch.uint8, device=dev); d_oref = torch.empty_like(d_ref)
def vcf_dev():
    rc = L.lib().swo_lift_vcf_points_dev(cf.handle, n, d_tc.data_ptr(), d_pos.data_ptr(), d_rl.data_ptr(), d_ro.data_ptr(),
                                         d_ref.data_ptr(), d_oq.data_ptr(), d_op.data_ptr(), d_fl.data_ptr(),
                                         d_orl.data_ptr(), d_oref.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    assert rc == 0
ms = time_dev(vcf_dev)
assert np.array_equal(d_fl.cpu().numpy(), r["flags"]) and np.array_equal(d_op.cpu().numpy(), r["pos"])
out["cfg4_vcf_points_kernel_only"] = dict(records=n, ms=ms, G_records_per_s=n / ms / 1e6,
                                          GBs=(n * (20 + 13) + 2 * len(refb) + n * 32) / ms / 1e6)
cf.close()
# ---- cfg5 ------------------------------------------------------------------------------
from test_gpu_fullsize import fragmented_chain
cf5 = ChainFile(text=fragmented_chain())
n = 2_000_000
iv = S.gen_intervals(cf5, n, 20240005, len_lo=10_000, len_hi=1_000_000, device=dev)
bb = S.BinaryLiftBuffers(n, device=dev, rows_factor=1.0); bb.cap = 0
S.lift_intervals_dev(cf5, iv, bb); torch.cuda.synchronize()
nr = int(bb.totals[0].item())
bb = S.BinaryLiftBuffers(n, device=dev, rows_factor=nr / n + 0.01)
ms = time_dev(lambda: S.lift_intervals_dev(cf5, iv, bb), reps=3)
out["cfg5_binary_fragmented_chain"] = dict(blocks=cf5.num_blocks, queries=n, rows=nr, fanout=nr / n, ms=ms,
                                           M_queries_per_s=n / ms / 1e3, G_rows_per_s=nr / ms / 1e6,
                                           GBs=(n * 16 + nr * 16) / ms / 1e6)
cf5.close()
print(json.dumps(out, indent=1))
