#!/usr/bin/env python
"""bench.py — lifted intervals/sec on B200 (kernel + end-to-end) beside the CPU baseline.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port, all host threads)

Workload at every N: BASELINE.json configs[1] per GPU — synthetic sorted BED3 intervals
over the hg19 contigs of hg19ToHg38.over.chain, 1e8 per GPU (weak scaling: the
input is sharded by byte range, every rank owns one contiguous sorted shard, the chain
table is replicated, no collective on the data path).  One "step" = one pass of
liftBED (text -> lifted text + unmatched text) over the rank's shard.

  value  kernel-only: BED text already resident in HBM (swo_lift_bed_text_dev), CUDA events
  e2e    swo_lift_bed_text with PINNED HOST input: H2D + kernels + D2H inside the timed region
  other_configs  BASELINE configs[2..4] at their stated sizes per GPU share (kernel-only, device-
         resident inputs generated on the device; an end-to-end figure for configs[2]); under
         torchrun every rank runs its share and the slowest rank's time counts
  e2e.link_ceiling  a bare bidirectional pinned copy of the same byte counts on every rank at
         once, right before the end-to-end loop: what the box's host<->device path allows
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import gzip
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
RES = os.path.join(ROOT, "tests", "golden", "resources")
SEED_CFG2 = 20240002
METRIC = "lifted intervals/sec (BED text -> BED text)"
UNIT = "intervals/s"


def chain_bytes():
    return gzip.open(os.path.join(RES, "hg19ToHg38.over.chain.gz"), "rb").read()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_ffi import Oracle, build_oracle
    build_oracle()
    return Oracle()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is not None:
            self.p.terminate()  # exact PID we started
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from swiftover_b200 import ChainFile, shard
    from swiftover_b200 import synth as S

    rank, local, world = shard.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard.init("nccl", device_id=dev)  # rendezvous only: the data path has no collective
    n = args.n
    cf = ChainFile(text=chain_bytes(), devices=[local])

    # ---- workload: this rank's contiguous sorted shard (generation untimed) -----------------
    iv = S.gen_intervals(cf, n, SEED_CFG2, len_lo=1, len_hi=10_000, first=rank * n, device=dev, sort=True)
    text = S.bed_text(cf, iv, ncols=3)
    in_bytes = text.numel()
    bufs = S.TextLiftBuffers(in_bytes, device=dev, lifted_factor=1.35)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        shard.barrier()
        torch.cuda.synchronize(dev)

    # ---- kernel-only: text resident in HBM ---------------------------------------------------
    for _ in range(args.warmup):
        S.lift_text_dev(cf, text, bufs, stream)
    torch.cuda.synchronize(dev)
    sizes = bufs.sizes.cpu().tolist()
    assert sizes[5] == 0 and sizes[2] == n, f"kernel reported sizes {sizes}"
    assert sizes[0] <= bufs.lifted.numel(), "lifted buffer too small"
    out_bytes = sizes[0] + sizes[1]
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = cf.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    for a, b in evs:
        a.record(stream)
        S.lift_text_dev(cf, text, bufs, stream)
        b.record(stream)
    t_all1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    kern_ms_total = t_all0.elapsed_time(t_all1)
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    kernel_launches = cf.launch_count - launches0

    # ---- end to end: pinned host input -> pinned host outputs -------------------------------
    host_in = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    host_in.copy_(text)
    torch.cuda.synchronize(dev)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):  # first call also sizes the pinned arenas
        out = cf.lift_bed_text_ptr(host_in.data_ptr(), in_bytes)
    assert out.lifted_len == sizes[0] and out.unmatched_len == sizes[1] and out.n_records == n
    link_ms = link_ceiling_ms(torch, dev, host_in, sizes[0] + sizes[1], barrier)
    launches1 = cf.launch_count
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = cf.lift_bed_text_ptr(host_in.data_ptr(), in_bytes)
        # ordered concat across ranks = one exchange of piece lengths per step: a rank's piece goes to its
        # offset of the output (shard.piece_offsets); the bytes stay where the D2H copies put them
        my_off, totals = shard.piece_offsets([out.lifted_len, out.unmatched_len], device=dev)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_launches = cf.launch_count - launches1
    clocks = sampler.stop() if sampler else None
    del host_in, bufs
    torch.cuda.empty_cache()
    peak, peak_src = measured_peak()
    binary = None if args.no_binary else binary_kernel_line(cf, S, n, dev, peak, shard, world, rank)
    others = None if args.no_other_configs else other_configs(cf, S, shard, torch, dev, rank, world, peak, args)

    # ---- max over ranks -------------------------------------------------------------------------
    kern_ms_total, e2e_ms_total, link_ms = shard.max_over_ranks([kern_ms_total, e2e_s * 1e3, link_ms], device=dev)
    ms_per_step = kern_ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3)
    e2e_value = world * n / (e2e_ms_total * 1e-3 / e2e_steps)

    result = None
    if rank == 0:
        alg_bytes = in_bytes + out_bytes  # text in + lifted text + unmatched text (no index arrays are moved)
        med_ms = statistics.median(kern_ms)
        achieved = alg_bytes / (med_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("lift_bed_text_kernel")
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: {n} sorted BED3 intervals per GPU over hg19 contigs, "
                                   "hg19ToHg38.over.chain (53,950 blocks), text in -> lifted+unmatched text out",
                       "intervals_per_gpu": n, "input_bytes_per_gpu": in_bytes, "output_bytes_per_gpu": out_bytes,
                       "unmatched_frac": sizes[4] / n, "rows_per_interval": sizes[3] / n,
                       "l2": "inputs (GBs) far larger than the 126 MB L2; no flush needed",
                       "sharding": "byte-range shard per rank, chain table replicated, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes + 48 * max(1, e2e_launches // e2e_steps), "steps": e2e_steps,
                    "ms_per_step": e2e_ms_total / e2e_steps, "api": "swo_lift_bed_text (pinned host in/out)",
                    "concat": "pieces stay in each rank's pinned arena; their offsets in the ordered output come from one "
                              "all_gather of the piece lengths per step (inside the timed region)",
                    "link_ceiling": {"ms_per_step": link_ms, "value": world * n / (link_ms * 1e-3), "unit": UNIT,
                                     "what": "cudaMemcpyAsync H2D of the input bytes and D2H of the output bytes at the "
                                             "same time on two streams, pinned memory, all ranks at once (max over ranks)"},
                    "frac_of_link": link_ms / (e2e_ms_total / e2e_steps)},
            "gpu_launches": int(kernel_launches + e2e_launches),
            "roofline": {"bound": "hbm", "kernel": "lift_bed_text_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_interval": alg_bytes / n,
                         "kernel_ms_median": med_ms},
            "clocks": clocks,
        }
        if binary is not None:
            result["binary_kernel"] = binary
        if others is not None:
            result["other_configs"] = others
        if not args.no_cpu_baseline and world == 1:
            result["cpu_baseline"] = cpu_baseline_from_text(text, n, args)
    cf.close()
    if world > 1:
        dist.destroy_process_group()
    return result


def time_dev(torch, dev, fn, reps, warm=2):
    """Median device time (ms) of fn(), CUDA events on the launching stream."""
    st = torch.cuda.current_stream(dev)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(st)
        fn()
        b.record(st)
    torch.cuda.synchronize(dev)
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def link_ceiling_ms(torch, dev, host_in, out_bytes, barrier):
    """What the host<->device path allows for this step's byte counts: H2D of the input and D2H of
    the output at the same time (two streams, pinned memory), on every rank at once."""
    d_in = torch.empty(host_in.numel(), dtype=torch.uint8, device=dev)
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    h_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for it in range(3):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if it:  # first round warms the path up
            best = dt if best is None else min(best, dt)
    return best


def binary_kernel_line(cf, S, n, dev, peak, shard, world, rank):
    """Secondary figure: the binary query kernels (swo_lift_intervals_dev) on the same records —
    tcid/start/end int32 in, row_offset + 16-byte rows out (SURVEY.md §8d contract B).  Sorted
    batches run merge_count_kernel + scan + merge_write_kernel (csrc/kernels_merge.cuh)."""
    import torch
    iv = S.gen_intervals(cf, n, SEED_CFG2, len_lo=1, len_hi=10_000, first=rank * n, device=dev, sort=True)
    bufs = S.BinaryLiftBuffers(n, device=dev)
    st = torch.cuda.current_stream(dev)
    ms = time_dev(torch, dev, lambda: S.lift_intervals_dev(cf, iv, bufs, st), reps=10, warm=3)
    rows, unm = bufs.totals.cpu().tolist()
    alg = n * (12 + 4) + rows * 16
    ms, = shard.max_over_ranks([ms], device=dev)
    return {"kernel": "merge_count_kernel + scan_tiles_kernel1/2 + merge_write_kernel + expand_small_kernel + expand_wide_kernel (sorted batch)",
            "intervals_per_s": world * n / (ms * 1e-3), "ms": ms, "intervals_per_gpu": n, "rows_rank0": rows,
            "unmatched_rank0": unm, "algorithmic_bytes_per_gpu": alg, "achieved_GBs_per_gpu": alg / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak,
            "dram_traffic_note": "two passes: start and end are read twice and one byte per query travels between them: DRAM traffic = algorithmic + 9.5 B/query (ncu: 4.06 GB)"}


def other_configs(cf, S, shard, torch, dev, rank, world, peak, args):
    """BASELINE.json configs[2..4] at their stated sizes per GPU share, kernel-only (device-resident
    inputs generated on the device), every rank its own share, slowest rank counts."""
    import numpy as np
    from swiftover_b200 import ChainFile, _lib as L
    res, ms_keys = {}, []

    def put(key, d):
        res[key] = d
        ms_keys.append(key)

    # ---- configs[2]: 1e9 BED8 (strand + thick), shuffled, over 8 GPUs = 1.25e8 per GPU ------------
    n3 = args.n_cfg3
    iv = S.gen_intervals(cf, n3, 20240003, len_hi=100_000, first=rank * n3, strand=True, thick=True, device=dev)
    bb = S.BinaryLiftBuffers(n3, device=dev, rows_factor=1.6, thick=True)
    ms = time_dev(torch, dev, lambda: S.lift_intervals_dev(cf, iv, bb), reps=5)
    rows = int(bb.totals[0].item())
    alg = n3 * (21 + 4) + rows * 24
    put("cfg3_binary", {"workload": f"configs[2]: {n3} BED8 records per GPU (strand + thickStart/End), i.i.d. order, "
                                    "lengths log-uniform [1,1e5]", "n_per_gpu": n3, "rows_rank0": rows, "ms": ms,
                        "algorithmic_bytes_per_gpu": alg, "kernel": "lift_intervals_kernel<THICK>"})
    del bb
    text = S.bed_text(cf, iv, ncols=8, first_index=rank * n3)
    del iv
    torch.cuda.empty_cache()
    tb = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=1.6)
    ms = time_dev(torch, dev, lambda: S.lift_text_dev(cf, text, tb), reps=3, warm=1)
    sz = tb.sizes.cpu().tolist()
    assert sz[5] == 0 and sz[2] == n3 and sz[0] <= tb.lifted.numel(), f"cfg3 text sizes {sz}"
    alg = text.numel() + sz[0] + sz[1]
    put("cfg3_text", {"workload": f"configs[2]: {n3} BED8 lines per GPU, text in -> lifted + unmatched text out",
                      "n_per_gpu": n3, "in_bytes": text.numel(), "out_bytes": sz[0] + sz[1], "rows_rank0": sz[3], "ms": ms,
                      "algorithmic_bytes_per_gpu": alg, "kernel": "lift_bed_text_kernel"})
    del tb
    # end to end on the first n_e2e lines (pinned host in -> pinned host out)
    ne = min(n3, args.n_cfg3_e2e)
    cut = text.numel() if ne == n3 else int(torch.nonzero(text[: int(text.numel() * (ne / n3) * 1.05) + 4096] == 10)[ne - 1].item()) + 1
    host = torch.empty(cut, dtype=torch.uint8, pin_memory=True)
    host.copy_(text[:cut])
    del text
    torch.cuda.empty_cache()
    torch.cuda.synchronize(dev)
    out = cf.lift_bed_text_ptr(host.data_ptr(), cut)
    assert out.n_records == ne
    shard.barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        out = cf.lift_bed_text_ptr(host.data_ptr(), cut)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 2
    put("cfg3_e2e", {"workload": f"configs[2] end to end: {ne} BED8 lines per GPU, pinned host text in -> pinned host text out "
                                 "(swo_lift_bed_text)", "n_per_gpu": ne, "h2d_bytes": cut,
                     "d2h_bytes": int(out.lifted_len + out.unmatched_len), "ms": e2e_ms})
    del host

    # ---- configs[3]: 2e8 VCF records, REF check against a synthetic 2-bit destination genome, 1 % refchg ----
    n4 = args.n_cfg4 // world
    names = cf.contigNames
    Ls = np.array([cf.qContigSizes[nm] for nm in names], np.int64)
    off = np.zeros(len(names), np.int64)
    off[1:] = np.cumsum((Ls[:-1] + 31) // 32 * 32)
    tot = int(off[-1] + (Ls[-1] + 31) // 32 * 32)
    g = torch.Generator(device=dev)
    g.manual_seed(20240004)
    packed = torch.randint(0, 256, (tot // 4,), dtype=torch.uint8, device=dev, generator=g)
    cf.load_genome_2bit(packed.cpu().numpy(), off, Ls)
    del packed
    tsz = torch.tensor([cf.source_contig_size(t) for t in range(len(cf.sourceContigs))], dtype=torch.float64, device=dev)
    cdf = torch.cumsum(tsz, 0) / tsz.sum()
    first = rank * n4
    tc = torch.searchsorted(cdf, S.uniform01(20240004, 0, first, n4, dev)).clamp_(max=len(tsz) - 1)
    pos = torch.floor(S.uniform01(20240004, 1, first, n4, dev) * tsz[tc]).to(torch.int32)
    kind = S.uniform01(20240004, 2, first, n4, dev)  # 80 % SNV, 10 % insertion (REF 1 base), 10 % deletion (REF 2-20)
    rl = torch.where(kind < 0.9, 1, 2 + torch.floor(S.uniform01(20240004, 3, first, n4, dev) * 19).to(torch.int64)).to(torch.int32)
    tc = tc.to(torch.int32)

    def vcf_case(tc, pos, rl, key, what):
        """Kernel-only figure for one order of the records; returns what the host-array leg needs."""
        ro = torch.cumsum(rl.to(torch.int64), 0) - rl
        nref = int((ro[-1] + rl[-1]).item())
        ref = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (nref,), device=dev, generator=g)]
        oq = torch.empty(n4, dtype=torch.int32, device=dev)
        op, orl = torch.empty_like(oq), torch.empty_like(oq)
        fl = torch.empty(n4, dtype=torch.uint8, device=dev)
        oref = torch.empty_like(ref)

        def vcf(refb):
            rc = L.lib().swo_lift_vcf_points_dev(cf.handle, n4, tc.data_ptr(), pos.data_ptr(), rl.data_ptr(), ro.data_ptr(),
                                                 refb.data_ptr(), oq.data_ptr(), op.data_ptr(), fl.data_ptr(), orl.data_ptr(),
                                                 oref.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            assert rc == 0, L.lib().swo_last_error(cf.handle)
        vcf(ref)  # REF := what the destination genome holds there (the "source genome" agrees with it) ...
        torch.cuda.synchronize(dev)
        ref2 = oref.clone()
        differ = torch.arange(0, n4, 100, device=dev)  # ... except for a fixed 1 % of the records
        ref2[ro[differ]] = torch.where(ref2[ro[differ]] == 65, 67, 65).to(torch.uint8)
        ms = time_dev(torch, dev, lambda: vcf(ref2), reps=5)
        matched, refchg = int((fl & 1).sum().item()), int(((fl >> 1) & 1).sum().item())
        gather = 32 if key == "cfg4_vcf" else 0
        alg = n4 * (20 + 13 + gather) + 2 * nref + (0 if gather else n4 // 4)
        put(key, {"workload": f"configs[3]: {args.n_cfg4} VCF records / {world} GPU(s) (80 % SNV, 10 % insertion, 10 % deletion "
                              f"with REF 2-20 bases), {what}; point lift + REF check against a 0.78 GB 2-bit destination genome",
                  "n_per_gpu": n4, "matched_rank0": matched, "refchg_rank0": refchg, "refchg_frac_of_matched": refchg / max(1, matched),
                  "ms": ms, "algorithmic_bytes_per_gpu": alg,
                  "bytes_note": ("20 B in + 13 B out + REF bytes in and out + one 32-byte genome sector per record (SURVEY.md 8d); the "
                                 "DRAM moves 64 bytes per gathered sector: ncu 149 B/record, 3.8 TB/s" if gather else
                                 "20 B in + 13 B out + REF bytes in and out + 2 bits of genome per record (the gather is sequential)"),
                  "kernel": "lift_vcf_points_kernel", "scaling": "strong"})
        return ro, nref, ref2, fl

    # records in file order of a real VCF — sorted by (contig, position) — and, the worst case, in i.i.d. order
    order = torch.sort(pos.to(torch.int64) + (tc.to(torch.int64) << 32)).indices
    vcf_case(tc[order].contiguous(), pos[order].contiguous(), rl[order].contiguous(), "cfg4_vcf_sorted",
             "sorted by (contig, position) as a VCF file is")
    del order
    ro, nref, ref2, fl = vcf_case(tc, pos, rl, "cfg4_vcf", "i.i.d. order (every record a random gather)")
    # the same records through the host-array call (swo_lift_vcf_points): pinned host arrays in, pinned host arrays out,
    # chunks pipelined over three streams inside the call
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
    h_tc, h_pos, h_rl, h_ro, h_ref = pin(tc), pin(pos), pin(rl), pin(ro), pin(ref2)
    h_oq, h_op, h_orl = (torch.empty(n4, dtype=torch.int32, pin_memory=True) for _ in range(3))
    h_fl = torch.empty(n4, dtype=torch.uint8, pin_memory=True)
    h_oref = torch.empty(nref, dtype=torch.uint8, pin_memory=True)
    fl_dev = fl.clone()
    del tc, pos, rl, ro, ref2, fl
    torch.cuda.empty_cache()

    def vcf_host():
        rc = L.lib().swo_lift_vcf_points(cf.handle, n4, h_tc.data_ptr(), h_pos.data_ptr(), h_rl.data_ptr(), h_ro.data_ptr(),
                                         h_ref.data_ptr(), nref, h_oq.data_ptr(), h_op.data_ptr(), h_fl.data_ptr(),
                                         h_orl.data_ptr(), h_oref.data_ptr())
        assert rc == 0, L.lib().swo_last_error(cf.handle)
    vcf_host()
    assert torch.equal(h_fl, fl_dev.cpu()), "host-array VCF call disagrees with the device-resident one"
    shard.barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        vcf_host()
    shard.barrier()
    ms = (time.perf_counter() - t0) * 1e3 / 2
    put("cfg4_e2e", {"workload": f"configs[3] end to end: {n4} records per GPU, pinned host arrays in -> pinned host arrays out "
                                 "(swo_lift_vcf_points)", "n_per_gpu": n4, "h2d_bytes": n4 * 20 + nref, "d2h_bytes": n4 * 13 + nref,
                     "ms": ms})
    del h_tc, h_pos, h_rl, h_ro, h_ref, h_oq, h_op, h_orl, h_fl, h_oref, fl_dev
    torch.cuda.empty_cache()

    # ---- configs[4]: fragmented (CanFam4-like) chain, 1e8 long intervals / N GPUs, in batches (rows < 2^32 per call) ----
    cf5 = ChainFile(text=S.fragmented_chain(chain_bytes()), devices=[dev.index])
    n5 = args.n_cfg5 // world
    nb = max(1, (n5 + args.cfg5_batch - 1) // args.cfg5_batch)
    per = (n5 + nb - 1) // nb
    iv0 = S.gen_intervals(cf5, per, 20240005, len_lo=10_000, len_hi=1_000_000, first=rank * n5, device=dev)
    bb = S.BinaryLiftBuffers(per, device=dev, rows_factor=1.0)
    bb.cap = 0
    S.lift_intervals_dev(cf5, iv0, bb)  # sizing pass: nothing is written past capacity 0
    torch.cuda.synchronize(dev)
    fan = int(bb.totals[0].item()) / per
    bb = S.BinaryLiftBuffers(per, device=dev, rows_factor=fan * 1.03 + 0.5)
    ivs = [iv0] + [S.gen_intervals(cf5, per, 20240005, len_lo=10_000, len_hi=1_000_000, first=rank * n5 + b * per, device=dev)
                   for b in range(1, nb)]
    rows5 = []

    def run5():
        rows5.clear()
        for ivb in ivs:
            S.lift_intervals_dev(cf5, ivb, bb)
    ms = time_dev(torch, dev, run5, reps=2, warm=1)
    nrows = 0
    for ivb in ivs:  # row counts (untimed)
        S.lift_intervals_dev(cf5, ivb, bb)
        nrows += int(bb.totals[0].item())
        assert int(bb.totals[0].item()) <= bb.cap
    alg = nb * per * 16 + nrows * 16
    put("cfg5_binary", {"workload": f"configs[4]: fragmented chain ({cf5.num_blocks} blocks), {args.n_cfg5} intervals / {world} GPU(s), "
                                    f"lengths log-uniform [1e4,1e6], i.i.d. order, {nb} batches of {per} per step",
                        "n_per_gpu": nb * per, "rows_rank0": nrows, "rows_per_interval": nrows / (nb * per), "ms": ms,
                        "algorithmic_bytes_per_gpu": alg, "kernel": "lift_intervals_kernel", "scaling": "strong"})
    del bb, ivs
    torch.cuda.empty_cache()
    # text kernel on the first n_cfg5_text lines of the share
    nt = min(n5, args.n_cfg5_text)
    text = S.bed_text(cf5, {k: v[:nt].contiguous() for k, v in iv0.items()}, ncols=3)
    del iv0
    tb = S.TextLiftBuffers(text.numel(), device=dev, lifted_factor=fan * 30.0 / 24.0 + 2.0)
    ms = time_dev(torch, dev, lambda: S.lift_text_dev(cf5, text, tb), reps=2, warm=1)
    sz = tb.sizes.cpu().tolist()
    assert sz[5] == 0 and sz[2] == nt and sz[0] <= tb.lifted.numel(), f"cfg5 text sizes {sz}"
    alg = text.numel() + sz[0] + sz[1]
    put("cfg5_text", {"workload": f"configs[4], text: {nt} BED3 lines per GPU on the fragmented chain", "n_per_gpu": nt,
                      "in_bytes": text.numel(), "out_bytes": sz[0] + sz[1], "rows_rank0": sz[3], "ms": ms,
                      "algorithmic_bytes_per_gpu": alg, "kernel": "lift_bed_text_kernel"})
    del text, tb
    cf5.close()
    torch.cuda.empty_cache()

    mx = shard.max_over_ranks([res[k]["ms"] for k in ms_keys], device=dev)
    for k, m in zip(ms_keys, mx):
        d = res[k]
        d["ms"] = m
        d["n_total"] = d["n_per_gpu"] * world
        d["per_s"] = d["n_total"] / (m * 1e-3)
        if "algorithmic_bytes_per_gpu" in d:
            d["achieved_GBs_per_gpu"] = d["algorithmic_bytes_per_gpu"] / (m * 1e-3) / 1e9
            d["frac_of_hbm_peak"] = d["achieved_GBs_per_gpu"] / peak
    return res


def cpu_baseline_from_text(text, n, args):
    """Oracle (CPU port of the reference algorithm) on a strided sample of the SAME workload."""
    import torch
    data = text.cpu().numpy()  # sample = contiguous line-aligned slices spread over the whole input
    want = min(args.cpu_sample, n)
    import numpy as np
    nslices = 64
    per = max(1, int(len(data) * (want / n) / nslices))
    parts = []
    for k in range(nslices):
        a = int(len(data) * k / nslices)
        if a:
            a = a + int(np.argmax(data[a:a + 4096] == 10)) + 1
        b = min(len(data), a + per)
        nl = np.nonzero(data[b:b + 4096] == 10)[0]
        b = b + int(nl[0]) + 1 if len(nl) else len(data)
        parts.append(data[a:b])
    sample = np.concatenate(parts)
    return time_oracle(sample, args, kind_note=f"{nslices} line-aligned slices spread over the sorted input")


def time_oracle(sample, args, kind_note):
    import numpy as np
    oc = load_oracle().chain_parse(chain_bytes())
    threads = args.cpu_threads or os.cpu_count() or 1
    lines = int((sample == 10).sum())
    ptr = sample.ctypes.data
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        r = oc.lift_bed_ptr(ptr, sample.nbytes, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t0 = time.perf_counter()
    oc.lift_bed_ptr(ptr, sample.nbytes, 1)
    dt1 = time.perf_counter() - t0
    return {"value": lines / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{lines} lines ({sample.nbytes} bytes): {kind_note}; best of 2",
            "value_1core": lines / dt1,
            "note": "oracle/ = C restatement of swiftover's algorithm; the D reference cannot be built here "
                    "(no D toolchain, un-vendored deps) and is single-threaded as shipped"}


# ------------------------------------------------------------------------------------------
def chain_tsizes(text):
    """tSize of every source contig in first-appearance order, from the chain headers (chain.d:298)."""
    order, sizes = [], {}
    for ln in text.split(b"\n"):
        if ln.startswith(b"chain "):
            f = ln.split(b" ")
            if f[2] not in sizes:
                sizes[f[2]] = int(f[3])
                order.append(f[2])
    return order, [sizes[k] for k in order]


def gen_sample_cpu(n_total, n_sample, seed):
    """numpy version of swiftover_b200.synth.gen_intervals (same RNG, same distribution): the SAME sorted
    stream of n_total records the GPU arm lifts, of which 64 contiguous runs (n_sample records in all)
    spread over the stream are kept — same record density per contig as the full workload."""
    import numpy as np
    chain = chain_bytes()
    oc = load_oracle().chain_parse(chain)
    names, tsz = chain_tsizes(chain)
    assert [oc.tcid(nm) for nm in names] == list(range(len(names)))
    sizes = np.array(tsz, np.float64)

    def u01(stream, n):
        with np.errstate(over="ignore"):
            x = np.arange(n, dtype=np.uint64) ^ np.uint64(seed ^ (stream << 56))
            z = x + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)

    n = n_total
    cdf = np.cumsum(sizes) / sizes.sum()
    tcid = np.minimum(np.searchsorted(cdf, u01(0, n)), len(sizes) - 1).astype(np.int32)
    ln = np.maximum(1, np.floor(np.exp(u01(1, n) * np.log(10_000.0)))).astype(np.int64)
    sz = sizes[tcid].astype(np.int64)
    ln = np.minimum(ln, np.maximum(sz, 1))
    start = np.floor(u01(2, n) * np.maximum(sz - ln, 0)).astype(np.int64)
    del sz
    # sort key = (contig, start, end) in one word: 7 + 28 + 14 bits (len <= 1e4); only the 64 runs are sorted
    key = (tcid.astype(np.uint64) << np.uint64(42)) | (start.astype(np.uint64) << np.uint64(14)) | ln.astype(np.uint64)
    del tcid, start, ln
    runs, per = 64, max(1, min(n_sample, n) // 64)
    lo = [int(n * k / runs) for k in range(runs)]
    kth = sorted({i for a in lo for i in (a, min(n - 1, a + per - 1))})
    key.partition(kth)
    keep = np.concatenate([np.sort(key[a: a + per]) for a in lo])
    del key
    k_tc = (keep >> np.uint64(42)).astype(np.int32)
    k_st = ((keep >> np.uint64(14)) & np.uint64((1 << 28) - 1)).astype(np.int64)
    k_ln = (keep & np.uint64((1 << 14) - 1)).astype(np.int64)
    return oc, oc.format_bed3(k_tc, k_st, k_st + k_ln), len(keep)


def run_reference(args):
    """CPU arm: the oracle port on all host threads over a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    oc, sample, n = gen_sample_cpu(args.n, min(args.cpu_sample, args.n), SEED_CFG2)
    threads = args.cpu_threads or os.cpu_count() or 1
    ptr = sample.ctypes.data
    for _ in range(args.warmup):
        oc.lift_bed_ptr(ptr, sample.nbytes, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = oc.lift_bed_ptr(ptr, sample.nbytes, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    t0 = time.perf_counter()
    import numpy as np
    cut8 = int(np.flatnonzero(sample[: sample.nbytes // 8] == 10)[-1]) + 1  # the reference as shipped is single-threaded:
    lines8 = int((sample[:cut8] == 10).sum())                               # 1/8 of the sample on one core
    t0 = time.perf_counter()
    oc.lift_bed_ptr(ptr, cut8, 1)
    value_1core = lines8 / (time.perf_counter() - t0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample_note = (f"{n} records = 64 contiguous runs spread over the same sorted stream of {args.n} BED3 intervals the GPU arm "
                   f"lifts (same generator/seed; {sample.nbytes} bytes of text), oracle port, {threads} threads over "
                   "line-aligned splits")
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1] (bounded sample): {n} sorted BED3 intervals over hg19 contigs, "
                                   "hg19ToHg38.over.chain, text in -> text out"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_note,
                             "value_1core": value_1core},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--n", type=int, default=100_000_000, help="intervals per GPU (configs[1]: 1e8)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=16_000_000, help="lines in the CPU baseline sample")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-binary", action="store_true", help="skip the secondary binary-kernel measurement")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs[2..4]")
    ap.add_argument("--n-cfg3", type=int, default=125_000_000, help="configs[2]: BED8 records per GPU (1e9 over 8 GPUs)")
    ap.add_argument("--n-cfg3-e2e", type=int, default=50_000_000, help="configs[2]: lines of the end-to-end leg per GPU")
    ap.add_argument("--n-cfg4", type=int, default=200_000_000, help="configs[3]: VCF records in all")
    ap.add_argument("--n-cfg5", type=int, default=100_000_000, help="configs[4]: long intervals in all")
    ap.add_argument("--cfg5-batch", type=int, default=25_000_000, help="configs[4]: intervals per call (rows < 2^32)")
    ap.add_argument("--n-cfg5-text", type=int, default=12_500_000, help="configs[4]: lines of the text-kernel leg per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else max(args.warmup, 1)
    res = run_reference(args) if args.impl == "reference" else run_cuda(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
