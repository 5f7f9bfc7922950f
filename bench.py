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
    launches1 = cf.launch_count
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = cf.lift_bed_text_ptr(host_in.data_ptr(), in_bytes)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_launches = cf.launch_count - launches1
    clocks = sampler.stop() if sampler else None

    # ---- max over ranks -------------------------------------------------------------------------
    kern_ms_total, e2e_ms_total = shard.max_over_ranks([kern_ms_total, e2e_s * 1e3], device=dev)
    ms_per_step = kern_ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3)
    e2e_value = world * n / (e2e_ms_total * 1e-3 / e2e_steps)

    result = None
    if rank == 0:
        peak, peak_src = measured_peak()
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
                    "ms_per_step": e2e_ms_total / e2e_steps, "api": "swo_lift_bed_text (pinned host in/out)"},
            "gpu_launches": int(kernel_launches + e2e_launches),
            "roofline": {"bound": "hbm", "kernel": "lift_bed_text_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_interval": alg_bytes / n,
                         "kernel_ms_median": med_ms},
            "clocks": clocks,
        }
        if world == 1 and not args.no_binary:
            result["binary_kernel"] = binary_kernel_line(cf, S, n, dev, peak)
        if not args.no_cpu_baseline and world == 1:
            result["cpu_baseline"] = cpu_baseline_from_text(text, n, args)
    cf.close()
    if world > 1:
        dist.destroy_process_group()
    return result


def binary_kernel_line(cf, S, n, dev, peak):
    """Secondary figure: the binary query kernel (swo_lift_intervals_dev) on the same records —
    tcid/start/end int32 in, row_offset + 16-byte rows out (SURVEY.md §8d contract B)."""
    import torch
    iv = S.gen_intervals(cf, n, SEED_CFG2, len_lo=1, len_hi=10_000, device=dev, sort=True)
    bufs = S.BinaryLiftBuffers(n, device=dev)
    st = torch.cuda.current_stream(dev)
    for _ in range(3):
        S.lift_intervals_dev(cf, iv, bufs, st)
    torch.cuda.synchronize(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        a.record(st)
        S.lift_intervals_dev(cf, iv, bufs, st)
        b.record(st)
    torch.cuda.synchronize(dev)
    ms = statistics.median(a.elapsed_time(b) for a, b in evs)
    rows, unm = bufs.totals.cpu().tolist()
    alg = n * (12 + 4) + rows * 16
    return {"kernel": "lift_intervals_kernel", "intervals_per_s": n / (ms * 1e-3), "ms": ms, "rows": rows,
            "unmatched": unm, "algorithmic_bytes": alg, "achieved_GBs": alg / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak}


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
def gen_sample_cpu(n, seed):
    """numpy version of swiftover_b200.synth.gen_intervals (same RNG, same distribution), sorted."""
    import numpy as np
    oc = load_oracle().chain_parse(chain_bytes())
    from swiftover_b200 import ChainFile
    cf = ChainFile(text=chain_bytes(), inspect_only=True)  # host chain parser only: contig sizes
    sizes = np.array([cf.source_contig_size(t) for t in range(len(cf.sourceContigs))], np.float64)
    cf.close()

    def u01(stream):
        with np.errstate(over="ignore"):
            x = np.arange(n, dtype=np.uint64) ^ np.uint64(seed ^ (stream << 56))
            z = x + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)

    cdf = np.cumsum(sizes) / sizes.sum()
    tcid = np.minimum(np.searchsorted(cdf, u01(0)), len(sizes) - 1)
    ln = np.maximum(1, np.floor(np.exp(u01(1) * np.log(10_000.0)))).astype(np.int64)
    sz = sizes[tcid].astype(np.int64)
    ln = np.minimum(ln, np.maximum(sz, 1))
    start = np.floor(u01(2) * np.maximum(sz - ln, 0)).astype(np.int64)
    end = start + ln
    order = np.lexsort((end, start, tcid))
    return oc, oc.format_bed3(tcid[order], start[order], end[order])


def run_reference(args):
    """CPU arm: the oracle port on all host threads over a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    n = min(args.cpu_sample, args.n)
    oc, sample = gen_sample_cpu(n, SEED_CFG2)
    threads = args.cpu_threads or os.cpu_count() or 1
    ptr = sample.ctypes.data
    for _ in range(args.warmup):
        oc.lift_bed_ptr(ptr, sample.nbytes, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = oc.lift_bed_ptr(ptr, sample.nbytes, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample_note = (f"{n} sorted BED3 intervals drawn with the same generator/seed as the GPU arm "
                   f"({sample.nbytes} bytes of text), oracle port, {threads} threads over line-aligned splits")
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1] (bounded sample): {n} sorted BED3 intervals over hg19 contigs, "
                                   "hg19ToHg38.over.chain, text in -> text out"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_note},
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else max(args.warmup, 1)
    res = run_reference(args) if args.impl == "reference" else run_cuda(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
