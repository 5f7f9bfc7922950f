#!/usr/bin/env python
"""Box probe (not a bench line): ONE process driving every GPU of the box through
swo_create(devices=[0..N-1]) — the product's in-process sharder — on a single input of
PROBE_GB gigabytes per device: byte equality with the 1-GPU output (sha256), end-to-end time with
and without the concatenation copy, and the bare pinned-copy ceiling of the same byte counts with
one host thread per device."""
import ctypes, gzip, hashlib, json, os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftover_b200 import ChainFile, synth as S

chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
ndev = torch.cuda.device_count()
per_dev = float(os.environ.get("PROBE_GB", "2.4")) * 1e9
n = int(per_dev * ndev / 23.9)
one = ChainFile(text=chain, devices=[0])
iv = S.gen_intervals(one, n, 20240002, device="cuda:0", sort=True)
text = S.bed_text(one, iv, ncols=3)
del iv
host = torch.empty(text.numel(), dtype=torch.uint8, pin_memory=True)
host.copy_(text)
del text
torch.cuda.empty_cache()
view = lambda p, k: (ctypes.c_char * k).from_address(p) if k else b""
sha = lambda p, k: hashlib.sha256(view(p, k)).hexdigest()
res = {"devices": ndev, "input_bytes": host.numel(), "intervals": n}

def timed(cf, reps=3):
    cf.lift_bed_text_ptr(host.data_ptr(), host.numel())
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = cf.lift_bed_text_ptr(host.data_ptr(), host.numel())
        ts.append(time.perf_counter() - t0)
    return out, min(ts) * 1e3

out, ms1 = timed(one, reps=2)
ref = (out.lifted_len, out.unmatched_len, out.n_records, sha(out.lifted, out.lifted_len), sha(out.unmatched, out.unmatched_len))
res["one_device"] = {"ms": ms1, "G_intervals_per_s": n / ms1 / 1e6, "lifted_bytes": out.lifted_len, "unmatched_bytes": out.unmatched_len}
out_bytes = out.lifted_len + out.unmatched_len
one.close()
if ndev > 1:
    cf = ChainFile(text=chain, devices=list(range(ndev)))
    out, ms = timed(cf)
    got = (out.lifted_len, out.unmatched_len, out.n_records, sha(out.lifted, out.lifted_len), sha(out.unmatched, out.unmatched_len))
    res["all_devices_concat"] = {"ms": ms, "G_intervals_per_s": n / ms / 1e6, "equals_one_device": got == ref}
    cf.set_option("concat", 0)
    out, ms = timed(cf)
    h_l, h_u = hashlib.sha256(), hashlib.sha256()
    for p, k in cf.text_segments(0):
        h_l.update(view(p, k))
    for p, k in cf.text_segments(1):
        h_u.update(view(p, k))
    got = (out.lifted_len, out.unmatched_len, out.n_records, h_l.hexdigest(), h_u.hexdigest())
    res["all_devices_segments"] = {"ms": ms, "G_intervals_per_s": n / ms / 1e6, "equals_one_device": got == ref}
    for mib in (16, 256):
        cf.set_option("chunk_bytes", mib << 20)
        _, ms = timed(cf, reps=2)
        res[f"all_devices_segments_chunk{mib}MiB"] = {"ms": ms}
    cf.close()
    # bare copies: every device moves its share in both directions at once, one host thread per device
    ib, ob = host.numel() // ndev, out_bytes // ndev
    bufs = []
    for d in range(ndev):
        with torch.cuda.device(d):
            bufs.append((torch.empty(ib, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(ob, dtype=torch.uint8, device=f"cuda:{d}"),
                         torch.empty(ob, dtype=torch.uint8, pin_memory=True), torch.cuda.Stream(d), torch.cuda.Stream(d)))
    def work(d):
        di, do, ho, s1, s2 = bufs[d]
        with torch.cuda.device(d):
            with torch.cuda.stream(s1):
                di.copy_(host[d * ib:(d + 1) * ib], non_blocking=True)
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
            s1.synchronize(); s2.synchronize()
    best = None
    for it in range(3):
        th = [threading.Thread(target=work, args=(d,)) for d in range(ndev)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = (time.perf_counter() - t0) * 1e3
        if it: best = dt if best is None else min(best, dt)
    res["bare_copy_ceiling"] = {"ms": best, "GBs_each_way_total": host.numel() / best / 1e6}
print(json.dumps(res, indent=1))
