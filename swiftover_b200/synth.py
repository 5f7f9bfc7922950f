"""Synthetic workloads of SURVEY.md §8d, generated ON THE DEVICE (bench/test tooling).

Counter-based RNG: value i of stream s = splitmix64(seed ^ (s << 56) ^ i).  Contigs are
drawn in proportion to their size from the chain's source contigs, lengths are
log-uniform, starts uniform in [0, size-len).  torch is used for device memory and the
element-wise arithmetic; the BED text itself is written by the library's
swo_synth_bed_* kernels so that 1e8-1e9 rows never exist on the host.
"""
import math

import torch

from . import _lib as L

_M64 = (1 << 64) - 1


def _s64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x):
    """splitmix64 finaliser on an int64 tensor holding uint64 bit patterns."""
    z = x + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def uniform01(seed, stream, first, n, device):
    i = torch.arange(first, first + n, dtype=torch.int64, device=device)
    z = splitmix64(i ^ _s64(seed ^ (stream << 56)))
    return _lsr(z, 11).to(torch.float64) * (1.0 / (1 << 53))


def gen_intervals(cf, n, seed, len_lo=1, len_hi=10_000, first=0, device="cuda", sort=False, strand=False, thick=False):
    """Returns dict of device tensors: tcid,start,end (int32) [+ strand uint8, thick_start/end int32]."""
    nt = len(cf.sourceContigs)
    sizes = torch.tensor([cf.source_contig_size(t) for t in range(nt)], dtype=torch.float64, device=device)
    cdf = torch.cumsum(sizes, 0) / sizes.sum()
    tcid = torch.searchsorted(cdf, uniform01(seed, 0, first, n, device)).clamp_(max=nt - 1)
    ln = torch.floor(len_lo * torch.exp(uniform01(seed, 1, first, n, device) * math.log(len_hi / len_lo)))
    ln = ln.clamp_(min=1).to(torch.int64)
    sz = sizes[tcid].to(torch.int64)
    ln = torch.minimum(ln, sz.clamp(min=1))
    start = torch.floor(uniform01(seed, 2, first, n, device) * (sz - ln).clamp(min=0).to(torch.float64)).to(torch.int64)
    end = start + ln
    out = dict(tcid=tcid, start=start, end=end)
    if strand or thick:
        out["strand"] = torch.where(uniform01(seed, 3, first, n, device) < 0.5, 43, 45).to(torch.uint8)  # '+' '-'
    if thick:
        q = ln // 4
        out["thick_start"] = start + q
        out["thick_end"] = end - q
    if sort:  # (contig in chain first-appearance order, start, end)
        order = torch.sort(end, stable=True).indices
        order = order[torch.sort(start[order], stable=True).indices]
        order = order[torch.sort(tcid[order], stable=True).indices]
        out = {k: v[order] for k, v in out.items()}
    for k in ("tcid", "start", "end", "thick_start", "thick_end"):
        if k in out:
            out[k] = out[k].to(torch.int32).contiguous()
    if "strand" in out:
        out["strand"] = out["strand"].contiguous()
    return out


def bed_text(cf, iv, ncols=3, first_index=0):
    """Device uint8 tensor with the BED text of the intervals `iv` (3, 6 or 8 columns)."""
    lib = L.lib()
    n = iv["tcid"].numel()
    dev = iv["tcid"].device
    p = lambda k: iv[k].data_ptr() if k in iv else None
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.swo_synth_bed_lengths_dev(cf.handle, n, ncols, first_index, p("tcid"), p("start"), p("end"), p("strand"),
                                       p("thick_start"), p("thick_end"), lens.data_ptr(), st)
    if rc:
        raise RuntimeError(f"swo_synth_bed_lengths_dev: {rc}")
    off = torch.cumsum(lens.to(torch.int64), 0)
    total = int(off[-1].item()) if n else 0
    off = (off - lens).contiguous()
    text = torch.empty(total + 64, dtype=torch.uint8, device=dev)  # slack keeps 16-byte tail reads in bounds
    rc = lib.swo_synth_bed_write_dev(cf.handle, n, ncols, first_index, p("tcid"), p("start"), p("end"), p("strand"),
                                     p("thick_start"), p("thick_end"), off.data_ptr(), text.data_ptr(), st)
    if rc:
        raise RuntimeError(f"swo_synth_bed_write_dev: {rc}")
    torch.cuda.current_stream(dev).synchronize()
    return text[:total]


class TextLiftBuffers:
    """Device output buffers for swo_lift_bed_text_dev, sized from the input."""

    def __init__(self, in_bytes, device="cuda", lifted_factor=1.5):
        self.lifted = torch.empty(int(in_bytes * lifted_factor) + (1 << 16), dtype=torch.uint8, device=device)
        self.unmatched = torch.empty(in_bytes + 64, dtype=torch.uint8, device=device)
        self.sizes = torch.zeros(6, dtype=torch.int64, device=device)


def lift_text_dev(cf, text, bufs, stream=None):
    """Launch the text kernel on a device-resident BED buffer (asynchronous)."""
    st = (stream or torch.cuda.current_stream(text.device)).cuda_stream
    rc = L.lib().swo_lift_bed_text_dev(cf.handle, text.data_ptr(), text.numel(), bufs.lifted.data_ptr(),
                                       bufs.lifted.numel(), bufs.unmatched.data_ptr(), bufs.unmatched.numel(),
                                       bufs.sizes.data_ptr(), st)
    if rc:
        raise RuntimeError(f"swo_lift_bed_text_dev: {rc} {L.lib().swo_last_error(cf.handle).decode()}")


class BinaryLiftBuffers:
    def __init__(self, n, device="cuda", rows_factor=1.25, thick=False):
        self.cap = int(n * rows_factor) + 4096
        self.row_offset = torch.empty(n + 1 + 4, dtype=torch.int32, device=device)
        self.rows = torch.empty((self.cap, 4), dtype=torch.int32, device=device)
        self.thick = torch.empty((self.cap, 2), dtype=torch.int32, device=device) if thick else None
        self.totals = torch.zeros(2, dtype=torch.int64, device=device)


def lift_intervals_dev(cf, iv, bufs, stream=None):
    """Launch the binary lift kernel on device-resident records (asynchronous)."""
    n = iv["tcid"].numel()
    st = (stream or torch.cuda.current_stream(iv["tcid"].device)).cuda_stream
    p = lambda k: iv[k].data_ptr() if k in iv else None
    rc = L.lib().swo_lift_intervals_dev(cf.handle, n, p("tcid"), p("start"), p("end"), p("strand"), p("thick_start"),
                                        p("thick_end"), bufs.row_offset.data_ptr(), bufs.rows.data_ptr(),
                                        bufs.thick.data_ptr() if bufs.thick is not None else None, bufs.cap,
                                        bufs.totals.data_ptr(), st)
    if rc:
        raise RuntimeError(f"swo_lift_intervals_dev: {rc} {L.lib().swo_last_error(cf.handle).decode()}")


def fragmented_chain(src, seed=20245005, target_blocks=540_000):
    """BASELINE configs[4] stand-in, derived from the chain file text `src`: a target-disjoint chain file with ~10x the blocks of hg19ToHg38 — short
    blocks separated by 1..50 bp gaps on both sides, 30 % of the chains on the '-' strand, 10 %
    re-targeted to another destination contig."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sizes = {}
    for ln in src.split(b"\n"):
        if ln.startswith(b"chain"):
            f = ln.split(b" ")
            sizes[f[2]] = int(f[3])
    names = list(sizes)
    total = float(sum(sizes.values()))
    out = []
    cid = 0
    qsize = 400_000_000
    for t in names:
        nb_t = max(4, int(target_blocks * sizes[t] / total))
        per_chain = 400
        pos_t = int(rng.integers(0, 2000))
        done = 0
        while done < nb_t:
            nb = min(per_chain, nb_t - done)
            avg = max(20, int(0.9 * sizes[t] / nb_t) - 25)
            bs = rng.integers(max(1, avg // 2), avg * 3 // 2 + 2, nb)
            dt = rng.integers(1, 51, nb - 1) if nb > 1 else np.zeros(0, np.int64)
            dq = rng.integers(1, 51, nb - 1) if nb > 1 else np.zeros(0, np.int64)
            span_t = int(bs.sum() + dt.sum())
            span_q = int(bs.sum() + dq.sum())
            if pos_t + span_t >= sizes[t]:
                break
            cid += 1
            q = t if rng.random() >= 0.10 else names[int(rng.integers(0, len(names)))]
            strand = b"-" if rng.random() < 0.30 else b"+"
            qs = int(rng.integers(0, qsize - span_q - 1))
            out.append(b"chain %d %s %d + %d %d %s %d %s %d %d %d" % (
                1_000_000 - cid, t, sizes[t], pos_t, pos_t + span_t, q, qsize, strand, qs, qs + span_q, cid))
            body = np.empty((nb, 3), np.int64)
            body[:, 0] = bs
            body[:-1, 1] = dt
            body[:-1, 2] = dq
            lines = [b"%d\t%d\t%d" % (r[0], r[1], r[2]) for r in body[:-1]]
            lines.append(b"%d" % body[-1, 0])
            out.extend(lines)
            out.append(b"")
            pos_t += span_t + int(rng.integers(1, 200))
            done += nb
    return b"\n".join(out) + b"\n"
