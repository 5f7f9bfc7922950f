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
