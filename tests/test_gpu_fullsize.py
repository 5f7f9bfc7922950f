"""BASELINE-size checks of the CUDA path (run with -m gpu on a B200).

The oracle cannot process 1e8 records in test time, so these tests pin the large runs with
size-independent properties and tie them back to the oracle on samples:

  cfg 2  1e8 sorted BED3 lines, text kernel: the whole file lifted at once == the
         concatenation of uneven line-aligned slices lifted separately (tile / look-back /
         sharding invariance, byte exact); several of the slices are small enough for the
         oracle and are compared with it byte for byte.
  cfg 3  BED8 (strand + thickStart/thickEnd), shuffled order: same invariance, oracle on
         sampled slices, host-buffer API == device API.
  cfg 4  VCF records against a synthetic 2-bit destination genome: refchg fires exactly
         where the REF was altered (or clipped), lifted REF equals the genome, oracle sample.
  cfg 5  fragmented chain (~10x blocks, '-' chains, re-targeted chains), long intervals with
         fan-out in the tens to hundreds: offsets / coverage properties, oracle sample, and
         the text kernel on a subset (stage overflow -> direct path).
"""
import gzip
import os

import numpy as np
import pytest

from conftest import resource_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg(oracle):
    from swiftover_b200 import ChainFile
    text = resource_bytes("hg19ToHg38.over.chain")
    cf = ChainFile(text=text)
    yield cf, oracle.chain_parse(text)
    cf.close()


def _line_cut(text, x):
    """Smallest line-aligned offset >= x (device tensor of bytes)."""
    import torch
    if x <= 0:
        return 0
    if x >= text.numel():
        return text.numel()
    w = text[x - 1: x - 1 + 4096]
    nl = (w == 10).nonzero()
    return text.numel() if nl.numel() == 0 else x + int(nl[0].item())


def _lift_dev(cf, text, lifted_factor=1.6):
    """Device API on a device tensor -> (lifted bytes tensor, unmatched bytes tensor, sizes list)."""
    import torch
    from swiftover_b200 import synth as S
    buf = torch.empty(text.numel() + 64, dtype=torch.uint8, device=text.device)  # aligned copy with slack
    buf[: text.numel()] = text
    bufs = S.TextLiftBuffers(text.numel(), device=text.device, lifted_factor=lifted_factor)
    S.lift_text_dev(cf, buf[: text.numel()], bufs)
    torch.cuda.synchronize()
    sz = bufs.sizes.cpu().tolist()
    assert sz[5] == 0, "kernel reported a malformed line"
    assert sz[0] <= bufs.lifted.numel() and sz[1] <= bufs.unmatched.numel()
    return bufs.lifted[: sz[0]], bufs.unmatched[: sz[1]], sz


def _slices_invariance(cf, oc, text, fracs, oracle_max_bytes=3_000_000):
    """whole == concat(slices); slices below oracle_max_bytes are also checked against the oracle."""
    import torch
    whole_l, whole_u, wsz = _lift_dev(cf, text)
    cuts = [_line_cut(text, int(text.numel() * f)) for f in fracs]
    cuts = [0] + cuts + [text.numel()]
    pos_l = pos_u = 0
    recs = rows = unm = 0
    checked = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            continue
        sl_l, sl_u, sz = _lift_dev(cf, text[a:b])
        assert torch.equal(whole_l[pos_l: pos_l + sz[0]], sl_l), f"lifted differs in slice [{a},{b})"
        assert torch.equal(whole_u[pos_u: pos_u + sz[1]], sl_u), f"unmatched differs in slice [{a},{b})"
        if b - a <= oracle_max_bytes:
            o_l, o_u, st = oc.lift_bed(bytes(text[a:b].cpu().numpy()))
            assert bytes(sl_l.cpu().numpy()) == o_l and bytes(sl_u.cpu().numpy()) == o_u
            assert st["n_records"] == sz[2] and st["n_rows"] == sz[3] and st["n_unmatched"] == sz[4]
            checked += 1
        pos_l += sz[0]
        pos_u += sz[1]
        recs += sz[2]
        rows += sz[3]
        unm += sz[4]
    assert pos_l == wsz[0] and pos_u == wsz[1]
    assert (recs, rows, unm) == (wsz[2], wsz[3], wsz[4])
    assert checked >= 3
    return wsz


def test_cfg2_text_full_size(hg):
    import torch
    from swiftover_b200 import synth as S
    cf, oc = hg
    n = 100_000_000
    iv = S.gen_intervals(cf, n, seed=20240002, sort=True)
    text = S.bed_text(cf, iv, ncols=3)
    del iv
    torch.cuda.empty_cache()
    # uneven cuts; the narrow ones (1e-3 of the file = ~2.4 MB) go to the oracle as well
    fracs = [0.0310, 0.0320, 0.21, 0.211, 0.4999, 0.5009, 0.77, 0.9985]
    sz = _slices_invariance(cf, oc, text, fracs)
    assert sz[2] == n
    # the line count of the lifted text equals the reported number of rows
    whole_l, whole_u, _ = _lift_dev(cf, text)
    assert int((whole_l == 10).sum().item()) == sz[3] and int((whole_u == 10).sum().item()) == sz[4]


def test_cfg3_bed8_shuffled(hg):
    import torch
    from swiftover_b200 import synth as S
    cf, oc = hg
    n = 20_000_000
    iv = S.gen_intervals(cf, n, seed=20240003, len_hi=100_000, strand=True, thick=True)  # i.i.d. = shuffled
    text = S.bed_text(cf, iv, ncols=8)
    fracs = [0.0500, 0.0515, 0.33, 0.3312, 0.8, 0.8011]
    sz = _slices_invariance(cf, oc, text, fracs, oracle_max_bytes=2_500_000)
    assert sz[2] == n
    # host-buffer API (chunked H2D / kernel / D2H pipeline) == device API
    whole_l, whole_u, _ = _lift_dev(cf, text)
    host = torch.empty(text.numel(), dtype=torch.uint8, pin_memory=True)
    host.copy_(text)
    torch.cuda.synchronize()
    out = cf.lift_bed_text_ptr(host.data_ptr(), host.numel())
    assert out.lifted_len == whole_l.numel() and out.unmatched_len == whole_u.numel() and out.n_records == n
    import ctypes as C
    got_l = np.ctypeslib.as_array(C.cast(C.c_void_p(out.lifted), C.POINTER(C.c_uint8)), shape=(out.lifted_len,))
    got_u = np.ctypeslib.as_array(C.cast(C.c_void_p(out.unmatched), C.POINTER(C.c_uint8)), shape=(out.unmatched_len,))
    assert torch.equal(torch.from_numpy(got_l), whole_l.cpu())
    assert torch.equal(torch.from_numpy(got_u), whole_u.cpu())
    # binary kernel on the same records: same fan-out per record as the text rows per line
    bufs = S.BinaryLiftBuffers(n, rows_factor=1.6, thick=True)
    S.lift_intervals_dev(cf, iv, bufs)
    torch.cuda.synchronize()
    assert int(bufs.totals[0].item()) == sz[3] and int(bufs.totals[1].item()) == sz[4]


def test_cfg4_vcf_points(hg):
    """REF check against a synthetic 2-bit destination genome at scale (vcf.d:115-169)."""
    import torch
    cf, oc = hg
    rng = np.random.default_rng(20240004)
    n = 10_000_000
    names = cf.contigNames
    Ls = np.array([cf.qContigSizes[nm] for nm in names], np.int64)
    off = np.zeros(len(names), np.int64)
    off[1:] = np.cumsum((Ls[:-1] + 31) // 32 * 32)
    tot = int(off[-1] + (Ls[-1] + 31) // 32 * 32)
    packed = rng.integers(0, 256, tot // 4, dtype=np.uint8)  # 2 bit/base, ~3.2 Gbp -> 0.8 GB
    cf.load_genome_2bit(packed, off, Ls)
    # records: 80 % SNV, 10 % insertion (REF 1), 10 % deletion (REF 2..20)
    tsz = np.array([cf.source_contig_size(t) for t in range(len(cf.sourceContigs))], np.float64)
    tc = rng.choice(len(tsz), n, p=tsz / tsz.sum()).astype(np.int32)
    pos = np.floor(rng.random(n) * tsz[tc]).astype(np.int32)
    rl = np.where(rng.random(n) < 0.9, 1, rng.integers(2, 21, n)).astype(np.int32)
    ro = np.zeros(n, np.uint64)
    ro[1:] = np.cumsum(rl[:-1].astype(np.uint64))
    nref = int(rl.sum())
    # first pass with a dummy REF gives the destination bases; copy them, then alter 1 %
    got0 = cf.lift_vcf_points(tc, pos, rl, ro, np.full(nref, ord("N"), np.uint8))
    matched = (got0["flags"] & 1).astype(bool)
    refb = got0["ref"].copy()
    alter = matched & (rng.random(n) < 0.01)
    idx = ro[alter].astype(np.int64)
    refb[idx] = np.where(refb[idx] == ord("A"), ord("C"), ord("A")).astype(np.uint8)
    got = cf.lift_vcf_points(tc, pos, rl, ro, refb)
    assert np.array_equal(got["flags"] & 1, got0["flags"] & 1)
    clipped = matched & (got["reflen"] != rl)
    changed = (got["flags"] >> 1).astype(bool)
    assert np.array_equal(changed, alter | clipped)  # refchg exactly where REF differs (vcf.d:144)
    assert np.array_equal(got["ref"][ro[matched].astype(np.int64)], got0["ref"][ro[matched].astype(np.int64)])
    oq, op, hit = cf.lift_points(tc, pos)
    assert np.array_equal(hit.astype(bool), matched) and np.array_equal(op[matched], got["pos"][matched])
    assert 0.85 < matched.mean() < 0.97 and alter.sum() > 50_000
    # oracle on a sample
    s = np.sort(rng.choice(n, 100_000, replace=False))
    rls = rl[s]
    ros = np.zeros(len(s), np.uint64)
    ros[1:] = np.cumsum(rls[:-1].astype(np.uint64))
    sub_ref = np.concatenate([refb[int(ro[i]): int(ro[i]) + int(rl[i])] for i in s])
    ref = oc.lift_vcf_points(packed, off, Ls, tc[s], pos[s], rls, ros, sub_ref)
    for k in ("qcid", "pos", "flags", "reflen"):
        assert np.array_equal(got[k][s], ref[k]), k
    sub_out = np.concatenate([got["ref"][int(ro[i]): int(ro[i]) + int(got["reflen"][i])] for i in s])
    exp_out = np.concatenate([ref["ref"][int(ros[j]): int(ros[j]) + int(ref["reflen"][j])] for j in range(len(s))])
    assert np.array_equal(sub_out, exp_out)


def fragmented_chain(seed=20245005, target_blocks=540_000):
    from swiftover_b200.synth import fragmented_chain as fc
    return fc(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "resources", "hg19ToHg38.over.chain.gz"), "rb").read(),
              seed, target_blocks)


def test_cfg5_fragmented_chain(oracle):
    import torch
    from swiftover_b200 import ChainFile
    from swiftover_b200 import synth as S
    chain = fragmented_chain()
    cf = ChainFile(text=chain)
    oc = oracle.chain_parse(chain)
    assert cf.num_blocks > 400_000 and cf.is_disjoint
    n = 1_000_000
    iv = S.gen_intervals(cf, n, seed=20240005, len_lo=10_000, len_hi=1_000_000)
    bufs = S.BinaryLiftBuffers(n, rows_factor=1.0)
    bufs.cap = 0
    S.lift_intervals_dev(cf, iv, bufs)  # sizing pass: nothing is written past capacity 0
    torch.cuda.synchronize()
    nr = int(bufs.totals[0].item())
    assert nr > 20 * n  # fan-out in the tens at least
    bufs = S.BinaryLiftBuffers(n, rows_factor=nr / n + 0.01)
    S.lift_intervals_dev(cf, iv, bufs)
    torch.cuda.synchronize()
    assert int(bufs.totals[0].item()) == nr
    off = bufs.row_offset[: n + 1].to(torch.int64) & 0xFFFFFFFF
    fan = off[1:] - off[:-1]
    assert int(off[-1].item()) == nr and bool((fan >= 0).all())
    rows = bufs.rows[:nr]
    assert bool((rows[:, 2] > rows[:, 1]).all())
    src = rows[:, 3].to(torch.int64) & 0xFFFFFFFF
    assert bool((src[1:] >= src[:-1]).all())
    lifted_len = torch.zeros(n, dtype=torch.int64, device="cuda").index_add_(0, src, (rows[:, 2] - rows[:, 1]).to(torch.int64))
    assert bool((lifted_len <= (iv["end"] - iv["start"]).to(torch.int64)).all())
    # oracle on a sample of whole records (all their rows, in order)
    idx = torch.randperm(n, device="cuda")[:3000].sort().values
    sub = {k: v[idx].cpu().numpy() for k, v in iv.items()}
    ref = oc.lift_intervals(sub["tcid"], sub["start"], sub["end"])
    f = fan[idx].cpu().numpy()
    assert np.array_equal(f, np.diff(ref["row_offset"].astype(np.int64)))
    first = off[idx].cpu().numpy()
    sel = np.concatenate([np.arange(a, a + c) for a, c in zip(first, f)])
    got_rows = rows[torch.from_numpy(sel).cuda()].cpu().numpy()
    assert np.array_equal(got_rows[:, 0] & 0xFFFF, ref["qcid"])
    assert np.array_equal(got_rows[:, 1], ref["start"]) and np.array_equal(got_rows[:, 2], ref["end"])
    # text kernel on a subset: every tile overflows its stage (fan-out) -> direct path
    m = 20_000
    sub_iv = {k: v[:m].contiguous() for k, v in iv.items()}
    text = S.bed_text(cf, sub_iv, ncols=3)
    lifted, unm, sz = _lift_dev(cf, text, lifted_factor=float(fan[:m].sum().item()) * 40 / text.numel() + 2)
    o_l, o_u, st = oc.lift_bed(bytes(text.cpu().numpy()))
    assert bytes(lifted.cpu().numpy()) == o_l and bytes(unm.cpu().numpy()) == o_u
    cf.close()
