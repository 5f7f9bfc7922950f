"""Pins the CPU oracle (oracle/swo_oracle.c).

1. the reference's own known answers: chain.d:204-221 (ChainLink KATs) and
   chain.d:451-481 (19-block chrX chain);
2. hand-derivable answers on resources/minimal.chain (a '-' strand chain) and the
   fixtures named in SURVEY.md §8c;
3. a literal brute-force Python model (tests/pyref.py) on seeded random chains/BEDs,
   including target-overlapping chains, BED3..BED12, odd white space and strands;
4. committed golden digests of the fixture outputs (tests/golden/expected.json).
"""
import hashlib
import json
import os

import numpy as np
import pytest

import pyref
import synth
from conftest import resource_bytes
from oracle_ffi import OrcLink, OracleError

GOLD = os.path.join(os.path.dirname(__file__), "golden", "expected.json")


# --- 1. reference unittests ---------------------------------------------------

def test_chainlink_kats_chain_d_204_221(oracle):
    L = oracle.L
    c1 = OrcLink(1000, 2000, 10_000, 0, 1)
    c2 = OrcLink(1000, 3000, 10_000, 0, 1)
    c3 = OrcLink(1500, 2500, 10_500, 0, -1)
    assert L.orc_link_cmp(c1, c2) < 0          # c1 < c2
    assert L.orc_link_cmp(c2, c3) < 0          # c2 < c3
    assert L.orc_link_cmp_scalar(c3, 1501) < 0  # c3 < 1501
    assert L.orc_link_cmp_scalar(c1, 999) > 0   # c1 > 999
    assert L.orc_link_qend(c1) == 11_000
    assert L.orc_link_qend(c2) == 12_000
    assert L.orc_link_qend(c3) == 9_500         # minus strand
    assert L.orc_link_size(c1) == 1000
    # chain.d:220 expects 1000 for the invert=-1 link, but size() as written
    # (chain.d:143) returns invert*(tEnd-tStart) = -1000: the unittest is stale
    # (SURVEY.md §4).  The oracle follows the code.
    assert abs(L.orc_link_size(c3)) == 1000


CHRX = b"""chain 267340 chrX 155270560 + 49204606 49241588 chrX 156040895 + 49338635 49547634 916
75      964     965
128     1047    1049
686     37      36
63      12      13
279     51      51
204     73      73
73      25      28
145     38      39
100     55      55
270     2613    2636
148     242     242
1292    11570   183551
304     3114    3099
55      43      43
51      5975    6008
104     3239    3227
40      182     182
1334    2239    2239
112"""


def test_chrx_chain_nlinks_chain_d_451_481(oracle):
    assert oracle.count_links(CHRX) == 19
    cf = oracle.chain_parse(CHRX)
    assert cf.nblocks() == 19
    b = cf.blocks(0)
    assert b[0][:3] == (49204606, 49204681, 49338635)
    # the last block must end exactly at the header's tEnd / qEnd
    assert b[-1][1] == 49241588 and b[-1][2] + 112 == 49547634


# --- 2. fixtures ----------------------------------------------------------------

def test_chain_fixture_shapes(oracle):
    cf = oracle.chain_parse(resource_bytes("hg19ToHg38.over.chain"))
    assert cf.nblocks() == 53_950
    assert len(cf.tcontigs) == 93 and len(cf.qcontigs) == 95 and len(cf.qsizes) == 95
    assert cf.tcontigs[0] == "chr1" and cf.qsizes["chr1"] == 248_956_422
    assert cf.nblocks(cf.tcid("chr9")) == 15_222 and cf.nblocks(cf.tcid("chr1")) == 9_716
    c1 = oracle.chain_parse(resource_bytes("chr1_first.chain"))
    assert c1.nblocks() == 4_115
    m = oracle.chain_parse(resource_bytes("minimal.chain"))
    assert m.nblocks() == 9
    # '-' query strand: qStart of block 0 is the exclusive forward upper end
    assert m.blocks(0)[0] == (206072707, 206072707 + 33374, 248956422 - 42687778, 0, -1)


def test_minimal_chain_point_and_interval(oracle):
    m = oracle.chain_parse(resource_bytes("minimal.chain"))
    assert m.lift("chr1", 206072807, 206072907) == [(206072807, 206072907, 206268544, 0, -1)]
    assert m.lift_coord_only("chr1", 206072817) == (1, 206268534)
    # 55-bp gap after block 6
    assert m.lift("chr1", 206276801, 206276856) == []
    assert m.lift_directly("chr1", 206072707) == (1, "chr1", 206268644)
    assert m.lift_directly("chr1", 206072706)[0] == 0


KAT_MINIMAL = [
    (b"chr1 206072807 206072907 x 0 + 206072817 206072897\n",
     b"chr1\t206268444\t206268544\tx\t0\t-\t206268454\t206268534\n", b""),
    # spans blocks 1|2 -> cumulative strand flip: '-' -> '+' -> '-'
    (b"chr1 206106000 206106100 y 0 - 206106000 206106100\n",
     b"chr1\t206235230\t206235249\ty\t0\t+\t206235230\t206235249\n"
     b"chr1\t206235270\t206235351\ty\t0\t-\t206235270\t206235351\n", b""),
    # '.' strand -> '!'; thickStart falls in the 1-bp gap after block 3 -> no thick
    (b"chr1 206207600 206207700 z 0 . 206207649 206207660\n",
     b"chr1\t206133630\t206133680\tz\t0\t!\t206133630\t206133630\n"
     b"chr1\t206133680\t206133729\tz\t0\t!\t206133680\t206133680\n", b""),
    (b"chr1\t206276801\t206276856\n", b"", b"chr1\t206276801\t206276856\n"),
    (b"# comment\ntrack name=x\nbrowser position chr1\nchr1 206276801 206276856 a\r\n", b"",
     b"chr1\t206276801\t206276856\ta\n"),
]


@pytest.mark.parametrize("inp,lifted,unm", KAT_MINIMAL)
def test_minimal_chain_bed_kats(oracle, inp, lifted, unm):
    m = oracle.chain_parse(resource_bytes("minimal.chain"))
    a, b, _ = m.lift_bed(inp)
    assert a == lifted and b == unm
    pa, pb = pyref.lift_bed(pyref.PyChainFile(resource_bytes("minimal.chain")), inp)
    assert (pa, pb) == (lifted, unm)


def test_hg38_fixture_kats(oracle):
    cf = oracle.chain_parse(resource_bytes("hg19ToHg38.over.chain"))
    a, b, _ = cf.lift_bed(resource_bytes("test.bed"))
    assert b == b""
    assert a == (b"chr1\t498221\t498253\tuc001aaq.2\t0\t-\t498221\t498221\t0\t1\t32,\t0,\n"
                 b"chr1\t498129\t498191\tuc001aar.2\t0\t-\t498129\t498129\t0\t1\t62,\t0,\n"
                 b"chr1\t493440\t495049\tuc021oeh.1\t0\t-\t493731\t494822\t0\t4\t58,248,406,514,\t0,151,431,1095,\n")
    a, b, _ = cf.lift_bed(resource_bytes("nomatch_hg19.bed"))
    assert a == b"chr10\t78468614\t78468656\nchr10\t78468665\t78468669\n" and b == b""
    a, b, st = cf.lift_bed(resource_bytes("empty.bed"))
    assert a == b"" and b == b"" and st["n_records"] == 0


def test_errors(oracle):
    m = oracle.chain_parse(resource_bytes("minimal.chain"))
    for bad in (b"chr1 5\n", b"\n", b"chr1 x 10\n", b"chr1 1 99999999999\n", b"chr1 1 2 a b + z 3\n"):
        with pytest.raises(OracleError):
            m.lift_bed(bad)
    with pytest.raises(OracleError):
        oracle.chain_parse(b"")
    with pytest.raises(OracleError):
        oracle.chain_parse(b"chain 1 chr1 10 + 0 5\n5\n")
    with pytest.raises(OracleError):
        oracle.chain_parse(b"chain 1 a 100 + 0 50 b 100 + 0 50 1\n10 5\n")
    # unknown contig -> unmatched (defined deviation; the reference crashes)
    a, b, _ = m.lift_bed(b"chrZ 1 2\n")
    assert a == b"" and b == b"chrZ\t1\t2\n"
    # no trailing newline on the last line
    a, b, _ = m.lift_bed(b"chr1 206276801 206276856")
    assert b == b"chr1\t206276801\t206276856\n"


# --- 3. brute-force model -------------------------------------------------------

@pytest.mark.parametrize("seed", range(12))
def test_against_bruteforce_model(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    overlap = seed % 3 == 2
    chain, tn = synth.random_chain(rng, n_chains=int(rng.integers(1, 12)), overlap=overlap,
                                   sep=b" " if seed % 2 else b"\t")
    ncols = [3, 4, 6, 8, 9, 12][seed % 6]
    bed = synth.random_bed(rng, tn, 400, ncols=ncols, sep=b"  " if seed % 4 == 1 else b"\t",
                           unknown_contig_frac=0.02)
    cf = oracle.chain_parse(chain)
    py = pyref.PyChainFile(chain)
    assert cf.qcontigs == [n.decode() for n in py.contig_names]
    assert cf.qsizes == {k.decode(): v for k, v in py.q_sizes.items()}
    for t in cf.tcontigs:
        assert [b for b in cf.blocks(cf.tcid(t))] == py.trees[t.encode()]
    a, b, st = cf.lift_bed(bed)
    pa, pb = pyref.lift_bed(py, bed)
    assert a == pa and b == pb
    assert st["n_rows"] == a.count(b"\n") and st["n_unmatched"] == b.count(b"\n")
    # threaded driver: byte-identical for any split
    for th in (2, 3, 7):
        a2, b2, _ = cf.lift_bed(bed, threads=th)
        assert a2 == a and b2 == b


def test_binary_rows_match_text(oracle):
    rng = np.random.default_rng(7)
    chain, tn = synth.random_chain(rng, n_chains=10, overlap=True)
    cf = oracle.chain_parse(chain)
    bed = synth.random_bed(rng, tn, 300, ncols=8)
    recs = [l.split() for l in bed.splitlines()]
    tc = [cf.tcid(r[0]) for r in recs]
    strand = [r[5][0] for r in recs]
    R = cf.lift_intervals(tc, [int(r[1]) for r in recs], [int(r[2]) for r in recs], strand,
                          [int(r[6]) for r in recs], [int(r[7]) for r in recs])
    a, b, _ = cf.lift_bed(bed)
    rows = [l.split(b"\t") for l in a.splitlines()]
    assert len(rows) == R["n_rows"]
    qn = cf.qcontigs
    for k, row in enumerate(rows):
        assert row[0].decode() == qn[R["qcid"][k]]
        assert (int(row[1]), int(row[2])) == (R["start"][k], R["end"][k])
        assert row[5][0] == R["strand"][k]
        assert (int(row[6]), int(row[7])) == (R["thick_start"][k], R["thick_end"][k])
        assert row[3] == recs[R["src"][k]][3]
    assert [recs[i][3] for i in R["unmatched_idx"]] == [l.split(b"\t")[3] for l in b.splitlines()]


def test_vcf_points_model(oracle):
    rng = np.random.default_rng(11)
    chain, tn = synth.random_chain(rng, n_chains=8)
    cf = oracle.chain_parse(chain)
    py = pyref.PyChainFile(chain)
    nq = len(cf.qcontigs)
    L = np.array([cf.qsizes[q] for q in cf.qcontigs], np.int64)
    off = np.zeros(nq, np.int64)
    off[1:] = np.cumsum((L[:-1] + 31) // 32 * 32)
    tot = int(off[-1] + (L[-1] + 31) // 32 * 32)
    codes = rng.integers(0, 4, tot, dtype=np.uint8)
    packed = (codes[0::4] | (codes[1::4] << 2) | (codes[2::4] << 4) | (codes[3::4] << 6)).astype(np.uint8)
    n = 500
    tsel = rng.integers(0, len(tn), n)
    tc = np.array([cf.tcid(tn[k]) for k in tsel], np.int32)
    pos = rng.integers(0, 25_000, n).astype(np.int32)
    rl = rng.integers(1, 8, n).astype(np.int32)
    ro = np.zeros(n, np.uint64)
    ro[1:] = np.cumsum(rl[:-1])
    refb = np.frombuffer(bytes(rng.choice(list(b"ACGT"), int(rl.sum())).astype(np.uint8)), np.uint8).copy()
    # make ~half of them agree with the destination genome
    out = cf.lift_vcf_points(packed, off, L, tc, pos, rl, ro, refb)
    for i in range(n):
        name = tn[tsel[i]]
        hits = py.lift(name, int(pos[i]), int(pos[i]) + 1)
        if not hits:
            assert out["flags"][i] == 0 and out["pos"][i] == pos[i]
            continue
        h = hits[0]
        assert out["flags"][i] & 1 and out["qcid"][i] == h[3] and out["pos"][i] == h[2]
        g = int(off[h[3]]) + h[2]
        m = max(0, min(int(rl[i]), int(L[h[3]]) - h[2]))
        new = bytes(b"ACGT"[codes[g + k]] for k in range(m))
        assert bytes(out["ref"][int(ro[i]): int(ro[i]) + m]) == new
        old = bytes(refb[int(ro[i]): int(ro[i]) + int(rl[i])])
        assert bool(out["flags"][i] & 2) == (new != old)


# --- 4. golden digests ----------------------------------------------------------

def test_golden_digests(oracle):
    gold = json.load(open(GOLD))
    for case in gold["cases"]:
        cf = oracle.chain_parse(resource_bytes(case["chain"]))
        a, b, st = cf.lift_bed(resource_bytes(case["bed"]))
        assert hashlib.md5(a).hexdigest() == case["lifted_md5"], case
        assert hashlib.md5(b).hexdigest() == case["unmatched_md5"], case
        assert st["n_rows"] == case["n_rows"] and st["n_unmatched"] == case["n_unmatched"]
