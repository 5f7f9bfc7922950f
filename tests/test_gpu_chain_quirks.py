"""Chain-parser quirks and an Ensembl-style chain through the CUDA engine (the driver's -m gpu run):
the table the device searches is built by the host parser (csrc/chain_table.cpp, replacing
chain.d:272-405 and :509-597), so each quirk the parser reproduces is lifted through on the GPU and
compared with the oracle byte for byte."""
import numpy as np
import pytest

import synth
from conftest import resource_bytes

pytestmark = pytest.mark.gpu


def ensembl_style(chain: bytes) -> bytes:
    """The UCSC fixture renamed the way Ensembl's GRCh37_to_GRCh38.chain names contigs (the
    reference's Makefile:2 fetches that file; its contigs are 1..22, X, Y, MT): 'chr' stripped,
    chrM -> MT, alt/random contigs left out."""
    out, keep = [], False
    for ln in chain.split(b"\n"):
        if ln.startswith(b"chain"):
            f = ln.split(b" ")
            keep = b"_" not in f[2] and b"_" not in f[7]
            if keep:
                ren = lambda n: b"MT" if n == b"chrM" else n[3:]
                f[2], f[7] = ren(f[2]), ren(f[7])
                out.append(b" ".join(f))
        elif keep:
            out.append(ln)
    return b"\n".join(out) + b"\n"


def ensembl_bed(bed: bytes) -> bytes:
    rows = []
    for ln in bed.splitlines():
        f = ln.split(b"\t")
        if b"_" in f[0]:
            continue
        f[0] = b"MT" if f[0] == b"chrM" else f[0][3:]
        rows.append(b"\t".join(f))
    return b"\n".join(rows) + b"\n"


def both(oracle, chain):
    from swiftover_b200 import ChainFile
    return ChainFile(text=chain), oracle.chain_parse(chain)


def test_ensembl_style_chain_bed_and_binary(oracle):
    chain = ensembl_style(resource_bytes("hg19ToHg38.over.chain"))
    cf, oc = both(oracle, chain)
    names = [n.decode() for n in cf.sourceContigs]
    assert "1" in names and "X" in names and "MT" in names and not any(n.startswith("chr") for n in names)
    for bed in (ensembl_bed(resource_bytes("hg19_col123.bed")), ensembl_bed(resource_bytes("chr1_hg19.bed"))):
        lifted, unm, _ = cf.lift_bed_text(bed)
        o_l, o_u, _ = oc.lift_bed(bed)
        assert lifted == o_l and unm == o_u and lifted.startswith((b"1\t", b"2\t", b"3\t", b"4\t", b"5\t", b"6\t", b"7\t", b"8\t", b"9\t", b"X\t", b"Y\t", b"MT\t"))
    recs = [l.split(b"\t") for l in ensembl_bed(resource_bytes("hg19_col123.bed")).splitlines()[:30000]]
    tc = np.array([cf.tcid(r[0]) for r in recs], np.int32)
    s, e = np.array([int(r[1]) for r in recs], np.int32), np.array([int(r[2]) for r in recs], np.int32)
    order = np.lexsort((e, s, tc))  # sorted batch: the merge kernels
    for idx in (np.arange(len(tc)), order):
        got, ref = cf.lift_intervals(tc[idx], s[idx], e[idx]), oc.lift_intervals(tc[idx], s[idx], e[idx])
        for k in ("row_offset", "qcid", "start", "end", "src"):
            assert np.array_equal(got[k], ref[k]), k
    cf.close()


def test_ensembl_style_vcf_header_order(oracle, tmp_path):
    """contigSort / numeric (vcf.d:181-215) on names without 'chr': 1, 2, 10, X, Y, MT."""
    rng = np.random.default_rng(9)
    qn = [b"10", b"2", b"MT", b"1", b"X", b"Y"]
    chain = b"".join(b"chain %d %s 5000 + 100 400 %s 4000 + 1000 1300 %d\n300\n\n" % (1000 - i, q, q, i + 1) for i, q in enumerate(qn))
    fasta = b"".join(b">" + q + b"\n" + bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 4000)) + b"\n" for q in qn)
    text = (b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n" +
            b"".join(b"%s\t%d\t.\tA\tC\t.\t.\t.\n" % (q, 150 + i) for i, q in enumerate(qn)))
    fa = tmp_path / "e.fa"
    fa.write_bytes(fasta)
    cf, oc = both(oracle, chain)
    cf.load_genome_fasta(str(fa))
    got = cf.lift_vcf_text(text, "e.chain", str(fa), "20240101")
    ref = oc.lift_vcf_text(fasta, text, "e.chain", str(fa), "20240101")
    assert got[0] == ref[0] and got[1] == ref[1]
    ids = [l.split(b"ID=")[1].split(b",")[0] for l in got[0].split(b"\n") if l.startswith(b"##contig")]
    assert ids == [b"1", b"2", b"10", b"X", b"Y", b"MT"]
    cf.close()


QUIRKS = {
    # one leading non-chain line is tolerated, the chain ending at line index 0 is dropped (chain.d:528-530)
    "leading_junk_line": b"junk\nchain 1 a 100 + 0 50 b 100 + 0 50 1\n50\n",
    "chain_at_line_0_dropped": b"chain 9 z 100 + 0 50 y 100 + 0 50 9\nchain 1 a 100 + 0 50 b 100 + 0 50 1\n50\n",
    # '-' on the TARGET side: start/end mirrored (chain.d:305-310) and the blocks then run backwards from there
    "minus_target_strand": b"chain 1 a 100 - 10 60 b 100 - 0 50 1\n20 5 5\n25\n\n",
    # space- and tab-separated data lines, blank lines inside, no trailing newline (chain.d:344-347)
    "mixed_separators": b"chain 5 a 1000 + 10 400 b 900 - 100 500 7\n100 20 30\n\n50\t10\t0\n210",
    # two chains with opCmp-equal blocks (chain.d:161-163: one tree node) and target overlaps
    "duplicate_and_overlapping_blocks": (b"chain 3 a 1000 + 0 300 b 900 + 0 300 1\n100 100 100\n100\n\n"
                                         b"chain 2 a 1000 + 0 300 b 900 + 0 300 2\n100 100 100\n100\n\n"
                                         b"chain 1 a 1000 + 50 250 c 500 - 0 200 3\n200\n\n"),
}


@pytest.mark.parametrize("name", sorted(QUIRKS))
def test_quirk_chains_lift_like_the_oracle(oracle, name):
    chain = QUIRKS[name]
    cf, oc = both(oracle, chain)
    assert [n.decode() for n in cf.sourceContigs] == oc.tcontigs and cf.num_blocks == oc.nblocks()
    rng = np.random.default_rng(3)
    rows = []
    for t in oc.tcontigs + ["nowhere"]:
        for _ in range(400):
            a = int(rng.integers(0, 1000))
            b = a + int(rng.integers(1, 400))
            strand = [b"+", b"-", b"."][int(rng.integers(0, 3))]
            rows.append(b"%s\t%d\t%d\tn\t0\t%s\t%d\t%d" % (t.encode(), a, b, strand, a + (b - a) // 4, b - (b - a) // 4))
    bed = b"\n".join(rows) + b"\n"
    lifted, unm, _ = cf.lift_bed_text(bed)
    o_l, o_u, _ = oc.lift_bed(bed)
    assert lifted == o_l and unm == o_u
    tc = np.array([cf.tcid(r.split(b"\t")[0]) for r in rows], np.int32)
    s = np.array([int(r.split(b"\t")[1]) for r in rows], np.int32)
    e = np.array([int(r.split(b"\t")[2]) for r in rows], np.int32)
    keep = tc >= 0
    for idx in (np.flatnonzero(keep), np.flatnonzero(keep)[np.lexsort((e[keep], s[keep], tc[keep]))]):
        got, ref = cf.lift_intervals(tc[idx], s[idx], e[idx]), oc.lift_intervals(tc[idx], s[idx], e[idx])
        for k in ("row_offset", "qcid", "start", "end", "src"):
            assert np.array_equal(got[k], ref[k]), (name, k)
    cf.close()


def test_random_chain_files_sorted_and_unsorted_binary(oracle):
    """Random chains (disjoint and target-overlapping) x sorted batches: disjoint tables take the merge
    kernels, overlapping ones must fall back to the per-query search; both equal the oracle."""
    for seed in range(6):
        rng = np.random.default_rng(300 + seed)
        chain, tn = synth.random_chain(rng, n_chains=int(rng.integers(2, 12)), overlap=bool(seed % 2))
        cf, oc = both(oracle, chain)
        n = 20_000
        tc = rng.integers(0, len(oc.tcontigs), n).astype(np.int32)
        s = rng.integers(0, 30_000, n).astype(np.int32)
        e = (s + rng.integers(1, 3_000, n)).astype(np.int32)
        order = np.lexsort((e, s, tc))
        tc, s, e = tc[order], s[order], e[order]
        strand = rng.choice(np.frombuffer(b"+-.", np.uint8), n)
        for st in (None, strand):
            got, ref = cf.lift_intervals(tc, s, e, strand=st), oc.lift_intervals(tc, s, e, strand=st)
            for k in ("row_offset", "qcid", "start", "end", "src") + (("strand",) if st is not None else ()):
                assert np.array_equal(got[k], ref[k]), (seed, k)
        cf.close()
