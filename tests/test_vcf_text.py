"""liftVCF on text (vcf.d:29-176): the oracle's restatement against hand-derived answers
(CPU), and the CUDA path (C ABI, FASTA -> 2-bit genome with case / N side channels, header
rewrite on the host, record body in lift_vcf_points_kernel) against the oracle (GPU)."""
import numpy as np
import pytest

import synth
from conftest import resource_bytes

MINIMAL_FASTA_LEN = 4000


def minimal_case():
    """resources/minimal.chain: chr1 -> chr1 on the '-' strand, first block tStart 206072707,
    link.qStart 206268644.  A tiny destination 'genome' cannot hold chr1, so the
    chain is re-expressed on small coordinates: one '+' chain and one '-' chain."""
    chain = (b"chain 1000 cA 5000 + 100 400 qA 4000 + 1000 1300 1\n100\t50\t50\n150\n\n"
             b"chain 900 cA 5000 + 1000 1200 qB 3000 - 500 700 2\n200\n\n")
    rng = np.random.default_rng(5)
    qa = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 4000))
    qb = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), 3000))
    qb[2400:2410] = b"acgtNNnnRY"  # soft-masked bases and non-ACGT symbols survive the packing
    fasta = b">qA some description\n" + b"\n".join(qa[i:i + 60] for i in range(0, 4000, 60)) + b"\n>qB\n" + bytes(qb) + b"\n>extra\nACGT\n"
    return chain, fasta, qa, bytes(qb)


def vcf_text(chain_qa, qb):
    # cA:151 (0-based 150) lies in block 1 [100,200) -> qA 1050; REF taken from qA so it matches
    # cA:1101 (0-based 1100) lies in the '-' chain: block [1000,1200), qStart = 3000-500 = 2500,
    #   liftDirectly gives 2500 - 100 = 2400 (one past the true forward coordinate; no strand handling)
    hdr = (b"##fileformat=VCFv4.2\n##contig=<ID=cA,length=5000>\n##contig=<ID=qB,length=1>\n"
           b"##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n")
    ref_ok = chain_qa[1050:1052]
    recs = [
        b"cA\t151\trs1\t" + ref_ok + b"\tA\t50\tPASS\tDP=3\tGT\t0/1",          # matched, REF agrees
        b"cA\t151\t.\tGG\tA\t.\t.\t.",                                          # REF differs (unless qA says GG)
        b"cA\t221\t.\tA\tC\t.\t.\tDP=9",                                        # gap [200,250): unmatched
        b"cA\t1101\t.\tACGTACGTAC\tT\t.\t.\tDP=1;X",                            # '-' chain, lower case + N in the genome
        b"zz\t5\t.\tA\tC\t.\t.\t.",                                             # unknown contig: unmatched
        b"cA\t400\t.\tA\tC\t.\t.\t.",                                           # last base of block 2 [250,400)
    ]
    return hdr + b"\n".join(recs) + b"\n"


def test_oracle_vcf_text_kats(oracle):
    chain, fasta, qa, qb = minimal_case()
    oc = oracle.chain_parse(chain)
    text = vcf_text(qa, qb)
    lifted, unm, st = oc.lift_vcf_text(fasta, text, "x.chain", "g.fa", "20240101")
    L = lifted.split(b"\n")
    assert L[0] == b"##fileformat=VCFv4.2"
    # additions sit right before #CHROM; qB already has a contig line (not added again), qA is added
    i = L.index(b"##INFO=<ID=refchg,Number=0,Type=Flag,Description=\"REF allele changed in new genome\">")
    assert L[i + 1] == b"##contig=<ID=qA,length=4000,source=x.chain>"
    assert L[i + 2:i + 7] == [b"##fileDate=20240101", b"##liftoverProgram=swiftover",
                              b"##liftoverProgramURL=https://github.com/blachlylab/swiftover",
                              b"##liftoverChainfile=x.chain", b"##liftoverGenome=g.fa"]
    assert L[i + 7].startswith(b"#CHROM")
    recs = L[i + 8:]
    ref_ok = qa[1050:1052]
    assert recs[0] == b"qA\t1051\trs1\t" + ref_ok + b"\tA\t50\tPASS\tDP=3\tGT\t0/1"
    if ref_ok != b"GG":
        assert recs[1] == b"qA\t1051\t.\t" + ref_ok + b"\tA\t.\t.\trefchg"
    assert recs[2] == b"qB\t2401\t.\tacgtNNnnRY\tT\t.\t.\tDP=1;X;refchg"
    assert recs[3].startswith(b"qA\t1300\t.\t")  # 0-based 399 -> 1150 + 149 = 1299
    U = unm.split(b"\n")
    assert U[0] == b"##fileformat=VCFv4.2" and b"refchg" not in unm
    assert U[-3:-1] == [b"cA\t221\t.\tA\tC\t.\t.\tDP=9", b"zz\t5\t.\tA\tC\t.\t.\t."]
    assert st == dict(n_records=6, n_matched=4, n_unmatched=2, n_refchg=st["n_refchg"]) and st["n_refchg"] >= 2


def random_vcf_case(seed):
    rng = np.random.default_rng(seed)
    chain, tn = synth.random_chain(rng, n_chains=12, max_blocks=30)
    qnames = sorted({ln.split(b" ")[7] for ln in chain.split(b"\n") if ln.startswith(b"chain")})
    fasta = b""
    seqs = {}
    for q in qnames:
        u = rng.random()  # length mismatches, shorter AND longer than the chain header says: no contig line, lifting goes on
        L = 3_000_000 if u < 0.6 else (2_999_000 if u < 0.8 else 3_000_777)  # (the reference only warns, vcf.d:83-86)
        codes = rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), L, p=[.2, .2, .2, .2, .04, .04, .04, .04, .04])
        s = bytes(codes)
        seqs[q] = s
        fasta += b">" + q + b"\n" + b"\n".join(s[i:i + 70] for i in range(0, L, 70)) + b"\n"
    hdr = b"##fileformat=VCFv4.2\n##contig=<ID=" + qnames[0] + b",length=7>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    recs = []
    for i in range(4000):
        t = tn[int(rng.integers(0, len(tn)))] if rng.random() < 0.97 else b"nope"
        pos = int(rng.integers(1, 30_000))
        rl = 1 if rng.random() < 0.8 else int(rng.integers(2, 21))
        ref = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), rl))
        info = [b".", b"DP=4", b"AF=0.5;DB"][int(rng.integers(0, 3))]
        tail = b"" if rng.random() < 0.5 else b"\tGT:DP\t0/1:7"
        recs.append(t + b"\t%d\t.\t" % pos + ref + b"\tT\t.\tPASS\t" + info + tail)
    return chain, fasta, hdr + b"\n".join(recs) + b"\n", seqs


def test_oracle_vcf_text_against_model(oracle):
    """Independent check of the record body: positions via orc_lift_directly, REF via plain slicing."""
    chain, fasta, text, seqs = random_vcf_case(3)
    oc = oracle.chain_parse(chain)
    lifted, unm, st = oc.lift_vcf_text(fasta, text, "c", "g")
    out = [l for l in lifted.split(b"\n") if l and not l.startswith(b"#")]
    inp = [l for l in text.split(b"\n") if l and not l.startswith(b"#")]
    k = 0
    nchg = 0
    for ln in inp:
        f = ln.split(b"\t")
        hit, q, c = oc.lift_directly(f[0], int(f[1]) - 1)
        if not hit:
            continue
        g = out[k].split(b"\t")
        k += 1
        q = q if isinstance(q, bytes) else q.encode()
        new_ref = seqs[q][c:c + len(f[3])]
        assert g[0] == q and int(g[1]) == c + 1 and g[2] == f[2] and g[4:7] == f[4:7] and g[8:] == f[8:]
        if new_ref != f[3]:
            nchg += 1
            assert g[3] == new_ref and g[7] == (b"refchg" if f[7] == b"." else f[7] + b";refchg")
        else:
            assert g[3] == f[3] and g[7] == f[7]
    assert k == len(out) == st["n_matched"] and nchg == st["n_refchg"] and st["n_unmatched"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2])
def test_gpu_vcf_text_matches_oracle(oracle, tmp_path, seed):
    from swiftover_b200 import ChainFile
    chain, fasta, text, _ = random_vcf_case(seed)
    fa = tmp_path / "dest.fa"
    fa.write_bytes(fasta)
    oc = oracle.chain_parse(chain)
    cf = ChainFile(text=chain)
    cf.load_genome_fasta(str(fa))
    got = cf.lift_vcf_text(text, "my.chain", str(fa), "20240101")
    ref = oc.lift_vcf_text(fasta, text, "my.chain", str(fa), "20240101")
    assert got[0] == ref[0] and got[1] == ref[1]
    assert got[2]["n_matched"] == ref[2]["n_matched"] and got[2]["n_unmatched"] == ref[2]["n_unmatched"]
    cf.close()


@pytest.mark.gpu
def test_gpu_vcf_text_kats_and_cli(oracle, tmp_path):
    import os
    import subprocess
    from swiftover_b200 import ChainFile, liftVCF
    chain, fasta, qa, qb = minimal_case()
    text = vcf_text(qa, qb)
    (tmp_path / "m.chain").write_bytes(chain)
    (tmp_path / "g.fa").write_bytes(fasta)
    (tmp_path / "in.vcf").write_bytes(text)
    oc = oracle.chain_parse(chain)
    cf = ChainFile(str(tmp_path / "m.chain"))
    cf.load_genome_fasta(str(tmp_path / "g.fa"))
    assert cf.genome_seq_len("qA") == 4000 and cf.genome_seq_len("extra") == 4 and cf.genome_seq_len("nope") == -1
    got = cf.lift_vcf_text(text, "x.chain", "g.fa", "20240101")
    ref = oc.lift_vcf_text(fasta, text, "x.chain", "g.fa", "20240101")
    assert got[0] == ref[0] and got[1] == ref[1]
    assert b"qB\t2401\t.\tacgtNNnnRY\tT\t.\t.\tDP=1;X;refchg" in got[0]
    cf.close()
    # Python driver and the C++ command line (same flags as the reference, main.d:41-52)
    st = liftVCF(str(tmp_path / "m.chain"), str(tmp_path / "g.fa"), str(tmp_path / "in.vcf"), str(tmp_path / "o.vcf"),
                 str(tmp_path / "u.vcf"), filedate="20240101")
    assert st["n_matched"] == 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "swiftover_b200", "host", "swiftover")
    if os.path.exists(exe):
        r = subprocess.run([exe, "-t", "vcf", "-c", str(tmp_path / "m.chain"), "-g", str(tmp_path / "g.fa"), "-i",
                            str(tmp_path / "in.vcf"), "-o", str(tmp_path / "o2.vcf"), "-u", str(tmp_path / "u2.vcf")],
                           capture_output=True)
        assert r.returncode == 0, r.stderr
        mask = lambda b: b"\n".join(l for l in b.split(b"\n") if not l.startswith((b"##fileDate", b"##liftover", b"##contig")))
        assert mask((tmp_path / "o2.vcf").read_bytes()) == mask((tmp_path / "o.vcf").read_bytes())
        assert (tmp_path / "u2.vcf").read_bytes() == (tmp_path / "u.vcf").read_bytes()
