"""liftMAF on text (maf.d:119-319): the oracle's restatement against hand-derived answers and an
independent Python model (CPU), and the CUDA path (C ABI: column split/join on the host, lift +
REF fetch/compare in lift_maf_records_kernel) against the oracle (GPU)."""
import numpy as np
import pytest

import synth
from oracle_ffi import OracleError
from test_vcf_text import minimal_case

HDR = b"#version 2.4\nHugo_Symbol\tEntrez_Gene_Id\tCenter\tNCBI_Build\tChromosome\tStart_Position\tEnd_Position\tStrand\tVariant_Classification\tVariant_Type\tReference_Allele\tTumor_Seq_Allele1\tTumor_Seq_Allele2\tdbSNP_RS\n"


def rec(contig, start, end, ref, t1, t2, vtype=b"SNP", extra=b"rs1", sep=b"\t"):
    return sep.join([b"GENE", b"0", b"ctr", b"hg19", contig, b"%d" % start, b"%d" % end, b"+", b"Missense", vtype, ref, t1, t2, extra])


def maf_kat_text(qa):
    ref_ok = qa[1050:1052]
    recs = [
        rec(b"cA", 151, 152, ref_ok, ref_ok, b"TT"),            # block 1 [100,200) -> qA 1051-1052, REF agrees
        rec(b"cA", 151, 152, b"NN", b"NN", b"GG", sep=b"  "),   # REF differs; space separated input is tab joined
        rec(b"cA", 221, 221, b"A", b"A", b"C", sep=b" \t"),     # gap [200,250): unmatched
        rec(b"cA", 1091, 1100, b"ACGTACGTAC", b"ACGTN", b"NA"), # '-' chain: qB 2401-2410, alleles reverse complemented
        rec(b"cA", 190, 260, b"A", b"A", b"C"),                 # blocks 1 and 2: dropped (maf.d:298-307)
        rec(b"cA", 151, 152, b"-", b"-", b"ACG", vtype=b"INS"), # insertion: "-" is never checked (maf.d:263)
        rec(b"zz", 5, 5, b"A", b"A", b"C"),                     # unknown contig: unmatched
        rec(b"cA", 151, 152, b"NN", b"NA", b"-", vtype=b"DEL"), # REF change on a DEL: not counted as "neither"
        rec(b"cA", 151, 152, b"NN", ref_ok, b"GG"),             # REF change, new REF equals tumor allele 1
    ]
    return HDR + b"\n".join(recs) + b"\n"


def test_oracle_reverse_complement(oracle):
    rc = oracle.reverse_complement
    assert rc(b"A") == b"T" and rc(b"ACGT") == b"ACGT" and rc(b"AACG") == b"CGTT" and rc(b"acgt-") == b"-acgt"
    assert rc(b"NA") == b"NA"            # maf.d:41-42
    assert rc(b"ACN") == b"\x00GT"       # NT_COMP_TABLE holds 0 for everything but ACGTacgt and '-'
    assert rc(b"N") == b"\x00" and rc(b"\xc3") == b"\x00"


def test_oracle_maf_kats(oracle):
    chain, fasta, qa, qb = minimal_case()
    oc = oracle.chain_parse(chain)
    lifted, unm, st = oc.lift_maf_text(fasta, maf_kat_text(qa), "GRCh38")
    L = lifted.split(b"\n")
    U = unm.split(b"\n")
    assert lifted.startswith(HDR) and unm.startswith(HDR)
    ref_ok = qa[1050:1052]
    base = [b"GENE", b"0", b"ctr", b"GRCh38"]
    assert L[2] == b"\t".join(base + [b"qA", b"1051", b"1052", b"+", b"Missense", b"SNP", ref_ok, ref_ok, b"TT", b"rs1"])
    assert L[3] == b"\t".join(base + [b"qA", b"1051", b"1052", b"+", b"Missense", b"SNP", ref_ok, b"NN", b"GG", b"rs1"])
    # '-' block [1000,1200) -> qB, link.qStart 2500: [1090,1100) -> qStart 2410, qEnd 2400 -> 2401..2410
    assert L[4] == b"\t".join(base + [b"qB", b"2401", b"2410", b"+", b"Missense", b"SNP", b"acgtNNnnRY", b"\x00ACGT", b"NA", b"rs1"])
    assert L[5] == b"\t".join(base + [b"qA", b"1051", b"1052", b"+", b"Missense", b"INS", b"-", b"-", b"ACG", b"rs1"])
    assert L[6] == b"\t".join(base + [b"qA", b"1051", b"1052", b"+", b"Missense", b"DEL", ref_ok, b"NA", b"-", b"rs1"])
    assert L[7] == b"\t".join(base + [b"qA", b"1051", b"1052", b"+", b"Missense", b"SNP", ref_ok, ref_ok, b"GG", b"rs1"])
    assert L[8:] == [b""]
    assert U[2] == rec(b"cA", 221, 221, b"A", b"A", b"C") and U[3] == rec(b"zz", 5, 5, b"A", b"A", b"C") and U[4:] == [b""]
    assert st == dict(n_records=9, n_matched=7, n_unmatched=2, n_multiple=1, n_refchg=4, n_neither=2)


def test_oracle_maf_headers_and_errors(oracle):
    chain, fasta, qa, _ = minimal_case()
    oc = oracle.chain_parse(chain)
    assert oc.lift_maf_text(fasta, b"")[:2] == (b"", b"")
    assert oc.lift_maf_text(fasta, b"only one line")[:2] == (b"only one line", b"only one line")
    # the first two lines are headers whatever they hold (readln x2, maf.d:166-167); no final newline
    two = rec(b"cA", 151, 152, b"A", b"A", b"C") + b"\n" + rec(b"cA", 151, 152, b"A", b"A", b"C")
    assert oc.lift_maf_text(fasta, two)[:2] == (two, two)
    last = HDR + rec(b"cA", 221, 221, b"A", b"A", b"C\r")  # unterminated last record, '\r' is white space
    assert oc.lift_maf_text(fasta, last)[1] == HDR + rec(b"cA", 221, 221, b"A", b"A", b"C") + b"\n"
    for bad in (b"a\tb\tc\n", b"\n", rec(b"cA", 5, 4, b"A", b"A", b"C") + b"\n",
                rec(b"cA", 5, 6, b"A", b"A", b"C").replace(b"\t5\t", b"\t5x\t") + b"\n"):
        with pytest.raises(OracleError):
            oc.lift_maf_text(fasta, HDR + bad)


def random_maf_case(seed, n=3000):
    rng = np.random.default_rng(seed)
    chain, tn = synth.random_chain(rng, n_chains=12, max_blocks=30)
    qnames = sorted({ln.split(b" ")[7] for ln in chain.split(b"\n") if ln.startswith(b"chain")})
    fasta, seqs = b"", {}
    for k, q in enumerate(qnames):
        if k == 1 and len(qnames) > 2:
            continue  # a destination contig the FASTA lacks: empty fetch
        Ln = 3_000_000 if k else 20_000  # a short sequence: fetches clipped at its end, or past it
        s = bytes(rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), Ln, p=[.2, .2, .2, .2, .04, .04, .04, .04, .04]))
        seqs[q] = s
        fasta += b">" + q + b"\n" + b"\n".join(s[i:i + 70] for i in range(0, Ln, 70)) + b"\n"
    alle = [b"A", b"C", b"G", b"T", b"-", b"NA", b"ACGT", b"acgtn", b"TTTTTTTTTTGA", b"R"]
    recs = []
    for _ in range(n):
        t = tn[int(rng.integers(0, len(tn)))] if rng.random() < 0.97 else b"nope"
        start = int(rng.integers(1, 30_000))
        ln = 1 if rng.random() < 0.6 else int(rng.integers(2, 40))
        r = rng.random()
        ref = b"-" if r < 0.1 else bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), ln))
        vt = [b"SNP", b"DEL", b"INS"][int(rng.integers(0, 3))]
        sep = b"\t" if rng.random() < 0.9 else b" "
        recs.append(rec(t, start, start + ln - 1, ref, alle[int(rng.integers(0, len(alle)))],
                        alle[int(rng.integers(0, len(alle)))], vtype=vt, sep=sep))
    return chain, fasta, HDR + b"\n".join(recs) + b"\n", seqs


def py_revcomp(a):
    if a == b"NA":
        return a
    comp = {ord(x): ord(y) for x, y in zip("ACGTacgt-", "TGCAtgca-")}
    return bytes(comp.get(c, 0) for c in reversed(a))


def test_oracle_maf_against_model(oracle):
    """Independent model of the record body: hits via orc_lift, REF via plain slicing of the FASTA."""
    chain, fasta, text, seqs = random_maf_case(11)
    oc = oracle.chain_parse(chain)
    lifted, unm, st = oc.lift_maf_text(fasta, text, "B38")
    exp_l, exp_u = [], []
    nmulti = nchg = 0
    for ln in text.split(b"\n")[2:-1]:
        f = ln.split()
        s, e = int(f[5]) - 1, int(f[6])
        hits = oc.lift(f[4], s, e) if oc.tcid(f[4]) >= 0 else []
        if not hits:
            exp_u.append(b"\t".join(f))
            continue
        if len(hits) > 1:
            nmulti += 1
            continue
        ts, te, qs, qcid, inv = hits[0]
        qe = qs + inv * (te - ts)
        a, b = min(qs, qe), max(qs, qe)
        q = oc.qcontigs[qcid]
        q = q if isinstance(q, bytes) else q.encode()
        f[3], f[4], f[5], f[6] = b"B38", q, b"%d" % (a + 1), b"%d" % b
        if inv < 0:
            f[11], f[12] = py_revcomp(f[11]), py_revcomp(f[12])
        new = b"-" if f[10] == b"-" else seqs.get(q, b"")[a:b]
        if new != f[10]:
            nchg += 1
            f[10] = new
        exp_l.append(b"\t".join(f))
    assert lifted == HDR + b"".join(x + b"\n" for x in exp_l)
    assert unm == HDR + b"".join(x + b"\n" for x in exp_u)
    assert st["n_multiple"] == nmulti > 0 and st["n_refchg"] == nchg > 0 and st["n_unmatched"] == len(exp_u) > 0
    assert st["n_matched"] == len(exp_l) + nmulti and st["n_records"] == len(exp_l) + len(exp_u) + nmulti


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_gpu_maf_text_matches_oracle(oracle, tmp_path, seed):
    from swiftover_b200 import ChainFile
    chain, fasta, text, _ = random_maf_case(seed)
    fa = tmp_path / "dest.fa"
    fa.write_bytes(fasta)
    oc = oracle.chain_parse(chain)
    cf = ChainFile(text=chain)
    cf.load_genome_fasta(str(fa))
    got = cf.lift_maf_text(text, "B38")
    ref = oc.lift_maf_text(fasta, text, "B38")
    assert got[0] == ref[0] and got[1] == ref[1] and got[2] == ref[2]
    cf.close()


@pytest.mark.gpu
def test_gpu_maf_kats_errors_and_cli(oracle, tmp_path):
    import os
    import subprocess
    from swiftover_b200 import ChainFile, SwiftoverError, liftMAF
    chain, fasta, qa, qb = minimal_case()
    text = maf_kat_text(qa)
    (tmp_path / "m.chain").write_bytes(chain)
    (tmp_path / "g.fa").write_bytes(fasta)
    (tmp_path / "in.maf").write_bytes(text)
    oc = oracle.chain_parse(chain)
    cf = ChainFile(str(tmp_path / "m.chain"))
    cf.load_genome_fasta(str(tmp_path / "g.fa"))
    got = cf.lift_maf_text(text, "GRCh38")
    ref = oc.lift_maf_text(fasta, text, "GRCh38")
    assert got == ref
    assert b"qB\t2401\t2410\t+\tMissense\tSNP\tacgtNNnnRY\t\x00ACGT\tNA\trs1\n" in got[0]
    # 64-bit coordinates: the reference parses to!long; a query reaching past 2^31 still finds its blocks
    big = HDR + rec(b"cA", 300, 1 << 40, b"A", b"A", b"C") + b"\n" + rec(b"cA", 1 << 40, 1 << 41, b"A", b"A", b"C") + b"\n" + \
        rec(b"cA", -(1 << 40), 120, b"A", b"A", b"C") + b"\n"
    assert cf.lift_maf_text(big, "b") == oc.lift_maf_text(fasta, big, "b")
    for t in (b"", b"one line", HDR, HDR + rec(b"cA", 221, 221, b"A", b"A", b"C\r")):
        assert cf.lift_maf_text(t, "b") == oc.lift_maf_text(fasta, t, "b")
    for bad in (b"a\tb\tc\n", b"\n", rec(b"cA", 5, 4, b"A", b"A", b"C") + b"\n",
                rec(b"cA", 5, 6, b"A", b"A", b"C").replace(b"\t5\t", b"\t5x\t") + b"\n"):
        with pytest.raises(SwiftoverError):
            cf.lift_maf_text(HDR + bad, "b")
    cf.close()
    # Python driver and the C++ command line (same flags as the reference, main.d:41-52, 123-128)
    st = liftMAF(str(tmp_path / "m.chain"), str(tmp_path / "g.fa"), "GRCh38", str(tmp_path / "in.maf"),
                 str(tmp_path / "o.maf"), str(tmp_path / "u.maf"))
    assert st == ref[2] and (tmp_path / "o.maf").read_bytes() == ref[0] and (tmp_path / "u.maf").read_bytes() == ref[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "swiftover_b200", "host", "swiftover")
    if os.path.exists(exe):
        args = [exe, "-t", "maf", "-c", str(tmp_path / "m.chain"), "-g", str(tmp_path / "g.fa"), "-i", str(tmp_path / "in.maf"),
                "-o", str(tmp_path / "o2.maf"), "-u", str(tmp_path / "u2.maf")]
        r = subprocess.run(args, capture_output=True)
        assert r.returncode == 1 and b"build name" in r.stderr  # main.d:126-127
        r = subprocess.run(args + ["-b", "GRCh38"], capture_output=True)
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "o2.maf").read_bytes() == ref[0] and (tmp_path / "u2.maf").read_bytes() == ref[1]
        assert b"Matched 7 records (1 multiply, skipped) and 2 unmatched" in r.stderr


@pytest.mark.gpu
def test_gpu_maf_records_non_disjoint(oracle):
    """lift_maf_records on a chain whose blocks overlap on the target: hit counting needs the tEnd filter."""
    from swiftover_b200 import ChainFile
    rng = np.random.default_rng(4)
    chain, tn = synth.random_chain(rng, n_chains=10, max_blocks=20, overlap=True)
    oc = oracle.chain_parse(chain)
    cf = ChainFile(text=chain)
    packed = np.zeros(1 << 20, np.uint8)
    nq = len(oc.qcontigs)
    cf.load_genome_2bit(packed, np.zeros(nq, np.int64), np.full(nq, 1 << 22, np.int64))
    n = 5000
    assert not cf.is_disjoint
    pick = rng.integers(0, len(tn), n)
    tc = np.array([cf.tcid(tn[k]) for k in pick], np.int32)
    s = rng.integers(0, 30_000, n).astype(np.int32)
    e = (s + rng.integers(1, 3000, n)).astype(np.int32)
    r = cf.lift_maf_records(tc, s, e, np.full(n, -1, np.int32), np.zeros(n, np.uint64), np.zeros(0, np.uint8),
                            np.zeros(n, np.uint64), 0)
    for i in range(n):
        hits = oc.lift(tn[pick[i]], int(s[i]), int(e[i]))
        assert r["nhits"][i] == min(len(hits), 2)
        if len(hits) == 1:
            ts, te, qs, qcid, inv = hits[0]
            qe = qs + inv * (te - ts)
            assert (r["qcid"][i], r["start"][i], r["end"][i]) == (qcid, min(qs, qe) + 1, max(qs, qe))
            assert r["flags"][i] == (1 | (4 if inv < 0 else 0))
    cf.close()
