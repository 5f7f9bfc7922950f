"""The C++ `swiftover` command line (swiftover_b200/host): flag set, checks and exit codes of the
reference's main.d (main.d:28-195).  CPU part: everything main() decides before touching the GPU;
GPU part: a BED liftover through the executable equals the oracle."""
import os
import subprocess

import pytest

from conftest import resource_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "swiftover_b200", "host", "swiftover")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "swiftover_b200", "csrc"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "swiftover_b200", "host"), "-s"])
    return EXE


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True)


def test_cli_exit_codes(exe, tmp_path):
    chain = tmp_path / "ok.chain"
    chain.write_bytes(resource_bytes("minimal.chain"))
    bed = tmp_path / "in.bed"
    bed.write_bytes(resource_bytes("nomatch_hg19.bed"))
    r = run(exe, "--help")  # main.d:74-77: help, return 1
    assert r.returncode == 1 and b"--chainfile" in r.stdout and b"--unmatched" in r.stdout
    assert run(exe).returncode == 1  # -t and -c are required (main.d:41-52)
    assert run(exe, "-t", "bed").returncode == 1
    assert run(exe, "-t", "bed", "-c", str(tmp_path / "none.chain"), "-u", str(tmp_path / "u")).returncode == 1  # main.d:91-92
    gz = tmp_path / "x.chain.gz"
    gz.write_bytes(b"")
    r = run(exe, "-t", "bed", "-c", str(gz), "-u", str(tmp_path / "u"))  # main.d:94-99: return -1
    assert r.returncode == 255 and b"uncompressed" in r.stderr
    r = run(exe, "-t", "bed", "-c", str(chain), "-i", str(tmp_path / "missing.bed"), "-u", str(tmp_path / "u"))
    assert r.returncode == 1  # main.d:100-101
    r = run(exe, "-t", "bed", "-c", str(chain), "-i", str(bed))  # main.d:114-115: -u is mandatory
    assert r.returncode == 1 and b"unmatched" in r.stderr
    r = run(exe, "-t", "maf", "-c", str(chain), "-i", str(bed), "-u", str(tmp_path / "u"))
    assert r.returncode == 1
    r = run(exe, "-t", "nope", "-c", str(chain), "-i", str(bed), "-u", str(tmp_path / "u"))  # main.d:136
    assert r.returncode == 1 and b"Unknown file type" in r.stderr
    r = run(exe, "-t", "vcf", "-c", str(chain), "-i", str(bed), "-o", str(tmp_path / "o"), "-u", str(tmp_path / "u"))
    assert r.returncode == 1 and b"Genome FASTA" in r.stderr  # main.d:126-127
    r = run(exe, "--type=bed", "--chainfile=" + str(chain), "--bogus", "1")
    assert r.returncode == 1


@pytest.mark.gpu
def test_cli_bed_liftover(exe, oracle, tmp_path):
    chain = resource_bytes("hg19ToHg38.over.chain")
    (tmp_path / "c.chain").write_bytes(chain)
    oc = oracle.chain_parse(chain)
    for name in ("chr1_hg19.bed", "test.bed", "nomatch_hg19.bed", "empty.bed"):
        data = resource_bytes(name)
        (tmp_path / "in.bed").write_bytes(data)
        r = run(exe, "-t", "bed", "-c", str(tmp_path / "c.chain"), "-i", str(tmp_path / "in.bed"), "-o",
                str(tmp_path / "out.bed"), "-u", str(tmp_path / "un.bed"))
        assert r.returncode == 0, r.stderr
        o_l, o_u, _ = oc.lift_bed(data)
        assert (tmp_path / "out.bed").read_bytes() == o_l and (tmp_path / "un.bed").read_bytes() == o_u
    # stdin -> stdout when -i / -o are omitted (bed.d:89-97)
    data = resource_bytes("test.bed")
    r = subprocess.run([exe, "-t", "bed", "-c", str(tmp_path / "c.chain"), "-u", str(tmp_path / "un2.bed")], input=data,
                       capture_output=True)
    assert r.returncode == 0 and r.stdout == oc.lift_bed(data)[0]
    # a malformed line: the records in front of it are written (the reference streams, bed.d:113-199), exit status 1
    good = resource_bytes("chr1_hg19.bed")[:50_000]
    good = good[: good.rfind(b"\n") + 1]
    (tmp_path / "bad.bed").write_bytes(good + b"chr1\t12\n" + resource_bytes("test.bed"))
    r = run(exe, "-t", "bed", "-c", str(tmp_path / "c.chain"), "-i", str(tmp_path / "bad.bed"), "-o", str(tmp_path / "o3.bed"),
            "-u", str(tmp_path / "u3.bed"))
    assert r.returncode == 1 and b"fewer than 3 fields" in r.stderr
    o_l, o_u, _ = oc.lift_bed(good)
    assert (tmp_path / "o3.bed").read_bytes() == o_l and (tmp_path / "u3.bed").read_bytes() == o_u
    # an output path that cannot be written: non-zero exit, not a silently truncated file
    (tmp_path / "in.bed").write_bytes(resource_bytes("chr1_hg19.bed"))
    r = run(exe, "-t", "bed", "-c", str(tmp_path / "c.chain"), "-i", str(tmp_path / "in.bed"), "-o", "/dev/full",
            "-u", str(tmp_path / "u4.bed"))
    assert r.returncode != 0
