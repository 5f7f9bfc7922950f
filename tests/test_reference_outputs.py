"""Pin to the REAL swiftover's outputs when somebody has produced them (tests/golden/ref_outputs/,
made by tools/make_reference_outputs.sh).  Skipped — loudly — while that directory holds no manifest:
then parity with the reference itself is UNPINNED and rests on the oracle's line-by-line restatement
of chain.d / bed.d (oracle/swo_oracle.c) and the reference's two unittests (tests/test_oracle.py)."""
import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref_outputs")
sys.path.insert(0, os.path.join(ROOT, "tools"))
from diff_against_reference import CASES, Ours, compare_modulo_ties, fixture  # noqa: E402

MANIFEST = os.path.join(REF, "MANIFEST.json")
needs_ref = pytest.mark.skipif(not os.path.exists(MANIFEST), reason=(
    "PARITY UNPINNED AGAINST THE REFERENCE BINARY: tests/golden/ref_outputs/MANIFEST.json is absent (no D toolchain / "
    "network in the build image). Run tools/make_reference_outputs.sh where docker exists and drop its output there."))


@needs_ref
@pytest.mark.parametrize("case,bed,chain", CASES)
def test_oracle_output_equals_reference(case, bed, chain):
    md5 = json.load(open(MANIFEST))["md5"]
    lifted, unm, keys = Ours("oracle").lift(fixture(bed), fixture(chain))
    assert hashlib.md5(unm).hexdigest() == md5[f"{case}.unmatched.bed"]
    if hashlib.md5(lifted).hexdigest() == md5[f"{case}.lifted.bed"]:
        return
    path = os.path.join(REF, f"{case}.lifted.bed")
    assert os.path.exists(path), "lifted output differs from the reference's md5 and the reference file is not here to compare modulo tie order"
    rep = compare_modulo_ties(lifted, open(path, "rb").read(), keys)
    assert rep["comparable"] and rep["records_different"] == 0, rep


def test_modulo_tie_comparison_itself():
    """The harness on a made-up pair: a permutation inside an equal-key group passes, anything else does not."""
    ours = b"c\t10\t20\nc\t10\t30\nc\t40\t50\n"
    ref_perm = b"c\t10\t30\nc\t10\t20\nc\t40\t50\n"
    ref_bad = b"c\t10\t30\nc\t40\t50\nc\t10\t20\n"
    keys = [[10, 10, 40]]
    r = compare_modulo_ties(ours, ref_perm, keys)
    assert r["comparable"] and r["records_different"] == 0 and r["tie_groups_permuted"] == 1 and r["records_with_ties"] == 1
    assert compare_modulo_ties(ours, ref_bad, keys)["records_different"] == 1
    assert compare_modulo_ties(ours, ours[:-9], keys)["comparable"] is False
