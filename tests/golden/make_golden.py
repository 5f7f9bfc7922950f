"""Regenerates tests/golden/expected.json: digests of the oracle's output on every
(chain, bed) fixture pair.  NOTE: the reference ships no expected outputs and cannot be
built here (no D toolchain), so these are ORACLE-derived regression digests, not
reference-derived golden vectors.  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import resource_bytes  # noqa: E402
from oracle_ffi import Oracle, build_oracle  # noqa: E402

build_oracle()
o = Oracle()
cases = []
for chain in ("hg19ToHg38.over.chain", "chr1_first.chain", "minimal.chain"):
    cf = o.chain_parse(resource_bytes(chain))
    for bed in ("hg19_col123.bed", "chr1_hg19.bed", "test.bed", "nomatch_hg19.bed", "empty.bed"):
        a, b, st = cf.lift_bed(resource_bytes(bed))
        cases.append(dict(chain=chain, bed=bed, lifted_md5=hashlib.md5(a).hexdigest(),
                          unmatched_md5=hashlib.md5(b).hexdigest(), n_rows=st["n_rows"],
                          n_unmatched=st["n_unmatched"], lifted_bytes=len(a), unmatched_bytes=len(b)))
json.dump(dict(note="oracle-derived regression digests (see make_golden.py)", cases=cases),
          open(os.path.join(HERE, "expected.json"), "w"), indent=1)
print(json.dumps(cases, indent=1))
