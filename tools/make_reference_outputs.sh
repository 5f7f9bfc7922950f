#!/bin/sh
# Produce the REAL swiftover's outputs for this repo's parity cases — to be run by a maintainer on a
# machine with docker and network (this build image has neither a D toolchain nor network).
#
#   tools/make_reference_outputs.sh /path/to/blachlylab/swiftover  [outdir]
#
# Builds the reference exactly as its Dockerfile does (ubuntu:20.04, ldc2 1.24.0, htslib 1.11,
# `dub build -b=release-nobounds`: Dockerfile:1-27; default = first dub configuration = splay tree,
# dub.json:19-22), runs every case of tools/diff_against_reference.py (CASES) through
# `swiftover -t bed -c X.chain -i in.bed -o out.bed -u un.bed`, and writes
#   <outdir>/<case>.lifted.bed, <outdir>/<case>.unmatched.bed, <outdir>/MANIFEST.json (md5 per file)
# Copy MANIFEST.json (and, if small enough, the outputs) to tests/golden/ref_outputs/: then
# tests/test_reference_outputs.py compares them with this repo's output, and
# tools/diff_against_reference.py --ref-dir <outdir> reports exact and modulo-tie-order differences.
set -eu
REF=${1:?path to a checkout of blachlylab/swiftover}
OUT=${2:-ref_outputs}
HERE=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$OUT"
OUT=$(cd "$OUT" && pwd)
docker build -t swiftover-ref "$REF"
WORK=$(mktemp -d)
python3 - "$HERE" "$WORK" <<'PY'
import gzip, os, sys
here, work = sys.argv[1], sys.argv[2]
sys.path.insert(0, os.path.join(here, "tools"))
from diff_against_reference import CASES, fixture
for case, bed, chain in CASES:
    open(os.path.join(work, case + ".bed"), "wb").write(fixture(bed))
    open(os.path.join(work, case + ".chain"), "wb").write(fixture(chain))
open(os.path.join(work, "cases.txt"), "w").write("\n".join(c[0] for c in CASES) + "\n")
PY
while read -r CASE; do
  docker run --rm -v "$WORK":/w -v "$OUT":/o swiftover-ref \
    swiftover -t bed -c "/w/$CASE.chain" -i "/w/$CASE.bed" -o "/o/$CASE.lifted.bed" -u "/o/$CASE.unmatched.bed"
done < "$WORK/cases.txt"
python3 - "$OUT" <<'PY'
import hashlib, json, os, sys
out = sys.argv[1]
m = {f: hashlib.md5(open(os.path.join(out, f), "rb").read()).hexdigest() for f in sorted(os.listdir(out)) if f.endswith(".bed")}
json.dump({"tool": "blachlylab/swiftover, Dockerfile build (ldc2 1.24.0, htslib 1.11, release-nobounds, splay)", "md5": m},
          open(os.path.join(out, "MANIFEST.json"), "w"), indent=1)
print("wrote", len(m), "md5 sums to", os.path.join(out, "MANIFEST.json"))
PY
