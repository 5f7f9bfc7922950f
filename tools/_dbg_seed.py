import os, sys
ROOT = "/root/repo" if os.path.exists("/root/repo/tests") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth, oracle_ffi
from swiftover_b200 import ChainFile
orc = oracle_ffi.Oracle()
seed = int(os.environ.get("SEED", 771))
rng = np.random.default_rng(50_000 + seed)
ntc = int(rng.integers(1, 5))
chain, tn = synth.random_chain(rng, n_chains=int(rng.integers(1, 16)), n_tcontigs=ntc, max_blocks=int(rng.choice([5, 40, 200])), overlap=bool(rng.random() < 0.3))
cf, oc = ChainFile(text=chain), orc.chain_parse(chain)
names = cf.sourceContigs
n = int(rng.choice([1, 7, 1023, 1025, 5000, 20_000, 60_000]))
tc = rng.integers(-1 if rng.random() < 0.2 else 0, len(names), n).astype(np.int32)
span = int(rng.choice([3_000, 30_000, 120_000]))
s = rng.integers(0, span, n).astype(np.int32)
heavy = rng.random(n) < float(rng.choice([0.0, 0.02, 0.3]))
ln = np.where(heavy, rng.integers(1_000, span + 2, n), rng.integers(1, int(rng.choice([20, 400, 4000])), n))
order = np.lexsort((s, tc))
tc, s, ln = tc[order], s[order], ln[order]
m = int(n * float(rng.choice([0.0, 0.05, 0.12, 0.2, 0.5, 0.9, 1.0])))
perm = np.concatenate([np.arange(n - m), (n - m) + rng.permutation(m)])
if rng.random() < 0.3: perm = perm[::-1].copy()
tc, s, ln = tc[perm], s[perm], ln[perm]
e = (s + ln).astype(np.int32)
print("n", n, "ntc", ntc, "names", len(names), "disjoint", cf.is_disjoint, "blocks", cf.num_blocks, "min tc", tc.min(), "max len", ln.max())
got, ref = cf.lift_intervals(tc, s, e), oc.lift_intervals(tc, s, e)
print("rows", got["n_rows"], ref["n_rows"], "unm", got["n_unmatched"], ref["n_unmatched"])
ro_g, ro_r = got["row_offset"].astype(np.int64), ref["row_offset"].astype(np.int64)
fg, fr = np.diff(ro_g), np.diff(ro_r)
bad = np.flatnonzero(fg != fr)
print("queries with wrong fan-out:", len(bad), bad[:20])
for i in bad[:8]:
    print(i, "tile", i // 128, "lane", (i % 128) // 4, "tc", tc[i], "s", s[i], "e", e[i], "got", fg[i], "ref", fr[i])
d = [(t, int(ro_g[128 * t] - ro_r[128 * t])) for t in range(0, (n + 127) // 128 + 1) if 128 * t <= n]
print("tile-boundary offset errors:", [x for x in d if x[1]][:10])
t = next((x[0] for x in d if x[1]), None)
if t:
    t -= 1
    q = np.arange(128 * t, 128 * t + 128)
    print("tile", t, "tc", np.unique(tc[q]), "ref fan-outs", fr[q].tolist())
    print("s", s[q].tolist())
    print("e-s", (e[q]-s[q]).tolist())
    i0 = cf.tcid(names[tc[q][0]]) if tc[q][0] >= 0 else -1
    ts, te, qs, qc, inv = cf.blocks(int(tc[q][0]))
    print("contig blocks", len(ts), "tStart", ts[:70].tolist(), "tEnd", te[:70].tolist())
if os.environ.get("SWO_DUMP_COUNTS") and os.path.exists(os.environ["SWO_DUMP_COUNTS"]):
    raw = np.fromfile(os.environ["SWO_DUMP_COUNTS"], dtype=np.int32)
    dbg, tot = raw[:4 * n].reshape(n, 4), raw[4 * n:].view(np.uint32)
    cnt = np.maximum(dbg[:, 1] - dbg[:, 0], 0)
    per_tile_ref = np.add.reduceat(fr, np.arange(0, n, 128))
    per_tile_cnt = np.add.reduceat(cnt, np.arange(0, n, 128))
    print("tiles where stored total != ref:", [(t, int(tot[t]), int(per_tile_ref[t]), int(per_tile_cnt[t])) for t in range(len(tot)) if tot[t] != per_tile_ref[t]])
    badq = np.flatnonzero(cnt != fr)
    print("queries where count-pass cnt != ref:", [(int(i), dbg[i].tolist(), int(fr[i])) for i in badq[:10]])
