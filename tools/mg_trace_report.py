#!/usr/bin/env python
"""Summarise the per-tile trace of lift_sorted_kernel (debug build: tools/build_variant.sh X -DMG_TRACE=1,
run with SWO_MG_TRACE=file): draw -> publish -> resolve timing, and how often a resolve had to wait."""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t0 = t[t[:, 1] > 0, 1].min()
draw, pub, r0, r1 = (t[:, i] - t0 for i in range(4))
flags = t[:, 4]
qarr, merged, rowsd = (t[:, i] - t0 for i in (5, 6, 7))
ok = (t[:, 0] > 0) & (t[:, 2] > 0)
pc = lambda a, q: tuple(np.percentile(a, q))
print("tiles", len(t), "span us %.1f" % (r1.max() / 1e3))
print("draw->publish us: median %.2f p90 %.2f p99 %.2f max %.2f" % pc((pub - draw)[ok] / 1e3, [50, 90, 99, 100]))
print("publish->resolve begin us: median %.2f p10 %.2f p1 %.2f min %.2f" % pc((r0 - pub)[ok] / 1e3, [50, 10, 1, 0]))
print("resolve dur us: median %.2f p90 %.2f p99 %.2f max %.2f" % pc((r1 - r0)[ok] / 1e3, [50, 90, 99, 100]))
cm = np.maximum.accumulate(pub)
late = cm[:-1] - r0[1:]
w = late > 0
print("resolve began before all predecessors published: frac %.3f, median lateness us %.2f p99 %.2f" %
      (w.mean(), np.median(late[w]) / 1e3 if w.any() else 0, np.percentile(late[w], 99) / 1e3 if w.any() else 0))
k = np.argsort(-(pub - draw) * ok)[:8]
for i in k:
    print("   flags", flags[i], "q-arrive +%.1f merged +%.1f rows +%.1f pub +%.1f" % tuple((x[i] - draw[i]) / 1e3 for x in (qarr, merged, rowsd, pub)))
    print(i, "draw %.1f pub %.1f (%.1f us) r0 %.1f r1 %.1f" % (draw[i] / 1e3, pub[i] / 1e3, (pub[i] - draw[i]) / 1e3, r0[i] / 1e3, r1[i] / 1e3))
print("tiles with generic warps %d, with window overflow %d" % ((flags & 1 > 0).sum(), (flags & 2 > 0).sum()))
for name, m in (("plain", flags == 0), ("overflow", flags & 2 > 0), ("generic", flags & 1 > 0)):
    m = m & ok
    if m.any():
        print(name, "n=%d draw->q %.2f q->merged %.2f merged->rows %.2f rows->pub %.2f (median us)" % (
            m.sum(), np.median((qarr - draw)[m]) / 1e3, np.median((merged - qarr)[m]) / 1e3,
            np.median((rowsd - merged)[m]) / 1e3, np.median((pub - rowsd)[m]) / 1e3))
