// lift_core.cuh — the per-record arithmetic of the liftover path as
// __host__ __device__ functions shared by the binary, text and VCF kernels.
//
// Reference semantics (paths relative to /root/reference/source/swiftover):
//   overlap set      intervaltree findOverlapsWith, half-open (chain.d:667)
//   intersect        chain.d:728-743
//   qEnd             chain.d:128-134
//   orderStartEnd    chain.d:787-793
//   STRAND_TABLE     bed.d:34-52, applied cumulatively in place (bed.d:173)
//   thick clamp      bed.d:175-196
//   point lifts      chain.d:604-654 (first hit in (tStart,tEnd,qStart) order)
#pragma once
#include <stdint.h>

#include "chain_table.hpp"

#if defined(__CUDACC__)
#define SWO_HD __host__ __device__ __forceinline__
#else
#define SWO_HD inline
#endif

namespace swo {

// Device view of the flattened block table (see chain_table.hpp).
struct TableView {
    const Block *__restrict__ blk;        // [B] payload (128-bit per block)
    const int32_t *__restrict__ tStartKey;  // [B]
    const int32_t *__restrict__ pmaxKey;    // [B] prefix max of tEnd per contig
    const int32_t *__restrict__ toff;       // [ntc+1]
    const int4 *__restrict__ cmeta;         // [ntc] {first block, end block, Eytzinger offset, samples}
    const int2 *__restrict__ ey;            // Eytzinger upper levels {key, block index} (global copy)
    int ey_n;                               // total samples (<= ChainTable::EY_CAP)
    int ey_shift;                           // group size S = 1 << ey_shift
    int ntc;
    int disjoint;  // pmaxKey == tEnd: candidates need no tEnd filter
};

#if defined(__CUDA_ARCH__)
#define SWO_LDG(p) __ldg(p)
#else
#define SWO_LDG(p) (*(p))
#endif

// first j in [a,b) with key[j] > x, else b
SWO_HD int lower_gt(const int32_t *__restrict__ key, int a, int b, int x) {
    while (a < b) {
        const int m = (int)(((unsigned)a + (unsigned)b) >> 1);
        if (SWO_LDG(key + m) > x) b = m;
        else a = m + 1;
    }
    return a;
}

// first j in [lo,b) with key[j] >= x, else b — exponential probe from lo
// (the answer is almost always lo, lo+1 or lo+2), then bisection.
SWO_HD int gallop_ge(const int32_t *__restrict__ key, int lo, int b, int x) {
    int a = lo, step = 1;
    while (a < b) {
        int p = a + step - 1;
        if (p >= b) p = b - 1;
        if (SWO_LDG(key + p) >= x) {
            int l = a, r = p;
            while (l < r) {
                const int m = (l + r) >> 1;
                if (SWO_LDG(key + m) >= x) r = m;
                else l = m + 1;
            }
            return l;
        }
        a = p + 1;
        step <<= 1;
    }
    return b;
}

// first j in [a,b) with pmaxKey[j] > x, else b.  Upper levels: branch-free
// descent of the contig's Eytzinger-ordered group maxima (`ey` = shared-memory
// or global copy of T.ey), then a bisection inside one group of S blocks.
SWO_HD int lower_gt_ey(const TableView &T, const int2 *__restrict__ ey, const int4 cm, int x) {
    const int a = cm.x, b = cm.y, m = cm.w;
    const int2 *__restrict__ e = ey + cm.z - 1;  // 1-based
    int k = 1;
    while (k <= m) k = 2 * k + (e[k].x <= x ? 1 : 0);
#if defined(__CUDA_ARCH__)
    k >>= __ffs(~k);
#else
    k >>= __builtin_ffs(~k);
#endif
    int lo, hi;  // answer in [lo,hi], hi == b means "none"
    if (k) {
        hi = e[k].y;  // pmaxKey[hi] > x
        lo = hi - ((1 << T.ey_shift) - 1);
    } else {
        lo = a + (m << T.ey_shift);
        hi = b;
    }
    // inside the group: three independent probes per round (the loads of a round overlap)
    while (lo < hi) {
        const int q = (hi - lo + 3) >> 2;
        const int i1 = lo + q - 1;
        const int i2 = i1 + q < hi ? i1 + q : hi - 1;
        const int i3 = i2 + q < hi ? i2 + q : hi - 1;
        const int k1 = SWO_LDG(T.pmaxKey + i1), k2 = SWO_LDG(T.pmaxKey + i2), k3 = SWO_LDG(T.pmaxKey + i3);
        if (k1 > x) {
            hi = i1;
        } else if (k2 > x) {
            lo = i1 + 1;
            hi = i2;
        } else if (k3 > x) {
            lo = i2 + 1;
            hi = i3;
        } else {
            lo = i3 + 1;
        }
    }
    return lo;
}
#if defined(__CUDACC__)
// N lower bounds in lockstep (unsorted batches): the probes of every step are issued for all N queries
// before any of them is used, so one round trip to shared memory / L2 serves N searches.  Same result
// as lower_gt_ey per query; ok[i] false (unknown contig): out[i] = 0.
template <int N>
__device__ __forceinline__ void lower_gt_ey_n(const TableView &T, const int2 *__restrict__ ey, const int4 (&cm)[N], const bool (&ok)[N],
                                              const int (&x)[N], int (&out)[N]) {
    int k[N], m[N], lo[N], hi[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        k[i] = 1;
        m[i] = ok[i] ? cm[i].w : 0;
    }
    for (;;) {
        bool more = false;
        int key[N];
#pragma unroll
        for (int i = 0; i < N; i++) key[i] = k[i] <= m[i] ? ey[cm[i].z - 1 + k[i]].x : 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (k[i] <= m[i]) {
                k[i] = 2 * k[i] + (key[i] <= x[i] ? 1 : 0);
                more = more || k[i] <= m[i];
            }
        }
        if (!more) break;
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int kk = k[i] >> __ffs(~k[i]);
        if (!ok[i]) {
            lo[i] = hi[i] = 0;
        } else if (kk) {
            hi[i] = ey[cm[i].z - 1 + kk].y;  // pmaxKey[hi] > x
            lo[i] = hi[i] - ((1 << T.ey_shift) - 1);
        } else {
            lo[i] = cm[i].x + (m[i] << T.ey_shift);
            hi[i] = cm[i].y;
        }
    }
    for (;;) {
        bool more = false;
        int i1[N], i2[N], i3[N], k1[N], k2[N], k3[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            const bool act = lo[i] < hi[i];
            const int q = (hi[i] - lo[i] + 3) >> 2;
            i1[i] = lo[i] + q - 1;
            i2[i] = i1[i] + q < hi[i] ? i1[i] + q : hi[i] - 1;
            i3[i] = i2[i] + q < hi[i] ? i2[i] + q : hi[i] - 1;
            k1[i] = act ? __ldg(T.pmaxKey + i1[i]) : 0;
            k2[i] = act ? __ldg(T.pmaxKey + i2[i]) : 0;
            k3[i] = act ? __ldg(T.pmaxKey + i3[i]) : 0;
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (lo[i] < hi[i]) {
                if (k1[i] > x[i]) {
                    hi[i] = i1[i];
                } else if (k2[i] > x[i]) {
                    lo[i] = i1[i] + 1;
                    hi[i] = i2[i];
                } else if (k3[i] > x[i]) {
                    lo[i] = i2[i] + 1;
                    hi[i] = i3[i];
                } else {
                    lo[i] = i3[i] + 1;
                }
                more = more || lo[i] < hi[i];
            }
        }
        if (!more) break;
    }
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = lo[i];
}
#endif

// (plain load: the text kernel keeps the per-contig table in shared memory)
SWO_HD int4 load_cmeta(const TableView &T, int tcid) { return T.cmeta[tcid]; }

// Candidate block range of the half-open query [start,end) on contig tcid:
// every block j in [lo,hi) has tStart < end and prefix-max(tEnd) > start; the
// overlap set is { j in [lo,hi) : tEnd[j] > start } (all of them when the
// table is target-disjoint).
struct Cand {
    int lo, hi;
};
SWO_HD Cand find_candidates(const TableView &T, const int2 *__restrict__ ey, int tcid, int start, int end) {
    Cand c{0, 0};
    if (tcid < 0 || tcid >= T.ntc) return c;
    const int4 cm = load_cmeta(T, tcid);
    c.lo = lower_gt_ey(T, ey, cm, start);
    c.hi = gallop_ge(T.tStartKey, c.lo, cm.y, end);
    return c;
}

SWO_HD Block load_block(const TableView &T, int j);
// Same, also returning the first candidate block (valid when hi > lo): blocks lo and
// lo+1 are fetched together, which settles the common 0/1-hit cases in one round trip.
SWO_HD Cand find_candidates_b0(const TableView &T, const int2 *__restrict__ ey, int tcid, int start, int end,
                               Block &b0);

SWO_HD Block load_block(const TableView &T, int j) {
#if defined(__CUDA_ARCH__)
    const int4 v = __ldg(reinterpret_cast<const int4 *>(T.blk) + j);
    return Block{v.x, v.y, v.z, v.w};
#else
    return T.blk[j];
#endif
}

// Same with the contig's cmeta entry given (a caller that keeps the per-contig table in shared
// memory saves the dependent global load at the head of the search).
SWO_HD Cand find_candidates_cm(const TableView &T, const int2 *__restrict__ ey, const int4 cm, int start, int end,
                               Block &b0) {
    Cand c{0, 0};
    b0 = Block{0, 0, 0, 0};
    const int lo = lower_gt_ey(T, ey, cm, start);
    c.lo = c.hi = lo;
    if (lo >= cm.y) return c;
    b0 = load_block(T, lo);
    const int s1 = lo + 1 < cm.y ? SWO_LDG(T.tStartKey + lo + 1) : 0x7fffffff;
    if (b0.tStart >= end) return c;
    c.hi = s1 >= end ? lo + 1 : gallop_ge(T.tStartKey, lo + 2, cm.y, end);
    return c;
}

SWO_HD Cand find_candidates_b0(const TableView &T, const int2 *__restrict__ ey, int tcid, int start, int end,
                               Block &b0) {
    Cand c{0, 0};
    b0 = Block{0, 0, 0, 0};
    if (tcid < 0 || tcid >= T.ntc) return c;
    const int4 cm = load_cmeta(T, tcid);
    const int lo = lower_gt_ey(T, ey, cm, start);
    c.lo = c.hi = lo;
    if (lo >= cm.y) return c;
    b0 = load_block(T, lo);
    const int s1 = lo + 1 < cm.y ? SWO_LDG(T.tStartKey + lo + 1) : 0x7fffffff;
    if (b0.tStart >= end) return c;
    c.hi = s1 >= end ? lo + 1 : gallop_ge(T.tStartKey, lo + 2, cm.y, end);
    return c;
}

// first j in [lo,b) with key[j] > x, else b — exponential probe from lo, then bisection
SWO_HD int gallop_gt(const int32_t *__restrict__ key, int lo, int b, int x) {
    int a = lo, step = 1;
    while (a < b) {
        int p = a + step - 1;
        if (p >= b) p = b - 1;
        if (SWO_LDG(key + p) > x) {
            int l = a, r = p;
            while (l < r) {
                const int m = (l + r) >> 1;
                if (SWO_LDG(key + m) > x) r = m;
                else l = m + 1;
            }
            return l;
        }
        a = p + 1;
        step <<= 1;
    }
    return b;
}
// Candidates of a query known to start at or after an earlier query of the same contig whose
// lower bound was lo_hint (sorted batches): the lower bound only moves forward, usually by 0-2.
SWO_HD Cand find_candidates_from(const TableView &T, int contig_end, int lo_hint, int start, int end, Block &b0) {
    Cand c;
    const int lo = gallop_gt(T.pmaxKey, lo_hint, contig_end, start);
    c.lo = c.hi = lo;
    b0 = Block{0, 0, 0, 0};
    if (lo >= contig_end) return c;
    b0 = load_block(T, lo);
    const int s1 = lo + 1 < contig_end ? SWO_LDG(T.tStartKey + lo + 1) : 0x7fffffff;
    if (b0.tStart >= end) return c;
    c.hi = s1 >= end ? lo + 1 : gallop_ge(T.tStartKey, lo + 2, contig_end, end);
    return c;
}

SWO_HD bool block_hits(const Block &b, int start) { return b.tEnd > start; }

// number of overlapping blocks in the candidate range
SWO_HD int count_hits(const TableView &T, const Cand &c, int start) {
    if (T.disjoint) return c.hi - c.lo;
    int n = 0;
    for (int j = c.lo; j < c.hi; j++) n += SWO_LDG(&T.blk[j].tEnd) > start;
    return n;
}

// One lifted row before text formatting.
struct RowCore {
    int key;   // raw qStart after intersect == the sort key of bed.d:156
    int s, e;  // orderStartEnd(qStart, qEnd)
    int inv;   // +1 | -1
    int qcid;
};
SWO_HD RowCore make_row(const Block &b, int start, int end) {
    RowCore r;
    const int ts = b.tStart > start ? b.tStart : start;  // chain.d:733-734
    const int te = b.tEnd < end ? b.tEnd : end;
    r.inv = (b.meta & 1) ? -1 : 1;
    r.qcid = b.meta >> 1;
    r.key = b.qStart + r.inv * (ts - b.tStart);  // chain.d:736-737
    const int qe = r.key + r.inv * (te - ts);    // chain.d:133
    r.s = r.key < qe ? r.key : qe;               // chain.d:787-793
    r.e = r.key < qe ? qe : r.key;
    return r;
}

// STRAND_TABLE[c + invert]  (bed.d:34-52,173)
SWO_HD unsigned char strand_flip(unsigned char c, int inv) {
    const int idx = (int)c + inv;
    return (idx == 0x2A || idx == 0x2E) ? '-' : (idx == 0x2C ? '+' : '!');
}
// Strand byte of the row at sorted position `rank` of a record, given the
// record's original byte, the invert of the rank-0 row and the number of
// invert==-1 rows at ranks <= rank (self included).  Equals applying
// strand_flip once per row in sorted order (the in-place update of bed.d:173).
SWO_HD unsigned char strand_at_rank(unsigned char orig, int inv_first, int rank, int negs_le) {
    const unsigned char f0 = strand_flip(orig, inv_first);
    if (rank == 0 || f0 == '!') return f0;
    const int negs_after_first = negs_le - (inv_first < 0 ? 1 : 0);
    if ((negs_after_first & 1) == 0) return f0;
    return f0 == '+' ? '-' : '+';
}

// liftCoordOnly / liftDirectly (chain.d:604-654): first block covering c.
// Returns block index or -1.
SWO_HD int point_hit(const TableView &T, const int2 *__restrict__ ey, int tcid, int c) {
    if (tcid < 0 || tcid >= T.ntc) return -1;
    const int4 cm = load_cmeta(T, tcid);
    const int b = cm.y;
    int j = lower_gt_ey(T, ey, cm, c);
    for (; j < b; j++) {
        if (SWO_LDG(T.tStartKey + j) > c) return -1;
        if (T.disjoint || SWO_LDG(&T.blk[j].tEnd) > c) return j;
    }
    return -1;
}
SWO_HD int point_coord(const Block &b, int c) {
    return b.qStart + ((b.meta & 1) ? -1 : 1) * (c - b.tStart);
}

// thickStart/thickEnd of a record (bed.d:134-148).
struct Thick {
    int has;  // hasThickInterval
    int ts, te;
};
SWO_HD Thick lift_thick(const TableView &T, const int2 *__restrict__ ey, int tcid, int ts, int te) {
    Thick k{0, ts, te};
    if (ts == te) return k;
    const int j0 = point_hit(T, ey, tcid, ts);
    if (j0 < 0) return k;  // short-circuit: thickEnd is not looked up
    const int j1 = point_hit(T, ey, tcid, te);
    if (j1 < 0) return k;
    const int a = point_coord(load_block(T, j0), ts);
    const int b = point_coord(load_block(T, j1), te);
    k.has = 1;
    k.ts = a < b ? a : b;  // orderStartEnd  bed.d:146
    k.te = a < b ? b : a;
    return k;
}
// Point query c >= start on the contig of an interval query whose lower bound was lo_hint: the
// lower bound of c lies at or after it (thickStart / thickEnd normally fall inside the interval),
// so a short gallop replaces a full search.
SWO_HD int point_hit_from(const TableView &T, int contig_end, int lo_hint, int c) {
    int j = gallop_gt(T.pmaxKey, lo_hint, contig_end, c);
    for (; j < contig_end; j++) {
        if (SWO_LDG(T.tStartKey + j) > c) return -1;
        if (T.disjoint || SWO_LDG(&T.blk[j].tEnd) > c) return j;
    }
    return -1;
}
// lift_thick for a record whose interval [start, ..) has candidates starting at block lo_hint
SWO_HD Thick lift_thick_from(const TableView &T, const int2 *__restrict__ ey, int tcid, int ts, int te, int start,
                             int lo_hint, int contig_end) {
    Thick k{0, ts, te};
    if (ts == te) return k;
    const int j0 = ts >= start ? point_hit_from(T, contig_end, lo_hint, ts) : point_hit(T, ey, tcid, ts);
    if (j0 < 0) return k;  // short-circuit: thickEnd is not looked up
    const int j1 = te >= start ? point_hit_from(T, contig_end, lo_hint, te) : point_hit(T, ey, tcid, te);
    if (j1 < 0) return k;
    const int a = point_coord(load_block(T, j0), ts);
    const int b = point_coord(load_block(T, j1), te);
    k.has = 1;
    k.ts = a < b ? a : b;  // orderStartEnd  bed.d:146
    k.te = a < b ? b : a;
    return k;
}
// columns 7 and 8 of one output row (bed.d:175-196)
SWO_HD void thick_clamp(const Thick &k, int s, int e, int &c7, int &c8) {
    if (!k.has || k.ts >= e || k.te <= s) {
        c7 = s;
        c8 = s;
    } else {
        c7 = k.ts < s ? s : k.ts;
        c8 = k.te > e ? e : k.te;
    }
}

}  // namespace swo
