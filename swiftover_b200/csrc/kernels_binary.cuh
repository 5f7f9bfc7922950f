// kernels_binary.cuh — single-pass interval lift on binary records.
//
// Replaces, for a whole batch at once, the reference's per-record sequence
//   cf.lift(contig,start,end)            chain.d:664-713 (+ intervaltree search)
//   liftCoordOnly x2 for thickStart/End  bed.d:134-148, chain.d:637-654
//   sort by qStart, orderStartEnd, strand table, thick clamp   bed.d:156-196
//
// One persistent kernel in two variants (sorted / unsorted batches, picked on the device by
// sample_sorted_kernel).  A tile is 1024 consecutive queries (256 threads x 4,
// 128-bit loads).  Each thread searches its four queries (Eytzinger upper levels
// in shared memory + quaternary bottom search; in sorted batches queries 2-4
// gallop forward from the previous lower bound), computes the rows of its
// single-hit queries, the per-query fan-out is block-scanned and the tile
// aggregate is chained to its predecessors with a decoupled look-back resolved
// block-wide (so the variable-length output needs no separate count and scan
// passes: every input byte is read once and every output byte is written
// once); then rows are stored in place.  Queries with >= 2 hits: on a target-disjoint
// table (every UCSC over.chain) and without thick columns they are appended to a global
// list and expanded behind the kernel by expand_wide_kernel (kernels_merge.cuh), a warp
// per query and without sorting (expand_warp_runs below: the candidates of one chain are
// in qStart order already); otherwise they go to a shared-memory work list, filled and
// expanded in rounds: a warp (or the whole block for heavy fan-out) writes one sort key
// per candidate to shared memory, sorts only if the keys are not already ascending, and
// stores every row at its sorted position with the cumulative strand byte.
//
// Sorted batches without thick columns take the merge kernels (kernels_merge.cuh) instead;
// the DEFER variant of this kernel (rows staged in shared memory, look-back resolved one
// tile later) is their predecessor and is no longer instantiated.
// Every block barrier / warp collective behind the divergent per-thread work is
// preceded by converge() (device_util.cuh: converge_warp).
#pragma once
#include "device_util.cuh"
#include "lift_core.cuh"
#include "swiftover_cuda.h"

namespace swo {

#ifndef LI_LOCKSTEP
#define LI_LOCKSTEP 1  // unsorted quadruples: the four searches of a thread in lockstep (lower_gt_ey_n)
#endif
#ifndef LI_MIN_CTAS
#define LI_MIN_CTAS 4
#endif
constexpr int LI_TPB = 256;
constexpr int LI_IPT = 4;
constexpr int LI_TILE = LI_TPB * LI_IPT;
constexpr int LI_NW = LI_TPB / 32;
#ifndef LI_WL_CAP_OVR
#define LI_WL_CAP_OVR 256
#endif
constexpr int LI_WL_CAP = LI_WL_CAP_OVR;      // multi-hit queries queued per tile
constexpr int LI_SMALL = 8;        // = MG_SMALL (kernels_merge.cuh)
constexpr int LI_HEAVY_MIN = 128;   // candidates above this: the whole block expands the query
constexpr int LI_XSCR = 1024;       // shared-memory sort scratch (= LI_NW warp regions of LI_HEAVY_MIN)
constexpr int LI_CM_CAP = 128;      // per-contig table entries kept in shared memory (more contigs: global loads)
constexpr int LI_STAGE = 1152;       // rows of a tile staged in shared memory for the deferred copy-out

struct LiftArgs {
    TableView T;
    uint32_t n;
    int ntiles;
    const int32_t *tcid, *start, *end;
    const uint8_t *strand;            // may be null
    const int32_t *thick_s, *thick_e; // may be null (both)
    uint32_t *row_offset;             // [n+1]
    swo_row *rows;
    swo_thick *thick;
    uint64_t rows_cap;
    uint64_t *desc;     // [ntiles] look-back descriptors, zeroed
    unsigned *ticket;   // zeroed
    uint64_t *totals;   // [0] rows, [1] unmatched   (zeroed)
    const uint32_t *sorted_flag;  // written by sample_sorted_kernel: which kernel variant does the work
    unsigned full_mask;           // 0xffffffff, as a run-time value: see converge()
    unsigned long long *gscr;     // sorted variant: per-warp sort scratch in global memory (kernels_merge.cuh)
    // multi-hit queries left for expand_wide_kernel (kernels_merge.cuh), two uint4 per entry:
    // {query, first candidate block, candidates, row position} {start, end, strand byte, tcid}
    uint4 *wide_ent;              // null: expand inside the kernel
    unsigned *wide_count;         // appended so far (may run past wide_cap: those were expanded on the spot)
    unsigned wide_cap;
    uint4 *small_ent;             // the same for queries with at most LI_SMALL candidates (expand_small_kernel)
    unsigned *small_count;
    unsigned small_cap;
};

// see converge_warp (device_util.cuh)
__device__ __forceinline__ void converge(const LiftArgs &A) { converge_warp(A.full_mask); }
// the same per call site, so that experiments can swap single sites for the plain intrinsic
// (-DSWO_PLAIN_SITES=<bit mask>; profiles/r02_converge_experiments.md)
#ifndef SWO_PLAIN_SITES
#define SWO_PLAIN_SITES 0
#endif
template <int SITE>
__device__ __forceinline__ void converge_at(const LiftArgs &A) {
    if ((SWO_PLAIN_SITES >> SITE) & 1) __syncwarp();
    else converge_warp(A.full_mask);
}

// Where a tile's rows go: global memory (positions are global row indices) or the tile's
// shared-memory stage (positions are offsets inside the tile).
struct RowSink {
    swo_row *rows;
    swo_thick *thick;
    uint64_t cap;
};

// Expand one multi-hit query: every candidate block computes its own sorted
// position (rank by (key, candidate index)), the cumulative strand byte and
// the thick clamp, and writes its row.  Work is strided by (first, stride) so
// the same code serves one thread, one warp or one block.
template <bool THICK>
__device__ __forceinline__ void expand_query(const LiftArgs &A, const RowSink &K, const int2 *ey, uint32_t q, int lo,
                                             int ncand, uint64_t rowbase, int first, int stride) {
    const TableView &T = A.T;
    const int start = __ldg(A.start + q), end = __ldg(A.end + q);
    const unsigned char orig = A.strand ? __ldg(A.strand + q) : 0;
    Thick th{0, 0, 0};
    if (THICK) th = lift_thick(T, ey, __ldg(A.tcid + q), __ldg(A.thick_s + q), __ldg(A.thick_e + q));
    for (int c = first; c < ncand; c += stride) {
        const Block bj = load_block(T, lo + c);
        if (!T.disjoint && !block_hits(bj, start)) continue;
        const RowCore rj = make_row(bj, start, end);
        int rank = 0, negs = rj.inv < 0 ? 1 : 0;
        int best_key = rj.key, best_k = c, best_inv = rj.inv;
        for (int k = 0; k < ncand; k++) {
            if (k == c) continue;
            const Block bk = load_block(T, lo + k);
            if (!T.disjoint && !block_hits(bk, start)) continue;
            const RowCore rk = make_row(bk, start, end);
            const bool before = rk.key < rj.key || (rk.key == rj.key && k < c);
            rank += before;
            negs += (before && rk.inv < 0) ? 1 : 0;
            if (rk.key < best_key || (rk.key == best_key && k < best_k)) {
                best_key = rk.key;
                best_k = k;
                best_inv = rk.inv;
            }
        }
        const uint64_t pos = rowbase + (uint64_t)rank;
        if (pos < K.cap) {
            const unsigned char sb = A.strand ? strand_at_rank(orig, best_inv, rank, negs) : 0;
            swo_row r;
            r.qcid_strand = rj.qcid | ((int)sb << 16);
            r.start = rj.s;
            r.end = rj.e;
            r.src = q;
            K.rows[pos] = r;
            if (THICK) {
                swo_thick t;
                thick_clamp(th, rj.s, rj.e, t.thick_start, t.thick_end);
                K.thick[pos] = t;
            }
        }
    }
}

// Cooperative, sort-based form for the queued multi-hit queries (one warp, or the block for heavy
// fan-out): one 64-bit key (qStart, candidate, invert) per candidate in shared memory, a bitonic
// sort only if the keys are not already ascending, then every sorted position writes its row;
// the cumulative strand byte comes from a running count of inverted rows (bed.d:173).
template <bool BLOCK, bool THICK>
__device__ __noinline__ void expand_query_coop(const LiftArgs &A, const RowSink &K, const int2 *ey, uint32_t q, int lo,
                                               int ncand, uint64_t rowbase, unsigned long long *scr, uint32_t *scan_scr,
                                               int first, int stride) {
    const TableView &T = A.T;
    const int start = __ldg(A.start + q), end = __ldg(A.end + q);
    const unsigned char orig = A.strand ? __ldg(A.strand + q) : 0;
    Thick th{0, 0, 0};
    if (THICK) th = lift_thick(T, ey, __ldg(A.tcid + q), __ldg(A.thick_s + q), __ldg(A.thick_e + q));
    int P = 32;
    while (P < ncand) P <<= 1;
    for (int c = first; c < P; c += stride) {
        unsigned long long key = ~0ull;  // padding and non-hits sort last
        if (c < ncand) {
            const Block b = load_block(T, lo + c);
            if (T.disjoint || block_hits(b, start)) {
                const RowCore r = make_row(b, start, end);
                key = ((unsigned long long)((uint32_t)r.key ^ 0x80000000u) << 32) | ((uint32_t)c << 1) | (r.inv < 0 ? 1u : 0u);
            }
        }
        scr[c] = key;
    }
    converge_at<0>(A);
    if (BLOCK) __syncthreads();
    bool mine = true;
    for (int c = first; c < P; c += stride)
        if (c > 0 && scr[c - 1] > scr[c]) mine = false;
    const bool sorted = BLOCK ? (__syncthreads_and(mine) != 0) : (__all_sync(FULL, mine) != 0);
    if (!sorted) {
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = first; i < P; i += stride) {
                    const int x = i ^ j;
                    if (x > i) {
                        const unsigned long long va = scr[i], vb = scr[x];
                        if ((va > vb) == ((i & k) == 0)) {
                            scr[i] = vb;
                            scr[x] = va;
                        }
                    }
                }
                // warp form: the out-of-line barrier (a plain __syncwarp() is dropped here and the lanes
                // then race on the keys: test_binary_* fail); block form: uniform trip counts behind
                // the reconvergence above, the block barrier alone is enough
                if (BLOCK) __syncthreads();
                else converge_at<1>(A);
            }
        }
    }
    const int inv_first = (scr[0] & 1ull) ? -1 : 1;  // position 0 is a hit: the query has >= 2 hits
    uint32_t carry = 0;                              // inverted rows at earlier positions
    for (int base = 0; base < P; base += stride) {
        const int j = base + first;
        const unsigned long long e = j < P ? scr[j] : ~0ull;
        const bool hit = e != ~0ull;
        const uint32_t neg = hit ? (uint32_t)(e & 1ull) : 0u;
        uint32_t ex, tot;
        if (BLOCK) {
            ex = block_exclusive_scan<uint32_t, LI_NW>(neg, scan_scr, &tot);
        } else {
            const uint32_t inc = warp_inclusive_scan(neg);
            ex = inc - neg;
            tot = __shfl_sync(FULL, inc, 31);
        }
        const uint64_t pos = rowbase + (uint64_t)j;
        if (hit && pos < K.cap) {
            const int c = (int)(((uint32_t)e) >> 1);
            const RowCore r = make_row(load_block(T, lo + c), start, end);
            const unsigned char sb = A.strand ? strand_at_rank(orig, inv_first, j, (int)(carry + ex + neg)) : 0;
            swo_row row;
            row.qcid_strand = r.qcid | ((int)sb << 16);
            row.start = r.s;
            row.end = r.e;
            row.src = q;
            K.rows[pos] = row;
            if (THICK) {
                swo_thick t;
                thick_clamp(th, r.s, r.e, t.thick_start, t.thick_end);
                K.thick[pos] = t;
            }
        }
        carry += tot;
    }
    if (BLOCK) __syncthreads();  // scratch is reused by the next heavy query
}


// A multi-hit query on a target-disjoint table, by one warp and without sorting.  Along the candidate
// blocks (tStart order) the rows of one chain come in qStart order already ('+' chains) or exactly
// reversed ('-' chains): the candidates form a few monotone RUNS, one per chain piece the interval
// crosses.  Pass 1 cuts the candidates into runs (run_continues) and leaves {first candidate, first and
// last destination coordinate, invert} per run in the warp's scratch; lane j then owns run j and adds up
// the sizes (and inverted rows) of the runs that lie entirely below its own; pass 2 gives every candidate
// its output position = run offset + place inside the run, the cumulative strand byte (bed.d:173) from
// the inverted rows in front of it, and stores the row (16 bytes per lane, contiguous inside a run).
// Runs that overlap or touch in the destination, or more than 32 of them (fragmented regions, many small
// chains): up to 32 candidates (R == 1) every candidate counts the rows in front of it — smaller
// qStart, or equal and an earlier candidate: the order of bed.d:156-157 as defined in
// oracle/swo_oracle.h — with shuffles; otherwise the function returns false, nothing useful written,
// and the caller does the same count through its scratch (expand_warp_allpairs).
// Every lane calls it with the same arguments; all loops have warp-uniform trip counts.
// `runs`: 128 ints of per-warp shared memory.
// Does candidate c continue the run of candidate c - 1?  Same destination contig and direction (pm == meta), and
// its piece starts where the previous one ended, give or take a gap of the chain: pv is the previous row's end
// ('+') or start ('-').  A nested chain inside another chain's gap lands somewhere else and starts a run of its
// own.  (Only a heuristic for where to cut: whether the cut runs can be placed is checked exactly below.)
constexpr unsigned RUN_SLACK = 65536;
__device__ __forceinline__ bool run_continues(int c, int pm, int pv, int meta, const RowCore &row) {
    const int g = row.inv > 0 ? row.s - pv : pv - row.e;
    return c > 0 && pm == meta && (unsigned)g <= RUN_SLACK;
}

// R > 0: the query has at most 32 R candidates and their rows stay in registers between the passes (one
// round trip to the block table, all loads in flight together); R == 0: any number, blocks loaded in both passes.
template <bool THICK, int R>
__device__ __forceinline__ bool expand_warp_runs_t(const LiftArgs &A, const RowSink &K, const Thick &th, uint32_t q, int lo, int cnt,
                                                   uint64_t rowbase, int start, int end, unsigned char orig, int *runs, const Block &pre) {
    const TableView &T = A.T;
    const int lane = lane_id();
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    const int rounds = (cnt + 31) >> 5;
    RowCore rows[R > 0 ? R : 1];
    if (R > 0) {
#pragma unroll
        for (int i = 0; i < R; i++) {
            const int c = 32 * i + lane;
            rows[i] = make_row(i == 0 ? pre : load_block(T, lo + (c < cnt ? c : 0)), start, end);
        }
    }
    auto fetch = [&](int i) -> RowCore {
        if (R > 0) return rows[R > 0 ? i : 0];
        const int c = 32 * i + lane;
        return make_row(i == 0 ? pre : load_block(T, lo + (c < cnt ? c : 0)), start, end);
    };
    auto store = [&](const RowCore &row, int rank, int negs_le, int inv_first) {
        const uint64_t pos = rowbase + (uint64_t)rank;
        if (pos < K.cap) {
            const unsigned char sb = A.strand ? strand_at_rank(orig, inv_first, rank, negs_le) : 0;
            *reinterpret_cast<int4 *>(K.rows + pos) = make_int4(row.qcid | ((int)sb << 16), row.s, row.e, (int)q);
            if (THICK) {
                swo_thick tk;
                thick_clamp(th, row.s, row.e, tk.thick_start, tk.thick_end);
                K.thick[pos] = tk;
            }
        }
    };
    // Pass 1: cut into runs; RECORD: leave the runs in the scratch.  Returns the number of runs.
    auto cut = [&](const bool record) -> int {
        int nruns = 0, cm = 0, cv = 0;
#pragma unroll
        for (int i = 0; i < (R > 0 ? R : rounds); i++) {
            if (R > 0 && i >= rounds) break;
            const int c = 32 * i + lane;
            const bool valid = c < cnt;
            const RowCore row = fetch(i);
            const int meta = (row.qcid << 1) | (row.inv < 0 ? 1 : 0), send = row.inv > 0 ? row.e : row.s;
            int pm = __shfl_up_sync(FULL, meta, 1), pv = __shfl_up_sync(FULL, send, 1);
            if (lane == 0) pm = cm, pv = cv;
            const bool brk = valid && !run_continues(c, pm, pv, meta, row);
            const unsigned bm = __ballot_sync(FULL, brk);
            if (record) {
                const int r = nruns + __popc(bm & le_mask) - 1;  // run of this candidate
                if (brk && r < 32) {
                    runs[4 * r] = c;
                    runs[4 * r + 1] = row.inv > 0 ? row.s : row.e;
                    runs[4 * r + 3] = row.inv;
                }
                if (brk && c > 0 && r <= 32) runs[4 * (r - 1) + 2] = pv;  // the previous run ended just before
                if (valid && c == cnt - 1 && r < 32) runs[4 * r + 2] = send;
            }
            nruns += __popc(bm);
            cm = __shfl_sync(FULL, meta, 31);
            cv = __shfl_sync(FULL, send, 31);
        }
        return nruns;
    };
    int nruns = cut(false);
    if (nruns == 1) {  // one chain piece (the usual case): candidate order, or exactly reversed; nothing goes through the scratch
        const int inv_first = __shfl_sync(FULL, fetch(0).inv, 0);
#pragma unroll
        for (int i = 0; i < (R > 0 ? R : rounds); i++) {
            if (R > 0 && i >= rounds) break;
            const int c = 32 * i + lane;
            const RowCore row = fetch(i);
            const int t = inv_first > 0 ? c : cnt - 1 - c;
            if (c < cnt) store(row, t, inv_first < 0 ? t + 1 : 0, inv_first);
        }
        return true;
    }
    cut(true);
    __syncwarp();
    // lane j = run j
    const bool have = lane < nruns && nruns <= 32;
    const int r_first = have ? runs[4 * lane] : 0, r_k0 = have ? runs[4 * lane + 1] : 0, r_k1 = have ? runs[4 * lane + 2] : 0;
    const int r_inv = have ? runs[4 * lane + 3] : 1;
    const int r_size = have ? (lane + 1 < nruns ? runs[4 * (lane + 1)] : cnt) - r_first : 0;
    int off = 0, negoff = 0;
    bool sort = nruns > 32;
    if (!sort) {  // warp-uniform
        const int mn = r_inv > 0 ? r_k0 : r_k1, mx = r_inv > 0 ? r_k1 : r_k0;
        bool bad = false;
        for (int i = 0; i < nruns; i++) {
            const int si = __shfl_sync(FULL, r_size, i), mni = __shfl_sync(FULL, mn, i), mxi = __shfl_sync(FULL, mx, i);
            const int invi = __shfl_sync(FULL, r_inv, i);
            if (have && i != lane) {
                if (mxi < mn) {
                    off += si;
                    negoff += invi < 0 ? si : 0;
                } else if (!(mx < mni)) {
                    bad = true;
                }
            }
        }
        sort = __any_sync(FULL, bad);
    }
    __syncwarp();  // the scratch may be rewritten by the next call
    if (sort) {
        // Runs that overlap or touch in the destination (a nested chain that lands next to its host), or more than
        // 32 of them: every candidate counts the rows in front of it — smaller qStart, or equal and an earlier candidate.
        if (R != 1) return false;  // (through the scratch the count costs the same and needs no registers)
        const RowCore &mine = rows[0];
        const bool valid = lane < cnt;
        const unsigned invm = __ballot_sync(FULL, valid && mine.inv < 0);
        int rank = 0, negs = valid && mine.inv < 0 ? 1 : 0;
        for (int l = 0; l < cnt; l++) {  // warp-uniform
            const int k = __shfl_sync(FULL, mine.key, l);
            const int before = (k < mine.key || (k == mine.key && l < lane)) ? 1 : 0;
            rank += before;
            negs += before & (int)((invm >> l) & 1u);
        }
        const unsigned m0 = __ballot_sync(FULL, valid && rank == 0);
        const int inv_first = __shfl_sync(FULL, mine.inv, __ffs(m0) - 1);
        if (valid) store(mine, rank, negs, inv_first);
        return true;
    }
    const unsigned fm = __ballot_sync(FULL, have && off == 0);
    const int inv_first = __shfl_sync(FULL, r_inv, __ffs(fm) - 1);
    // Pass 2: every candidate to its place
    int cm = 0, cv = 0;
    nruns = 0;
#pragma unroll
    for (int i = 0; i < (R > 0 ? R : rounds); i++) {
        if (R > 0 && i >= rounds) break;
        const int c = 32 * i + lane;
        const bool valid = c < cnt;
        const RowCore row = fetch(i);
        const int meta = (row.qcid << 1) | (row.inv < 0 ? 1 : 0), send = row.inv > 0 ? row.e : row.s;
        int pm = __shfl_up_sync(FULL, meta, 1), pv = __shfl_up_sync(FULL, send, 1);
        if (lane == 0) pm = cm, pv = cv;
        const bool brk = valid && !run_continues(c, pm, pv, meta, row);
        const unsigned bm = __ballot_sync(FULL, brk);
        const int r = (nruns + __popc(bm & le_mask) - 1) & 31;
        const int first = __shfl_sync(FULL, r_first, r), sz = __shfl_sync(FULL, r_size, r);
        const int o = __shfl_sync(FULL, off, r), no = __shfl_sync(FULL, negoff, r);
        const int t = row.inv > 0 ? c - first : first + sz - 1 - c;
        if (valid) store(row, o + t, no + (row.inv < 0 ? t + 1 : 0), inv_first);
        nruns += __popc(bm);
        cm = __shfl_sync(FULL, meta, 31);
        cv = __shfl_sync(FULL, send, 31);
    }
    return true;
}

// `pre`: the lane's block of the first round, lo + min(lane, cnt - 1), loaded by the caller (expand_wide_kernel
// loads it while the previous query is expanded, which takes the table's latency out of the chain).
template <bool THICK>
__device__ __forceinline__ bool expand_warp_runs(const LiftArgs &A, const RowSink &K, const Thick &th, uint32_t q, int lo, int cnt,
                                                 uint64_t rowbase, int start, int end, unsigned char orig, int *runs, const Block &pre) {
    if (cnt <= 32) return expand_warp_runs_t<THICK, 1>(A, K, th, q, lo, cnt, rowbase, start, end, orig, runs, pre);
    if (cnt <= 64) return expand_warp_runs_t<THICK, 2>(A, K, th, q, lo, cnt, rowbase, start, end, orig, runs, pre);
    return expand_warp_runs_t<THICK, 0>(A, K, th, q, lo, cnt, rowbase, start, end, orig, runs, pre);
}

// Any multi-hit query of a target-disjoint table by one warp: one sort key (qStart, candidate, invert) per
// candidate goes to `scr` (cnt entries of the warp's global scratch; the loads below are warp-uniform, i.e.
// one L1 transaction each), then every candidate counts the keys in front of its own.  O(cnt^2 / 32) per
// lane, but with six instructions per pair and no barriers or dependent steps: for the few hundred
// candidates of a query in a fragmented region this beats a bitonic sort through memory by a wide margin.
template <bool THICK>
__device__ __forceinline__ void expand_warp_allpairs(const LiftArgs &A, const RowSink &K, const Thick &th, uint32_t q, int lo, int cnt,
                                                     uint64_t rowbase, int start, int end, unsigned char orig,
                                                     unsigned long long *scr) {
    const TableView &T = A.T;
    const int lane = lane_id();
    unsigned long long best = ~0ull;
    for (int base = 0; base < cnt; base += 32) {
        const int c = base + lane;
        if (c < cnt) {
            const RowCore r = make_row(load_block(T, lo + c), start, end);
            const unsigned long long key = ((unsigned long long)((uint32_t)r.key ^ 0x80000000u) << 32) | ((uint32_t)c << 1) | (r.inv < 0 ? 1u : 0u);
            scr[c] = key;
            best = key < best ? key : best;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, best, d);
        best = o < best ? o : best;
    }
    const int inv_first = (best & 1ull) ? -1 : 1;
    __syncwarp();
    for (int base = 0; base < cnt; base += 32) {
        const int c = base + lane;
        const bool valid = c < cnt;
        const unsigned long long my = valid ? scr[c] : 0ull;
        int rank = 0, negs = (int)(my & 1ull);
#pragma unroll 4
        for (int k = 0; k < cnt; k++) {
            const unsigned long long v = scr[k];
            const int before = v < my ? 1 : 0;
            rank += before;
            negs += before & (int)(v & 1ull);
        }
        const uint64_t pos = rowbase + (uint64_t)rank;
        if (valid && pos < K.cap) {
            const RowCore row = make_row(load_block(T, lo + c), start, end);
            const unsigned char sb = A.strand ? strand_at_rank(orig, inv_first, rank, negs) : 0;
            *reinterpret_cast<int4 *>(K.rows + pos) = make_int4(row.qcid | ((int)sb << 16), row.s, row.e, (int)q);
            if (THICK) {
                swo_thick tk;
                thick_clamp(th, row.s, row.e, tk.thick_start, tk.thick_end);
                K.thick[pos] = tk;
            }
        }
    }
    __syncwarp();  // the scratch may be rewritten by the next call
}

// A query across exactly one gap (two candidate blocks of a target-disjoint table, the usual multi-hit
// case): both rows, in (qStart, block) order, with the running strand flip (bed.d:156-173).  Straight-line.
template <bool THICK>
__device__ __forceinline__ void expand_pair(const LiftArgs &A, const RowSink &K, const Thick &th, uint32_t q, int lo, int start, int end,
                                            unsigned char strand_byte, uint64_t pos) {
    const TableView &T = A.T;
    const RowCore a = make_row(load_block(T, lo), start, end), b = make_row(load_block(T, lo + 1), start, end);
    const bool swap = b.key < a.key;
    const RowCore &r0 = swap ? b : a, &r1 = swap ? a : b;
    unsigned char s0 = 0, s1 = 0;
    if (A.strand) {
        s0 = strand_flip(strand_byte, r0.inv);
        s1 = strand_flip(s0, r1.inv);  // bed.d:173: the flips accumulate
    }
    if (pos < K.cap) {
        *reinterpret_cast<int4 *>(K.rows + pos) = make_int4(r0.qcid | ((int)s0 << 16), r0.s, r0.e, (int)q);
        if (THICK) thick_clamp(th, r0.s, r0.e, K.thick[pos].thick_start, K.thick[pos].thick_end);
    }
    if (pos + 1 < K.cap) {
        *reinterpret_cast<int4 *>(K.rows + pos + 1) = make_int4(r1.qcid | ((int)s1 << 16), r1.s, r1.e, (int)q);
        if (THICK) thick_clamp(th, r1.s, r1.e, K.thick[pos + 1].thick_start, K.thick[pos + 1].thick_end);
    }
}

// Shared memory of one CTA (dynamic: > 48 KiB with the stage).  The 16-byte aligned arrays come first.
template <bool THICK, bool DEFER>
struct LiftSmem {
    int4 st_rows[DEFER ? LI_STAGE : 1];  // staged rows of the pending tile (index = row offset inside the tile)
    uint32_t st_off[DEFER ? LI_TILE : 4];  //   and the tile-local row offset of each of its queries
    int2 st_thick[DEFER && THICK ? LI_STAGE : 1];
    int4 cmeta[LI_CM_CAP];               // per-contig {first block, end block, ey offset, ey samples}
    int2 ey[ChainTable::EY_CAP];         // upper search levels (Eytzinger order)
    unsigned long long xscr[LI_XSCR];    // sort scratch of the multi-hit expansion
    unsigned long long lb_part[LI_NW];
    uint32_t scan[LI_NW + 1];            // scratch of the block scans
    uint32_t wl_q[LI_WL_CAP];
    int wl_lo[LI_WL_CAP];
    int wl_cnt[LI_WL_CAP];
    uint32_t wl_off[LI_WL_CAP];
    int lb_found[LI_NW];
    uint32_t unm;                        // unmatched queries of the tile (shared-memory atomic)
    int tile, wl_n, wl_heavy;
    int p_tile;                          // pending tile (staged, not yet copied out), -1 = none
    uint32_t p_total, p_nq;
};

// Resolve the pending tile's look-back and copy its staged rows and row offsets out (coalesced
// 128-bit stores).  The tile published its aggregate a whole tile earlier, so its predecessors
// have long published theirs: the walk does not spin.  Block-uniform; ends with a barrier.
template <bool THICK, bool DEFER>
__device__ __forceinline__ void flush_pending(const LiftArgs &A, LiftSmem<THICK, DEFER> &S) {
    if (!DEFER) return;
    const int pt = S.p_tile;
    if (pt < 0) return;
    const int tid = threadIdx.x;
    const uint32_t total = S.p_total, nq = S.p_nq;
    const uint64_t gb = lookback_resolve_block<LI_NW>(A.desc, pt, (uint64_t)total, S.lb_part, S.lb_found);
    for (uint32_t i = tid; i < total; i += LI_TPB) {
        const uint64_t pos = gb + i;
        if (pos < A.rows_cap) {
            st_stream_v4(A.rows + pos, S.st_rows[i]);
            if (THICK) *reinterpret_cast<int2 *>(A.thick + pos) = S.st_thick[i];
        }
    }
    const uint32_t q0 = (uint32_t)pt * LI_TILE + (uint32_t)tid * LI_IPT;
    const uint32_t g32 = (uint32_t)gb;  // wraps only if rows >= 2^32 (host rejects)
    if ((uint32_t)tid * LI_IPT + LI_IPT <= nq) {
        const uint32_t *o = S.st_off + tid * LI_IPT;
        st_stream_v4(A.row_offset + q0, make_int4((int)(g32 + o[0]), (int)(g32 + o[1]), (int)(g32 + o[2]), (int)(g32 + o[3])));
    } else {
        for (int k = 0; k < LI_IPT; k++)
            if ((uint32_t)tid * LI_IPT + k < nq) A.row_offset[q0 + k] = g32 + S.st_off[tid * LI_IPT + k];
    }
    if (pt == A.ntiles - 1 && tid == 0) {
        A.row_offset[A.n] = (uint32_t)(gb + total);
        A.totals[0] = gb + total;
    }
    converge_at<2>(A);
    __syncthreads();  // the stage is free again
    if (tid == 0) S.p_tile = -1;
}

// Is the batch (mostly) sorted by position?  512 threads look at 1024 adjacent pairs spread over
// the batch; a pair counts when both queries name the same contig and the start does not move
// back.  *flag = 1 when at least 7 of 8 pairs do (and the caller allows the sorted variant at all).
// Decides which kernel variant does the work.
__global__ void __launch_bounds__(512) sample_sorted_kernel(const int32_t *__restrict__ tcid, const int32_t *__restrict__ start,
                                                             uint32_t n, uint32_t *flag, uint32_t allow) {
    __shared__ unsigned s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint32_t pairs = n > 1 ? n - 1 : 0, want = pairs < 1024u ? pairs : 1024u;
    unsigned ok = 0;
    const uint32_t j0 = threadIdx.x, j1 = threadIdx.x + 512u;
    const uint32_t i0 = j0 < want ? (uint32_t)(((unsigned long long)j0 * pairs) / want) : 0u;
    const uint32_t i1 = j1 < want ? (uint32_t)(((unsigned long long)j1 * pairs) / want) : 0u;
    int a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    if (j0 < want) a0 = __ldg(tcid + i0), a1 = __ldg(tcid + i0 + 1), b0 = __ldg(start + i0), b1 = __ldg(start + i0 + 1);
    if (j1 < want) c0 = __ldg(tcid + i1), c1 = __ldg(tcid + i1 + 1), d0 = __ldg(start + i1), d1 = __ldg(start + i1 + 1);
    ok += (j0 < want && a0 == a1 && b0 <= b1) ? 1u : 0u;
    ok += (j1 < want && c0 == c1 && d0 <= d1) ? 1u : 0u;
    ok = warp_sum(ok);
    if (lane_id() == 0 && ok) atomicAdd(&s_cnt, ok);
    __syncthreads();
    if (threadIdx.x == 0) *flag = (allow && s_cnt * 8u >= want * 7u) ? 1u : 0u;
}

// DEFER = the variant for sorted batches without thick columns (A.sorted_flag says which variant
// works; the other one returns at once; batches with thick columns always take the immediate one): a tile's rows and row offsets are staged in shared memory and its look-back
// is resolved one tile later.  The other variant resolves every tile right after its scan and
// keeps the shared memory for L1 (random probes of the block table).
template <bool THICK, bool DEFER>
__global__ void __launch_bounds__(LI_TPB, LI_MIN_CTAS) lift_intervals_kernel(const __grid_constant__ LiftArgs A) {
    extern __shared__ __align__(16) unsigned char li_smem_raw[];
    LiftSmem<THICK, DEFER> &S = *reinterpret_cast<LiftSmem<THICK, DEFER> *>(li_smem_raw);
    if ((!THICK && __ldg(A.sorted_flag) != 0u) != DEFER) return;

    const int tid = threadIdx.x;
    const TableView &T = A.T;
    for (int i = tid; i < T.ey_n; i += LI_TPB) S.ey[i] = __ldg(T.ey + i);
    const int2 *const ey = S.ey;
    const bool cm_smem = T.ntc <= LI_CM_CAP;  // the search then starts without a dependent global load
    if (cm_smem)
        for (int i = tid; i < T.ntc; i += LI_TPB) S.cmeta[i] = __ldg(T.cmeta + i);
    if (tid == 0) {
        S.p_tile = -1;
        S.unm = 0;
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            S.tile = (int)atomicAdd(A.ticket, 1u);
            S.wl_n = 0;
            S.wl_heavy = 0;
        }
        __syncthreads();
        const int tile = S.tile;
        if (tile >= A.ntiles) {
            flush_pending<THICK, DEFER>(A, S);
            break;
        }

        const uint32_t q0 = (uint32_t)tile * LI_TILE + (uint32_t)tid * LI_IPT;
        int tc[LI_IPT], st[LI_IPT], en[LI_IPT];
        if (q0 + LI_IPT <= A.n) {
            const int4 a = ld_stream_v4(A.tcid + q0), b = ld_stream_v4(A.start + q0), c = ld_stream_v4(A.end + q0);
            tc[0] = a.x, tc[1] = a.y, tc[2] = a.z, tc[3] = a.w;
            st[0] = b.x, st[1] = b.y, st[2] = b.z, st[3] = b.w;
            en[0] = c.x, en[1] = c.y, en[2] = c.z, en[3] = c.w;
        } else {
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) {
                const bool ok = q0 + k < A.n;
                tc[k] = ok ? A.tcid[q0 + k] : -1;
                st[k] = ok ? A.start[q0 + k] : 0;
                en[k] = ok ? A.end[q0 + k] : 0;
            }
        }
        Cand cd[LI_IPT];
        int nh[LI_IPT];
        // row of a single-hit query
        int r_meta[LI_IPT], r_s[LI_IPT], r_e[LI_IPT];
        uint32_t tsum = 0, unm = 0;
        constexpr bool LOCKSTEP = LI_LOCKSTEP && !THICK;  // (with thick columns the lockstep form was not faster)
        bool tc_ok[LI_IPT];
        int4 cm[LI_IPT];
        auto contig = [&](int k) {
            tc_ok[k] = tc[k] >= 0 && tc[k] < T.ntc;
            cm[k] = !tc_ok[k] ? make_int4(0, 0, 0, 0) : cm_smem ? S.cmeta[tc[k]] : load_cmeta(T, tc[k]);
        };
        if (LOCKSTEP) {
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) contig(k);
        }
        // A thread whose four queries are in order (same contig, starts not moving back: sorted stretches) searches
        // the first and gallops forward for the others; otherwise the four searches run in lockstep, so that every
        // round trip to the block table serves all four.
        bool chained = true;
#pragma unroll
        for (int k = 1; k < LI_IPT; k++) chained = chained && tc[k] == tc[k - 1] && st[k] >= st[k - 1];
        const bool lockstep = LOCKSTEP && !chained;
        int lo4[LI_IPT];
        if (lockstep) lower_gt_ey_n<LI_IPT>(T, ey, cm, tc_ok, st, lo4);
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            Block b0;
            if (!LOCKSTEP) contig(k);
            if (!tc_ok[k]) {
                cd[k] = Cand{0, 0};
            } else if (lockstep) {
                // (find_candidates_cm behind its search)
                const int lo = lo4[k];
                cd[k] = Cand{lo, lo};
                b0 = Block{0, 0, 0, 0};
                if (lo < cm[k].y) {
                    b0 = load_block(T, lo);
                    const int s1 = lo + 1 < cm[k].y ? __ldg(T.tStartKey + lo + 1) : 0x7fffffff;
                    if (b0.tStart < en[k]) cd[k].hi = s1 >= en[k] ? lo + 1 : gallop_ge(T.tStartKey, lo + 2, cm[k].y, en[k]);
                }
            } else if (k > 0 && tc[k] == tc[k - 1] && st[k] >= st[k - 1]) {
                cd[k] = find_candidates_from(T, cm[k].y, cd[k - 1].lo, st[k], en[k], b0);
            } else {
                cd[k] = find_candidates_cm(T, ey, cm[k], st[k], en[k], b0);
            }
            nh[k] = count_hits(T, cd[k], st[k]);
            tsum += (uint32_t)nh[k];
            unm += (q0 + k < A.n && nh[k] == 0) ? 1u : 0u;
            r_meta[k] = r_s[k] = r_e[k] = 0;
            if (nh[k] == 1) {
                int j = cd[k].lo;
                if (!T.disjoint)
                    while (!block_hits(b0, st[k])) b0 = load_block(T, ++j);
                const RowCore r = make_row(b0, st[k], en[k]);
                const unsigned char sb = A.strand ? strand_flip(__ldg(A.strand + q0 + k), r.inv) : 0;
                r_meta[k] = r.qcid | ((int)sb << 16);
                r_s[k] = r.s;
                r_e[k] = r.e;
            }
        }
        converge_at<3>(A);
        // Block scan of the per-thread fan-out (the warp was reconverged just above: the scan's
        // shuffles and barriers need every lane there); the unmatched count goes through a
        // shared-memory atomic.
        if (unm) atomicAdd(&S.unm, unm);
        uint32_t tile_total;
        const uint32_t excl = block_exclusive_scan<uint32_t, LI_NW>(tsum, S.scan, &tile_total);
        if (warp_id() == 0) lookback_publish(A.desc, tile, (uint64_t)tile_total);
        if (tid == 0 && S.unm) {  // every thread's atomicAdd is behind the scan's barriers
            atomicAdd((unsigned long long *)(A.totals + 1), (unsigned long long)S.unm);
            S.unm = 0;
        }

        // The previous tile of this CTA is resolved and copied out now, one whole search phase after
        // it published its aggregate.
        flush_pending<THICK, DEFER>(A, S);

        // A tile whose rows fit the stage is deferred the same way: rows and row offsets go to shared
        // memory under tile-local offsets and the look-back is resolved after the NEXT tile's search.
        // Other tiles (heavy fan-out) resolve now and write to global memory directly.
        const bool defer = DEFER && tile_total <= (uint32_t)LI_STAGE;
        uint64_t gbase = 0;
        RowSink sink{reinterpret_cast<swo_row *>(S.st_rows), reinterpret_cast<swo_thick *>(S.st_thick), (uint64_t)LI_STAGE};
        if (!defer) {
            gbase = lookback_resolve_block<LI_NW>(A.desc, tile, (uint64_t)tile_total, S.lb_part, S.lb_found);
            sink = RowSink{A.rows, A.thick, A.rows_cap};
        }

        // row_offset: four consecutive values per thread
        {
            uint32_t off[LI_IPT];
            uint32_t run = (uint32_t)(gbase + excl);  // wraps only if rows >= 2^32 (host rejects)
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) {
                off[k] = run;
                run += (uint32_t)nh[k];
            }
            if (defer) {
                *reinterpret_cast<int4 *>(S.st_off + tid * LI_IPT) = make_int4((int)off[0], (int)off[1], (int)off[2], (int)off[3]);
            } else if (q0 + LI_IPT <= A.n) {
                st_stream_v4(A.row_offset + q0, make_int4((int)off[0], (int)off[1], (int)off[2], (int)off[3]));
            } else {
#pragma unroll
                for (int k = 0; k < LI_IPT; k++)
                    if (q0 + k < A.n) A.row_offset[q0 + k] = off[k];
            }
        }
        if (!defer && tile == A.ntiles - 1 && tid == 0) {
            const uint64_t total = gbase + tile_total;
            A.row_offset[A.n] = (uint32_t)total;
            A.totals[0] = total;
        }

        // rows
        uint64_t run64 = gbase + excl;
        uint32_t pending = 0;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            const uint32_t q = q0 + k;
            if (nh[k] == 1) {
                if (run64 < sink.cap) {
                    const int4 row = make_int4(r_meta[k], r_s[k], r_e[k], (int)q);
                    if (defer) S.st_rows[run64] = row;
                    else st_stream_v4(A.rows + run64, row);
                    if (THICK) {
                        const Thick th = lift_thick_from(T, ey, tc[k], __ldg(A.thick_s + q), __ldg(A.thick_e + q), st[k],
                                                         cd[k].lo, (cm_smem ? S.cmeta[tc[k]] : load_cmeta(T, tc[k])).y);
                        swo_thick t;
                        thick_clamp(th, r_s[k], r_e[k], t.thick_start, t.thick_end);
                        sink.thick[run64] = t;
                    }
                }
            } else if (nh[k] >= 2) {
                // expand_wide_kernel's list, or queued below over several rounds of the work list.  (Expanding two-hit
                // queries here, by their thread, was measured slower on unsorted batches — configs[2] 9.8 -> 11.3 ms:
                // they are scattered, so nearly every warp runs the extra code for one or two of its lanes.)
                pending |= 1u << k;
            }
            run64 += (uint64_t)nh[k];
        }
        converge_at<4>(A);
        // Multi-hit queries on a target-disjoint table are appended to the global list (one atomic per
        // warp) and expanded by expand_wide_kernel, a warp per query with the rows coming straight from
        // registers: no work-list rounds and no block barriers in here, and the block's other warps do
        // not wait for the one with the heaviest query.
        if (!THICK && !defer && T.disjoint && A.wide_ent) {  // (batches with thick columns: see launch_lift_intervals)
            unsigned small = 0;  // queries with at most LI_SMALL candidates go to the list of expand_small_kernel
#pragma unroll
            for (int k = 0; k < LI_IPT; k++)
                if (((pending >> k) & 1u) && nh[k] <= LI_SMALL) small |= 1u << k;
            // (both counts in one scan: big in the low, small in the high half)
            const uint32_t mine = (uint32_t)__popc(pending & ~small) | ((uint32_t)__popc(small) << 16);
            const uint32_t inc = warp_inclusive_scan(mine);
            const uint32_t wtot = __shfl_sync(FULL, inc, 31);
            if (wtot) {  // warp-uniform
                unsigned slot_b = 0, slot_s = 0;
                if (lane_id() == 0) {
                    if (wtot & 0xFFFFu) slot_b = atomicAdd(A.wide_count, wtot & 0xFFFFu);
                    if (wtot >> 16) slot_s = atomicAdd(A.small_count, wtot >> 16);
                }
                slot_b = __shfl_sync(FULL, slot_b, 0) + ((inc - mine) & 0xFFFFu);
                slot_s = __shfl_sync(FULL, slot_s, 0) + ((inc - mine) >> 16);
                uint64_t rb = gbase + excl;
#pragma unroll
                for (int k = 0; k < LI_IPT; k++) {
                    if ((pending >> k) & 1u) {
                        const bool sm = (small >> k) & 1u;
                        const unsigned slot = sm ? slot_s : slot_b;
                        if (slot < (sm ? A.small_cap : A.wide_cap)) {
                            uint4 *const ent = (sm ? A.small_ent : A.wide_ent) + 2 * (size_t)slot;
                            ent[0] = make_uint4(q0 + k, (unsigned)cd[k].lo, (unsigned)nh[k], (uint32_t)rb);
                            ent[1] = make_uint4((unsigned)st[k], (unsigned)en[k], A.strand ? __ldg(A.strand + q0 + k) : 0u, (unsigned)tc[k]);
                            pending &= ~(1u << k);
                        }
                        if (sm) slot_s++;
                        else slot_b++;
                    }
                    rb += (uint64_t)nh[k];
                }
            }
        }
        // the rest (other tables, a full list): rounds of (fill the work list, expand it cooperatively)
        // until none is pending
        if (__syncthreads_or(pending != 0)) {
            for (;;) {
                {
                    uint64_t rb = excl;
#pragma unroll
                    for (int k = 0; k < LI_IPT; k++) {
                        if ((pending >> k) & 1u) {
                            const int slot = atomicAdd(&S.wl_n, 1);
                            if (slot < LI_WL_CAP) {
                                S.wl_q[slot] = q0 + k;
                                S.wl_lo[slot] = cd[k].lo;
                                S.wl_cnt[slot] = cd[k].hi - cd[k].lo;
                                S.wl_off[slot] = (uint32_t)rb;
                                if (cd[k].hi - cd[k].lo > LI_HEAVY_MIN) S.wl_heavy = 1;
                                pending &= ~(1u << k);
                            }
                        }
                        rb += (uint64_t)nh[k];
                    }
                }
                converge_at<5>(A);
                __syncthreads();
                const int nwl = min(S.wl_n, LI_WL_CAP);
                const bool heavy = S.wl_heavy != 0;
                // light and medium fan-out: one warp per query
                for (int e = warp_id(); e < nwl; e += LI_NW) {
                    if (S.wl_cnt[e] <= LI_HEAVY_MIN)
                        expand_query_coop<false, THICK>(A, sink, ey, S.wl_q[e], S.wl_lo[e], S.wl_cnt[e], gbase + S.wl_off[e],
                                                        S.xscr + warp_id() * LI_HEAVY_MIN, nullptr, lane_id(), 32);
                }
                if (heavy) {
                    __syncthreads();
                    // heavy fan-out: the whole block per query
                    for (int e = 0; e < nwl; e++) {
                        if (S.wl_cnt[e] <= LI_HEAVY_MIN) continue;
                        if (S.wl_cnt[e] <= LI_XSCR)
                            expand_query_coop<true, THICK>(A, sink, ey, S.wl_q[e], S.wl_lo[e], S.wl_cnt[e], gbase + S.wl_off[e],
                                                           S.xscr, S.scan, tid, LI_TPB);
                        else
                            expand_query<THICK>(A, sink, ey, S.wl_q[e], S.wl_lo[e], S.wl_cnt[e], gbase + S.wl_off[e], tid, LI_TPB);
                    }
                }
                converge_at<6>(A);
                if (!__syncthreads_or(pending != 0)) break;
                if (tid == 0) {
                    S.wl_n = 0;
                    S.wl_heavy = 0;
                }
                __syncthreads();
            }
        }
        if (defer && tid == 0) {
            S.p_tile = tile;
            S.p_total = tile_total;
            S.p_nq = min((uint32_t)LI_TILE, A.n - (uint32_t)tile * LI_TILE);
        }
    }
}

// liftCoordOnly / liftDirectly for a batch of points (chain.d:604-654).
__global__ void __launch_bounds__(256) lift_points_kernel(const TableView T, uint32_t n, const int32_t *__restrict__ tcid,
                                                          const int32_t *__restrict__ pos, int32_t *__restrict__ out_qcid,
                                                          int32_t *__restrict__ out_pos, uint8_t *__restrict__ hit) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = pos[i];
        const int j = point_hit(T, T.ey, tcid[i], c);
        if (j < 0) {
            out_qcid[i] = -1;
            out_pos[i] = c;
            hit[i] = 0;
        } else {
            const Block b = load_block(T, j);
            out_qcid[i] = b.meta >> 1;
            out_pos[i] = point_coord(b, c);
            hit[i] = 1;
        }
    }
}

}  // namespace swo
