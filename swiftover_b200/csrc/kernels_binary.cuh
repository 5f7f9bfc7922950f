// kernels_binary.cuh — single-pass interval lift on binary records.
//
// Replaces, for a whole batch at once, the reference's per-record sequence
//   cf.lift(contig,start,end)            chain.d:664-713 (+ intervaltree search)
//   liftCoordOnly x2 for thickStart/End  bed.d:134-148, chain.d:637-654
//   sort by qStart, orderStartEnd, strand table, thick clamp   bed.d:156-196
//
// One persistent kernel.  A tile is 1024 consecutive queries (256 threads x 4,
// 128-bit loads).  Each thread searches its four queries (Eytzinger upper levels
// in shared memory + quaternary bottom search; in sorted batches queries 2-4
// gallop forward from the previous lower bound), computes the rows of its
// single-hit queries, the per-query fan-out is block-scanned and the tile
// aggregate is chained to its predecessors with a decoupled look-back resolved
// block-wide (so the variable-length output needs no separate count and scan
// passes: every input byte is read once and every output byte is written
// once); then rows are stored in place.  Queries with >= 2 hits go to a
// shared-memory work list, filled and expanded in rounds: a warp (or the whole
// block for heavy fan-out) writes one sort key per candidate to shared memory,
// sorts only if the keys are not already ascending, and stores every row at its
// sorted position with the cumulative strand byte.
#pragma once
#include "device_util.cuh"
#include "lift_core.cuh"
#include "swiftover_cuda.h"

namespace swo {

constexpr int LI_TPB = 256;
constexpr int LI_IPT = 4;
constexpr int LI_TILE = LI_TPB * LI_IPT;
constexpr int LI_NW = LI_TPB / 32;
constexpr int LI_WL_CAP = 256;      // multi-hit queries queued per tile
constexpr int LI_HEAVY_MIN = 128;   // candidates above this: the whole block expands the query
constexpr int LI_XSCR = 1024;       // shared-memory sort scratch (= LI_NW warp regions of LI_HEAVY_MIN)

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
};

// Expand one multi-hit query: every candidate block computes its own sorted
// position (rank by (key, candidate index)), the cumulative strand byte and
// the thick clamp, and writes its row.  Work is strided by (first, stride) so
// the same code serves one thread, one warp or one block.
__device__ __forceinline__ void expand_query(const LiftArgs &A, const int2 *ey, uint32_t q, int lo, int ncand,
                                             uint64_t rowbase, int first, int stride) {
    const TableView &T = A.T;
    const int start = __ldg(A.start + q), end = __ldg(A.end + q);
    const unsigned char orig = A.strand ? __ldg(A.strand + q) : 0;
    Thick th{0, 0, 0};
    if (A.thick_s) th = lift_thick(T, ey, __ldg(A.tcid + q), __ldg(A.thick_s + q), __ldg(A.thick_e + q));
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
        if (pos < A.rows_cap) {
            const unsigned char sb = A.strand ? strand_at_rank(orig, best_inv, rank, negs) : 0;
            swo_row r;
            r.qcid_strand = rj.qcid | ((int)sb << 16);
            r.start = rj.s;
            r.end = rj.e;
            r.src = q;
            A.rows[pos] = r;
            if (A.thick_s) {
                swo_thick t;
                thick_clamp(th, rj.s, rj.e, t.thick_start, t.thick_end);
                A.thick[pos] = t;
            }
        }
    }
}

// Cooperative, sort-based form for the queued multi-hit queries (one warp, or the block for heavy
// fan-out): one 64-bit key (qStart, candidate, invert) per candidate in shared memory, a bitonic
// sort only if the keys are not already ascending, then every sorted position writes its row;
// the cumulative strand byte comes from a running count of inverted rows (bed.d:173).
template <bool BLOCK>
__device__ __noinline__ void expand_query_coop(const LiftArgs &A, const int2 *ey, uint32_t q, int lo, int ncand,
                                               uint64_t rowbase, unsigned long long *scr, uint32_t *scan_scr, int first,
                                               int stride) {
    const TableView &T = A.T;
    const int start = __ldg(A.start + q), end = __ldg(A.end + q);
    const unsigned char orig = A.strand ? __ldg(A.strand + q) : 0;
    Thick th{0, 0, 0};
    if (A.thick_s) th = lift_thick(T, ey, __ldg(A.tcid + q), __ldg(A.thick_s + q), __ldg(A.thick_e + q));
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
    if (BLOCK) __syncthreads();
    else __syncwarp();
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
                if (BLOCK) __syncthreads();
                else __syncwarp();
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
        if (hit && pos < A.rows_cap) {
            const int c = (int)(((uint32_t)e) >> 1);
            const RowCore r = make_row(load_block(T, lo + c), start, end);
            const unsigned char sb = A.strand ? strand_at_rank(orig, inv_first, j, (int)(carry + ex + neg)) : 0;
            swo_row row;
            row.qcid_strand = r.qcid | ((int)sb << 16);
            row.start = r.s;
            row.end = r.e;
            row.src = q;
            A.rows[pos] = row;
            if (A.thick_s) {
                swo_thick t;
                thick_clamp(th, r.s, r.e, t.thick_start, t.thick_end);
                A.thick[pos] = t;
            }
        }
        carry += tot;
    }
    if (BLOCK) __syncthreads();  // scratch is reused by the next heavy query
}

__global__ void __launch_bounds__(LI_TPB, 4) lift_intervals_kernel(const __grid_constant__ LiftArgs A) {
    __shared__ int s_tile;
    __shared__ uint32_t s_scan[LI_NW + 1];
    __shared__ unsigned long long s_lb_part[LI_NW];
    __shared__ int s_lb_found[LI_NW];
    __shared__ int s_wl_n;
    __shared__ uint32_t s_wl_q[LI_WL_CAP];
    __shared__ int s_wl_lo[LI_WL_CAP];
    __shared__ int s_wl_cnt[LI_WL_CAP];
    __shared__ uint32_t s_wl_off[LI_WL_CAP];
    __shared__ int2 s_ey[ChainTable::EY_CAP];  // upper search levels (Eytzinger order)
    __shared__ unsigned long long s_xscr[LI_XSCR];

    const int tid = threadIdx.x;
    const TableView &T = A.T;
    for (int i = tid; i < T.ey_n; i += LI_TPB) s_ey[i] = __ldg(T.ey + i);
    const int2 *const ey = s_ey;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_tile = (int)atomicAdd(A.ticket, 1u);
            s_wl_n = 0;
        }
        __syncthreads();
        const int tile = s_tile;
        if (tile >= A.ntiles) break;

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
        // row of a single-hit query, computed before the look-back is resolved
        int r_meta[LI_IPT], r_s[LI_IPT], r_e[LI_IPT];
        uint32_t tsum = 0, unm = 0;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            Block b0;
            // sorted batches: same contig and a start that did not move back -> the lower bound only moves
            // forward from the previous query's (a short gallop instead of a full search)
            if (k > 0 && tc[k] == tc[k - 1] && st[k] >= st[k - 1] && tc[k] >= 0 && tc[k] < T.ntc)
                cd[k] = find_candidates_from(T, load_cmeta(T, tc[k]).y, cd[k - 1].lo, st[k], en[k], b0);
            else
                cd[k] = find_candidates_b0(T, ey, tc[k], st[k], en[k], b0);
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
        uint32_t tile_total;
        const uint32_t excl = block_exclusive_scan<uint32_t, LI_NW>(tsum, s_scan, &tile_total);
        if (warp_id() == 0) lookback_publish(A.desc, tile, (uint64_t)tile_total);
        unm = warp_sum(unm);
        if (lane_id() == 0 && unm) atomicAdd((unsigned long long *)(A.totals + 1), (unsigned long long)unm);
        const uint64_t gbase = lookback_resolve_block<LI_NW>(A.desc, tile, (uint64_t)tile_total, s_lb_part, s_lb_found);

        // row_offset: four consecutive values per thread
        uint32_t off[LI_IPT];
        {
            uint32_t run = (uint32_t)(gbase + excl);  // wraps only if rows >= 2^32 (host rejects)
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) {
                off[k] = run;
                run += (uint32_t)nh[k];
            }
            if (q0 + LI_IPT <= A.n) {
                st_stream_v4(A.row_offset + q0, make_int4((int)off[0], (int)off[1], (int)off[2], (int)off[3]));
            } else {
#pragma unroll
                for (int k = 0; k < LI_IPT; k++)
                    if (q0 + k < A.n) A.row_offset[q0 + k] = off[k];
            }
        }
        if (tile == A.ntiles - 1 && tid == 0) {
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
                if (run64 < A.rows_cap) {
                    st_stream_v4(A.rows + run64, make_int4(r_meta[k], r_s[k], r_e[k], (int)q));
                    if (A.thick_s) {
                        const Thick th = lift_thick_from(T, ey, tc[k], __ldg(A.thick_s + q), __ldg(A.thick_e + q), st[k],
                                                         cd[k].lo, load_cmeta(T, tc[k]).y);
                        swo_thick t;
                        thick_clamp(th, r_s[k], r_e[k], t.thick_start, t.thick_end);
                        A.thick[run64] = t;
                    }
                }
            } else if (nh[k] >= 2) {
                pending |= 1u << k;  // queued below, possibly over several rounds of the work list
            }
            run64 += (uint64_t)nh[k];
        }
        // multi-hit queries: rounds of (fill the work list, expand it cooperatively) until none is
        // pending — batches where most queries have many hits need more than one round
        for (;;) {
            {
                uint64_t rb = gbase + excl;
#pragma unroll
                for (int k = 0; k < LI_IPT; k++) {
                    if ((pending >> k) & 1u) {
                        const int slot = atomicAdd(&s_wl_n, 1);
                        if (slot < LI_WL_CAP) {
                            s_wl_q[slot] = q0 + k;
                            s_wl_lo[slot] = cd[k].lo;
                            s_wl_cnt[slot] = cd[k].hi - cd[k].lo;
                            s_wl_off[slot] = (uint32_t)(rb - gbase);
                            pending &= ~(1u << k);
                        }
                    }
                    rb += (uint64_t)nh[k];
                }
            }
        __syncthreads();
        const int nwl = min(s_wl_n, LI_WL_CAP);
        // light and medium fan-out: one warp per query
        for (int e = warp_id(); e < nwl; e += LI_NW) {
            if (s_wl_cnt[e] <= LI_HEAVY_MIN)
                expand_query_coop<false>(A, ey, s_wl_q[e], s_wl_lo[e], s_wl_cnt[e], gbase + s_wl_off[e],
                                         s_xscr + warp_id() * LI_HEAVY_MIN, nullptr, lane_id(), 32);
        }
        if (nwl > 0) __syncthreads();
        // heavy fan-out: the whole block per query
        for (int e = 0; e < nwl; e++) {
            if (s_wl_cnt[e] <= LI_HEAVY_MIN) continue;
            if (s_wl_cnt[e] <= LI_XSCR)
                expand_query_coop<true>(A, ey, s_wl_q[e], s_wl_lo[e], s_wl_cnt[e], gbase + s_wl_off[e], s_xscr, s_scan, tid,
                                        LI_TPB);
            else
                expand_query(A, ey, s_wl_q[e], s_wl_lo[e], s_wl_cnt[e], gbase + s_wl_off[e], tid, LI_TPB);
        }
            if (!__syncthreads_or(pending != 0)) break;
            if (tid == 0) s_wl_n = 0;
            __syncthreads();
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
