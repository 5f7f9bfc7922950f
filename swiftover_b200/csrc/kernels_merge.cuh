// kernels_merge.cuh — the lift kernel for batches sorted by (contig, start): a warp-level
// merge of queries against blocks instead of one tree search per query.
//
// Replaces, like kernels_binary.cuh, the reference's per-record cf.lift(contig,start,end)
// (chain.d:664-713) + sort by qStart + orderStartEnd + strand table (bed.d:156-173) for a whole
// batch.  The reference's own answer to sorted input is the splay tree's "implicit predictive
// caching" (README.md:12-20: the last hit is splayed to the root); here sortedness is used
// explicitly:
//
//   * A warp owns 128 consecutive queries (4 per lane, 128-bit loads).  When they all name the
//     same contig, ONE cooperative search finds the group of 32 blocks that holds the lower bound
//     of the warp's smallest start: 32-ary rounds (one probe per lane, a ballot per round) over
//     the contig's sorted group maxima in shared memory.
//   * That group and the next one (64 blocks) are fetched with coalesced loads of the two key
//     arrays and merged against the 128 queries: block j is broadcast (shuffle) and every lane
//     counts, for its four queries, the blocks with tEnd <= start (they lie before the query: lo)
//     and the blocks with tStart < end (hi).  The loop ends as soon as no query reaches block j,
//     i.e. after (blocks spanned by the warp's queries) + 1 rounds — 2 on BASELINE configs[1],
//     where a block holds ~1,850 queries.  A query that runs past the window is finished by its
//     thread (four lockstep gallops); a warp with several contigs (contig boundaries, the batch
//     tail) takes the per-thread search of kernels_binary.cuh.  Any order inside the warp is fine.
//   * The first candidate's row is computed without branches for every query (four block loads in
//     flight per lane) and kept for single-hit queries.
//   * Variable-length output without a count pass: a tile is MG_NW warps x 128 queries.  Every warp
//     adds its row total to the tile's and the group's look-back words right after its search
//     (device_util.cuh: two-level look-back with per-warp publication), parks its results in a
//     shared-memory ring, and the CTA stores a tile MG_DEPTH iterations later, when its exclusive
//     prefix is available without waiting (the first warp to finish its search resolves it in front
//     of the one barrier at which the warps exchange their totals).
//   * Tiles are dealt round-robin (tile = CTA + k * grid) and the kernel is launched COOPERATIVELY
//     (all CTAs resident: a CTA may wait for any other's publication).  A ticket counter would
//     not need that, but it puts an atomic round trip and a second barrier on every iteration
//     and — worse — any ticket drawn before its search can start at once delays that tile's
//     publication, every later tile's look-back waits for it, and the waiting CTAs hold
//     tickets too: a convoy (measured 2-4x slower whenever a ticket was drawn ahead of a barrier,
//     of the stores or of a look-back).  With the static deal the next tile is known, so its
//     queries are in shared memory before they are needed: one thread arms an mbarrier and issues
//     three 1-D bulk copies (TMA, cp.async.bulk: UBLKCP in the SASS) two iterations ahead.
//   * Queries with 2..MG_SMALL candidate blocks are expanded by their own thread at store time
//     (sequentially when their rows are already in qStart order, forwards or backwards).  Larger
//     ones — whose cost would throw the static deal off balance — are appended to a list
//     {query, first block, candidates, row position} and expanded by expand_wide_kernel, a warp
//     per query, after this kernel (list full: by their warp, on the spot).
// Only for target-disjoint tables (every UCSC over.chain; the host sends other tables to the
// per-query-search variant): candidates == hits, and pmaxKey[j] == tEnd[j].
#pragma once
#include "kernels_binary.cuh"

namespace swo {

#ifdef MG_TRACE
// debug build: per tile {draw, last warp published, resolve begin, resolve end} in ns (globaltimer)
__device__ unsigned long long *mg_trace;
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MG_TR(tile, slot) do { if (mg_trace) mg_trace[(size_t)(tile) * 8 + (slot)] = gtimer(); } while (0)
#define MG_TRMAX(tile, slot) do { if (mg_trace) atomicMax(&mg_trace[(size_t)(tile) * 8 + (slot)], gtimer()); } while (0)
#define MG_TROR(tile, slot, v) do { if (mg_trace) atomicOr(&mg_trace[(size_t)(tile) * 8 + (slot)], (unsigned long long)(v)); } while (0)
#else
#define MG_TROR(tile, slot, v) do { } while (0)
#define MG_TR(tile, slot) do { } while (0)
#define MG_TRMAX(tile, slot) do { } while (0)
#endif

// A pending query's three words: w0 = code << 24 | row meta; code = hits (0, 1, ... 254; 255 = "255 or
// more": the count is in w2).  One hit: w0 also holds the row's {qcid, strand byte}, w1/w2 its
// {start, end}.  Several hits: w1 = first candidate block, w2 = candidates (= hits).
constexpr int MG_SMALL = 16;   // candidates up to which a multi-hit query is expanded by its own thread
constexpr int MG_XS = 128;     // per-warp sort scratch in shared memory (entries)
constexpr int MG_GXS = 2048;   // per-warp sort scratch in global memory (entries)
#ifndef MG_CTAS_OVR
#define MG_CTAS_OVR 6
#endif
constexpr int MG_CTAS = MG_CTAS_OVR;  // resident CTAs per SM the register budget is set for
#ifndef MG_DEPTH_OVR
#define MG_DEPTH_OVR 2
#endif
constexpr int MG_DEPTH = MG_DEPTH_OVR;  // a tile is stored MG_DEPTH iterations after it was searched and published
constexpr int MG_RING = MG_DEPTH + 1;
static_assert(MG_DEPTH >= 1, "ring depth");
constexpr int MG_QPW = 32 * LI_IPT;  // queries per warp and tile
#ifndef MG_NW_OVR
#define MG_NW_OVR 4
#endif
constexpr int MG_NW = MG_NW_OVR;        // warps per CTA
constexpr int MG_TPB = MG_NW * 32;
constexpr int MG_TILE = MG_NW * MG_QPW;  // queries per tile

struct MergeSmem {
    int4 q[2][3][MG_TPB];                  // the queries of the tile in work / the one after the next: {tcid}, {start}, {end}
    int4 carry[MG_RING][3][MG_TPB];        // a thread's four pending queries per tile in flight: {w0}, {w1}, {w2}
    uint32_t poff[MG_RING][MG_TPB];        //   and the tile-local offset of its first row
    int4 cmeta[LI_CM_CAP];                 // per-contig {first block, end block, sample offset, samples}
    int32_t smp[ChainTable::EY_CAP];       // group maxima of pmaxKey in SORTED order (per contig)
    unsigned long long xscr[MG_NW][MG_XS]; // sort scratch of a warp's cooperative expansion
    unsigned long long gbase[2];           // exclusive prefix of the tile being stored (alternating)
    uint64_t qbar[2];                      // mbarriers of the two query stages
    uint32_t wtot[2][MG_NW];               // rows of each warp's 128 queries (alternating)
    unsigned arrive;                       // warps that have finished their searches (running count)
};

// Multi-hit queries left for expand_wide_kernel: {query, first candidate block, candidates, row position}
struct WideList {
    unsigned *count;  // appended so far (may run past cap: those were expanded on the spot)
    uint4 *ent;
    unsigned cap;
};

// first i in [lo, hi] with key[i] > x, where hi is either a position known to hold a key > x or
// the end of the array ("none").  All 32 lanes call it with the same arguments; one probe per
// lane and round.
__device__ __forceinline__ int coop_lower_gt(const int32_t *__restrict__ key, int lo, int hi, int x) {
    const int lane = lane_id();
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) >> 5;
        int p = lo + (lane + 1) * step - 1;
        p = p < hi ? p : hi - 1;
        const unsigned b = __ballot_sync(FULL, key[p] > x);
        if (b == 0u) return hi;
        const int f = __ffs(b) - 1;
        const int top = lo + (f + 1) * step - 1;
        hi = top < hi ? top : hi - 1;
        lo += f * step;
    }
    const int n = hi - lo;
    const unsigned b = __ballot_sync(FULL, lane < n && key[lo + lane] > x);
    return lo + (b ? __ffs(b) - 1 : n);
}

// One block of the window against the lane's four queries: lo counts the blocks that end at or
// before the query's start, hi the blocks that start before its end.  Returns whether any query
// of the WARP still reaches this block.  Inline PTX: the compiler turns `if (c) x++` into a
// select plus register copies (39 SASS instructions per round instead of 24).
__device__ __forceinline__ bool merge_step(int pmj, int tsj, const int (&st)[LI_IPT], const int (&en)[LI_IPT], int (&lo)[LI_IPT],
                                           int (&hi)[LI_IPT]) {
    unsigned any;
    asm volatile(
        "{\n"
        ".reg .pred p0, p1, p2, p3, q0, q1, q2, q3, a;\n"
        "setp.le.s32 p0, %9, %11;\n"
        "setp.le.s32 p1, %9, %12;\n"
        "setp.le.s32 p2, %9, %13;\n"
        "setp.le.s32 p3, %9, %14;\n"
        "setp.lt.s32 q0, %10, %15;\n"
        "setp.lt.s32 q1, %10, %16;\n"
        "setp.lt.s32 q2, %10, %17;\n"
        "setp.lt.s32 q3, %10, %18;\n"
        "@p0 add.s32 %0, %0, 1;\n"
        "@p1 add.s32 %1, %1, 1;\n"
        "@p2 add.s32 %2, %2, 1;\n"
        "@p3 add.s32 %3, %3, 1;\n"
        "@q0 add.s32 %4, %4, 1;\n"
        "@q1 add.s32 %5, %5, 1;\n"
        "@q2 add.s32 %6, %6, 1;\n"
        "@q3 add.s32 %7, %7, 1;\n"
        "or.pred a, q0, q1;\n"
        "or.pred a, a, q2;\n"
        "or.pred a, a, q3;\n"
        "vote.sync.any.pred a, a, 0xffffffff;\n"
        "selp.u32 %8, 1, 0, a;\n"
        "}\n"
        : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "=r"(any)
        : "r"(pmj), "r"(tsj), "r"(st[0]), "r"(st[1]), "r"(st[2]), "r"(st[3]), "r"(en[0]), "r"(en[1]), "r"(en[2]), "r"(en[3]));
    return any != 0u;
}

// first j in [base, end) with key[j] > x (GE: >= x), else end — for up to four searches of one
// thread at once (bit k of `act`): an exponential probe from base, then bisection, with the loads
// of the four searches issued together in every round (one dependent-load chain instead of four).
struct Int4x {
    int v[LI_IPT];
};
template <bool GE>
__device__ __noinline__ Int4x gallop4(const int32_t *__restrict__ key, int end, Int4x base, Int4x x, unsigned act) {
    int a[LI_IPT], b[LI_IPT];
    unsigned need = 0;
#pragma unroll
    for (int k = 0; k < LI_IPT; k++) {
        a[k] = base.v[k];
        b[k] = end;
        if (((act >> k) & 1u) && a[k] < end) need |= 1u << k;
    }
    for (int step = 8; need; step <<= 1) {
        int p[LI_IPT], v[LI_IPT];
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            p[k] = a[k] + step - 1 < end ? a[k] + step - 1 : end - 1;
            v[k] = ((need >> k) & 1u) ? __ldg(key + p[k]) : 0;
        }
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            if ((need >> k) & 1u) {
                if (GE ? v[k] >= x.v[k] : v[k] > x.v[k]) {
                    b[k] = p[k];
                    need &= ~(1u << k);
                } else {
                    a[k] = p[k] + 1;
                    if (a[k] >= end) need &= ~(1u << k);
                }
            }
        }
    }
    for (;;) {
        unsigned open = 0;
        int m[LI_IPT], v[LI_IPT];
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            const bool o = ((act >> k) & 1u) && a[k] < b[k];
            m[k] = (int)(((unsigned)a[k] + (unsigned)b[k]) >> 1);
            v[k] = o ? __ldg(key + m[k]) : 0;
            open |= o ? 1u << k : 0u;
        }
        if (!open) break;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            if ((open >> k) & 1u) {
                if (GE ? v[k] >= x.v[k] : v[k] > x.v[k]) b[k] = m[k];
                else a[k] = m[k] + 1;
            }
        }
    }
    Int4x r;
#pragma unroll
    for (int k = 0; k < LI_IPT; k++) r.v[k] = a[k];
    return r;
}

// The per-query search of kernels_binary.cuh for a thread's four queries (re-read from global
// memory: this is the rare path — warps that straddle a contig boundary and the batch tail).
struct LoHi4 {
    int lo[LI_IPT], hi[LI_IPT];
};
__device__ __noinline__ LoHi4 search_each(const LiftArgs &A, uint32_t q0) {
    const TableView &T = A.T;
    LoHi4 r;
    int ptc = -1, pst = 0;
#pragma unroll
    for (int k = 0; k < LI_IPT; k++) {
        const bool ok = q0 + k < A.n;
        const int tc = ok ? __ldg(A.tcid + q0 + k) : -1;
        const int st = ok ? __ldg(A.start + q0 + k) : 0, en = ok ? __ldg(A.end + q0 + k) : 0;
        Block b0;
        Cand cd{0, 0};
        if (tc >= 0 && tc < T.ntc) {
            const int4 cm = load_cmeta(T, tc);
            if (k > 0 && tc == ptc && st >= pst) cd = find_candidates_from(T, cm.y, r.lo[k - 1], st, en, b0);
            else cd = find_candidates_cm(T, T.ey, cm, st, en, b0);
        }
        r.lo[k] = cd.lo;
        r.hi[k] = cd.hi;
        ptc = tc;
        pst = st;
    }
    return r;
}
// A multi-hit query with few candidates, expanded by its own thread.  The candidates of a query are
// consecutive blocks, nearly always of one chain: their sort keys (raw qStart, bed.d:156) then run
// ascending ('+' chains) or strictly descending ('-' chains) and the rows are written in one
// sequential pass, forwards or backwards, with the running strand flip (bed.d:173); any other
// order takes the rank-by-comparison form.
__device__ __noinline__ void expand_small(const LiftArgs &A, uint32_t q, int lo, int cnt, uint64_t pos) {
    const TableView &T = A.T;
    const RowSink sink{A.rows, A.thick, A.rows_cap};
    const int start = __ldg(A.start + q), end = __ldg(A.end + q);
    if (cnt == 2 && T.disjoint) {  // the usual case: an interval across one gap
        const RowCore a = make_row(load_block(T, lo), start, end), b = make_row(load_block(T, lo + 1), start, end);
        const bool swap = b.key < a.key;
        const RowCore &r0 = swap ? b : a, &r1 = swap ? a : b;
        unsigned char s0 = 0, s1 = 0;
        if (A.strand) {
            s0 = strand_flip(__ldg(A.strand + q), r0.inv);
            s1 = strand_flip(s0, r1.inv);  // bed.d:173: the flips accumulate
        }
        if (pos < A.rows_cap) st_stream_v4(A.rows + pos, make_int4(r0.qcid | ((int)s0 << 16), r0.s, r0.e, (int)q));
        if (pos + 1 < A.rows_cap) st_stream_v4(A.rows + pos + 1, make_int4(r1.qcid | ((int)s1 << 16), r1.s, r1.e, (int)q));
        return;
    }
    bool asc = true, desc = true, first = true;
    int prev = 0, nhit = 0;
    for (int c = 0; c < cnt; c++) {
        const Block b = load_block(T, lo + c);
        if (!T.disjoint && !block_hits(b, start)) continue;
        const int key = make_row(b, start, end).key;
        if (!first) {
            asc &= key >= prev;
            desc &= key < prev;
        }
        first = false;
        prev = key;
        nhit++;
    }
    if (!asc && !desc) {
        expand_query<false>(A, sink, T.ey, q, lo, cnt, pos, 0, 1);
        return;
    }
    const unsigned char orig = A.strand ? __ldg(A.strand + q) : 0;
    int rank = 0, negs = 0, inv_first = 1;
    for (int i = 0; i < cnt; i++) {
        const int c = asc ? i : cnt - 1 - i;
        const Block b = load_block(T, lo + c);
        if (!T.disjoint && !block_hits(b, start)) continue;
        const RowCore r = make_row(b, start, end);
        if (rank == 0) inv_first = r.inv;
        negs += r.inv < 0 ? 1 : 0;
        const uint64_t p = pos + (uint64_t)rank;
        if (p < A.rows_cap) {
            const unsigned char sb = A.strand ? strand_at_rank(orig, inv_first, rank, negs) : 0;
            st_stream_v4(A.rows + p, make_int4(r.qcid | ((int)sb << 16), r.s, r.e, (int)q));
        }
        rank++;
    }
}

// one multi-hit query by a whole warp: sort keys in shared memory, in the warp's global scratch, or
// (beyond that) every candidate ranks itself against all others
__device__ __forceinline__ void expand_wide(const LiftArgs &A, const RowSink &sink, uint32_t q, int lo, int cnt, uint64_t pos,
                                            unsigned long long *xs, unsigned long long *gxs) {
    if (cnt <= MG_XS) expand_query_coop<false, false>(A, sink, A.T.ey, q, lo, cnt, pos, xs, nullptr, lane_id(), 32);
    else if (cnt <= MG_GXS) expand_query_coop<false, false>(A, sink, A.T.ey, q, lo, cnt, pos, gxs, nullptr, lane_id(), 32);
    else expand_query<false>(A, sink, A.T.ey, q, lo, cnt, pos, lane_id(), 32);
}

template <bool STRAND>
__global__ void __launch_bounds__(MG_TPB, MG_CTAS) lift_sorted_kernel(const __grid_constant__ LiftArgs A, const __grid_constant__ Lb2 L,
                                                                      const __grid_constant__ WideList W) {
    extern __shared__ __align__(16) unsigned char li_smem_raw[];
    MergeSmem &S = *reinterpret_cast<MergeSmem *>(li_smem_raw);
    if (__ldg(A.sorted_flag) == 0u) return;

    const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const int G = (int)gridDim.x;
    const TableView &T = A.T;
    const bool cm_smem = T.ntc <= LI_CM_CAP;
    if (cm_smem)
        for (int i = tid; i < T.ntc; i += MG_TPB) S.cmeta[i] = __ldg(T.cmeta + i);
    // group maxima in sorted order: sample i of contig c is the last key of its i-th full group
    for (int c = w; c < T.ntc; c += MG_NW) {
        const int4 cm = __ldg(T.cmeta + c);
        for (int i = lane; i < cm.w; i += 32) S.smp[cm.z + i] = __ldg(T.pmaxKey + cm.x + ((i + 1) << T.ey_shift) - 1);
    }
    const int last_block = __ldg(&T.cmeta[T.ntc - 1].y) - 1;
    // a tile's queries come through shared memory when the whole tile exists (not the batch's last, partial tile)
    const auto tile_staged = [&](long long t) { return (t + 1) * MG_TILE <= (long long)A.n; };
    const auto stage_tile = [&](int t, int st) {  // one thread: arm the stage's barrier, issue the three copies
        mbar_expect_tx(&S.qbar[st], 3u * MG_TILE * 4u);
        bulk_g2s(S.q[st][0], A.tcid + (size_t)t * MG_TILE, MG_TILE * 4u, &S.qbar[st]);
        bulk_g2s(S.q[st][1], A.start + (size_t)t * MG_TILE, MG_TILE * 4u, &S.qbar[st]);
        bulk_g2s(S.q[st][2], A.end + (size_t)t * MG_TILE, MG_TILE * 4u, &S.qbar[st]);
    };
    if (tid == 0) {
        S.arrive = 0;
        mbar_init(&S.qbar[0], 1);
        mbar_init(&S.qbar[1], 1);
        mbar_init_fence();
        const long long t0 = blockIdx.x, t1 = t0 + G;
        if (t0 < A.ntiles && tile_staged(t0)) stage_tile((int)t0, 0);
        if (t1 < A.ntiles && tile_staged(t1)) stage_tile((int)t1, 1);
    }
    __syncthreads();

    const RowSink sink{A.rows, A.thick, A.rows_cap};
    unsigned long long *const gxs = A.gscr + ((size_t)blockIdx.x * MG_NW + w) * MG_GXS;
    const uint32_t cap32 = A.rows_cap > 0xffffffffull ? 0xffffffffu : (uint32_t)A.rows_cap;
    int slot = 0;  // ring slot of the current tile; the tile stored in this iteration sits in slot+1 (mod MG_RING)
    uint32_t unm_acc = 0;

    // iteration k: search tile blockIdx.x + k*G (while there is one) and store the tile of iteration k - MG_DEPTH
    const int my_tiles = (int)blockIdx.x < A.ntiles ? (A.ntiles - (int)blockIdx.x + G - 1) / G : 0;
    for (int k = 0; k < my_tiles + MG_DEPTH; k++) {
        const long long tile_ll = (long long)blockIdx.x + (long long)k * G;
        const bool live = k < my_tiles;
        const int tile = live ? (int)tile_ll : A.ntiles;
        const int par = k & 1;
        const int o_tile = k >= MG_DEPTH ? (int)(tile_ll - (long long)MG_DEPTH * G) : -1;
        uint32_t tsum = 0;
        if (live) {
            if (tid == 0) MG_TR(tile, 0);
            const uint32_t q0 = (uint32_t)tile * MG_TILE + (uint32_t)tid * LI_IPT;
            int tc[LI_IPT], st[LI_IPT], en[LI_IPT];
            const bool staged = tile_staged(tile_ll);
            const bool warp_full = staged;
            if (staged) {
                mbar_wait(&S.qbar[par], (k >> 1) & 1);  // this stage's (k/2)-th use
                const int4 a = S.q[par][0][tid], b = S.q[par][1][tid], c = S.q[par][2][tid];
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
                    unm_acc -= ok ? 0u : 1u;  // counted as "no hit" below
                }
            }
            const int tc0 = __shfl_sync(FULL, tc[0], 0);
            const bool same = __all_sync(FULL, tc[0] == tc0 && tc[1] == tc0 && tc[2] == tc0 && tc[3] == tc0) &&
                              tc0 >= 0 && tc0 < T.ntc;
            int lo[LI_IPT], hi[LI_IPT];
            if (lane == 0) MG_TRMAX(tile, 5);  // queries have arrived
            if (lane == 0 && !same) MG_TROR(tile, 4, 1);
            if (same) {
                const int4 cm = cm_smem ? S.cmeta[tc0] : load_cmeta(T, tc0);
                const int m01 = st[0] < st[1] ? st[0] : st[1], m23 = st[2] < st[3] ? st[2] : st[3];
                const int smin = __reduce_min_sync(FULL, m01 < m23 ? m01 : m23);
                // The warp's window: 64 consecutive blocks that contain the lower bound lo_w of its smallest
                // start.  Groups of 32 blocks (every table up to 65,536 blocks): the group found in the
                // shared-memory samples and the one behind it, no probe of the key array in between;
                // larger groups: the window starts at lo_w itself (one more search round trip).
                const int gi = coop_lower_gt(S.smp + cm.z, 0, cm.w, smin);
                const int g_lo = cm.x + (gi << T.ey_shift);
                int wb = g_lo;
                if (T.ey_shift != 5) wb = coop_lower_gt(T.pmaxKey, g_lo, gi < cm.w ? g_lo + (1 << T.ey_shift) - 1 : cm.y, smin);
                const int ja = wb + lane, jb = wb + 32 + lane;
                const int pmA = ja < cm.y ? __ldg(T.pmaxKey + ja) : 0x7fffffff, tsA = ja < cm.y ? __ldg(T.tStartKey + ja) : 0x7fffffff;
                const int pmB = jb < cm.y ? __ldg(T.pmaxKey + jb) : 0x7fffffff, tsB = jb < cm.y ? __ldg(T.tStartKey + jb) : 0x7fffffff;
                // blocks of the window that end at or before the smallest start lie before every query
                int jj = __popc(__ballot_sync(FULL, pmA <= smin));
                if (jj == 32) jj += __popc(__ballot_sync(FULL, pmB <= smin));
                const int w_end = wb + 64;
#pragma unroll
                for (int k = 0; k < LI_IPT; k++) lo[k] = hi[k] = wb + jj;
#pragma unroll 1
                for (; jj < 64; jj++) {
                    const int pmj = __shfl_sync(FULL, jj < 32 ? pmA : pmB, jj), tsj = __shfl_sync(FULL, jj < 32 ? tsA : tsB, jj);
                    if (!merge_step(pmj, tsj, st, en, lo, hi)) break;
                }
                if (lane == 0) MG_TRMAX(tile, 6);  // window merged
                if (jj == 64) {  // some query runs past the window: its thread finishes it (four searches in lockstep)
                    if (lane == 0) MG_TROR(tile, 4, 2);
                    unsigned act_lo = 0, act_hi = 0;
#pragma unroll
                    for (int k = 0; k < LI_IPT; k++) {
                        act_hi |= hi[k] == w_end ? 1u << k : 0u;
                        act_lo |= lo[k] == w_end ? 1u << k : 0u;  // implies the same bit of act_hi
                    }
                    if (act_lo) {
                        const Int4x r = gallop4<false>(T.pmaxKey, cm.y, Int4x{{lo[0], lo[1], lo[2], lo[3]}},
                                                       Int4x{{st[0], st[1], st[2], st[3]}}, act_lo);
#pragma unroll
                        for (int k = 0; k < LI_IPT; k++) lo[k] = r.v[k];
                    }
                    if (act_hi) {
                        Int4x base;
#pragma unroll
                        for (int k = 0; k < LI_IPT; k++) base.v[k] = lo[k] > w_end ? lo[k] : w_end;
                        const Int4x r = gallop4<true>(T.tStartKey, cm.y, base, Int4x{{en[0], en[1], en[2], en[3]}}, act_hi);
#pragma unroll
                        for (int k = 0; k < LI_IPT; k++)
                            if ((act_hi >> k) & 1u) hi[k] = r.v[k];
                    }
                    __syncwarp();
                }
            } else {
                // several contigs in the warp (or the batch tail): one search per query
                const LoHi4 r = search_each(A, q0);
#pragma unroll
                for (int k = 0; k < LI_IPT; k++) lo[k] = r.lo[k], hi[k] = r.hi[k];
                converge(A);
            }
            // Hits, and the row of the first candidate — computed without branches for every query (four
            // independent block loads in flight); it is only kept for single-hit queries.
            uchar4 sb4 = make_uchar4(0, 0, 0, 0);
            if (STRAND) {
                if (warp_full) sb4 = __ldg(reinterpret_cast<const uchar4 *>(A.strand + q0));
                else {
                    sb4.x = q0 + 0 < A.n ? A.strand[q0 + 0] : 0;
                    sb4.y = q0 + 1 < A.n ? A.strand[q0 + 1] : 0;
                    sb4.z = q0 + 2 < A.n ? A.strand[q0 + 2] : 0;
                    sb4.w = q0 + 3 < A.n ? A.strand[q0 + 3] : 0;
                }
            }
            const unsigned char sbk[LI_IPT] = {sb4.x, sb4.y, sb4.z, sb4.w};
            Block bk[LI_IPT];
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) bk[k] = load_block(T, lo[k] < last_block ? lo[k] : last_block);
            int w0[LI_IPT], w1[LI_IPT], w2[LI_IPT];
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) {
                const int cnt = hi[k] > lo[k] ? hi[k] - lo[k] : 0;
                const RowCore r = make_row(bk[k], st[k], en[k]);
                const int meta = STRAND ? (r.qcid | ((int)strand_flip(sbk[k], r.inv) << 16)) : r.qcid;
                const bool one = cnt == 1;
                w0[k] = one ? ((1 << 24) | meta) : ((cnt < 255 ? cnt : 255) << 24);
                w1[k] = one ? r.s : lo[k];
                w2[k] = one ? r.e : cnt;
                tsum += (uint32_t)cnt;
                unm_acc += cnt == 0 ? 1u : 0u;
            }
            if (lane == 0) MG_TRMAX(tile, 7);  // rows computed
            // the thread's results wait in its ring slot until the tile is stored
            S.carry[slot][0][tid] = make_int4(w0[0], w0[1], w0[2], w0[3]);
            S.carry[slot][1][tid] = make_int4(w1[0], w1[1], w1[2], w1[3]);
            S.carry[slot][2][tid] = make_int4(w2[0], w2[1], w2[2], w2[3]);
        }
        // the warp's rows: published at once (no tile waits for this CTA any more); hand-over at the barrier
        const uint32_t inc = warp_inclusive_scan(tsum);
        if (lane == 31) {
            S.wtot[par][w] = inc;
            if (live) lookback2_add(L, tile, (uint64_t)inc);
            if (live) MG_TRMAX(tile, 1);
        }
        // the first warp to get here resolves the old tile's prefix while the others finish their searches
        int resolver = 0;
        if (o_tile >= 0) {
            if (lane == 0) resolver = (atomicAdd(&S.arrive, 1u) & (MG_NW - 1)) == 0u;
            resolver = __shfl_sync(FULL, resolver, 0);
        }
        if (resolver) {
            if (lane == 0) MG_TR(o_tile, 2);
            const uint64_t gb = lookback2_resolve<MG_NW>(L, o_tile);
            if (lane == 0) MG_TR(o_tile, 3);
            if (lane == 0) S.gbase[par] = gb;
        }
        __syncthreads();  // the only barrier of an iteration
        // every thread has its queries of this stage in registers: the tile after the next goes in
        if (tid == 0) {
            const long long t2 = tile_ll + 2ll * G;
            if (t2 < A.ntiles && tile_staged(t2)) {
                fence_proxy_async();
                stage_tile((int)t2, par);
            }
        }
        uint32_t wbase = 0;
#pragma unroll
        for (int i = 0; i < MG_NW; i++) wbase += i < w ? S.wtot[par][i] : 0u;
        const int oslot = slot + 1 == MG_RING ? 0 : slot + 1;  // the tile searched MG_DEPTH iterations ago

        // store that tile: row offsets, single-hit rows, multi-hit queries
        if (o_tile >= 0) {
            const uint64_t gb = S.gbase[par];
            const uint32_t p_q0 = (uint32_t)o_tile * MG_TILE + (uint32_t)tid * LI_IPT;
            const int4 c0 = S.carry[oslot][0][tid], c1 = S.carry[oslot][1][tid], c2 = S.carry[oslot][2][tid];
            const int p0[LI_IPT] = {c0.x, c0.y, c0.z, c0.w}, p1[LI_IPT] = {c1.x, c1.y, c1.z, c1.w},
                      p2[LI_IPT] = {c2.x, c2.y, c2.z, c2.w};
            // row indices are 32-bit (the host rejects batches with 2^32 rows or more)
            uint32_t pos[LI_IPT + 1];
            pos[0] = (uint32_t)gb + S.poff[oslot][tid];
            unsigned multi = 0;
#pragma unroll
            for (int k2 = 0; k2 < LI_IPT; k2++) {
                const uint32_t code = (uint32_t)p0[k2] >> 24;
                pos[k2 + 1] = pos[k2] + (code == 255u ? (uint32_t)p2[k2] : code);
                multi |= code >= 2u ? 1u << k2 : 0u;
                if (code == 1u && pos[k2] < cap32)
                    st_stream_v4(A.rows + pos[k2], make_int4(p0[k2] & 0xffffff, p1[k2], p2[k2], (int)(p_q0 + k2)));
            }
            if (p_q0 + LI_IPT <= A.n) {
                st_stream_v4(A.row_offset + p_q0, make_int4((int)pos[0], (int)pos[1], (int)pos[2], (int)pos[3]));
            } else {
#pragma unroll
                for (int k2 = 0; k2 < LI_IPT; k2++)
                    if (p_q0 + k2 < A.n) A.row_offset[p_q0 + k2] = pos[k2];
            }
            if (o_tile == A.ntiles - 1 && tid == MG_TPB - 1) {  // the batch's last thread: its last position is the total
                const uint64_t all = gb + (uint64_t)(uint32_t)(pos[LI_IPT] - (uint32_t)gb);
                A.row_offset[A.n] = (uint32_t)all;
                A.totals[0] = all;
            }
            if (__any_sync(FULL, multi != 0u)) {
                unsigned wide = 0;  // list full: queries for the whole warp
#pragma unroll
                for (int k2 = 0; k2 < LI_IPT; k2++) {
                    if ((multi >> k2) & 1u) {
                        if (p2[k2] <= MG_SMALL) {
                            expand_small(A, p_q0 + k2, p1[k2], p2[k2], gb + (uint64_t)(uint32_t)(pos[k2] - (uint32_t)gb));
                        } else {
                            const unsigned e = atomicAdd(W.count, 1u);
                            if (e < W.cap) W.ent[e] = make_uint4(p_q0 + k2, (unsigned)p1[k2], (unsigned)p2[k2], pos[k2]);
                            else wide |= 1u << k2;
                        }
                    }
                }
                converge(A);
                if (__any_sync(FULL, wide != 0u)) {
#pragma unroll 1
                    for (int k2 = 0; k2 < LI_IPT; k2++) {
                        unsigned m = __ballot_sync(FULL, (wide >> k2) & 1u);
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t q = __shfl_sync(FULL, p_q0, src) + (uint32_t)k2;
                            const int qlo = __shfl_sync(FULL, p1[k2], src), qcnt = __shfl_sync(FULL, p2[k2], src);
                            const uint64_t qpos = gb + (uint64_t)(uint32_t)(__shfl_sync(FULL, pos[k2], src) - (uint32_t)gb);
                            expand_wide(A, sink, q, qlo, qcnt, qpos, S.xscr[w], gxs);
                            converge(A);
                        }
                    }
                }
            }
        }
        // the current tile enters the ring
        S.poff[slot][tid] = wbase + inc - tsum;
        __syncwarp();
        slot = oslot;
    }
    const uint32_t unm = warp_sum(unm_acc);
    if (lane == 0 && unm) atomicAdd((unsigned long long *)(A.totals + 1), (unsigned long long)unm);
}

// The queries lift_sorted_kernel left on the list: a warp per query.  Row positions are 32-bit
// (the host rejects batches with 2^32 rows or more).
__global__ void __launch_bounds__(256) expand_wide_kernel(const __grid_constant__ LiftArgs A, const __grid_constant__ WideList W) {
    __shared__ unsigned long long xs[8][MG_XS];
    if (__ldg(A.sorted_flag) == 0u) return;
    const unsigned n = min(*W.count, W.cap);
    const int w = warp_id();
    const RowSink sink{A.rows, A.thick, A.rows_cap};
    unsigned long long *const gxs = A.gscr + ((size_t)blockIdx.x * 8 + w) * MG_GXS;
    for (unsigned e = blockIdx.x * 8 + w; e < n; e += gridDim.x * 8) {
        const uint4 v = W.ent[e];
        expand_wide(A, sink, v.x, (int)v.y, (int)v.z, (uint64_t)v.w, xs[w], gxs);
        converge(A);
    }
}

}  // namespace swo
