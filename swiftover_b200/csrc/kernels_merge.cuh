// kernels_merge.cuh — the lift kernels for batches sorted by (contig, start): a warp-level
// merge of queries against blocks instead of one tree search per query, as a count pass, a
// prefix scan and a write pass.
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
//     where a block holds ~1,850 queries.  A query that runs past the window is finished by the
//     whole warp (32-ary rounds over the key arrays); a warp with several contigs (contig
//     boundaries, the batch tail) merges contig by contig.  Any order inside the warp is fine.
//   * The 64-block window stays in the warp's registers from tile to tile (struct Window): tiles are
//     dealt in chunks (struct Deal), a warp's next tile is 1,024 queries further on and usually
//     still inside the window, or one 32-block slide away — no search, no key loads.
//   * Variable-length output, three launches (the count / scan / write scheme of the task's
//     north star): merge_count_kernel stores every warp tile's row total, scan_tiles_kernel{1,2}
//     turn the totals into exclusive prefixes, merge_write_kernel repeats the merge and stores row
//     offsets and rows at their final places.  The queries are read twice (24 instead of 12 bytes
//     per query of DRAM traffic) — measured against single-pass forms of the same merge with a
//     chained look-back (block tiles or warp tiles, tickets or a static deal, 1-3 tiles of
//     deferral; all in this file's history) the two-pass form wins: there every tile's store waits
//     for the slowest of the ~500-10,000 tiles in flight before it, a ticket held for anything but
//     its own search starts a convoy, and the per-tile chain ticket -> load -> search -> publish ->
//     resolve -> store leaves the SMs idle; here no warp ever waits for another.
//   * The first candidate's row is computed without branches for every query (four block loads in
//     flight per lane) and kept for single-hit queries.  A query with two candidate blocks (an
//     interval across one gap) is expanded by its own thread, straight-line.  Larger ones are
//     appended to a list {query, first block, candidates, row position, start, end, strand} and
//     expanded by expand_wide_kernel, a warp per query and without sorting (expand_warp_runs,
//     kernels_binary.cuh; list full: by their warp, on the spot).  lift_intervals_kernel appends the
//     multi-hit queries of unsorted batches to the same list.
//   * Every loop of the merge kernels has a warp-uniform trip count (see merge_window for why).
// Only for target-disjoint tables (every UCSC over.chain; the host sends other tables to the
// per-query-search variant): candidates == hits, and pmaxKey[j] == tEnd[j].
#pragma once
#include "kernels_binary.cuh"

namespace swo {


constexpr int MG_XS = 128;     // per-warp sort scratch in shared memory (entries)
constexpr int MG_GXS = 2048;   // per-warp sort scratch in global memory (entries)
constexpr int MG_SMALL = 8;    // candidates up to which a listed query is expanded by eight lanes (expand_small_kernel)
constexpr int MG_PAIRS = 256;  // expand_wide_kernel: candidates up to which all-pairs counting beats the bitonic sort
constexpr int MG_QPW = 32 * LI_IPT;  // queries per warp pass ("tile")
constexpr int MG_NW = 8;             // warps per CTA
constexpr int MG_TPB = MG_NW * 32;
#ifndef MG_WRITE_CTAS
#define MG_WRITE_CTAS 3
#endif
constexpr int MG_SCAN_ITEMS = 16, MG_SCAN_TILE = 256 * MG_SCAN_ITEMS;  // scan_tiles_kernel1: tiles per CTA

// Multi-hit queries left for expand_wide_kernel: {query, first candidate block, candidates, row position}
struct WideList {
    unsigned *count;  // appended so far (may run past cap: those were expanded on the spot)
    uint4 *ent;       // two per entry: {query, first block, candidates, row position} {start, end, strand byte, -}
    unsigned cap;
    // the same for queries with at most MG_SMALL candidates (expand_small_kernel: eight lanes per query)
    unsigned *count_small;
    uint4 *ent_small;
    unsigned cap_small;
    unsigned gscr_ctas;  // CTAs of expand_wide_kernel that own a region of the global sort scratch
};
// per warp tile, between the passes
struct MergeTiles {
    uint32_t *total;          // [ntw] rows of the tile; after the scan: exclusive prefix inside its scan tile
    // What the count pass found, for the write pass: per query one byte {first candidate block - lo0 of its tile,
    // candidates << 4}, a nibble of 15 = "does not fit, merge this tile again"; the write pass then needs neither
    // the contig column nor the merge (it reads 9 instead of 12 bytes per query and goes straight to the blocks).
    uint32_t *packed;         // [ntw * 32] the four bytes of a lane's queries
    int32_t *lo0;             // [ntw]
    unsigned long long *sum;  // [ntw / MG_SCAN_TILE + 1] rows per scan tile; after the scan: exclusive prefix
#ifdef MG_DEBUG_COUNTS
    int32_t *dbg;             // [4 * n]: per query {lo, hi, path, -} as the count pass saw them
#endif
};

// The deal of warp tiles in the count and write passes: chunks of MG_CHUNK consecutive tiles go round-robin
// to the CTAs, a chunk's tiles round-robin to the CTA's warps.  A warp's next tile is then 8 tiles (1,024
// queries) further on — in a sorted batch still inside its 64-block window most of the time — while all
// CTAs together stay inside a few dozen megabytes of each array.  (One contiguous range per CTA was
// measured 1.5-1.8x slower in both passes: ~1,800 streams, each on its own 2 MiB page.)
#ifndef MG_CHUNK_STEPS_V
#define MG_CHUNK_STEPS_V 2
#endif
constexpr int MG_CHUNK_STEPS = MG_CHUNK_STEPS_V, MG_CHUNK = MG_NW * MG_CHUNK_STEPS;
// (tile, step inside the chunk) -> the warp's next tile; 32-bit on purpose (an int64 division per tile and use
// showed in the instruction count: 224 M -> 247 M in the count pass).  Tile indices stay below 2^26 (n < 2^32).
struct Deal {
    int tile, step;
    __device__ __forceinline__ Deal next() const {
        Deal d{tile + MG_NW, step + 1};
        if (d.step == MG_CHUNK_STEPS) {
            d.step = 0;
            d.tile += ((int)gridDim.x - 1) * MG_CHUNK;
        }
        return d;
    }
};
__device__ __forceinline__ Deal deal_first() { return Deal{(int)blockIdx.x * MG_CHUNK + warp_id(), 0}; }

struct MergeSmem {
    int4 q[MG_NW][2][3][32];               // per warp, two stages: the queries {tcid}, {start}, {end} of a tile (TMA-filled)
    uint64_t qbar[MG_NW][2];               //   and their mbarriers
    int4 cmeta[LI_CM_CAP];                 // per-contig {first block, end block, sample offset, samples}
    int32_t smp[ChainTable::EY_CAP];       // group maxima of pmaxKey in SORTED order (per contig)
    unsigned long long xscr[MG_NW][MG_XS]; // sort scratch of a warp's cooperative expansion (write pass)
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

// one multi-hit query by a whole warp: sort keys in shared memory, in the warp's global scratch, or
// (beyond that) every candidate ranks itself against all others
__device__ __forceinline__ void expand_wide(const LiftArgs &A, const RowSink &sink, uint32_t q, int lo, int cnt, uint64_t pos,
                                            unsigned long long *xs, unsigned long long *gxs) {
    if (cnt <= MG_XS) expand_query_coop<false, false>(A, sink, A.T.ey, q, lo, cnt, pos, xs, nullptr, lane_id(), 32);
    else if (cnt <= MG_GXS) expand_query_coop<false, false>(A, sink, A.T.ey, q, lo, cnt, pos, gxs, nullptr, lane_id(), 32);
    else expand_query<false>(A, sink, A.T.ey, q, lo, cnt, pos, lane_id(), 32);
}

// A warp's 64-block window of the key arrays, kept in registers from tile to tile: lane l holds blocks
// wb + l (A) and wb + 32 + l (B); positions at or behind the contig's end read as INT_MAX.  In a sorted
// batch the tiles of a warp move slowly through the table (a block of hg19ToHg38 holds ~1,850 queries of
// BASELINE configs[1]), so most tiles find their window already there: no search and no key loads.
struct Window {
    int tc;      // contig the window is on, -1: none yet
    int wb;      // first block
    int cend;    // end block of the contig
    int before;  // every block of the contig in front of wb ends at or before this position
    int pmA, tsA, pmB, tsB;
};
__device__ __forceinline__ void window_load_half(const TableView &T, int j0, int cend, int &pm, int &ts) {
    const int j = j0 + lane_id();
    pm = j < cend ? __ldg(T.pmaxKey + j) : 0x7fffffff;
    ts = j < cend ? __ldg(T.tStartKey + j) : 0x7fffffff;
}

// The window merge of one warp: the 64 blocks of window W against those of the lane's four queries that
// are on its contig (bit k of `act`), none of which starts before smin.
//
// EVERY LOOP IN THE MERGE KERNELS HAS A WARP-UNIFORM TRIP COUNT.  An earlier version finished queries
// that ran past the window, and warps with several contigs, with per-thread searches (data-dependent
// loops per lane).  Behind such loops ptxas 12.9 (sm_100a) takes the warp for converged: it emits the
// following SHFL / REDUX without their BRA.DIV + WARPSYNC.COLLECTIVE guard and reduces __syncwarp()
// — also an out-of-line one with a run-time mask — to a NOP; on the GPU a lane was occasionally
// not there (tools/fuzz_parity.py seed 771: one lane's rows missing from a tile total, in one build
// and not in another with a printf added).  profiles/r02_converge_experiments.md has the SASS.
// With uniform loops there is nothing to reconverge.
__device__ __forceinline__ void merge_window(const TableView &T, const Window &W, int smin, unsigned act, const int (&st)[LI_IPT],
                                             const int (&en)[LI_IPT], int (&lo)[LI_IPT], int (&hi)[LI_IPT]) {
    const int lane = lane_id();
    const int wb = W.wb, cend = W.cend;
    const int pmA = W.pmA, tsA = W.tsA, pmB = W.pmB, tsB = W.tsB;
    // queries of other contigs take no part: no block ends at or before INT_MIN, none starts before it
    int s2[LI_IPT], e2[LI_IPT], l2[LI_IPT], h2[LI_IPT];
#pragma unroll
    for (int k = 0; k < LI_IPT; k++) {
        const bool on = (act >> k) & 1u;
        s2[k] = on ? st[k] : (int)0x80000000;
        e2[k] = on ? en[k] : (int)0x80000000;
    }
    // blocks of the window that end at or before the smallest start lie before every query
    int jj = __popc(__ballot_sync(FULL, pmA <= smin));
    if (jj == 32) jj += __popc(__ballot_sync(FULL, pmB <= smin));
    const int w_end = wb + 64;
#pragma unroll
    for (int k = 0; k < LI_IPT; k++) l2[k] = h2[k] = wb + jj;
#pragma unroll 1
    for (; jj < 64; jj++) {
        const int pmj = __shfl_sync(FULL, jj < 32 ? pmA : pmB, jj), tsj = __shfl_sync(FULL, jj < 32 ? tsA : tsB, jj);
        if (!merge_step(pmj, tsj, s2, e2, l2, h2)) break;
    }
    // Queries that run past the window (all 64 blocks start before their end): one at a time, the
    // whole warp searches for it — 32-ary rounds over the key arrays.
    if (jj == 64) {
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            unsigned m = __ballot_sync(FULL, ((act >> k) & 1u) && h2[k] == w_end);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int x_st = __shfl_sync(FULL, s2[k], src), x_en = __shfl_sync(FULL, e2[k], src), x_lo = __shfl_sync(FULL, l2[k], src);
                int nlo = x_lo;
                if (x_lo == w_end) nlo = coop_lower_gt(T.pmaxKey, w_end, cend, x_st);  // its lower bound lies behind the window too
                const int from = nlo > w_end ? nlo : w_end;
                const int nhi = x_en == (int)0x80000000 ? from : coop_lower_gt(T.tStartKey, from, cend, x_en - 1);  // first tStart >= end
                if (lane == src) {
                    l2[k] = nlo;
                    h2[k] = nhi;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < LI_IPT; k++)
        if ((act >> k) & 1u) lo[k] = l2[k], hi[k] = h2[k];
}

// The first block of the 64-block window of a contig {first block, end block, sample offset, samples} for
// a smallest start smin: groups of 32 blocks (every table up to 65,536 blocks): the group found in the
// shared-memory samples (the next group follows it in the window), no probe of the key array; larger
// groups: the window starts at the lower bound itself (one more search round trip).
__device__ __forceinline__ int find_window(const TableView &T, const MergeSmem &S, const int4 cm, int smin) {
    const int gi = coop_lower_gt(S.smp + cm.z, 0, cm.w, smin);
    const int g_lo = cm.x + (gi << T.ey_shift);
    if (T.ey_shift == 5) return g_lo;
    return coop_lower_gt(T.pmaxKey, g_lo, gi < cm.w ? g_lo + (1 << T.ey_shift) - 1 : cm.y, smin);
}

// Candidate block ranges [lo, hi) of the lane's four queries.  C: the warp's window from its previous tile.
__device__ __forceinline__ bool warp_search(const TableView &T, const MergeSmem &S, bool cm_smem, Window &C, const int (&tc)[LI_IPT],
                                            const int (&st)[LI_IPT], const int (&en)[LI_IPT], int (&lo)[LI_IPT], int (&hi)[LI_IPT]) {
    const int tc0 = __shfl_sync(FULL, tc[0], 0);
    const bool same = __all_sync(FULL, tc[0] == tc0 && tc[1] == tc0 && tc[2] == tc0 && tc[3] == tc0) && tc0 >= 0 && tc0 < T.ntc;
    if (same) {
        const int m01 = st[0] < st[1] ? st[0] : st[1], m23 = st[2] < st[3] ? st[2] : st[3];
        const int smin = __reduce_min_sync(FULL, m01 < m23 ? m01 : m23);  // (redux.sync: the warp is converged here, all loops are uniform)
        bool hit = C.tc == tc0 && smin >= C.before;  // (warp-uniform)
        if (hit && __all_sync(FULL, C.pmA <= smin)) {
            // the first half of the window lies in front of every query: slide by one group
            C.before = __shfl_sync(FULL, C.pmA, 31);
            C.pmA = C.pmB, C.tsA = C.tsB;
            C.wb += 32;
            window_load_half(T, C.wb + 32, C.cend, C.pmB, C.tsB);
            hit = !__all_sync(FULL, C.pmA <= smin);  // still in front: the tile is further away, search
        }
        if (!hit) {
            const int4 cm = cm_smem ? S.cmeta[tc0] : load_cmeta(T, tc0);
            C.tc = tc0;
            C.cend = cm.y;
            C.wb = find_window(T, S, cm, smin);
            C.before = smin;
            window_load_half(T, C.wb, C.cend, C.pmA, C.tsA);
            window_load_half(T, C.wb + 32, C.cend, C.pmB, C.tsB);
        }
        merge_window(T, C, smin, 0xFu, st, en, lo, hi);
        return true;
    }
    // Several contigs in the tile (contig boundaries, the batch tail, unsorted stretches): one round of
    // the same merge per distinct contig, smallest id first; queries of other contigs sit the round out.
    unsigned done = 0;
#pragma unroll
    for (int k = 0; k < LI_IPT; k++) lo[k] = hi[k] = 0;
    for (;;) {
        int mine = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++)
            if (!((done >> k) & 1u) && tc[k] < mine) mine = tc[k];
        const int c = warp_min(mine);
        if (c == 0x7fffffff) break;  // warp-uniform
        unsigned act = 0;
        int smine = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            if (!((done >> k) & 1u) && tc[k] == c) {
                act |= 1u << k;
                smine = st[k] < smine ? st[k] : smine;
            }
        }
        done |= act;
        if (c < 0 || c >= T.ntc) continue;  // unknown contig: no candidates (lo = hi = 0)
        const int4 cm = cm_smem ? S.cmeta[c] : load_cmeta(T, c);
        const int smin = warp_min(smine);
        Window W;
        W.tc = c;
        W.cend = cm.y;
        W.wb = find_window(T, S, cm, smin);
        W.before = smin;
        window_load_half(T, W.wb, W.cend, W.pmA, W.tsA);
        window_load_half(T, W.wb + 32, W.cend, W.pmB, W.tsB);
        merge_window(T, W, smin, act, st, en, lo, hi);
    }
    return false;
}

// A warp's queries come through shared memory: lane 0 arms the stage's mbarrier and issues three
// 1-D bulk copies (TMA; UBLKCP in the SASS) for the warp's NEXT tile — tiles are dealt round-robin,
// so it is known — while the current one is merged; the load latency is off the critical path.
// The batch's last, partial tile is read directly (tcid -1 / empty interval behind its end).
struct QueryStager {
    MergeSmem &S;
    const LiftArgs &A;
    int w, lane;
    unsigned uses[2];  // how often each stage has been waited for (mbarrier phase parity)
    __device__ QueryStager(MergeSmem &S_, const LiftArgs &A_) : S(S_), A(A_), w(warp_id()), lane(lane_id()) {
        uses[0] = uses[1] = 0;
        if (lane == 0) {
            mbar_init(&S.qbar[w][0], 1);
            mbar_init(&S.qbar[w][1], 1);
            mbar_init_fence();
        }
        __syncwarp();
    }
    __device__ bool whole(int tile) const { return (uint32_t)(tile + 1) * MG_QPW <= A.n; }
    // lane 0 issues; every lane of the warp has finished reading stage `st` (the caller's __syncwarp)
    __device__ void issue(int tile, int st) {
        if (lane == 0 && whole(tile)) {
            fence_proxy_async();
            mbar_expect_tx(&S.qbar[w][st], 2u * MG_QPW * 4u);
            bulk_g2s(S.q[w][st][1], A.start + (size_t)tile * MG_QPW, MG_QPW * 4u, &S.qbar[w][st]);
            bulk_g2s(S.q[w][st][2], A.end + (size_t)tile * MG_QPW, MG_QPW * 4u, &S.qbar[w][st]);
        }
    }
    __device__ void fetch(int tile, int st, int (&sta)[LI_IPT], int (&en)[LI_IPT]) {
        if (whole(tile)) {
            mbar_wait(&S.qbar[w][st], uses[st] & 1u);
            uses[st]++;
            const int4 b = S.q[w][st][1][lane], c = S.q[w][st][2][lane];
            sta[0] = b.x, sta[1] = b.y, sta[2] = b.z, sta[3] = b.w;
            en[0] = c.x, en[1] = c.y, en[2] = c.z, en[3] = c.w;
        } else {
            const uint32_t q0 = (uint32_t)tile * MG_QPW + (uint32_t)lane * LI_IPT;
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) {
                const bool ok = q0 + k < A.n;
                sta[k] = ok ? A.start[q0 + k] : 0;
                en[k] = ok ? A.end[q0 + k] : 0;
            }
        }
    }
};

__device__ __forceinline__ void merge_tables(const TableView &T, MergeSmem &S, bool samples) {
    const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
    if (T.ntc <= LI_CM_CAP)
        for (int i = tid; i < T.ntc; i += MG_TPB) S.cmeta[i] = __ldg(T.cmeta + i);
    // group maxima in sorted order: sample i of contig c is the last key of its i-th full group
    if (samples)
        for (int c = w; c < T.ntc; c += MG_NW) {
            const int4 cm = __ldg(T.cmeta + c);
            for (int i = lane; i < cm.w; i += 32) S.smp[cm.z + i] = __ldg(T.pmaxKey + cm.x + ((i + 1) << T.ey_shift) - 1);
        }
    __syncthreads();
}

// Pass 1: rows and window position of every warp tile; unmatched count.
__global__ void __launch_bounds__(MG_TPB, 4) merge_count_kernel(const __grid_constant__ LiftArgs A, const __grid_constant__ MergeTiles M) {
    extern __shared__ __align__(16) unsigned char li_smem_raw[];
    MergeSmem &S = *reinterpret_cast<MergeSmem *>(li_smem_raw);
    if (__ldg(A.sorted_flag) == 0u) return;
    const TableView &T = A.T;
    merge_tables(T, S, true);
    const int lane = lane_id();
    const bool cm_smem = T.ntc <= LI_CM_CAP;
    uint32_t unm_acc = 0;
    // (This pass is short per tile: plain streaming loads with the tile after the next prefetched into L2 keep
    // more bytes in flight than the write pass's one-tile-ahead shared-memory stages.)
    Window C;
    C.tc = -1;
    C.wb = C.cend = C.before = C.pmA = C.tsA = C.pmB = C.tsB = 0;
    // (lanes 0-11 prefetch the four 128-byte lines of each of the three arrays)
    const int32_t *const pf_base = (lane < 4 ? A.tcid : lane < 8 ? A.start : A.end) + (lane & 3) * 32;
    for (Deal dl = deal_first(); dl.tile < A.ntiles; dl = dl.next()) {
        const int tile = dl.tile;
        const uint32_t q0 = (uint32_t)tile * MG_QPW + (uint32_t)lane * LI_IPT;
        const bool full = (uint32_t)(tile + 1) * MG_QPW <= A.n;
        int tc[LI_IPT], st[LI_IPT], en[LI_IPT], lo[LI_IPT], hi[LI_IPT];
        if (full) {
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
        {
            const uint32_t pt = (uint32_t)dl.next().next().tile;
            if (lane < 12 && (pt + 1u) * MG_QPW <= A.n) prefetch_l2(pf_base + (size_t)pt * MG_QPW);
        }
        const bool one = warp_search(T, S, cm_smem, C, tc, st, en, lo, hi);
        const int m01 = lo[0] < lo[1] ? lo[0] : lo[1], m23 = lo[2] < lo[3] ? lo[2] : lo[3];
        const int lo0 = __reduce_min_sync(FULL, m01 < m23 ? m01 : m23);
        uint32_t tsum = 0, pk = 0;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            const int cnt = hi[k] > lo[k] ? hi[k] - lo[k] : 0;
            tsum += (uint32_t)cnt;
            unm_acc += cnt == 0 ? 1u : 0u;  // (the empty slots behind the batch's end are taken off below)
            const uint32_t d = (uint32_t)(lo[k] - lo0);
            pk |= ((d < 15u ? d : 15u) | ((uint32_t)(cnt < 15 ? cnt : 15) << 4)) << (8 * k);
#ifdef MG_DEBUG_COUNTS
            if (M.dbg && q0 + k < A.n) {
                M.dbg[4 * (size_t)(q0 + k)] = lo[k];
                M.dbg[4 * (size_t)(q0 + k) + 1] = hi[k];
                M.dbg[4 * (size_t)(q0 + k) + 2] = C.wb;
                M.dbg[4 * (size_t)(q0 + k) + 3] = (int)tsum;
            }
#endif
        }
#ifdef MG_DEBUG_COUNTS
        {
            const unsigned am = __activemask();
            if (M.dbg && am != FULL) printf("tile %d lane %d: active mask %08x in front of the tile total\n", tile, lane, am);
        }
#endif
        // (redux.sync is safe again: no per-lane loop is left in front of it — an earlier version with a divergent
        // per-thread search lost a lane's rows here, profiles/r02_converge_experiments.md)
        const uint32_t total = __reduce_add_sync(FULL, tsum);
        M.packed[(size_t)tile * 32 + lane] = one ? pk : 0xFFFFFFFFu;  // several contigs: the write pass merges again
        if (lane == 0) {
            M.total[tile] = total;
            M.lo0[tile] = lo0;
        }
    }
    uint32_t unm = warp_sum(unm_acc);
    if (lane == 0 && unm) {
        // the warp that took the batch's last, partial tile counted its empty slots as unmatched queries
        const int last = A.ntiles - 1;
        const uint32_t slots = (uint32_t)A.ntiles * MG_QPW - A.n;
        const bool mine = slots && (last / MG_CHUNK) % (int)gridDim.x == (int)blockIdx.x && last % MG_NW == warp_id();
        atomicAdd((unsigned long long *)(A.totals + 1), (unsigned long long)(unm - (mine ? slots : 0u)));
    }
}

// Scan, step 1: every CTA turns MG_SCAN_TILE tile totals into exclusive prefixes (in place) and stores their sum.
__global__ void __launch_bounds__(256) scan_tiles_kernel1(const uint32_t *flag, MergeTiles M, int ntw) {
    __shared__ uint32_t scr[9];
    if (__ldg(flag) == 0u) return;
    const int base = (int)blockIdx.x * MG_SCAN_TILE + (int)threadIdx.x * MG_SCAN_ITEMS;
    uint32_t v[MG_SCAN_ITEMS], run = 0;
#pragma unroll
    for (int i = 0; i < MG_SCAN_ITEMS; i += 4) {
        int4 x = make_int4(0, 0, 0, 0);
        if (base + i + 3 < ntw) x = *reinterpret_cast<const int4 *>(M.total + base + i);
        else {
            x.x = base + i < ntw ? (int)M.total[base + i] : 0;
            x.y = base + i + 1 < ntw ? (int)M.total[base + i + 1] : 0;
            x.z = base + i + 2 < ntw ? (int)M.total[base + i + 2] : 0;
        }
        v[i] = (uint32_t)x.x, v[i + 1] = (uint32_t)x.y, v[i + 2] = (uint32_t)x.z, v[i + 3] = (uint32_t)x.w;
    }
#pragma unroll
    for (int i = 0; i < MG_SCAN_ITEMS; i++) {
        const uint32_t t = v[i];
        v[i] = run;
        run += t;
    }
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan<uint32_t, 8>(run, scr, &tot);
#pragma unroll
    for (int i = 0; i < MG_SCAN_ITEMS; i++)
        if (base + i < ntw) M.total[base + i] = ex + v[i];
    if (threadIdx.x == 0) M.sum[blockIdx.x] = tot;  // a scan tile holds < 2^32 rows (the host rejects larger batches)
}
// Scan, step 2: one CTA scans the per-scan-tile sums; the grand total goes to totals[0] and row_offset[n].
__global__ void __launch_bounds__(1024) scan_tiles_kernel2(const uint32_t *flag, MergeTiles M, int nsum, uint64_t *totals,
                                                            uint32_t *row_offset_n) {
    __shared__ unsigned long long scr[33];
    if (__ldg(flag) == 0u) return;
    unsigned long long carry = 0;
    for (int b = 0; b < nsum; b += 1024) {
        const int i = b + (int)threadIdx.x;
        const unsigned long long v = i < nsum ? M.sum[i] : 0ull;
        unsigned long long tot;
        const unsigned long long ex = block_exclusive_scan<unsigned long long, 32>(v, scr, &tot);
        if (i < nsum) M.sum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        totals[0] = carry;
        *row_offset_n = (uint32_t)carry;
    }
}

// Pass 2: the merge again, from the stored window; row offsets and rows to their final places.
template <bool STRAND>
__global__ void __launch_bounds__(MG_TPB, MG_WRITE_CTAS) merge_write_kernel(const __grid_constant__ LiftArgs A, const __grid_constant__ MergeTiles M,
                                                                const __grid_constant__ WideList W) {
    extern __shared__ __align__(16) unsigned char li_smem_raw[];
    MergeSmem &S = *reinterpret_cast<MergeSmem *>(li_smem_raw);
    if (__ldg(A.sorted_flag) == 0u) return;
    const TableView &T = A.T;
    merge_tables(T, S, true);
    const int lane = lane_id(), w = warp_id();
    const bool cm_smem = T.ntc <= LI_CM_CAP;
    const int last_block = __ldg(&T.cmeta[T.ntc - 1].y) - 1;
    const RowSink sink{A.rows, A.thick, A.rows_cap};
    const uint32_t cap32 = A.rows_cap > 0xffffffffull ? 0xffffffffu : (uint32_t)A.rows_cap;
    unsigned long long *const gxs = A.gscr + ((size_t)blockIdx.x * MG_NW + w) * MG_GXS;
    Window C;
    C.tc = -1;
    C.wb = C.cend = C.before = C.pmA = C.tsA = C.pmB = C.tsB = 0;
    QueryStager Q(S, A);
    if (deal_first().tile < A.ntiles) Q.issue(deal_first().tile, 0);
    int stage = 0;
    for (Deal dl = deal_first(); dl.tile < A.ntiles; dl = dl.next(), stage ^= 1) {
        const int tile = dl.tile;
        const uint32_t q0 = (uint32_t)tile * MG_QPW + (uint32_t)lane * LI_IPT;
        const bool full = (uint32_t)(tile + 1) * MG_QPW <= A.n;
        int st[LI_IPT], en[LI_IPT], lo[LI_IPT], cnt[LI_IPT];
        __syncwarp();
        if (dl.next().tile < A.ntiles) Q.issue(dl.next().tile, stage ^ 1);
        // what the count pass found for this tile (and, further down, the blocks) before the queries are needed
        const uint32_t pk = __ldg(M.packed + (size_t)tile * 32 + lane);
        const int lo0 = __ldg(M.lo0 + tile);
        const uint64_t tbase = M.sum[tile / MG_SCAN_TILE] + (uint64_t)M.total[tile];
        Q.fetch(tile, stage, st, en);
        uchar4 sb4 = make_uchar4(0, 0, 0, 0);
        if (STRAND) {
            if (full) sb4 = __ldg(reinterpret_cast<const uchar4 *>(A.strand + q0));
            else {
                sb4.x = q0 + 0 < A.n ? A.strand[q0 + 0] : 0;
                sb4.y = q0 + 1 < A.n ? A.strand[q0 + 1] : 0;
                sb4.z = q0 + 2 < A.n ? A.strand[q0 + 2] : 0;
                sb4.w = q0 + 3 < A.n ? A.strand[q0 + 3] : 0;
            }
        }
        // a nibble of 15 anywhere in the tile: a delta or a count did not fit (or the tile has several contigs) —
        // the whole warp merges this tile again (warp-uniform)
        const uint32_t esc = (((pk & 0x0F0F0F0Fu) + 0x01010101u) | (((pk >> 4) & 0x0F0F0F0Fu) + 0x01010101u)) & 0x10101010u;
        if (__any_sync(FULL, esc != 0u)) {
            int tc[LI_IPT], hi[LI_IPT];
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) tc[k] = q0 + k < A.n ? __ldg(A.tcid + q0 + k) : -1;
            warp_search(T, S, cm_smem, C, tc, st, en, lo, hi);
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) cnt[k] = hi[k] > lo[k] ? hi[k] - lo[k] : 0;
        } else {
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) {
                lo[k] = lo0 + (int)((pk >> (8 * k)) & 15u);
                cnt[k] = (int)((pk >> (8 * k + 4)) & 15u);
            }
        }
        // Hits, and the row of the first candidate — computed without branches for every query (four
        // independent block loads in flight); it is only stored for single-hit queries.
        Block bk[LI_IPT];
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) bk[k] = load_block(T, lo[k] < last_block ? lo[k] : last_block);
        const unsigned char sbk[LI_IPT] = {sb4.x, sb4.y, sb4.z, sb4.w};
        uint32_t tsum = 0;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) tsum += (uint32_t)cnt[k];
        const uint32_t inc = warp_inclusive_scan(tsum);
        const uint64_t pos64 = tbase + (inc - tsum);
        uint32_t pos = (uint32_t)pos64;  // row indices are 32-bit (the host rejects batches with 2^32 rows or more)
        const uint32_t o0 = pos;
        const uint32_t o1 = o0 + (uint32_t)cnt[0], o2 = o1 + (uint32_t)cnt[1], o3 = o2 + (uint32_t)cnt[2];
        if (full) {
            st_stream_v4(A.row_offset + q0, make_int4((int)o0, (int)o1, (int)o2, (int)o3));
        } else {
            const uint32_t o[LI_IPT] = {o0, o1, o2, o3};
#pragma unroll
            for (int k = 0; k < LI_IPT; k++)
                if (q0 + k < A.n) A.row_offset[q0 + k] = o[k];
        }
        unsigned multi = 0;
#pragma unroll
        for (int k = 0; k < LI_IPT; k++) {
            // intersect + qEnd + orderStartEnd (chain.d:728-743, 128-134, 787-793) for a +/-1 invert
            const Block &b = bk[k];
            const int ts = max(b.tStart, st[k]), te = min(b.tEnd, en[k]);
            const bool neg = b.meta & 1;
            const int key = neg ? b.qStart - (ts - b.tStart) : b.qStart + (ts - b.tStart);
            const int len = te - ts;
            const int rs = neg ? key - len : key, re = neg ? key : key + len;  // cnt == 1: len > 0
            const int meta = STRAND ? ((b.meta >> 1) | ((int)strand_flip(sbk[k], neg ? -1 : 1) << 16)) : (b.meta >> 1);
            if (cnt[k] == 1 && pos < cap32) st_stream_v4(A.rows + pos, make_int4(meta, rs, re, (int)(q0 + k)));
            multi |= cnt[k] >= 2 ? 1u << k : 0u;
            pos += (uint32_t)cnt[k];
        }
        if (__any_sync(FULL, multi != 0u)) {
            // Two candidates (an interval across one gap): by the thread, straight-line.  More: onto the list for
            // expand_wide_kernel — or, list full, by the whole warp one query at a time.  (No per-thread loops
            // in here: see merge_window.)
            unsigned wide = 0;
            uint64_t p = pos64;
#pragma unroll
            for (int k = 0; k < LI_IPT; k++) {
                if ((multi >> k) & 1u) {
                    if (cnt[k] == 2) {
                        expand_pair<false>(A, sink, Thick{0, 0, 0}, q0 + k, lo[k], st[k], en[k], sbk[k], p);
                    } else {
                        const bool small = cnt[k] <= MG_SMALL;
                        const unsigned e = atomicAdd(small ? W.count_small : W.count, 1u);
                        if (e < (small ? W.cap_small : W.cap)) {
                            uint4 *const ent = (small ? W.ent_small : W.ent) + 2 * (size_t)e;
                            ent[0] = make_uint4(q0 + k, (unsigned)lo[k], (unsigned)cnt[k], (uint32_t)p);
                            ent[1] = make_uint4((unsigned)st[k], (unsigned)en[k], sbk[k], 0u);
                        }
                        else wide |= 1u << k;
                    }
                }
                p += (uint64_t)cnt[k];
            }
            if (__any_sync(FULL, wide != 0u)) {
                p = pos64;
#pragma unroll
                for (int k = 0; k < LI_IPT; k++) {
                    unsigned m = __ballot_sync(FULL, (wide >> k) & 1u);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t q = __shfl_sync(FULL, q0, src) + (uint32_t)k;
                        const int qlo = __shfl_sync(FULL, lo[k], src), qcnt = __shfl_sync(FULL, cnt[k], src);
                        const uint64_t qpos = __shfl_sync(FULL, p, src);
                        expand_wide(A, sink, q, qlo, qcnt, qpos, S.xscr[w], gxs);
                    }
                    p += (uint64_t)cnt[k];
                }
            }
        }
    }
}

// Listed queries with 3..MG_SMALL candidates — three quarters of all multi-hit queries on hg19ToHg38 — by EIGHT lanes
// each, four queries per warp: a lane per candidate, every candidate counts the keys in front of its own (eight
// width-8 shuffles: no runs, no scratch, any order), the cumulative strand byte from the inverted rows among them.
// The loop is warp-uniform: a group whose entry index is past the list's end carries an empty query.
__global__ void __launch_bounds__(256) expand_small_kernel(const __grid_constant__ LiftArgs A, const __grid_constant__ WideList W) {
    const TableView &T = A.T;
    const unsigned n = min(*W.count_small, W.cap_small);
    const int lane = lane_id(), g = lane & 7, gshift = lane & ~7;
    const unsigned step = gridDim.x * 32u;
    for (unsigned e0 = (blockIdx.x * 8u + (unsigned)warp_id()) * 4u; e0 < n; e0 += step) {
        const unsigned e = e0 + (unsigned)(lane >> 3);
        uint4 v = make_uint4(0, 0, 0, 0), x = v;
        if (e < n) v = __ldg(W.ent_small + 2 * (size_t)e), x = __ldg(W.ent_small + 2 * (size_t)e + 1);
        const int cnt = (int)v.z, start = (int)x.x, end = (int)x.y;
        const bool valid = g < cnt;
        const RowCore row = make_row(load_block(T, (int)v.y + (valid ? g : 0)), start, end);
        const unsigned invb = (__ballot_sync(FULL, valid && row.inv < 0) >> gshift) & 0xFFu;
        int rank = 0, negs = valid && row.inv < 0 ? 1 : 0;
#pragma unroll
        for (int l = 0; l < MG_SMALL; l++) {
            const int k = __shfl_sync(FULL, row.key, l, 8);
            const int before = (l < cnt && (k < row.key || (k == row.key && l < g))) ? 1 : 0;
            rank += before;
            negs += before & (int)((invb >> l) & 1u);
        }
        const unsigned m0 = (__ballot_sync(FULL, valid && rank == 0) >> gshift) & 0xFFu;
        const int inv_first = __shfl_sync(FULL, row.inv, m0 ? __ffs(m0) - 1 : 0, 8);
        const uint64_t pos = (uint64_t)v.w + (uint64_t)rank;
        if (valid && pos < A.rows_cap) {
            const unsigned char sb = A.strand ? strand_at_rank((unsigned char)x.z, inv_first, rank, negs) : 0;
            *reinterpret_cast<int4 *>(A.rows + pos) = make_int4(row.qcid | ((int)sb << 16), row.s, row.e, (int)v.x);
        }
    }
}

// The multi-hit queries merge_write_kernel (sorted batches) or lift_intervals_kernel (other batches) left
// on the list: a warp per query, rows from registers (expand_warp_runs; the entry carries the query, so
// nothing but the blocks is loaded); what that form declines is sorted in shared memory / the warp's
// global scratch.  Row positions are 32-bit (the host rejects batches with 2^32 rows or more).
template <bool THICK>
__global__ void __launch_bounds__(256, 4) expand_wide_kernel(const __grid_constant__ LiftArgs A, const __grid_constant__ WideList W) {
    __shared__ unsigned long long xs[8][MG_XS];
    const unsigned n = min(*W.count, W.cap);
    const int w = warp_id();
    const RowSink sink{A.rows, A.thick, A.rows_cap};
    unsigned long long *const gxs = A.gscr + ((size_t)blockIdx.x * 8 + w) * MG_GXS;  // (the host sizes the scratch for this grid)
    constexpr int STRIDE = THICK ? 3 : 2;  // (a THICK instantiation would read the lifted thickStart/End from a third uint4;
                                           //  batches with thick columns expand inside lift_intervals_kernel for now)
    // Software pipeline, two entries deep: while entry e is expanded, entry e + 2 steps is on its way from the
    // list and the first-round blocks of entry e + 1 step from the table — the warp never waits for either.
    const unsigned step = gridDim.x * 8;
    const int lane = lane_id();
    unsigned e = blockIdx.x * 8 + w;
    const uint4 z4 = make_uint4(0, 0, 1, 0);
    uint4 v0 = z4, x0 = z4, y0 = z4, v1 = z4, x1 = z4, y1 = z4;
    auto load_entry = [&](unsigned i, uint4 &v, uint4 &x, uint4 &y) {
        if (i < n) {
            v = __ldg(W.ent + STRIDE * (size_t)i), x = __ldg(W.ent + STRIDE * (size_t)i + 1);
            if (THICK) y = __ldg(W.ent + STRIDE * (size_t)i + 2);
        }
    };
    load_entry(e, v0, x0, y0);
    load_entry(e + step, v1, x1, y1);
    Block b0 = load_block(A.T, (int)v0.y + min(lane, (int)v0.z - 1));
    while (e < n) {
        uint4 v2 = z4, x2 = z4, y2 = z4;
        load_entry(e + 2 * step, v2, x2, y2);
        const Block b1 = load_block(A.T, (int)v1.y + min(lane, (int)v1.z - 1));
        Thick th{0, 0, 0};
        if (THICK) th.ts = (int)y0.x, th.te = (int)y0.y, th.has = (int)y0.z;
        if (!expand_warp_runs<THICK>(A, sink, th, v0.x, (int)v0.y, (int)v0.z, (uint64_t)v0.w, (int)x0.x, (int)x0.y, (unsigned char)x0.z,
                                     reinterpret_cast<int *>(xs[w]), b0)) {
            // what the run form declined: up to MG_PAIRS candidates every one counts the keys in front of it (n^2 / 32
            // steps per lane, six instructions each); larger ones are sorted (bitonic, n log^2 n) in the warp's
            // global scratch; beyond that scratch all pairs again, straight from the block table
            if ((int)v0.z <= MG_PAIRS)
                expand_warp_allpairs<THICK>(A, sink, th, v0.x, (int)v0.y, (int)v0.z, (uint64_t)v0.w, (int)x0.x, (int)x0.y, (unsigned char)x0.z, gxs);
            else if ((int)v0.z <= MG_GXS) {
                expand_query_coop<false, THICK>(A, sink, A.T.ey, v0.x, (int)v0.y, (int)v0.z, (uint64_t)v0.w, gxs, nullptr, lane, 32);
                converge(A);
            } else {
                expand_query<THICK>(A, sink, A.T.ey, v0.x, (int)v0.y, (int)v0.z, (uint64_t)v0.w, lane, 32);
            }
        }
        e += step;
        v0 = v1, x0 = x1, y0 = y1, b0 = b1;
        v1 = v2, x1 = x2, y1 = y2;
    }
}

}  // namespace swo
