// device_util.cuh — CUB-free warp/block scans, the decoupled look-back chained
// scan used by the single-pass lift kernels, and memory-access helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swo {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// ---- cache-hinted accesses -------------------------------------------------
// streaming (read-once) 128-bit load that does not allocate in L1
__device__ __forceinline__ int4 ld_stream_v4(const void *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(void *p, const int4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// relaxed gpu-scope 64-bit accesses for the look-back descriptors (bypass L1)
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- mbarrier + 1-D bulk copy (TMA) ------------------------------------------------
// One thread arms the barrier with the byte count and issues the copies; consumers wait on the
// barrier's phase parity.  SASS: UBLKCP (copy), SYNCS (barrier).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- warp reconvergence ---------------------------------------------------
// Behind long divergent per-thread work (per-query searches: loops with a per-lane trip count) a
// warp can reach the next warp collective or block barrier SPLIT although the code is straight-line
// there: ptxas 12.9 (sm_100a) takes the warp for converged behind its BSYNC.RECONVERGENT, emits
// the collective without the BRA.DIV / WARPSYNC.COLLECTIVE guard and drops a __syncwarp() placed
// there — in merge_count_kernel even this out-of-line form with a run-time mask became
// "LDC; NOP; RET" — and on the GPU a lane was occasionally not part of the collective (a short
// block scan, "illegal instruction" at a bar.sync, a tile total without one lane's rows).
// Experiments and SASS: profiles/r02_converge_experiments.md (compute-sanitizer is closed on this
// pool, so there is no synccheck log).  The kernels of kernels_binary.cuh and kernels_text.cuh call
// this in front of every barrier / collective behind their searches (ptxas keeps the WARPSYNC in
// those kernels; cost 1-5 %); the merge kernels (kernels_merge.cuh) avoid the situation instead:
// every loop there has a warp-uniform trip count.
#if defined(SWO_CONVERGE_NONE)  // experiments (profiles/r02_converge_*): nothing at all ...
__device__ __forceinline__ void converge_warp(unsigned) {}
#elif defined(SWO_CONVERGE_PLAIN)  // ... or the plain intrinsic, inlined
__device__ __forceinline__ void converge_warp(unsigned) { __syncwarp(); }
#else
__device__ __noinline__ void converge_warp(unsigned mask) { __syncwarp(mask); }
#endif

// ---- warp / block scans ----------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= d) v += o;
    }
    return v;
}
// (shuffle-based on purpose, not redux.sync: see the note at converge_warp)
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// Exclusive scan of one value per thread over a block of NWARPS*32 threads.
// `scratch` is NWARPS+1 elements of shared memory.  Returns the exclusive prefix;
// *total receives the block sum (same value in every thread).
// Contains two __syncthreads(); safe to call repeatedly with the same scratch.
// Warp shuffles and two block barriers: the warp must arrive converged (converge_warp above).
template <typename T, int NWARPS>
__device__ __forceinline__ T block_exclusive_scan(T v, T *scratch, T *total) {
    const T inc = warp_inclusive_scan(v);
    __syncthreads();  // previous users of scratch are done
    if (lane_id() == 31) scratch[warp_id()] = inc;
    __syncthreads();
    T wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; w++) {
        const T s = scratch[w];
        if (w < warp_id()) wbase += s;
        tot += s;
    }
    *total = tot;
    return wbase + inc - v;
}

// ---- decoupled look-back ---------------------------------------------------
// One 64-bit descriptor per tile: bits 63..62 = state, bits 61..0 = value.
constexpr uint64_t LB_EMPTY = 0ull, LB_AGG = 1ull << 62, LB_INC = 2ull << 62;
constexpr uint64_t LB_STATE = 3ull << 62, LB_VALUE = ~LB_STATE;

// Called by ONE full warp, in two steps so that a tile can publish its aggregate
// as soon as it is known and resolve its own prefix later (successors then never
// wait on this tile's remaining work).  Tiles must have been handed out by an
// atomic ticket so that every predecessor is already running (forward progress).
__device__ __forceinline__ void lookback_publish(uint64_t *desc, int tile, uint64_t aggregate) {
    if (lane_id() == 0) st_relaxed_u64(desc + tile, (tile == 0 ? LB_INC : LB_AGG) | aggregate);
}
// Walks the predecessors' descriptors 32 at a time until an inclusive prefix is
// found, publishes this tile's inclusive prefix and returns the EXCLUSIVE prefix
// (all lanes).
__device__ __forceinline__ uint64_t lookback_resolve(uint64_t *desc, int tile, uint64_t aggregate) {
    const int lane = lane_id();
    if (tile == 0) return 0;
    uint64_t excl = 0;
    int look = tile - 1;
    for (;;) {
        const int idx = look - lane;
        uint64_t v = LB_INC;  // virtual tile before tile 0: inclusive prefix 0
        if (idx >= 0) {
            v = ld_relaxed_u64(desc + idx);
            while ((v & LB_STATE) == LB_EMPTY) {
                __nanosleep(20);
                v = ld_relaxed_u64(desc + idx);
            }
        }
        const unsigned inc_mask = __ballot_sync(FULL, (v & LB_STATE) == LB_INC);
        const int first = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        const uint64_t contrib = (lane <= first) ? (v & LB_VALUE) : 0ull;
        excl += warp_sum(contrib);
        if (inc_mask) break;
        look -= 32;
    }
    if (lane == 0) st_relaxed_u64(desc + tile, LB_INC | (excl + aggregate));
    return excl;
}
// Block-wide resolve: every thread of an NW-warp block calls it; NW*32 predecessors are
// inspected per round (the walk is latency bound, so fewer rounds = shorter critical path).
// part/found: NW words of shared memory each.  Returns the exclusive prefix (all threads).
template <int NW>
__device__ __forceinline__ uint64_t lookback_resolve_block(uint64_t *desc, int tile, uint64_t aggregate,
                                                           unsigned long long *part, int *found) {
    uint64_t excl = 0;
    bool done = tile == 0;
    int look = tile - 1;
    while (!done) {
        const int idx = look - (int)threadIdx.x;
        uint64_t v = LB_INC;  // virtual tile before tile 0: inclusive prefix 0
        if (idx >= 0) {
            v = ld_relaxed_u64(desc + idx);
            while ((v & LB_STATE) == LB_EMPTY) {
                __nanosleep(20);
                v = ld_relaxed_u64(desc + idx);
            }
        }
        const unsigned m = __ballot_sync(FULL, (v & LB_STATE) == LB_INC);
        const int f = m ? (__ffs(m) - 1) : 32;
        const uint64_t s = warp_sum((lane_id() <= f) ? (v & LB_VALUE) : 0ull);
        __syncthreads();  // previous round's readers are done
        if (lane_id() == 0) {
            part[warp_id()] = s;
            found[warp_id()] = m != 0;
        }
        __syncthreads();
        for (int w = 0; w < NW; w++) {  // warps are ordered by distance from the tile
            excl += part[w];
            if (found[w]) {
                done = true;
                break;
            }
        }
        look -= NW * 32;
    }
    if (threadIdx.x == 0 && tile > 0) st_relaxed_u64(desc + tile, LB_INC | (excl + aggregate));
    return excl;
}

__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t *desc, int tile, uint64_t aggregate) {
    lookback_publish(desc, tile, aggregate);
    return lookback_resolve(desc, tile, aggregate);
}

}  // namespace swo
