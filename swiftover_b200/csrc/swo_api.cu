// swo_api.cu — C ABI of libswiftover_cuda.so (include/swiftover_cuda.h).
//
// Host-side plumbing around the sm_100a kernels: context, chain-table upload
// (replacing the tree build of chain.d:509-597), device/pinned buffers,
// byte-range sharding of text inputs over devices and the chunked
// H2D -> kernel -> D2H pipeline.  No CPU compute path exists here: every lift
// goes through a CUDA kernel or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "chain_table.hpp"
#include "genome.hpp"
#include "kernels_binary.cuh"
#include "kernels_merge.cuh"
#include "kernels_text.cuh"
#include "kernels_maf.cuh"
#include "kernels_vcf.cuh"
#include "swiftover_cuda.h"
#include "swo_internal.hpp"

using namespace swo;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + (n >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n, bool keep = false, size_t keep_bytes = 0) {
        if (n <= cap) return cudaSuccess;
        size_t want = n + (n >> 2) + 4096;
        void *np = nullptr;
        cudaError_t e = cudaHostAlloc(&np, want, cudaHostAllocPortable);
        if (e != cudaSuccess) return e;
        if (keep && p && keep_bytes) memcpy(np, p, keep_bytes);
        if (p) cudaFreeHost(p);
        p = np;
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr int N_SLOTS = 8;  // chunks in flight per device (H2D / kernel / D2H overlap)

// Everything one device (or logical shard) owns.
struct Device {
    int ordinal = 0;
    int sms = 148;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr, s_d2h2 = nullptr;  // two D2H streams: alternate chunks
    // chain table replica
    DevBuf blk, tkey, pkey, toff, cmeta, ey, thash, tchars, tnoff, qoff, qchars;
    TableView T{};
    uint32_t thash_mask = 0;
    int nq = 0;  // destination contigs
    // genome replica
    DevBuf g_packed, g_off, g_len, g_lower, g_chunk, g_estart, g_elen, g_ebyte;
    int g_exc_n = 0;
    bool g_has_lower = false;
    int g_n = 0;
    // scratch for the look-back chains
    DevBuf scratch, gscr, wide;
    // The scratch (ticket, look-back descriptors, tile totals, sorted flag) is per device, not per call: a launch
    // that uses it first waits for the previous one on ANY stream to be done with it (callers of the *_dev entry
    // points may use streams of their own).
    cudaEvent_t ev_scratch = nullptr;
    bool scratch_used = false;
    // binary path buffers
    DevBuf b_in, b_off, b_rows, b_thick, b_tot;
    // text path: per-slot chunk buffers
    DevBuf t_in[N_SLOTS], t_out[N_SLOTS], t_unm[N_SLOTS], t_sizes[N_SLOTS];
    cudaEvent_t ev_h2d[N_SLOTS]{}, ev_comp[N_SLOTS]{}, ev_d2h[N_SLOTS]{};
    PinBuf h_sizes;  // N_SLOTS * 6 u64
    // per-shard host output arenas (multi-shard runs)
    PinBuf a_lifted, a_unm;
    int text_grid = 0, bin_grid[5] = {0, 0, 0, 0, 0};
};

}  // namespace

struct swo_ctx {
    std::vector<Device> dev;
    ChainTable tab;
    bool have_chain = false;
    std::string err;
    uint64_t launches = 0;
    GenomeStore genome;        // FASTA-loaded destination genome (host copy of the tables, vcf header needs lengths)
    bool have_fasta = false;
    std::string vcf_l, vcf_u;  // outputs of swo_lift_vcf_text
    // pinned result arenas handed to the caller
    PinBuf r_lifted, r_unm, r_off, r_rows, r_thick;
    size_t chunk_bytes = 64u << 20;
    // read pinned input from the kernel over PCIe instead of staging it with H2D copies: as fast as the
    // copy when nothing else uses the link, but slower than staging once D2H copies of the output overlap
    bool zero_copy_input = false;
    int d2h_streams = 1;  // 2: alternate the chunks' D2H copies over two streams
    // multi-device runs: 1 = hand back ONE contiguous buffer per output (the shards' arenas are copied into it, all
    // shards at once); 0 = no copy: the caller walks the shards' arenas in order with swo_text_segment
    bool concat = true;
    std::vector<std::pair<const char *, size_t>> seg_l, seg_u;
    // text tile size: TX_TILE, or 8192 / 4096 for inputs with large fan-out (a tile's lifted text should fit
    // its 20 KiB stage); 0 = choose per chunk from the previous chunk's output/input ratio
    uint32_t tile_bytes = TX_TILE;
    bool tile_auto = true;
};

namespace {

int fail(swo_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CK(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? SWO_E_OOM : SWO_E_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                              \
    } while (0)

GenomeView genome_view(Device &d) {
    GenomeView G;
    G.packed = d.g_packed.as<uint8_t>();
    G.g_off = d.g_off.as<int64_t>();
    G.g_len = d.g_len.as<int64_t>();
    G.g_n = d.g_n;
    G.lower = d.g_has_lower ? d.g_lower.as<uint8_t>() : nullptr;
    G.exc_n = d.g_exc_n;
    G.exc_chunk = d.g_exc_n ? d.g_chunk.as<uint8_t>() : nullptr;
    G.exc_start = d.g_estart.as<int64_t>();
    G.exc_len = d.g_elen.as<int32_t>();
    G.exc_byte = d.g_ebyte.as<uint8_t>();
    return G;
}

template <typename T>
int upload(swo_ctx *c, DevBuf &b, const T *src, size_t n, cudaStream_t s) {
    CK(c, b.reserve(std::max<size_t>(n * sizeof(T), 16)));
    if (n) CK(c, cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return SWO_OK;
}

int upload_table(swo_ctx *c, Device &d) {
    const ChainTable &t = c->tab;
    CK(c, cudaSetDevice(d.ordinal));
    int rc;
    if ((rc = upload(c, d.blk, t.blocks.data(), t.blocks.size(), d.s_comp))) return rc;
    if ((rc = upload(c, d.tkey, t.tStartKey.data(), t.tStartKey.size(), d.s_comp))) return rc;
    if ((rc = upload(c, d.pkey, t.pmaxKey.data(), t.pmaxKey.size(), d.s_comp))) return rc;
    if ((rc = upload(c, d.toff, t.toff.data(), t.toff.size(), d.s_comp))) return rc;
    std::vector<int4> cmeta(t.ntc());
    for (int i = 0; i < t.ntc(); i++)
        cmeta[i] = make_int4(t.toff[i], t.toff[i + 1], t.eyoff[i], t.eyoff[i + 1] - t.eyoff[i]);
    std::vector<int2> ey(t.eyKey.size());
    for (size_t i = 0; i < ey.size(); i++) ey[i] = make_int2(t.eyKey[i], t.eyIdx[i]);
    if ((rc = upload(c, d.cmeta, cmeta.data(), cmeta.size(), d.s_comp))) return rc;
    if ((rc = upload(c, d.ey, ey.data(), ey.size(), d.s_comp))) return rc;
    // source-contig name table (device text parser)
    uint32_t hs = 16;
    while (hs < (uint32_t)t.ntc() * 4u) hs <<= 1;
    std::vector<NameEnt> ents(hs, NameEnt{0, -1, 0, 0, {0, 0, 0, 0}});
    std::string chars;
    std::vector<uint32_t> tnoff(1, 0);
    for (int i = 0; i < t.ntc(); i++) {
        const std::string &nm = t.tnames[i];
        NameEnt e{name_hash(nm.data(), nm.size()), i, (uint32_t)chars.size(), (uint32_t)nm.size(), {0, 0, 0, 0}};
        name_words(nm.data(), nm.size(), e.w);
        chars += nm;
        tnoff.push_back((uint32_t)chars.size());
        uint32_t s = e.hash & (hs - 1);
        while (ents[s].id >= 0) s = (s + 1) & (hs - 1);
        ents[s] = e;
    }
    chars.push_back('\0');
    d.thash_mask = hs - 1;
    if ((rc = upload(c, d.thash, ents.data(), ents.size(), d.s_comp))) return rc;
    if ((rc = upload(c, d.tchars, chars.data(), chars.size(), d.s_comp))) return rc;
    if ((rc = upload(c, d.tnoff, tnoff.data(), tnoff.size(), d.s_comp))) return rc;
    std::vector<uint32_t> qoff(1, 0);
    std::string qchars;
    for (const std::string &nm : t.qnames) {
        qchars += nm;
        qoff.push_back((uint32_t)qchars.size());
    }
    qchars.push_back('\0');
    if ((rc = upload(c, d.qoff, qoff.data(), qoff.size(), d.s_comp))) return rc;
    if ((rc = upload(c, d.qchars, qchars.data(), qchars.size(), d.s_comp))) return rc;
    CK(c, cudaStreamSynchronize(d.s_comp));
    d.T.blk = d.blk.as<Block>();
    d.T.tStartKey = d.tkey.as<int32_t>();
    d.T.pmaxKey = d.pkey.as<int32_t>();
    d.T.toff = d.toff.as<int32_t>();
    d.T.cmeta = d.cmeta.as<int4>();
    d.T.ey = d.ey.as<int2>();
    d.T.ey_n = (int)t.eyKey.size();
    d.T.ey_shift = t.eyShift;
    d.T.ntc = t.ntc();
    d.nq = (int)t.qnames.size();
    d.T.disjoint = t.disjoint ? 1 : 0;
    return SWO_OK;
}

// Zeroing of the per-launch scratch (ticket + look-back descriptors) and totals.  A kernel,
// not cudaMemsetAsync: memsets can be routed through a copy engine and then queue behind the
// multi-MiB H2D/D2H transfers of the pipeline (seen as multi-ms stalls of single chunks).
__global__ void zero_words_kernel(uint32_t *__restrict__ a, size_t na, uint32_t *__restrict__ b, size_t nb) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < na; i += stride) a[i] = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nb; i += stride) b[i] = 0;
}
int zero_async(swo_ctx *c, void *a, size_t abytes, void *b, size_t bbytes, cudaStream_t st) {
    const size_t na = (abytes + 3) / 4, nb = (bbytes + 3) / 4;
    const int grid = (int)std::min<size_t>(std::max<size_t>((std::max(na, nb) + 255) / 256, 1), 296);
    zero_words_kernel<<<grid, 256, 0, st>>>((uint32_t *)a, na, (uint32_t *)b, nb);
    c->launches++;
    CK(c, cudaGetLastError());
    return SWO_OK;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- binary path ---------------------------------------------------------------
int launch_lift_intervals(swo_ctx *c, Device &d, size_t n, const int32_t *tcid, const int32_t *start,
                          const int32_t *end, const uint8_t *strand, const int32_t *ths, const int32_t *the,
                          uint32_t *row_offset, swo_row *rows, swo_thick *thick, size_t rows_cap, uint64_t *totals,
                          cudaStream_t st) {
    const int ntiles = (int)((n + LI_TILE - 1) / LI_TILE);
    // scratch: [ticket, wide-list count, sorted flag, small-list count] [ntiles tile descriptors (per-query-search variants)]
    // sorted variant (count / scan / write over warp tiles of MG_QPW queries): [ntw totals] [nsum sums] [ntw lo0] [ntw x 32 packed words]
    const int ntw = (int)((n + MG_QPW - 1) / MG_QPW), nsum = (ntw + MG_SCAN_TILE - 1) / MG_SCAN_TILE;
    const size_t w_total = (16 + (size_t)ntiles * 8 + 15) & ~(size_t)15;
    const size_t w_sum = (w_total + (size_t)ntw * 4 + 15) & ~(size_t)15;
    const size_t w_lo0 = (w_sum + (size_t)(nsum + 1) * 8 + 15) & ~(size_t)15, w_packed = (w_lo0 + (size_t)ntw * 4 + 15) & ~(size_t)15;
    const size_t scratch_bytes = w_packed + (size_t)ntw * 32 * 4;  // (one byte per query between the two passes)
    const size_t zero_bytes = w_total;  // header + tile descriptors; the sorted variant's arrays are written before they are read
    if (d.scratch_used) CK(c, cudaStreamWaitEvent(st, d.ev_scratch, 0));
    CK(c, d.scratch.reserve((scratch_bytes + 3) & ~(size_t)3));
    if (int rc = zero_async(c, d.scratch.p, zero_bytes, totals, 16, st)) return rc;
    if (n == 0) {
        if (int rc = zero_async(c, row_offset, 4, nullptr, 0, st)) return rc;
        return SWO_OK;
    }
    LiftArgs A;
    A.T = d.T;
    A.n = (uint32_t)n;
    A.ntiles = ntiles;
    A.tcid = tcid;
    A.start = start;
    A.end = end;
    A.strand = strand;
    A.thick_s = ths;
    A.thick_e = the;
    A.row_offset = row_offset;
    A.rows = rows;
    A.thick = thick;
    A.rows_cap = rows_cap;
    A.ticket = d.scratch.as<unsigned>();
    A.desc = reinterpret_cast<uint64_t *>(d.scratch.as<char>() + 16);
    MergeTiles M;
    M.total = reinterpret_cast<uint32_t *>(d.scratch.as<char>() + w_total);
    M.sum = reinterpret_cast<unsigned long long *>(d.scratch.as<char>() + w_sum);
    M.lo0 = reinterpret_cast<int32_t *>(d.scratch.as<char>() + w_lo0);
    M.packed = reinterpret_cast<uint32_t *>(d.scratch.as<char>() + w_packed);
    A.gscr = nullptr;
    A.totals = totals;
    // sortedness sample -> flag word in the scratch header; both variants of the kernel
    // are launched and the one the flag does not name returns at once (no host round trip).
    // Sorted batches: the warp-merge kernel (kernels_merge.cuh).  Batches with thick columns always
    // take the immediate per-query-search variant (the two extra point searches per row dominate
    // there), so no sample for them.
    uint32_t *flag = reinterpret_cast<uint32_t *>(d.scratch.as<char>() + 8);
    A.sorted_flag = flag;
    A.full_mask = 0xffffffffu;
    const bool th = ths != nullptr;
    const bool merge_ok = d.T.disjoint != 0;  // the merge kernel and expand_warp_runs assume candidates == hits
    // Multi-hit queries of a target-disjoint table go to a list and are expanded by expand_wide_kernel behind
    // the main kernel (whichever variant ran).  A full list is not an error: the rest is expanded on the spot.
    WideList W;
    W.ent = nullptr;
    W.count = reinterpret_cast<unsigned *>(d.scratch.as<char>() + 4);
    W.cap = 0;
    W.count_small = reinterpret_cast<unsigned *>(d.scratch.as<char>() + 12);
    W.ent_small = nullptr;
    W.cap_small = 0;
    W.gscr_ctas = (unsigned)std::max(1, d.sms * 8);  // every CTA of expand_wide_kernel (and of merge_write_kernel) owns a region
    A.wide_ent = nullptr;
    A.wide_count = W.count;
    A.wide_cap = 0;
    A.small_ent = nullptr;
    A.small_count = W.count_small;
    A.small_cap = 0;
    if (merge_ok) {
        W.cap = (unsigned)std::min<size_t>(n, (size_t)32 << 20);
        W.cap_small = (unsigned)std::min<size_t>(n, (size_t)16 << 20);
        CK(c, d.wide.reserve(((size_t)W.cap + W.cap_small) * 2 * sizeof(uint4)));
        W.ent = d.wide.as<uint4>();
        W.ent_small = W.ent + 2 * (size_t)W.cap;
        CK(c, d.gscr.reserve((size_t)W.gscr_ctas * MG_NW * MG_GXS * sizeof(unsigned long long)));
        A.gscr = d.gscr.as<unsigned long long>();
        // (batches with thick columns keep their multi-hit queries inside lift_intervals_kernel: measured, BASELINE
        // configs[2], 9.8 ms against 11.4 ms through the list — there the list's entries are rare and heavy)
        if (!th) {
            A.wide_ent = W.ent;
            A.wide_cap = W.cap;
            A.small_ent = W.ent_small;
            A.small_cap = W.cap_small;
        }
    }
    if (!th) {
        sample_sorted_kernel<<<1, 512, 0, st>>>(tcid, start, (uint32_t)n, flag, merge_ok ? 1u : 0u);
        c->launches++;
    }
    for (int v = th ? 2 : 0; v < (th ? 3 : 2); v++) {
        if (v == 1) {
            // Sorted batches (flag set on the device; otherwise each of these returns at once): count pass, scan of the
            // warp tiles' totals, write pass.
            if (!merge_ok) continue;
            const void *fw = strand ? (const void *)merge_write_kernel<true> : (const void *)merge_write_kernel<false>;
            const size_t smem = sizeof(MergeSmem);
            if (!d.bin_grid[1]) {
                int per = 0;
                CK(c, cudaFuncSetAttribute((const void *)merge_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, (const void *)merge_count_kernel, MG_TPB, smem));
                d.bin_grid[1] = std::max(1, per) * d.sms;
            }
            const int gw = strand ? 4 : 3;
            if (!d.bin_grid[gw]) {
                int per = 0;
                CK(c, cudaFuncSetAttribute(fw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fw, MG_TPB, smem));
                d.bin_grid[gw] = std::min(std::max(1, per), 8) * d.sms;  // (the global sort scratch has regions for 8 CTAs per SM)
            }
            LiftArgs Aw = A;
            Aw.ntiles = ntw;
            const int need = (ntw + MG_NW - 1) / MG_NW;
            const int grid_c = std::min(need, d.bin_grid[1]), grid_w = std::min(need, d.bin_grid[gw]);
#ifdef MG_DEBUG_COUNTS
            M.dbg = nullptr;
            if (getenv("SWO_DUMP_COUNTS")) cudaMalloc(&M.dbg, n * 16);
#endif
            merge_count_kernel<<<grid_c, MG_TPB, smem, st>>>(Aw, M);
#ifdef MG_DEBUG_COUNTS
            if (M.dbg) {
                cudaStreamSynchronize(st);
                std::vector<int32_t> h(n * 4);
                std::vector<uint32_t> tot(ntw);
                cudaMemcpy(h.data(), M.dbg, n * 16, cudaMemcpyDeviceToHost);
                cudaMemcpy(tot.data(), M.total, (size_t)ntw * 4, cudaMemcpyDeviceToHost);
                FILE *f = fopen(getenv("SWO_DUMP_COUNTS"), "wb");
                if (f) { fwrite(h.data(), 4, h.size(), f); fwrite(tot.data(), 4, tot.size(), f); fclose(f); }
                cudaFree(M.dbg);
                M.dbg = nullptr;
            }
#endif
            scan_tiles_kernel1<<<nsum, 256, 0, st>>>(flag, M, ntw);
            scan_tiles_kernel2<<<1, 1024, 0, st>>>(flag, M, nsum, totals, row_offset + n);
            void *wargs[] = {(void *)&Aw, (void *)&M, (void *)&W};
            CK(c, cudaLaunchKernel(fw, dim3(grid_w), dim3(MG_TPB), wargs, smem, st));
            c->launches += 4;
            continue;
        }
        const void *fn = v == 0 ? (const void *)lift_intervals_kernel<false, false> : (const void *)lift_intervals_kernel<true, false>;
        const size_t smem = v == 0 ? sizeof(LiftSmem<false, false>) : sizeof(LiftSmem<true, false>);
        if (!d.bin_grid[v]) {
            int per = 0;
            CK(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fn, LI_TPB, smem));
            d.bin_grid[v] = std::max(1, per) * d.sms;
        }
        const int grid = std::min(ntiles, d.bin_grid[v]);
        void *args[] = {(void *)&A};
        CK(c, cudaLaunchKernel(fn, dim3(grid), dim3(LI_TPB), args, smem, st));
        c->launches++;
    }
    if (merge_ok && !th) {  // all the warps the SMs hold: a warp per listed query
        expand_small_kernel<<<d.sms * 8, 256, 0, st>>>(A, W);
        expand_wide_kernel<false><<<d.sms * 8, 256, 0, st>>>(A, W);
        c->launches += 2;
    }
    CK(c, cudaGetLastError());
    CK(c, cudaEventRecord(d.ev_scratch, st));
    d.scratch_used = true;
    return SWO_OK;
}

// ---- text path -------------------------------------------------------------------
// d_in: 16-byte aligned; the text is d_in[skip, in_len) and starts with a line
int launch_lift_text(swo_ctx *c, Device &d, const char *d_in, size_t in_len, char *d_lifted, size_t lifted_cap,
                     char *d_unm, size_t unm_cap, uint64_t *d_sizes, cudaStream_t st, uint32_t skip = 0,
                     uint32_t tile_bytes = 0) {
    if (!tile_bytes) tile_bytes = c->tile_bytes;
    const int ntiles = (int)((in_len + tile_bytes - 1) / tile_bytes);
    const size_t scratch_bytes = 16 + (size_t)ntiles * 16;
    if (d.scratch_used) CK(c, cudaStreamWaitEvent(st, d.ev_scratch, 0));
    CK(c, d.scratch.reserve((scratch_bytes + 3) & ~(size_t)3));
    if (int rc = zero_async(c, d.scratch.p, scratch_bytes, d_sizes, 48, st)) return rc;
    if (in_len == 0) return SWO_OK;
    TextArgs A;
    A.T = d.T;
    A.full_mask = 0xffffffffu;
    A.thash = d.thash.as<NameEnt>();
    A.thash_mask = d.thash_mask;
    A.tname_chars = d.tchars.as<char>();
    A.qname_off = d.qoff.as<uint32_t>();
    A.qname_chars = d.qchars.as<char>();
    A.nq = d.nq;
    A.in = d_in;
    A.skip = skip;
    A.tile_bytes = tile_bytes;
    A.in_len = in_len;
    A.ntiles = ntiles;
    A.lifted = d_lifted;
    A.lifted_cap = lifted_cap;
    A.unm = d_unm;
    A.unm_cap = unm_cap;
    A.ticket = d.scratch.as<unsigned>();
    A.desc_l = reinterpret_cast<uint64_t *>(d.scratch.as<char>() + 16);
    A.desc_u = A.desc_l + ntiles;
    A.sizes = d_sizes;
    const size_t smem = sizeof(TextSmem);
    if (!d.text_grid) {
        CK(c, cudaFuncSetAttribute(lift_bed_text_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per = 0;
        CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, lift_bed_text_kernel, TX_TPB, smem));
        d.text_grid = std::max(1, per) * d.sms;
        if (getenv("SWO_TRACE")) fprintf(stderr, "[swo] text kernel: %zu B shared memory, %d CTAs/SM x %d threads\n", smem, per, TX_TPB);
    }
    const int grid = std::min(ntiles, d.text_grid);
    CK(c, d.gscr.reserve((size_t)d.text_grid * (TX_GSCR + TX_GSCR / 2) * sizeof(unsigned long long)));
    A.gscr = d.gscr.as<unsigned long long>();
    lift_bed_text_kernel<<<grid, TX_TPB, smem, st>>>(A);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaEventRecord(d.ev_scratch, st));
    d.scratch_used = true;
    return SWO_OK;
}

struct ShardResult {
    size_t lifted = 0, unm = 0;
    uint64_t rec = 0, rows = 0, nunm = 0;
    int rc = SWO_OK;
    uint64_t err_off = 0;
    bool has_err = false;
};

// Lift the byte range [a,b) of `in` (line aligned) on device d, streaming chunks
// H2D -> kernel -> D2H.  Outputs are appended to (*pl)[base_l...] / (*pu)[base_u...].
int run_shard(swo_ctx *c, Device &d, const char *in, size_t a, size_t b, PinBuf *pl, size_t base_l, PinBuf *pu,
              size_t base_u, ShardResult &res) {
    CK(c, cudaSetDevice(d.ordinal));
    CK(c, d.h_sizes.reserve(N_SLOTS * 48));
    uint64_t *hs = d.h_sizes.as<uint64_t>();
    // chunk boundaries (line aligned).  Sizes ramp up from chunk/16 so that the first kernel and the
    // first D2H start early, and ramp down at the end so that little is left to drain.
    std::vector<size_t> cuts(1, a);
    {
        const size_t chunk = c->chunk_bytes, small = chunk / 16 + 1;
        size_t step = small;
        while (cuts.back() < b) {
            const size_t pos = cuts.back(), left = b - pos;
            size_t sz = std::min(step, chunk);
            if (left < 2 * chunk) sz = std::min(sz, std::max(left / 2, small));
            step = std::min(step * 2, chunk);
            size_t e = std::min(b, pos + sz);
            if (e < b) {
                const void *nl = memchr(in + e, '\n', b - e);
                e = nl ? (size_t)((const char *)nl - in) + 1 : b;
            }
            cuts.push_back(e);
        }
    }
    const int nch = (int)cuts.size() - 1;
    size_t out_l = base_l, out_u = base_u;
    // Pinned (or registered) host input is read by the kernel itself over PCIe (UVA zero-copy:
    // measured at the full H2D memcpy rate), so there is no H2D copy and no staging buffer;
    // pageable input is staged through t_in with cudaMemcpyAsync.
    const char *zc_in = nullptr;
    {
        cudaPointerAttributes at;
        if (c->zero_copy_input && cudaPointerGetAttributes(&at, in) == cudaSuccess && at.type == cudaMemoryTypeHost &&
            at.devicePointer)
            zc_in = static_cast<const char *>(at.devicePointer);
        else
            cudaGetLastError();
    }
    std::vector<uint32_t> skips(nch, 0);
    // tile size of the next launches: follows the output/input ratio of the chunks seen so far
    uint32_t tile_now = c->tile_bytes;
    // SWO_TRACE=1: per-chunk timeline (CUDA events on the three streams) printed to stderr
    static const bool trace = getenv("SWO_TRACE") != nullptr;
    struct Tr { cudaEvent_t h0, h1, k0, k1, o0, o1; };
    std::vector<Tr> tr(trace ? nch : 0);
    auto mark = [&](cudaEvent_t *e, cudaStream_t st) {
        if (!trace) return;
        cudaEventCreate(e);
        cudaEventRecord(*e, st);
    };
    auto launch_chunk = [&](int k) -> int {
        const int s = k % N_SLOTS;
        const size_t off = cuts[k], len = cuts[k + 1] - cuts[k];
        if (zc_in) {
            const uint32_t skip = (uint32_t)(reinterpret_cast<uintptr_t>(zc_in + off) & 15u);
            skips[k] = skip;
            return launch_lift_text(c, d, zc_in + off - skip, len + skip, d.t_out[s].as<char>(), d.t_out[s].cap,
                                    d.t_unm[s].as<char>(), d.t_unm[s].cap, d.t_sizes[s].as<uint64_t>(), d.s_comp, skip,
                                    tile_now);
        }
        return launch_lift_text(c, d, d.t_in[s].as<char>(), len, d.t_out[s].as<char>(), d.t_out[s].cap,
                                d.t_unm[s].as<char>(), d.t_unm[s].cap, d.t_sizes[s].as<uint64_t>(), d.s_comp, 0, tile_now);
    };
    auto sizes_to_host = [&](int s) -> int {
        copy_sizes_kernel<<<1, 32, 0, d.s_comp>>>(d.t_sizes[s].as<uint64_t>(), hs + 6 * s);
        c->launches++;
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(d.ev_comp[s], d.s_comp));
        return SWO_OK;
    };
    auto issue = [&](int k) -> int {  // (H2D +) kernel + sizes readback for chunk k
        const int s = k % N_SLOTS;
        const size_t off = cuts[k], len = cuts[k + 1] - cuts[k];
        CK(c, d.t_out[s].reserve(len + (len >> 1) + 65536));
        CK(c, d.t_unm[s].reserve(len + 64));
        CK(c, d.t_sizes[s].reserve(48));
        if (zc_in) {
            CK(c, cudaStreamWaitEvent(d.s_comp, d.ev_d2h[s], 0));  // slot's previous outputs are on the host
        } else {
            CK(c, d.t_in[s].reserve(len + 32));
            CK(c, cudaStreamWaitEvent(d.s_h2d, d.ev_d2h[s], 0));
            if (trace) mark(&tr[k].h0, d.s_h2d);
            CK(c, cudaMemcpyAsync(d.t_in[s].p, in + off, len, cudaMemcpyHostToDevice, d.s_h2d));
            if (trace) mark(&tr[k].h1, d.s_h2d);
            CK(c, cudaEventRecord(d.ev_h2d[s], d.s_h2d));
            CK(c, cudaStreamWaitEvent(d.s_comp, d.ev_h2d[s], 0));
        }
        if (trace) mark(&tr[k].k0, d.s_comp);
        int rc = launch_chunk(k);
        if (rc) return rc;
        if (trace) mark(&tr[k].k1, d.s_comp);
        return sizes_to_host(s);
    };
    int rc;
    int next_issue = 0;
    for (int k = 0; k < nch; k++) {
        const int s = k % N_SLOTS;
        // keep N_SLOTS chunks in flight; chunk j reuses the slot of chunk j - N_SLOTS, whose
        // D2H was queued in an earlier iteration
        while (next_issue < nch && next_issue < k + N_SLOTS)
            if ((rc = issue(next_issue++))) return rc;
        CK(c, cudaEventSynchronize(d.ev_comp[s]));
        uint64_t sz[6];
        memcpy(sz, hs + 6 * s, 48);
        const size_t len = cuts[k + 1] - cuts[k];
        if (sz[0] > d.t_out[s].cap) {
            // fan-out larger than the guess: grow this slot's output and redo the chunk
            CK(c, cudaStreamSynchronize(d.s_comp));
            CK(c, d.t_out[s].reserve((size_t)sz[0] + 4096));
            if ((rc = launch_chunk(k))) return rc;
            if ((rc = sizes_to_host(s))) return rc;
            CK(c, cudaEventSynchronize(d.ev_comp[s]));
            memcpy(sz, hs + 6 * s, 48);
        }
        if (c->tile_auto && len > 65536) {  // lifted bytes per input byte of this chunk -> tile size of later launches
            const double ratio = (double)sz[0] / (double)len;
            // measured on the reference's fixtures: 4 KiB tiles win for moderate fan-out (2.3 bytes out per byte
            // in: 570 vs 380 M lines/s); for very large fan-out every tile overflows anyway and big tiles win
            tile_now = (ratio >= 1.15 && ratio < 4.0) ? 4096u : (uint32_t)TX_TILE;
        }
        if (sz[5] && !res.has_err) {
            res.has_err = true;
            res.err_off = cuts[k] + (UINT64_MAX - sz[5]) - skips[k];
        }
        // grow the host arenas if needed (keeps what is already there)
        if (out_l + sz[0] > pl->cap) {
            CK(c, cudaStreamSynchronize(d.s_d2h));
            CK(c, cudaStreamSynchronize(d.s_d2h2));
            CK(c, pl->reserve(out_l + sz[0], true, out_l));
        }
        if (out_u + sz[1] > pu->cap) {
            CK(c, cudaStreamSynchronize(d.s_d2h));
            CK(c, cudaStreamSynchronize(d.s_d2h2));
            CK(c, pu->reserve(out_u + sz[1], true, out_u));
        }
        cudaStream_t sd = (c->d2h_streams > 1 && (k & 1)) ? d.s_d2h2 : d.s_d2h;
        CK(c, cudaStreamWaitEvent(sd, d.ev_comp[s], 0));
        if (trace) mark(&tr[k].o0, sd);
        if (sz[0]) CK(c, cudaMemcpyAsync(pl->as<char>() + out_l, d.t_out[s].p, sz[0], cudaMemcpyDeviceToHost, sd));
        if (sz[1]) CK(c, cudaMemcpyAsync(pu->as<char>() + out_u, d.t_unm[s].p, sz[1], cudaMemcpyDeviceToHost, sd));
        if (trace) mark(&tr[k].o1, sd);
        CK(c, cudaEventRecord(d.ev_d2h[s], sd));
        out_l += sz[0];
        out_u += sz[1];
        res.rec += sz[2];
        res.rows += sz[3];
        res.nunm += sz[4];
    }
    CK(c, cudaStreamSynchronize(d.s_d2h));
    CK(c, cudaStreamSynchronize(d.s_d2h2));
    if (trace && nch && !zc_in) {
        auto ms = [&](cudaEvent_t e) {
            float f = 0;
            cudaEventElapsedTime(&f, tr[0].h0, e);
            return f;
        };
        fprintf(stderr, "chunk   h2d_start h2d_end  k_start  k_end    d2h_start d2h_end (ms)\n");
        for (int k = 0; k < nch; k++) {
            fprintf(stderr, "%5d %9.3f %8.3f %8.3f %8.3f %9.3f %8.3f\n", k, ms(tr[k].h0), ms(tr[k].h1), ms(tr[k].k0),
                    ms(tr[k].k1), ms(tr[k].o0), ms(tr[k].o1));
        }
        for (int k = 0; k < nch; k++)
            for (cudaEvent_t e : {tr[k].h0, tr[k].h1, tr[k].k0, tr[k].k1, tr[k].o0, tr[k].o1}) cudaEventDestroy(e);
    }
    res.lifted = out_l - base_l;
    res.unm = out_u - base_u;
    return SWO_OK;
}

}  // namespace

// =================================================================================
extern "C" {

const char *swo_version(void) { return "swiftover_cuda 0.1 sm_100a"; }

int swo_create(swo_ctx **out, const int *devices, int ndev) {
    if (!out) return SWO_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return SWO_E_CUDA;  // no CPU fallback
    std::unique_ptr<swo_ctx> c(new swo_ctx);
    std::vector<int> ords;
    if (!devices || ndev <= 0) ords.push_back(0);
    else ords.assign(devices, devices + ndev);
    for (int o : ords) {
        if (o < 0 || o >= count) return SWO_E_ARG;
        Device d;
        d.ordinal = o;
        if (cudaSetDevice(o) != cudaSuccess) return SWO_E_CUDA;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, o) != cudaSuccess) return SWO_E_CUDA;
        d.sms = p.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&d.s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.s_comp, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.s_d2h2, cudaStreamNonBlocking) != cudaSuccess)
            return SWO_E_CUDA;
        for (int s = 0; s < N_SLOTS; s++) {
            if (!d.ev_scratch) cudaEventCreateWithFlags(&d.ev_scratch, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&d.ev_h2d[s], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&d.ev_comp[s], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&d.ev_d2h[s], cudaEventDisableTiming);
        }
        c->dev.push_back(d);
    }
    *out = c.release();
    return SWO_OK;
}

int swo_create_inspect(swo_ctx **out) {
    if (!out) return SWO_E_ARG;
    *out = new swo_ctx;  // no devices: chain parsing + table inspection only
    return SWO_OK;
}

void swo_destroy(swo_ctx *c) {
    if (!c) return;
    for (Device &d : c->dev) {
        cudaSetDevice(d.ordinal);
        cudaDeviceSynchronize();
        for (DevBuf *b : {&d.blk, &d.tkey, &d.pkey, &d.toff, &d.cmeta, &d.ey, &d.thash, &d.tchars, &d.tnoff, &d.qoff, &d.qchars, &d.g_packed,
                          &d.g_off, &d.g_len, &d.g_lower, &d.g_chunk, &d.g_estart, &d.g_elen, &d.g_ebyte, &d.scratch, &d.gscr, &d.wide, &d.b_in, &d.b_off, &d.b_rows, &d.b_thick, &d.b_tot})
            b->release();
        if (d.ev_scratch) cudaEventDestroy(d.ev_scratch);
        for (int s = 0; s < N_SLOTS; s++) {
            d.t_in[s].release();
            d.t_out[s].release();
            d.t_unm[s].release();
            d.t_sizes[s].release();
            cudaEventDestroy(d.ev_h2d[s]);
            cudaEventDestroy(d.ev_comp[s]);
            cudaEventDestroy(d.ev_d2h[s]);
        }
        d.h_sizes.release();
        d.a_lifted.release();
        d.a_unm.release();
        cudaStreamDestroy(d.s_h2d);
        cudaStreamDestroy(d.s_comp);
        cudaStreamDestroy(d.s_d2h);
        cudaStreamDestroy(d.s_d2h2);
    }
    for (PinBuf *b : {&c->r_lifted, &c->r_unm, &c->r_off, &c->r_rows, &c->r_thick}) b->release();
    delete c;
}

const char *swo_last_error(const swo_ctx *c) { return c ? c->err.c_str() : "null ctx"; }
uint64_t swo_launch_count(const swo_ctx *c) { return c ? c->launches : 0; }

void *swo_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void swo_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int swo_set_option(swo_ctx *c, const char *key, int64_t value) {
    if (!c || !key) return SWO_E_ARG;
    if (!strcmp(key, "chunk_bytes")) {
        if (value < 64) return fail(c, SWO_E_ARG, "chunk_bytes must be >= 64");
        c->chunk_bytes = (size_t)value;
        return SWO_OK;
    }
    if (!strcmp(key, "tile_bytes")) {
        if (value == 0) {
            c->tile_auto = true;
            c->tile_bytes = TX_TILE;
            return SWO_OK;
        }
        if ((value != 4096 && value != 8192 && value != 16384) || value > TX_TILE)
            return fail(c, SWO_E_ARG, "tile_bytes must be 0 (auto), 4096, 8192 or 16384");
        c->tile_bytes = (uint32_t)value;
        c->tile_auto = false;
        return SWO_OK;
    }
    if (!strcmp(key, "concat")) {
        c->concat = value != 0;
        return SWO_OK;
    }
    if (!strcmp(key, "d2h_streams")) {
        c->d2h_streams = value >= 2 ? 2 : 1;
        return SWO_OK;
    }
    if (!strcmp(key, "zero_copy_input")) {
        c->zero_copy_input = value != 0;
        return SWO_OK;
    }
    return fail(c, SWO_E_ARG, "unknown option %s", key);
}

// ---- chain ------------------------------------------------------------------------
int swo_load_chain_mem(swo_ctx *c, const char *text, size_t len) {
    if (!c || (!text && len)) return SWO_E_ARG;
    std::string err;
    ChainTable t;
    if (!parse_chain_text(std::string_view(text, len), t, err)) return fail(c, SWO_E_PARSE, "%s", err.c_str());
    c->tab = std::move(t);
    c->have_chain = false;
    for (Device &d : c->dev) {
        int rc = upload_table(c, d);
        if (rc) return rc;
    }
    c->have_chain = true;
    return SWO_OK;
}

int swo_load_chain(swo_ctx *c, const char *path) {
    if (!c || !path) return SWO_E_ARG;
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(c, SWO_E_IO, "File does not exist: %s", path);  // chain.d:513-514
    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return swo_load_chain_mem(c, text.data(), text.size());
}

int swo_num_tcontigs(const swo_ctx *c) { return c ? c->tab.ntc() : 0; }
const char *swo_tcontig_name(const swo_ctx *c, int i) {
    return (c && i >= 0 && i < c->tab.ntc()) ? c->tab.tnames[i].c_str() : nullptr;
}
int swo_tcontig_id(const swo_ctx *c, const char *name, size_t len) {
    return c ? c->tab.tcontig_id(std::string_view(name, len)) : -1;
}
int64_t swo_tcontig_nblocks(const swo_ctx *c, int i) {
    return (c && i >= 0 && i < c->tab.ntc()) ? (int64_t)c->tab.toff[i + 1] - c->tab.toff[i] : 0;
}
int64_t swo_tcontig_size(const swo_ctx *c, int i) {
    return (c && i >= 0 && i < c->tab.ntc()) ? c->tab.tsizes[i] : -1;
}
int64_t swo_num_blocks(const swo_ctx *c) { return c ? c->tab.nblocks() : 0; }
int swo_table_is_disjoint(const swo_ctx *c) { return c && c->tab.disjoint; }
int swo_num_qcontigs(const swo_ctx *c) { return c ? (int)c->tab.qnames.size() : 0; }
const char *swo_qcontig_name(const swo_ctx *c, int i) {
    return (c && i >= 0 && i < (int)c->tab.qnames.size()) ? c->tab.qnames[i].c_str() : nullptr;
}
int swo_num_qsizes(const swo_ctx *c) { return c ? (int)c->tab.qsizes.size() : 0; }
const char *swo_qsize_name(const swo_ctx *c, int i) {
    return (c && i >= 0 && i < (int)c->tab.qsizes.size()) ? c->tab.qsizes[i].first.c_str() : nullptr;
}
int64_t swo_qsize_value(const swo_ctx *c, int i) {
    return (c && i >= 0 && i < (int)c->tab.qsizes.size()) ? c->tab.qsizes[i].second : -1;
}

int swo_get_blocks(const swo_ctx *c, int tcid, int64_t first, int64_t n, int32_t *ts, int32_t *te, int32_t *qs,
                   int32_t *qcid, int8_t *inv) {
    if (!c || tcid < 0 || tcid >= c->tab.ntc()) return SWO_E_ARG;
    const int64_t a = c->tab.toff[tcid], b = c->tab.toff[tcid + 1];
    if (first < 0 || n < 0 || a + first + n > b) return SWO_E_ARG;
    for (int64_t i = 0; i < n; i++) {
        const Block &k = c->tab.blocks[a + first + i];
        if (ts) ts[i] = k.tStart;
        if (te) te[i] = k.tEnd;
        if (qs) qs[i] = k.qStart;
        if (qcid) qcid[i] = k.meta >> 1;
        if (inv) inv[i] = (k.meta & 1) ? -1 : 1;
    }
    return SWO_OK;
}

// ---- binary lift --------------------------------------------------------------------
int swo_lift_intervals_dev(swo_ctx *c, size_t n, const int32_t *d_tcid, const int32_t *d_start, const int32_t *d_end,
                           const uint8_t *d_strand, const int32_t *d_ths, const int32_t *d_the,
                           uint32_t *d_row_offset, swo_row *d_rows, swo_thick *d_thick, size_t rows_cap,
                           uint64_t *d_totals, void *stream) {
    if (!c) return SWO_E_ARG;
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    if (n > 0xFFFFF000ull) return fail(c, SWO_E_ARG, "n must be at most 2^32 - 4096 per call");  // (tile index arithmetic is 32-bit)
    if ((d_ths == nullptr) != (d_the == nullptr) || (d_ths && !d_thick)) return fail(c, SWO_E_ARG, "thick arrays");
    if (!aligned16(d_tcid) || !aligned16(d_start) || !aligned16(d_end) || !aligned16(d_row_offset) ||
        !aligned16(d_rows))
        return fail(c, SWO_E_ARG, "device arrays must be 16-byte aligned");
    Device &d = c->dev[0];
    CK(c, cudaSetDevice(d.ordinal));
    return launch_lift_intervals(c, d, n, d_tcid, d_start, d_end, d_strand, d_ths, d_the, d_row_offset, d_rows,
                                 d_thick, rows_cap, d_totals, (cudaStream_t)stream);
}

int swo_lift_intervals(swo_ctx *c, size_t n, const int32_t *tcid, const int32_t *start, const int32_t *end,
                       const uint8_t *strand, const int32_t *ths, const int32_t *the, swo_lifted *out) {
    if (!c || !out) return SWO_E_ARG;
    memset(out, 0, sizeof *out);
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    if (n > 0xFFFFF000ull) return fail(c, SWO_E_ARG, "n must be at most 2^32 - 4096 per call");  // (tile index arithmetic is 32-bit)
    if ((ths == nullptr) != (the == nullptr)) return fail(c, SWO_E_ARG, "thick_start and thick_end go together");
    Device &d = c->dev[0];
    CK(c, cudaSetDevice(d.ordinal));
    cudaStream_t st = d.s_comp;
    // device input block: [tcid][start][end][thick_s][thick_e][strand], each 16-byte aligned
    const size_t na = (n * 4 + 15) & ~(size_t)15;
    const size_t total_in = na * 5 + ((n + 15) & ~(size_t)15) + 16;
    CK(c, d.b_in.reserve(total_in));
    char *base = d.b_in.as<char>();
    int32_t *d_tcid = (int32_t *)base, *d_start = (int32_t *)(base + na), *d_end = (int32_t *)(base + 2 * na);
    int32_t *d_ths = ths ? (int32_t *)(base + 3 * na) : nullptr, *d_the = ths ? (int32_t *)(base + 4 * na) : nullptr;
    uint8_t *d_strand = strand ? (uint8_t *)(base + 5 * na) : nullptr;
    if (n) {
        CK(c, cudaMemcpyAsync(d_tcid, tcid, n * 4, cudaMemcpyHostToDevice, st));
        CK(c, cudaMemcpyAsync(d_start, start, n * 4, cudaMemcpyHostToDevice, st));
        CK(c, cudaMemcpyAsync(d_end, end, n * 4, cudaMemcpyHostToDevice, st));
        if (ths) {
            CK(c, cudaMemcpyAsync(d_ths, ths, n * 4, cudaMemcpyHostToDevice, st));
            CK(c, cudaMemcpyAsync(d_the, the, n * 4, cudaMemcpyHostToDevice, st));
        }
        if (strand) CK(c, cudaMemcpyAsync(d_strand, strand, n, cudaMemcpyHostToDevice, st));
    }
    CK(c, d.b_off.reserve((n + 1) * 4 + 16));
    CK(c, d.b_tot.reserve(16));
    size_t cap = std::max(d.b_rows.cap / sizeof(swo_row), n + (n >> 2) + 1024);
    uint64_t tot[2] = {0, 0};
    for (int attempt = 0; attempt < 2; attempt++) {
        CK(c, d.b_rows.reserve(cap * sizeof(swo_row)));
        if (ths) CK(c, d.b_thick.reserve(cap * sizeof(swo_thick)));
        int rc = launch_lift_intervals(c, d, n, d_tcid, d_start, d_end, d_strand, d_ths, d_the, d.b_off.as<uint32_t>(),
                                       d.b_rows.as<swo_row>(), ths ? d.b_thick.as<swo_thick>() : nullptr, cap,
                                       d.b_tot.as<uint64_t>(), st);
        if (rc) return rc;
        CK(c, cudaMemcpyAsync(tot, d.b_tot.p, 16, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        if (tot[0] <= cap) break;
        if (tot[0] > 0xFFFFFFFFull) return fail(c, SWO_E_ARG, "more than 2^32 rows in one call; split the batch");
        cap = (size_t)tot[0];
    }
    CK(c, c->r_off.reserve((n + 1) * 4));
    CK(c, c->r_rows.reserve(std::max<size_t>(tot[0], 1) * sizeof(swo_row)));
    if (ths) CK(c, c->r_thick.reserve(std::max<size_t>(tot[0], 1) * sizeof(swo_thick)));
    CK(c, cudaMemcpyAsync(c->r_off.p, d.b_off.p, (n + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (tot[0]) {
        CK(c, cudaMemcpyAsync(c->r_rows.p, d.b_rows.p, tot[0] * sizeof(swo_row), cudaMemcpyDeviceToHost, st));
        if (ths) CK(c, cudaMemcpyAsync(c->r_thick.p, d.b_thick.p, tot[0] * sizeof(swo_thick), cudaMemcpyDeviceToHost, st));
    }
    CK(c, cudaStreamSynchronize(st));
    out->n_rows = tot[0];
    out->n_unmatched = tot[1];
    out->row_offset = c->r_off.as<uint32_t>();
    out->rows = c->r_rows.as<swo_row>();
    out->thick = ths ? c->r_thick.as<swo_thick>() : nullptr;
    return SWO_OK;
}

int swo_lift_points(swo_ctx *c, size_t n, const int32_t *tcid, const int32_t *pos, int32_t *out_qcid,
                    int32_t *out_pos, uint8_t *hit) {
    if (!c || (n && (!tcid || !pos || !out_qcid || !out_pos || !hit))) return SWO_E_ARG;
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    if (n >= 0xFFFFFFFFull) return fail(c, SWO_E_ARG, "n must be < 2^32-1 per call");
    if (!n) return SWO_OK;
    Device &d = c->dev[0];
    CK(c, cudaSetDevice(d.ordinal));
    cudaStream_t st = d.s_comp;
    const size_t na = (n * 4 + 15) & ~(size_t)15;
    CK(c, d.b_in.reserve(na * 4 + n + 16));
    char *base = d.b_in.as<char>();
    int32_t *d_tcid = (int32_t *)base, *d_pos = (int32_t *)(base + na), *d_oq = (int32_t *)(base + 2 * na),
            *d_op = (int32_t *)(base + 3 * na);
    uint8_t *d_hit = (uint8_t *)(base + 4 * na);
    CK(c, cudaMemcpyAsync(d_tcid, tcid, n * 4, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(d_pos, pos, n * 4, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)d.sms * 8);
    lift_points_kernel<<<grid, 256, 0, st>>>(d.T, (uint32_t)n, d_tcid, d_pos, d_oq, d_op, d_hit);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out_qcid, d_oq, n * 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(out_pos, d_op, n * 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(hit, d_hit, n, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return SWO_OK;
}

// ---- text lift ------------------------------------------------------------------------
int swo_lift_bed_text_dev(swo_ctx *c, const char *d_in, size_t in_len, char *d_lifted, size_t lifted_cap,
                          char *d_unm, size_t unm_cap, uint64_t *d_sizes, void *stream) {
    if (!c || !d_sizes || (in_len && (!d_in || !d_lifted || !d_unm))) return SWO_E_ARG;
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    if (!aligned16(d_in)) return fail(c, SWO_E_ARG, "d_in must be 16-byte aligned");
    Device &d = c->dev[0];
    CK(c, cudaSetDevice(d.ordinal));
    return launch_lift_text(c, d, d_in, in_len, d_lifted, lifted_cap, d_unm, unm_cap, d_sizes, (cudaStream_t)stream);
}

int swo_lift_bed_text(swo_ctx *c, const char *in, size_t in_len, swo_text_out *out) {
    if (!c || !out || (in_len && !in)) return SWO_E_ARG;
    memset(out, 0, sizeof *out);
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    const int nd = (int)c->dev.size();
    // byte-range shards; a shard owns the lines that start inside it
    std::vector<size_t> cut(nd + 1, in_len);
    cut[0] = 0;
    for (int k = 1; k < nd; k++) {
        size_t e = (size_t)((unsigned __int128)in_len * (unsigned)k / (unsigned)nd);
        e = std::max(e, cut[k - 1]);
        while (e < in_len && e > 0 && in[e - 1] != '\n') e++;
        cut[k] = e;
    }
    std::vector<ShardResult> res(nd);
    CK(c, c->r_lifted.reserve(in_len + (in_len >> 2) + 4096));
    CK(c, c->r_unm.reserve(in_len / 4 + 4096));
    if (nd == 1) {
        int rc = run_shard(c, c->dev[0], in, 0, in_len, &c->r_lifted, 0, &c->r_unm, 0, res[0]);
        if (rc) return rc;
    } else {
        // One host thread per shard.  Shard 0 streams straight into the result arenas; the others into
        // their own pinned arenas.  A shard's place in the result is known once every shard before it
        // has finished, so when a thread is done it waits for its predecessors' sizes and copies its
        // arena to its place itself: the copies of all shards run at the same time (the first version
        // of this did them one after the other on the calling thread: 19 GB at 8 x 2.4 GB).
        std::vector<std::thread> th;
        std::vector<std::string> emsg(nd);
        std::vector<int> fin(nd, 0);
        std::mutex mu;
        std::condition_variable cv;
        const bool concat = c->concat;
        for (int k = 0; k < nd; k++) {
            th.emplace_back([&, k]() {
                swo_ctx local;  // private error sink; run_shard only touches err/launches/chunk_bytes
                local.chunk_bytes = c->chunk_bytes;
                local.zero_copy_input = c->zero_copy_input;
                local.d2h_streams = c->d2h_streams;
                local.tile_bytes = c->tile_bytes;
                local.tile_auto = c->tile_auto;
                Device &d = c->dev[k];
                PinBuf *pl = k == 0 ? &c->r_lifted : &d.a_lifted, *pu = k == 0 ? &c->r_unm : &d.a_unm;
                res[k].rc = run_shard(&local, d, in, cut[k], cut[k + 1], pl, 0, pu, 0, res[k]);
                emsg[k] = local.err;
                __atomic_fetch_add(&c->launches, local.launches, __ATOMIC_RELAXED);
                size_t ol = 0, ou = 0;
                bool ok = res[k].rc == SWO_OK;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    fin[k] = 1;
                    cv.notify_all();
                    if (k > 0 && concat) {
                        cv.wait(lk, [&] {
                            for (int j = 0; j < k; j++)
                                if (!fin[j]) return false;
                            return true;
                        });
                        for (int j = 0; j < k; j++) {
                            ok = ok && res[j].rc == SWO_OK;
                            ol += res[j].lifted;
                            ou += res[j].unm;
                        }
                    }
                }
                if (k > 0 && concat && ok) {
                    if (ol + res[k].lifted <= c->r_lifted.cap && ou + res[k].unm <= c->r_unm.cap) {
                        memcpy(c->r_lifted.as<char>() + ol, d.a_lifted.p, res[k].lifted);
                        memcpy(c->r_unm.as<char>() + ou, d.a_unm.p, res[k].unm);
                    } else {
                        res[k].rc = SWO_E_CAPACITY;  // redone below, single-threaded, after growing the result
                    }
                }
            });
        }
        for (auto &t : th) t.join();
        size_t tl = 0, tu = 0;
        bool redo = false;
        for (int k = 0; k < nd; k++) {
            if (res[k].rc == SWO_E_CAPACITY) redo = true;
            else if (res[k].rc) return fail(c, res[k].rc, "%s", emsg[k].c_str());
            tl += res[k].lifted;
            tu += res[k].unm;
        }
        c->seg_l.clear();
        c->seg_u.clear();
        for (int k = 0; k < nd; k++) {
            c->seg_l.emplace_back(k == 0 ? c->r_lifted.as<char>() : c->dev[k].a_lifted.as<char>(), res[k].lifted);
            c->seg_u.emplace_back(k == 0 ? c->r_unm.as<char>() : c->dev[k].a_unm.as<char>(), res[k].unm);
        }
        if (redo) {  // fan-out beyond the reserve: grow (keeps shard 0's bytes) and copy again
            CK(c, c->r_lifted.reserve(tl + 1, true, res[0].lifted));
            CK(c, c->r_unm.reserve(tu + 1, true, res[0].unm));
            c->seg_l[0].first = c->r_lifted.as<char>();
            c->seg_u[0].first = c->r_unm.as<char>();
            size_t ol = res[0].lifted, ou = res[0].unm;
            for (int k = 1; k < nd; k++) {
                memcpy(c->r_lifted.as<char>() + ol, c->dev[k].a_lifted.p, res[k].lifted);
                memcpy(c->r_unm.as<char>() + ou, c->dev[k].a_unm.p, res[k].unm);
                ol += res[k].lifted;
                ou += res[k].unm;
            }
        }
    }
    for (int k = 0; k < nd; k++) {
        if (res[k].has_err) {
            // The reference streams (bed.d:113-199): every record in front of the malformed line has been
            // written when it dies.  Same here: lift the input up to that line once more (it parses) and
            // hand those outputs back together with the error.
            const uint64_t off = res[k].err_off;
            if (off > 0 && off <= in_len) {
                const int rc2 = swo_lift_bed_text(c, in, (size_t)off, out);
                if (rc2 != SWO_OK) memset(out, 0, sizeof *out);
            }
            return fail(c, SWO_E_PARSE, "BED line at byte %llu has fewer than 3 fields or a non-integer coordinate",
                        (unsigned long long)off);
        }
        out->lifted_len += res[k].lifted;
        out->unmatched_len += res[k].unm;
        out->n_records += res[k].rec;
        out->n_rows += res[k].rows;
        out->n_unmatched += res[k].nunm;
    }
    out->lifted = c->r_lifted.as<char>();
    out->unmatched = c->r_unm.as<char>();
    if (nd == 1) {
        c->seg_l.assign(1, {out->lifted, out->lifted_len});
        c->seg_u.assign(1, {out->unmatched, out->unmatched_len});
    } else if (!c->concat) {
        out->lifted = out->unmatched = nullptr;  // the lengths are the totals; the bytes are in the segments
    }
    return SWO_OK;
}

int swo_text_num_segments(const swo_ctx *c) { return c ? (int)c->seg_l.size() : 0; }
int swo_text_segment(const swo_ctx *c, int which, int i, const char **ptr, size_t *len) {
    if (!c || !ptr || !len || i < 0 || i >= (int)c->seg_l.size() || (which != 0 && which != 1)) return SWO_E_ARG;
    const auto &sg = which == 0 ? c->seg_l[i] : c->seg_u[i];
    *ptr = sg.first;
    *len = sg.second;
    return SWO_OK;
}

// ---- VCF --------------------------------------------------------------------------------
int swo_load_genome_2bit(swo_ctx *c, const uint8_t *packed, size_t packed_bytes, const int64_t *contig_off,
                         const int64_t *contig_len, int n_contigs) {
    if (!c || !packed || !contig_off || !contig_len || n_contigs <= 0) return SWO_E_ARG;
    for (Device &d : c->dev) {
        CK(c, cudaSetDevice(d.ordinal));
        int rc;
        CK(c, d.g_packed.reserve(packed_bytes + 16));
        CK(c, cudaMemcpyAsync(d.g_packed.p, packed, packed_bytes, cudaMemcpyHostToDevice, d.s_comp));
        if ((rc = upload(c, d.g_off, contig_off, (size_t)n_contigs, d.s_comp))) return rc;
        if ((rc = upload(c, d.g_len, contig_len, (size_t)n_contigs, d.s_comp))) return rc;
        CK(c, cudaStreamSynchronize(d.s_comp));
        d.g_n = n_contigs;
        d.g_exc_n = 0;
        d.g_has_lower = false;
    }
    c->have_fasta = false;
    return SWO_OK;
}

int swo_load_genome_fasta(swo_ctx *c, const char *path) {
    if (!c || !path) return SWO_E_ARG;
    if (!c->tab.nblocks()) return fail(c, SWO_E_STATE, "load the chain before the genome (sequences are keyed by its destination contigs)");
    std::vector<int64_t> qs(c->tab.qnames.size(), 0);
    for (size_t i = 0; i < qs.size(); i++)
        for (const auto &kv : c->tab.qsizes)
            if (kv.first == c->tab.qnames[i]) qs[i] = kv.second;
    std::string err;
    if (!genome_from_fasta(path, c->tab.qnames, qs, c->genome, err)) return fail(c, SWO_E_IO, "%s", err.c_str());
    const GenomeStore &g = c->genome;
    for (Device &d : c->dev) {
        CK(c, cudaSetDevice(d.ordinal));
        int rc;
        if ((rc = upload(c, d.g_packed, g.packed.data(), g.packed.size(), d.s_comp))) return rc;
        if ((rc = upload(c, d.g_off, g.off.data(), g.off.size(), d.s_comp))) return rc;
        if ((rc = upload(c, d.g_len, g.len.data(), g.len.size(), d.s_comp))) return rc;
        d.g_has_lower = g.any_lower;
        if (g.any_lower && (rc = upload(c, d.g_lower, g.lower.data(), g.lower.size(), d.s_comp))) return rc;
        d.g_exc_n = (int)g.exc_start.size();
        if ((rc = upload(c, d.g_chunk, g.exc_chunk.data(), g.exc_chunk.size(), d.s_comp))) return rc;
        if ((rc = upload(c, d.g_estart, g.exc_start.data(), g.exc_start.size(), d.s_comp))) return rc;
        if ((rc = upload(c, d.g_elen, g.exc_len.data(), g.exc_len.size(), d.s_comp))) return rc;
        if ((rc = upload(c, d.g_ebyte, g.exc_byte.data(), g.exc_byte.size(), d.s_comp))) return rc;
        CK(c, cudaStreamSynchronize(d.s_comp));
        d.g_n = (int)g.off.size();
    }
    c->have_fasta = true;
    return SWO_OK;
}

int swo_genome_has_seq(const swo_ctx *c, const char *name) { return c && c->have_fasta && name && c->genome.has_seq(name); }
int64_t swo_genome_seq_len(const swo_ctx *c, const char *name) {
    return (c && c->have_fasta && name) ? c->genome.seq_len(name) : -1;
}

int swo_lift_vcf_points(swo_ctx *c, size_t n, const int32_t *tcid, const int32_t *pos, const int32_t *reflen,
                        const uint64_t *ref_off, const char *ref_bytes, size_t ref_bytes_len, int32_t *out_qcid,
                        int32_t *out_pos, uint8_t *flags, int32_t *out_reflen, char *out_ref) {
    if (!c || (n && (!tcid || !pos || !reflen || !ref_off || !out_qcid || !out_pos || !flags || !out_reflen)))
        return SWO_E_ARG;
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    Device &d = c->dev[0];
    if (!d.g_n) return fail(c, SWO_E_STATE, "no genome loaded");
    if (n >= 0xFFFFFFFFull) return fail(c, SWO_E_ARG, "n must be < 2^32-1 per call");
    if (!n) return SWO_OK;
    CK(c, cudaSetDevice(d.ordinal));
    cudaStream_t st = d.s_comp;
    const size_t na = (n * 4 + 15) & ~(size_t)15, n8 = (n * 8 + 15) & ~(size_t)15,
                 nr = (ref_bytes_len + 15) & ~(size_t)15;
    CK(c, d.b_in.reserve(na * 6 + n8 + 2 * nr + n + 64));
    char *base = d.b_in.as<char>();
    int32_t *d_tcid = (int32_t *)base, *d_pos = (int32_t *)(base + na), *d_rl = (int32_t *)(base + 2 * na),
            *d_oq = (int32_t *)(base + 3 * na), *d_op = (int32_t *)(base + 4 * na), *d_orl = (int32_t *)(base + 5 * na);
    uint64_t *d_ro = (uint64_t *)(base + 6 * na);
    char *d_ref = base + 6 * na + n8, *d_oref = d_ref + nr;
    uint8_t *d_fl = (uint8_t *)(d_oref + nr);
    // Chunks of records flow through three streams (H2D / kernel / D2H), so that both directions of the link
    // and the kernel overlap; with caller arrays in pinned memory (swo_host_alloc) the copies run at link
    // speed, with pageable arrays the driver stages them and the pipeline still hides the kernel.  The REF
    // bytes travel whole: in before the first chunk, out behind the last.
    (void)st;
    const size_t chunk = (size_t)8 << 20;
    const int nch = (int)((n + chunk - 1) / chunk);
    std::vector<cudaEvent_t> ev((size_t)nch * 2, nullptr);
    int rc = SWO_OK;
    auto ck = [&](cudaError_t e) {
        if (e != cudaSuccess && rc == SWO_OK) rc = fail(c, SWO_E_CUDA, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    for (auto &e : ev) ck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // (the staging buffer's previous user is done: every host-array entry point ends with a synchronise)
    if (rc == SWO_OK && ref_bytes_len) ck(cudaMemcpyAsync(d_ref, ref_bytes, ref_bytes_len, cudaMemcpyHostToDevice, d.s_h2d));
    for (int k = 0; k < nch && rc == SWO_OK; k++) {
        const size_t i0 = (size_t)k * chunk, m = std::min(chunk, n - i0);
        ck(cudaMemcpyAsync(d_tcid + i0, tcid + i0, m * 4, cudaMemcpyHostToDevice, d.s_h2d));
        ck(cudaMemcpyAsync(d_pos + i0, pos + i0, m * 4, cudaMemcpyHostToDevice, d.s_h2d));
        ck(cudaMemcpyAsync(d_rl + i0, reflen + i0, m * 4, cudaMemcpyHostToDevice, d.s_h2d));
        ck(cudaMemcpyAsync(d_ro + i0, ref_off + i0, m * 8, cudaMemcpyHostToDevice, d.s_h2d));
        ck(cudaEventRecord(ev[2 * k], d.s_h2d));
        ck(cudaStreamWaitEvent(d.s_comp, ev[2 * k], 0));
        VcfArgs A;
        A.T = d.T;
        A.G = genome_view(d);
        A.n = (uint32_t)m;
        A.tcid = d_tcid + i0;
        A.pos = d_pos + i0;
        A.reflen = d_rl + i0;
        A.ref_off = d_ro + i0;  // (offsets into the whole REF buffer)
        A.ref = d_ref;
        A.out_qcid = d_oq + i0;
        A.out_pos = d_op + i0;
        A.flags = d_fl + i0;
        A.out_reflen = d_orl + i0;
        A.out_ref = d_oref;
        const int grid = (int)std::min<size_t>((m + 255) / 256, (size_t)d.sms * 8);
        lift_vcf_points_kernel<<<grid, 256, 0, d.s_comp>>>(A);
        c->launches++;
        ck(cudaGetLastError());
        ck(cudaEventRecord(ev[2 * k + 1], d.s_comp));
        ck(cudaStreamWaitEvent(d.s_d2h, ev[2 * k + 1], 0));
        ck(cudaMemcpyAsync(out_qcid + i0, d_oq + i0, m * 4, cudaMemcpyDeviceToHost, d.s_d2h));
        ck(cudaMemcpyAsync(out_pos + i0, d_op + i0, m * 4, cudaMemcpyDeviceToHost, d.s_d2h));
        ck(cudaMemcpyAsync(out_reflen + i0, d_orl + i0, m * 4, cudaMemcpyDeviceToHost, d.s_d2h));
        ck(cudaMemcpyAsync(flags + i0, d_fl + i0, m, cudaMemcpyDeviceToHost, d.s_d2h));
    }
    if (rc == SWO_OK && out_ref && ref_bytes_len)  // behind the last kernel: s_d2h already waits for it
        ck(cudaMemcpyAsync(out_ref, d_oref, ref_bytes_len, cudaMemcpyDeviceToHost, d.s_d2h));
    ck(cudaStreamSynchronize(d.s_h2d));
    ck(cudaStreamSynchronize(d.s_comp));
    ck(cudaStreamSynchronize(d.s_d2h));
    for (auto &e : ev)
        if (e) cudaEventDestroy(e);
    return rc;
}

int swo_lift_vcf_points_dev(swo_ctx *c, size_t n, const int32_t *d_tcid, const int32_t *d_pos, const int32_t *d_rl,
                            const uint64_t *d_ro, const char *d_ref, int32_t *d_oq, int32_t *d_op, uint8_t *d_fl,
                            int32_t *d_orl, char *d_oref, void *stream) {
    if (!c || (n && (!d_tcid || !d_pos || !d_rl || !d_ro || !d_oq || !d_op || !d_fl || !d_orl))) return SWO_E_ARG;
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    Device &d = c->dev[0];
    if (!d.g_n) return fail(c, SWO_E_STATE, "no genome loaded");
    if (n >= 0xFFFFFFFFull) return fail(c, SWO_E_ARG, "n must be < 2^32-1 per call");
    if (!n) return SWO_OK;
    CK(c, cudaSetDevice(d.ordinal));
    VcfArgs A;
    A.T = d.T;
    A.G = genome_view(d);
    A.n = (uint32_t)n;
    A.tcid = d_tcid;
    A.pos = d_pos;
    A.reflen = d_rl;
    A.ref_off = d_ro;
    A.ref = d_ref;
    A.out_qcid = d_oq;
    A.out_pos = d_op;
    A.flags = d_fl;
    A.out_reflen = d_orl;
    A.out_ref = d_oref;
    const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)d.sms * 8);
    lift_vcf_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A);
    c->launches++;
    CK(c, cudaGetLastError());
    return SWO_OK;
}

// ---- liftVCF on text ---------------------------------------------------------------------
namespace {
// vcf.d:200-215
int contig_numeric(std::string_view z) {
    if (z.size() >= 3 && z.substr(0, 3) == "chr") z.remove_prefix(3);
    if (z.empty()) return 200;
    if (z[0] == 'X') return 100;
    if (z[0] == 'Y') return 101;
    if (z[0] == 'M') return 102;
    size_t k = 0;
    long v = 0;
    while (k < z.size() && z[k] >= '0' && z[k] <= '9') v = v * 10 + (z[k++] - '0');
    return k == 0 ? 200 : (int)v;
}
// vcf.d:181-196
bool contig_sort(const std::string &x, const std::string &y) {
    const bool ux = x.find('_') != std::string::npos, uy = y.find('_') != std::string::npos;
    if (ux != uy) return !ux;
    const int a = contig_numeric(x), b = contig_numeric(y);
    return a != b ? a < b : x < y;
}
void put_num(std::string &o, long long v) {
    char b[24];
    o.append(b, (size_t)snprintf(b, sizeof b, "%lld", v));
}
}  // namespace

int swo_lift_vcf_text(swo_ctx *c, const char *in, size_t in_len, const char *chainfile, const char *genomefile,
                      const char *filedate, swo_text_out *out) {
    if (!c || !out || (in_len && !in) || !chainfile || !genomefile) return SWO_E_ARG;
    memset(out, 0, sizeof *out);
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    if (!c->dev[0].g_n) return fail(c, SWO_E_STATE, "no genome loaded");
    char today[16];
    if (!filedate) {  // addFiledate (vcf.d:99)
        time_t t = time(nullptr);
        struct tm tmv;
        localtime_r(&t, &tmv);
        strftime(today, sizeof today, "%Y%m%d", &tmv);
        filedate = today;
    }
    std::string &L = c->vcf_l, &U = c->vcf_u;
    L.clear();
    U.clear();
    // pass 1: header, and the columns of every record
    struct Rec {
        size_t a, n;       // line
        uint32_t c[9];     // column starts, relative to a; c[8] = start of the pass-through tail or n+1
    };
    std::vector<Rec> recs;
    std::vector<int32_t> tcid, pos, rl;
    std::vector<uint64_t> ro;
    std::string refb;
    std::vector<std::string> hdr_contigs;
    bool header_done = false;
    for (size_t i = 0; i < in_len;) {
        const char *nlp = (const char *)memchr(in + i, '\n', in_len - i);
        const size_t e = nlp ? (size_t)(nlp - in) : in_len;
        size_t n = e - i;
        const char *ln = in + i;
        const size_t a = i;
        i = e + 1;
        if (n && ln[n - 1] == '\r') n--;
        if (!n) continue;
        if (ln[0] == '#') {
            if (n >= 2 && ln[1] == '#') {
                L.append(ln, n).push_back('\n');
                U.append(ln, n).push_back('\n');
                if (n > 13 && !memcmp(ln, "##contig=<ID=", 13)) {
                    size_t k = 13;
                    while (k < n && ln[k] != ',' && ln[k] != '>') k++;
                    hdr_contigs.emplace_back(ln + 13, k - 13);
                }
                continue;
            }
            if (!header_done) {  // vcf.d:62-103: additions go in front of the #CHROM line
                L += "##INFO=<ID=refchg,Number=0,Type=Flag,Description=\"REF allele changed in new genome\">\n";
                std::vector<std::string> keys;
                for (const auto &kv : c->tab.qsizes) keys.push_back(kv.first);
                std::stable_sort(keys.begin(), keys.end(), contig_sort);
                for (const std::string &k : keys) {
                    int64_t qsz = -1, glen = -1;
                    for (const auto &kv : c->tab.qsizes)
                        if (kv.first == k) qsz = kv.second;
                    if (c->have_fasta) {
                        glen = c->genome.seq_len(k);
                    } else {  // 2-bit genome given directly: lengths are per destination contig id
                        for (size_t q = 0; q < c->tab.qnames.size(); q++)
                            if (c->tab.qnames[q] == k && (int)q < c->dev[0].g_n) glen = qsz;
                    }
                    if (glen < 0 || glen != qsz) continue;  // vcf.d:79-87
                    if (std::find(hdr_contigs.begin(), hdr_contigs.end(), k) != hdr_contigs.end()) continue;
                    L += "##contig=<ID=" + k + ",length=";
                    put_num(L, qsz);
                    L += std::string(",source=") + chainfile + ">\n";
                }
                L += std::string("##fileDate=") + filedate + "\n##liftoverProgram=swiftover\n";
                L += "##liftoverProgramURL=https://github.com/blachlylab/swiftover\n";
                L += std::string("##liftoverChainfile=") + chainfile + "\n##liftoverGenome=" + genomefile + "\n";
                header_done = true;
            }
            L.append(ln, n).push_back('\n');
            U.append(ln, n).push_back('\n');
            continue;
        }
        Rec r;
        r.a = a;
        r.n = n;
        int nc = 0;
        size_t st0 = 0;
        for (size_t k = 0; k <= n && nc < 8; k++)
            if (k == n || ln[k] == '\t') {
                r.c[nc++] = (uint32_t)st0;
                st0 = k + 1;
            }
        if (nc < 8) return fail(c, SWO_E_PARSE, "VCF record with fewer than 8 columns at byte %zu", a);
        r.c[8] = (uint32_t)st0;  // one past the tab that ends INFO (n + 1 if INFO is the last column)
        long long p1 = 0;
        bool ok = r.c[2] - 1 > r.c[1];
        for (size_t k = r.c[1]; k + 1 < r.c[2]; k++) {
            if (ln[k] < '0' || ln[k] > '9') ok = false;
            else p1 = p1 * 10 + (ln[k] - '0');
            if (p1 > 2147483647ll) ok = false;
        }
        if (!ok) return fail(c, SWO_E_PARSE, "bad POS at byte %zu", a);
        recs.push_back(r);
        tcid.push_back(c->tab.tcontig_id(std::string_view(ln, r.c[1] - 1)));
        pos.push_back((int32_t)(p1 - 1));  // rec.pos is 0-based
        const size_t rlen = r.c[4] - 1 - r.c[3];
        rl.push_back((int32_t)rlen);
        ro.push_back(refb.size());
        refb.append(ln + r.c[3], rlen);
    }
    // the record body on the GPU (vcf.d:115-169)
    const size_t n = recs.size();
    std::vector<int32_t> oq(n), op(n), orl(n);
    std::vector<uint8_t> fl(n);
    std::string oref(refb.size(), '\0');
    if (n) {
        int rc = swo_lift_vcf_points(c, n, tcid.data(), pos.data(), rl.data(), ro.data(), refb.data(), refb.size(),
                                     oq.data(), op.data(), fl.data(), orl.data(), oref.data());
        if (rc) return rc;
    }
    // pass 2: join
    uint64_t nm = 0, nu = 0;
    for (size_t k = 0; k < n; k++) {
        const Rec &r = recs[k];
        const char *ln = in + r.a;
        if (!(fl[k] & 1)) {  // vcf.d:134-137
            U.append(ln, r.n).push_back('\n');
            nu++;
            continue;
        }
        nm++;
        L += c->tab.qnames[oq[k]];
        L.push_back('\t');
        put_num(L, (long long)op[k] + 1);
        L.push_back('\t');
        L.append(ln + r.c[2], r.c[3] - r.c[2]);  // ID + tab
        const bool chg = (fl[k] & 2) != 0;
        if (chg) L.append(oref.data() + ro[k], (size_t)orl[k]);  // vcf.d:156-157
        else L.append(ln + r.c[3], r.c[4] - 1 - r.c[3]);
        L.push_back('\t');
        L.append(ln + r.c[4], r.c[7] - r.c[4]);  // ALT QUAL FILTER + tab
        const size_t info_end = r.c[8] - 1;         // INFO = [c[7], info_end)
        if (chg) {                                   // vcf.d:161
            if (info_end - r.c[7] == 1 && ln[r.c[7]] == '.') L += "refchg";
            else L.append(ln + r.c[7], info_end - r.c[7]).append(";refchg");
        } else {
            L.append(ln + r.c[7], info_end - r.c[7]);
        }
        if (info_end < r.n) L.append(ln + info_end, r.n - info_end);  // FORMAT and samples, with their tab
        L.push_back('\n');
    }
    out->lifted = L.data();
    out->lifted_len = L.size();
    out->unmatched = U.data();
    out->unmatched_len = U.size();
    out->n_records = n;
    out->n_rows = nm;
    out->n_unmatched = nu;
    return SWO_OK;
}

// ---- liftMAF ------------------------------------------------------------------------------
int swo_lift_maf_records(swo_ctx *c, size_t n, const int32_t *tcid, const int32_t *start, const int32_t *end,
                         const int32_t *reflen, const uint64_t *ref_off, const char *ref_bytes, size_t ref_bytes_len,
                         const uint64_t *out_off, size_t out_ref_cap, uint8_t *nhits, int32_t *out_qcid,
                         int32_t *out_start, int32_t *out_end, uint8_t *flags, int32_t *out_reflen, char *out_ref) {
    if (!c || (n && (!tcid || !start || !end || !reflen || !ref_off || !out_off || !nhits || !out_qcid || !out_start ||
                     !out_end || !flags || !out_reflen)) ||
        (ref_bytes_len && !ref_bytes) || (out_ref_cap && !out_ref))
        return SWO_E_ARG;
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    Device &d = c->dev[0];
    if (!d.g_n) return fail(c, SWO_E_STATE, "no genome loaded");
    if (n >= 0xFFFFFFFFull) return fail(c, SWO_E_ARG, "n must be < 2^32-1 per call");
    if (!n) return SWO_OK;
    CK(c, cudaSetDevice(d.ordinal));
    cudaStream_t st = d.s_comp;
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t n4 = al(n * 4), n8 = al(n * 8), n1 = al(n);
    CK(c, d.b_in.reserve(8 * n4 + 2 * n8 + 2 * n1 + al(ref_bytes_len) + al(out_ref_cap) + 64));
    char *p = d.b_in.as<char>();
    auto take = [&p](size_t b) {
        char *r = p;
        p += b;
        return r;
    };
    int32_t *d_tcid = (int32_t *)take(n4), *d_s = (int32_t *)take(n4), *d_e = (int32_t *)take(n4),
            *d_rl = (int32_t *)take(n4), *d_oq = (int32_t *)take(n4), *d_os = (int32_t *)take(n4),
            *d_oe = (int32_t *)take(n4), *d_orl = (int32_t *)take(n4);
    uint64_t *d_ro = (uint64_t *)take(n8), *d_oo = (uint64_t *)take(n8);
    uint8_t *d_nh = (uint8_t *)take(n1), *d_fl = (uint8_t *)take(n1);
    char *d_ref = take(al(ref_bytes_len)), *d_oref = take(al(out_ref_cap));
    CK(c, cudaMemcpyAsync(d_tcid, tcid, n * 4, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(d_s, start, n * 4, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(d_e, end, n * 4, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(d_rl, reflen, n * 4, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(d_ro, ref_off, n * 8, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(d_oo, out_off, n * 8, cudaMemcpyHostToDevice, st));
    if (ref_bytes_len) CK(c, cudaMemcpyAsync(d_ref, ref_bytes, ref_bytes_len, cudaMemcpyHostToDevice, st));
    MafArgs A;
    A.T = d.T;
    A.G = genome_view(d);
    A.n = (uint32_t)n;
    A.tcid = d_tcid;
    A.start = d_s;
    A.end = d_e;
    A.reflen = d_rl;
    A.ref_off = d_ro;
    A.ref = d_ref;
    A.out_off = d_oo;
    A.nhits = d_nh;
    A.out_qcid = d_oq;
    A.out_start = d_os;
    A.out_end = d_oe;
    A.flags = d_fl;
    A.out_reflen = d_orl;
    A.out_ref = d_oref;
    const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)d.sms * 8);
    lift_maf_records_kernel<<<grid, 256, 0, st>>>(A);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(nhits, d_nh, n, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(flags, d_fl, n, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(out_qcid, d_oq, n * 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(out_start, d_os, n * 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(out_end, d_oe, n * 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(out_reflen, d_orl, n * 4, cudaMemcpyDeviceToHost, st));
    if (out_ref_cap) CK(c, cudaMemcpyAsync(out_ref, d_oref, out_ref_cap, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return SWO_OK;
}

namespace {
// std.conv.to!long (maf.d:180-181): optional sign, digits only
bool token_to_i64(const char *s, size_t n, int64_t &out) {
    size_t i = 0;
    bool neg = false;
    if (n && (s[0] == '-' || s[0] == '+')) {
        neg = s[0] == '-';
        i = 1;
    }
    if (i == n) return false;
    uint64_t v = 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        const uint64_t dg = (uint64_t)(s[i] - '0');
        if (v > (UINT64_MAX - dg) / 10) return false;
        v = v * 10 + dg;
    }
    if (neg ? v > (uint64_t)INT64_MAX + 1 : v > (uint64_t)INT64_MAX) return false;
    out = neg ? (int64_t)(0 - v) : (int64_t)v;
    return true;
}
// NT_COMP_TABLE (maf.d:21-32): bytes outside ACGTacgt and '-' become NUL
char nt_comp(unsigned char ch) {
    switch (ch) {
    case '-': return '-';
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case 'a': return 't';
    case 'c': return 'g';
    case 'g': return 'c';
    case 't': return 'a';
    default: return 0;
    }
}
// reverse_complement (maf.d:36-54)
void append_revcomp(std::string &o, const char *a, size_t n) {
    if (n == 2 && a[0] == 'N' && a[1] == 'A') {
        o.append("NA");
        return;
    }
    for (size_t k = n; k-- > 0;) o.push_back(nt_comp((unsigned char)a[k]));
}
bool maf_space(char ch) { return ch == ' ' || (ch >= 9 && ch <= 13); }
}  // namespace

int swo_lift_maf_text(swo_ctx *c, const char *in, size_t in_len, const char *genomebuild, swo_text_out *out,
                      swo_maf_stats *stats) {
    enum { M_BUILD = 3, M_CONTIG = 4, M_START = 5, M_END = 6, M_VTYPE = 9, M_REF = 10, M_T1 = 11, M_T2 = 12 };  // maf.d:57-72
    if (!c || !out || (in_len && !in) || !genomebuild) return SWO_E_ARG;
    memset(out, 0, sizeof *out);
    if (stats) memset(stats, 0, sizeof *stats);
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    if (!c->dev[0].g_n) return fail(c, SWO_E_STATE, "no genome loaded");
    std::string &L = c->vcf_l, &U = c->vcf_u;
    L.clear();
    U.clear();
    size_t i = 0;
    for (int h = 0; h < 2 && i < in_len; h++) {  // the two header lines, terminators kept (maf.d:166-171)
        const char *nlp = (const char *)memchr(in + i, '\n', in_len - i);
        const size_t e = nlp ? (size_t)(nlp - in) + 1 : in_len;
        L.append(in + i, e - i);
        U.append(in + i, e - i);
        i = e;
    }
    int64_t longest = 0;  // no single hit is longer than the longest block
    for (const Block &b : c->tab.blocks) longest = std::max<int64_t>(longest, (int64_t)b.tEnd - b.tStart);
    // pass 1: fields of every record (splitter(): runs of white space, maf.d:175)
    struct Rec {
        size_t line, nf, f;  // byte offset of the line, field count, first entry in fld
    };
    std::vector<Rec> recs;
    std::vector<std::pair<uint32_t, uint32_t>> fld;  // [begin,end) relative to the line
    std::vector<int32_t> tcid, st0, en, rl;
    std::vector<uint64_t> ro, oo;
    std::string refb;
    uint64_t ocap = 0;
    while (i < in_len) {
        const char *nlp = (const char *)memchr(in + i, '\n', in_len - i);
        const size_t e = nlp ? (size_t)(nlp - in) : in_len;
        const char *ln = in + i;
        const size_t n = e - i, at = i;
        i = e + 1;
        if (n > 0xFFFFFFF0u) return fail(c, SWO_E_PARSE, "MAF line longer than 4 GiB at byte %zu", at);
        Rec r{at, 0, fld.size()};
        for (size_t k = 0; k < n;) {
            while (k < n && maf_space(ln[k])) k++;
            if (k == n) break;
            const size_t a = k;
            while (k < n && !maf_space(ln[k])) k++;
            fld.emplace_back((uint32_t)a, (uint32_t)k);
            r.nf++;
        }
        if (r.nf < 13) return fail(c, SWO_E_PARSE, "MAF record with fewer than 13 fields at byte %zu", at);
        const auto *F = &fld[r.f];
        int64_t s64, e64;
        if (!token_to_i64(ln + F[M_START].first, F[M_START].second - F[M_START].first, s64) ||
            !token_to_i64(ln + F[M_END].first, F[M_END].second - F[M_END].first, e64))
            return fail(c, SWO_E_PARSE, "bad start/end at byte %zu", at);
        if (s64 > e64) return fail(c, SWO_E_PARSE, "start > end at byte %zu", at);  // assert, maf.d:183
        // to 32 bits: every block lies in [0, 2^31), so clamping the query keeps the hit set and the trimmed hit
        const int64_t s0 = s64 == INT64_MIN ? INT64_MIN : s64 - 1;
        const int32_t s32 = (int32_t)std::min<int64_t>(std::max<int64_t>(s0, INT32_MIN), INT32_MAX);
        const int32_t e32 = (int32_t)std::min<int64_t>(std::max<int64_t>(e64, INT32_MIN), INT32_MAX);
        recs.push_back(r);
        tcid.push_back(c->tab.tcontig_id(std::string_view(ln + F[M_CONTIG].first, F[M_CONTIG].second - F[M_CONTIG].first)));
        st0.push_back(s32);
        en.push_back(e32);
        const size_t rlen = F[M_REF].second - F[M_REF].first;
        const bool dash = rlen == 1 && ln[F[M_REF].first] == '-';  // maf.d:263
        if (rlen > 0x7FFFFFFFu) return fail(c, SWO_E_PARSE, "Reference_Allele longer than 2 GiB at byte %zu", at);
        rl.push_back(dash ? -1 : (int32_t)rlen);
        ro.push_back(refb.size());
        if (!dash) refb.append(ln + F[M_REF].first, rlen);
        oo.push_back(ocap);
        if (!dash) ocap += (uint64_t)std::min<int64_t>(std::max<int64_t>((int64_t)e32 - s32, 0), longest);
    }
    // the record body on the GPU (maf.d:189-291)
    const size_t n = recs.size();
    std::vector<uint8_t> nh(n), fl(n);
    std::vector<int32_t> oq(n), os(n), oe(n), orl(n);
    std::string oref(ocap, '\0');
    if (n) {
        int rc = swo_lift_maf_records(c, n, tcid.data(), st0.data(), en.data(), rl.data(), ro.data(), refb.data(),
                                      refb.size(), oo.data(), oref.size(), nh.data(), oq.data(), os.data(), oe.data(),
                                      fl.data(), orl.data(), oref.data());
        if (rc) return rc;
    }
    // pass 2: join
    swo_maf_stats S;
    memset(&S, 0, sizeof S);
    S.n_records = n;
    std::string t1, t2;
    for (size_t k = 0; k < n; k++) {
        const Rec &r = recs[k];
        const char *ln = in + r.line;
        const auto *F = &fld[r.f];
        if (nh[k] == 0) {  // maf.d:192-195
            S.n_unmatched++;
            for (size_t j = 0; j < r.nf; j++) {
                if (j) U.push_back('\t');
                U.append(ln + F[j].first, F[j].second - F[j].first);
            }
            U.push_back('\n');
            continue;
        }
        S.n_matched++;
        if (nh[k] > 1) {  // maf.d:298-307
            S.n_multiple++;
            continue;
        }
        const bool neg = (fl[k] & 4) != 0, chg = (fl[k] & 2) != 0;
        const char *a1 = ln + F[M_T1].first, *a2 = ln + F[M_T2].first;
        size_t l1 = F[M_T1].second - F[M_T1].first, l2 = F[M_T2].second - F[M_T2].first;
        if (neg) {  // maf.d:220-224
            t1.clear();
            t2.clear();
            append_revcomp(t1, a1, l1);
            append_revcomp(t2, a2, l2);
            a1 = t1.data();
            a2 = t2.data();
        }
        const char *nr = oref.data() + oo[k];
        const size_t nrl = (size_t)orl[k];
        if (chg) {
            S.n_refchg++;
            const bool del = F[M_VTYPE].second - F[M_VTYPE].first == 3 && !memcmp(ln + F[M_VTYPE].first, "DEL", 3);
            const bool eq1 = nrl == l1 && !memcmp(nr, a1, l1), eq2 = nrl == l2 && !memcmp(nr, a2, l2);
            if (!del && !eq1 && !eq2) S.n_neither++;  // maf.d:273-277
        }
        for (size_t j = 0; j < r.nf; j++) {  // maf.d:296
            if (j) L.push_back('\t');
            if (j == M_BUILD) L += genomebuild;
            else if (j == M_CONTIG) L += c->tab.qnames[oq[k]];
            else if (j == M_START) put_num(L, os[k]);
            else if (j == M_END) put_num(L, oe[k]);
            else if (j == M_REF && chg) L.append(nr, nrl);
            else if (j == M_T1) L.append(a1, l1);
            else if (j == M_T2) L.append(a2, l2);
            else L.append(ln + F[j].first, F[j].second - F[j].first);
        }
        L.push_back('\n');
    }
    out->lifted = L.data();
    out->lifted_len = L.size();
    out->unmatched = U.data();
    out->unmatched_len = U.size();
    out->n_records = n;
    out->n_rows = S.n_matched - S.n_multiple;
    out->n_unmatched = S.n_unmatched;
    if (stats) *stats = S;
    return SWO_OK;
}

}  // extern "C"

int swo_internal_tname_table(swo_ctx *c, const uint32_t **d_off, const char **d_chars) {
    if (!c) return SWO_E_ARG;
    if (c->dev.empty()) return fail(c, SWO_E_CUDA, "inspection-only context: no CUDA device, and there is no CPU fallback");
    if (!c->have_chain) return fail(c, SWO_E_STATE, "no chain loaded");
    *d_off = c->dev[0].tnoff.as<uint32_t>();
    *d_chars = c->dev[0].tchars.as<char>();
    return SWO_OK;
}
