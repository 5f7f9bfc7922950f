// kernels_vcf.cuh — liftVCF record body on device.
//
// Replaces, per record (/root/reference/source/swiftover/vcf.d:115-169):
//   liftDirectly(chrom, pos)                   vcf.d:125 -> chain.d:604-627
//   fa.fetchSequence(chrom, pos, pos+refLen)   vcf.d:143  (here: gather from a
//                                              2-bit packed genome held in HBM)
//   REF != fetched -> replace REF, set refchg  vcf.d:144-161
// The reference does no strand handling: on '-' blocks the
// position is qStart - delta and alleles are not reverse-complemented; both are
// reproduced.
#pragma once
#include "device_util.cuh"
#include "lift_core.cuh"

namespace swo {

// Destination genome in HBM (genome.hpp): what IndexedFastaFile.fetchSequence reads (vcf.d:143, maf.d:264)
struct GenomeView {
    const uint8_t *packed;     // 2 bit/base, A=0 C=1 G=2 T=3
    const uint8_t *lower;      // 1 bit/base: soft-masked (lower case); null if the genome has none
    const uint8_t *exc_chunk;  // 1 bit per 4096 bases: chunk holds a non-ACGT symbol; null if none
    const int64_t *exc_start;  // runs of non-ACGT symbols, sorted by start
    const int32_t *exc_len;
    const uint8_t *exc_byte;
    int exc_n;
    const int64_t *g_off;  // [g_n] first base of each destination contig
    const int64_t *g_len;  // [g_n]
    int g_n;
};

struct VcfArgs {
    TableView T;
    GenomeView G;
    uint32_t n;
    const int32_t *tcid, *pos, *reflen;
    const uint64_t *ref_off;
    const char *ref;
    int32_t *out_qcid, *out_pos;
    uint8_t *flags;
    int32_t *out_reflen;
    char *out_ref;
};

// base g of the packed destination genome, exactly as the FASTA spells it (genome.hpp)
__device__ __forceinline__ char genome_base(const GenomeView &A, int64_t g) {
    char base = "ACGT"[(__ldg(A.packed + (g >> 2)) >> ((g & 3) * 2)) & 3];
    if (A.lower && ((__ldg(A.lower + (g >> 3)) >> (g & 7)) & 1)) base |= 0x20;
    if (A.exc_n && ((__ldg(A.exc_chunk + (g >> 15)) >> ((g >> 12) & 7)) & 1)) {
        int lo = 0, hi = A.exc_n;  // last run starting at or before g
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(A.exc_start + mid) <= g) lo = mid + 1;
            else hi = mid;
        }
        if (lo > 0 && g < __ldg(A.exc_start + lo - 1) + __ldg(A.exc_len + lo - 1)) base = (char)__ldg(A.exc_byte + lo - 1);
    }
    return base;
}

// liftDirectly for one position with the upper search levels and the per-contig table in shared
// memory (point_hit of lift_core.cuh walks the global copies: ~10 dependent L2 round trips per record,
// which made this kernel L2-throughput bound).  Returns the block index or -1.
__device__ __forceinline__ int point_hit_smem(const TableView &T, const int2 *ey, const int4 *cm_s, int tcid, int c) {
    if (tcid < 0 || tcid >= T.ntc) return -1;
    const int4 cm = cm_s ? cm_s[tcid] : load_cmeta(T, tcid);
    int j = lower_gt_ey(T, ey, cm, c);
    for (; j < cm.y; j++) {
        if (__ldg(T.tStartKey + j) > c) return -1;
        if (T.disjoint || __ldg(&T.blk[j].tEnd) > c) return j;
    }
    return -1;
}

constexpr int VCF_CM_CAP = 128;

__global__ void __launch_bounds__(256) lift_vcf_points_kernel(const __grid_constant__ VcfArgs A) {
    __shared__ int2 s_ey[ChainTable::EY_CAP];
    __shared__ int4 s_cm[VCF_CM_CAP];
    for (int i = threadIdx.x; i < A.T.ey_n; i += 256) s_ey[i] = __ldg(A.T.ey + i);
    const bool cm_smem = A.T.ntc <= VCF_CM_CAP;
    if (cm_smem)
        for (int i = threadIdx.x; i < A.T.ntc; i += 256) s_cm[i] = __ldg(A.T.cmeta + i);
    __syncthreads();
    const bool plain = !A.G.lower && !A.G.exc_n;  // an ACGT-only, upper-case genome: 16 bases per 32-bit word
    const uint32_t *const gw = reinterpret_cast<const uint32_t *>(A.G.packed);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
        const int c0 = A.pos[i];
        const int rl = A.reflen[i];
        const uint64_t ro = A.ref_off[i];
        const int j = point_hit_smem(A.T, s_ey, cm_smem ? s_cm : nullptr, A.tcid[i], c0);
        if (j < 0) {  // vcf.d:134-137: record goes to the unmatched file unchanged
            A.out_qcid[i] = -1;
            A.out_pos[i] = c0;
            A.flags[i] = 0;
            A.out_reflen[i] = rl;
            for (int k = 0; k < rl; k++) A.out_ref[ro + k] = A.ref[ro + k];
            continue;
        }
        const Block b = load_block(A.T, j);
        const int q = b.meta >> 1;
        const int c = point_coord(b, c0);
        const int64_t L = q < A.G.g_n ? A.G.g_len[q] : 0;
        int m = 0;  // faidx clips the fetch at the contig end
        if (c >= 0 && (int64_t)c < L) m = (int64_t)c + rl > L ? (int)(L - c) : rl;
        bool diff = m != rl;
        const int64_t g0 = q < A.G.g_n ? A.G.g_off[q] + c : 0;
        if (plain) {
            for (int k0 = 0; k0 < m; k0 += 16) {  // one or two aligned words cover 16 bases from g0 + k0
                const int64_t g = g0 + k0;
                const uint32_t w0 = __ldg(gw + (g >> 4)), w1 = (g & 15) ? __ldg(gw + (g >> 4) + 1) : 0u;
                uint32_t bits = __funnelshift_r(w0, w1, (unsigned)(g & 15) * 2u);
                const int cnt = m - k0 < 16 ? m - k0 : 16;
                for (int k = 0; k < cnt; k++, bits >>= 2) {
                    const char base = (char)(0x54474341u >> ((bits & 3u) * 8u));  // "ACGT"
                    A.out_ref[ro + k0 + k] = base;
                    diff |= base != A.ref[ro + k0 + k];  // vcf.d:144 byte-exact compare
                }
            }
        } else {
            for (int k = 0; k < m; k++) {
                const char base = genome_base(A.G, g0 + k);
                A.out_ref[ro + k] = base;
                diff |= base != A.ref[ro + k];  // vcf.d:144 byte-exact compare
            }
        }
        A.out_qcid[i] = q;
        A.out_pos[i] = c;
        A.out_reflen[i] = m;
        A.flags[i] = diff ? 3 : 1;
    }
}

}  // namespace swo
