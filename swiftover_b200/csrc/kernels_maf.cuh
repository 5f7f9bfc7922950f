// kernels_maf.cuh — liftMAF record body on device.
//
// Replaces, per record (/root/reference/source/swiftover/maf.d:172-310):
//   cf.lift(contig, start-1, end)                maf.d:189 -> chain.d:664-713
//   0 hits -> unmatched, >1 hits -> dropped      maf.d:192-195, 298-307
//   1 hit: orderStartEnd, start+1                maf.d:208-214
//          fa.fetchSequence!obc(contig,start,end) maf.d:263-264 (gather from the packed genome)
//          REF != fetched -> replace REF          maf.d:265-291
// The strand of the hit is reported so that the host can reverse-complement the tumor
// alleles (maf.d:220-224); the text split/join stays on the host like liftVCF's.
#pragma once
#include "kernels_vcf.cuh"

namespace swo {

struct MafArgs {
    TableView T;
    GenomeView G;
    uint32_t n;
    const int32_t *tcid, *start, *end;  // zero-based half-open [start,end) on the source contig
    const int32_t *reflen;              // length of Ref_Allele, or -1 for "-" (no check, maf.d:263)
    const uint64_t *ref_off;            // [n] into ref
    const char *ref;
    const uint64_t *out_off;            // [n] into out_ref; record i owns min(end-start, longest block) bytes
    uint8_t *nhits;                     // 0 | 1 | 2 (= more than one)
    int32_t *out_qcid, *out_start, *out_end;  // one-based closed, destination
    uint8_t *flags;                     // bit0 single hit, bit1 REF changed, bit2 '-' block
    int32_t *out_reflen;
    char *out_ref;
};

__global__ void __launch_bounds__(256) lift_maf_records_kernel(const MafArgs A) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
        const int s = A.start[i], e = A.end[i];
        Block b0;
        const Cand cd = find_candidates_b0(A.T, A.T.ey, A.tcid[i], s, e, b0);
        int nh = 0, first = -1;
        if (A.T.disjoint) {
            nh = cd.hi - cd.lo;
            first = cd.lo;
        } else {
            for (int j = cd.lo; j < cd.hi && nh < 2; j++)
                if (__ldg(&A.T.blk[j].tEnd) > s) {
                    if (!nh) first = j;
                    nh++;
                }
        }
        A.nhits[i] = (uint8_t)(nh > 2 ? 2 : nh);
        if (nh != 1) {
            A.flags[i] = 0;
            A.out_qcid[i] = -1;
            A.out_start[i] = 0;
            A.out_end[i] = 0;
            A.out_reflen[i] = 0;
            continue;
        }
        const Block b = first == cd.lo ? b0 : load_block(A.T, first);
        const RowCore r = make_row(b, s, e);
        uint8_t fl = 1 | (r.inv < 0 ? 4 : 0);
        const int rl = A.reflen[i];
        int m = 0;
        if (rl >= 0) {  // fetch [r.s, r.e), clipped at the end of the sequence like faidx
            const int q = r.qcid;
            const int64_t L = q < A.G.g_n ? A.G.g_len[q] : 0;
            if (r.s >= 0 && (int64_t)r.s < L) m = (int64_t)r.e > L ? (int)(L - r.s) : r.e - r.s;
            bool diff = m != rl;
            const int64_t g0 = q < A.G.g_n ? A.G.g_off[q] + r.s : 0;
            const uint64_t ro = A.ref_off[i], oo = A.out_off[i];
            for (int k = 0; k < m; k++) {
                const char base = genome_base(A.G, g0 + k);
                A.out_ref[oo + k] = base;
                if (k < rl) diff |= base != A.ref[ro + k];  // maf.d:265 byte-exact compare
            }
            if (diff) fl |= 2;
        }
        A.out_qcid[i] = r.qcid;
        A.out_start[i] = r.s + 1;  // maf.d:212
        A.out_end[i] = r.e;
        A.out_reflen[i] = m;
        A.flags[i] = fl;
    }
}

}  // namespace swo
