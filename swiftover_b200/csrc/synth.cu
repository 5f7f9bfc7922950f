// synth.cu — device-side BED text writer for synthetic workloads (bench/test tooling,
// see include/swiftover_synth.h).  Not on the product path.
#include <cuda_runtime.h>

#include "swiftover_synth.h"
#include "swo_internal.hpp"

namespace {

__device__ __forceinline__ int ndig_u64(unsigned long long u) {
    int n = 1;
    while (u >= 10ull) {
        u /= 10ull;
        n++;
    }
    return n;
}
__device__ __forceinline__ int ndig_i32(int v) {
    return ndig_u64(v < 0 ? (unsigned long long)(-(long long)v) : (unsigned long long)v) + (v < 0);
}
__device__ __forceinline__ char *put_u64(char *d, unsigned long long u) {
    const int n = ndig_u64(u);
    char *p = d + n;
    do {
        *--p = (char)('0' + (int)(u % 10ull));
        u /= 10ull;
    } while (u);
    return d + n;
}
__device__ __forceinline__ char *put_int(char *d, int v) {
    if (v < 0) {
        *d++ = '-';
        return put_u64(d, (unsigned long long)(-(long long)v));
    }
    return put_u64(d, (unsigned long long)v);
}

struct SynthArgs {
    size_t n;
    int ncols;
    uint64_t first_index;
    const int32_t *tcid, *start, *end;
    const uint8_t *strand;
    const int32_t *ths, *the;
    const uint32_t *tname_off;
    const char *tname_chars;
};

__global__ void synth_len_kernel(SynthArgs A, uint32_t *len) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < A.n; i += (size_t)gridDim.x * blockDim.x) {
        const int t = A.tcid[i];
        uint32_t n = A.tname_off[t + 1] - A.tname_off[t] + 1 + ndig_i32(A.start[i]) + 1 + ndig_i32(A.end[i]) + 1;
        if (A.ncols >= 6) n += 6 + ndig_u64(A.first_index + i);  // "\tr<i>\t0\t<strand>"
        if (A.ncols >= 8) n += 1 + ndig_i32(A.ths[i]) + 1 + ndig_i32(A.the[i]);
        len[i] = n;
    }
}

__global__ void synth_write_kernel(SynthArgs A, const uint64_t *off, char *out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < A.n; i += (size_t)gridDim.x * blockDim.x) {
        char *d = out + off[i];
        const int t = A.tcid[i];
        for (uint32_t k = A.tname_off[t]; k < A.tname_off[t + 1]; k++) *d++ = A.tname_chars[k];
        *d++ = '\t';
        d = put_int(d, A.start[i]);
        *d++ = '\t';
        d = put_int(d, A.end[i]);
        if (A.ncols >= 6) {
            *d++ = '\t';
            *d++ = 'r';
            d = put_u64(d, A.first_index + i);
            *d++ = '\t';
            *d++ = '0';
            *d++ = '\t';
            *d++ = (char)A.strand[i];
        }
        if (A.ncols >= 8) {
            *d++ = '\t';
            d = put_int(d, A.ths[i]);
            *d++ = '\t';
            d = put_int(d, A.the[i]);
        }
        *d++ = '\n';
    }
}

int fill(swo_ctx *c, SynthArgs &A, size_t n, int ncols, uint64_t first_index, const int32_t *tcid,
         const int32_t *start, const int32_t *end, const uint8_t *strand, const int32_t *ths, const int32_t *the) {
    if (!c || !tcid || !start || !end) return SWO_E_ARG;
    if (ncols != 3 && ncols != 6 && ncols != 8) return SWO_E_ARG;
    if (ncols >= 6 && !strand) return SWO_E_ARG;
    if (ncols >= 8 && (!ths || !the)) return SWO_E_ARG;
    const uint32_t *toff;
    const char *tchars;
    int rc = swo_internal_tname_table(c, &toff, &tchars);
    if (rc) return rc;
    A = SynthArgs{n, ncols, first_index, tcid, start, end, strand, ths, the, toff, tchars};
    return SWO_OK;
}

}  // namespace

extern "C" {

int swo_synth_bed_lengths_dev(swo_ctx *c, size_t n, int ncols, uint64_t first_index, const int32_t *tcid,
                              const int32_t *start, const int32_t *end, const uint8_t *strand, const int32_t *ths,
                              const int32_t *the, uint32_t *d_len, void *stream) {
    SynthArgs A;
    int rc = fill(c, A, n, ncols, first_index, tcid, start, end, strand, ths, the);
    if (rc || !n) return rc;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    synth_len_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, d_len);
    return cudaGetLastError() == cudaSuccess ? SWO_OK : SWO_E_CUDA;
}

int swo_synth_bed_write_dev(swo_ctx *c, size_t n, int ncols, uint64_t first_index, const int32_t *tcid,
                            const int32_t *start, const int32_t *end, const uint8_t *strand, const int32_t *ths,
                            const int32_t *the, const uint64_t *d_off, char *d_out, void *stream) {
    SynthArgs A;
    int rc = fill(c, A, n, ncols, first_index, tcid, start, end, strand, ths, the);
    if (rc || !n) return rc;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    synth_write_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, d_off, d_out);
    return cudaGetLastError() == cudaSuccess ? SWO_OK : SWO_E_CUDA;
}
}
