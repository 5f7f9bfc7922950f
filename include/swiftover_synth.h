/*
 * swiftover_synth.h — workload-generation helpers exported by
 * libswiftover_cuda.so for bench.py and the GPU tests (NOT part of the
 * reference-facing interface).  They turn device-resident binary records into
 * BED text on the device, so the synthetic inputs of SURVEY.md §8d (1e8-1e9
 * rows, 2.5-60 GB of text) never have to be produced on the host.
 */
#ifndef SWIFTOVER_SYNTH_H
#define SWIFTOVER_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#include "swiftover_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ncols = 3: "tname\tstart\tend\n"
 * ncols = 6: "...\tr<i>\t0\t<strand>\n"           (i = first_index + row)
 * ncols = 8: "...\tr<i>\t0\t<strand>\t<thickStart>\t<thickEnd>\n"
 * Pass 1 writes the byte length of every row to d_len[n]. */
int swo_synth_bed_lengths_dev(swo_ctx *ctx, size_t n, int ncols, uint64_t first_index, const int32_t *d_tcid,
                              const int32_t *d_start, const int32_t *d_end, const uint8_t *d_strand,
                              const int32_t *d_thick_start, const int32_t *d_thick_end, uint32_t *d_len,
                              void *stream);
/* Pass 2 writes row i at d_out + d_off[i] (d_off = exclusive prefix of d_len). */
int swo_synth_bed_write_dev(swo_ctx *ctx, size_t n, int ncols, uint64_t first_index, const int32_t *d_tcid,
                            const int32_t *d_start, const int32_t *d_end, const uint8_t *d_strand,
                            const int32_t *d_thick_start, const int32_t *d_thick_end, const uint64_t *d_off,
                            char *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
