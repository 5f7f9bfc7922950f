/*
 * swiftover_cuda.h — C ABI of libswiftover_cuda.so, the B200 (sm_100a) engine
 * for swiftover's liftover hot path.
 *
 * The seam it replaces is `struct ChainFile` of the reference
 * (/root/reference/source/swiftover/chain.d:485-714) together with the
 * per-record loop bodies that call it (bed.d:113-202, vcf.d:115-169).  The
 * reference interface is one record per call; this one is batched: the host
 * (D via extern(C), or the bundled C++ host) hands over whole buffers.
 *
 * Conventions
 *   - every function returns SWO_OK (0) or a negative swo_status; the message
 *     is available from swo_last_error().  Nothing throws, nothing exits.
 *   - inputs are borrowed for the duration of the call only.
 *   - host outputs live in library-owned PINNED host memory that stays valid
 *     until the next lift call on the same ctx (or swo_destroy).
 *   - a ctx is used by one host thread at a time; distinct ctxs are independent.
 *   - coordinates are 32-bit (the reference parses them with to!int,
 *     chain.d:347, bed.d:125-126).
 *   - there is NO CPU fallback: without a CUDA device swo_create fails.
 */
#ifndef SWIFTOVER_CUDA_H
#define SWIFTOVER_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum swo_status {
    SWO_OK = 0,
    SWO_E_IO = -1,        /* FileException (chain.d:513-514, main.d:91-101)        */
    SWO_E_PARSE = -2,     /* ConvException / malformed chain or record              */
    SWO_E_CUDA = -3,      /* CUDA runtime error or no device                        */
    SWO_E_OOM = -4,       /* host or device allocation failed                       */
    SWO_E_ARG = -5,       /* bad argument                                           */
    SWO_E_STATE = -6,     /* e.g. lift before swo_load_chain                        */
    SWO_E_CAPACITY = -7   /* caller-provided device output buffer too small         */
} swo_status;

typedef struct swo_ctx swo_ctx;

/* ---- lifetime --------------------------------------------------------- */

/* devices: CUDA ordinals the ctx may use (NULL/0 -> device 0).  The chain table
 * and genome are replicated on each; text inputs are sharded by byte range
 * over them.  The same ordinal may be listed several times
 * ("logical shards") — used by the tests to exercise the sharder on one GPU. */
int swo_create(swo_ctx **out, const int *devices, int ndev);
/* A context WITHOUT a device: swo_load_chain* and the table accessors work (host
 * parser only); every lift call fails with SWO_E_CUDA.  Used by CPU-side tests
 * of the chain parser. */
int swo_create_inspect(swo_ctx **out);
void swo_destroy(swo_ctx *ctx);
const char *swo_last_error(const swo_ctx *ctx);
/* library build/ABI version, e.g. "swiftover_cuda 0.1 sm_100a" */
const char *swo_version(void);

/* ---- ChainFile.this(fn)  (chain.d:509-597) ----------------------------- */

int swo_load_chain(swo_ctx *ctx, const char *path);
int swo_load_chain_mem(swo_ctx *ctx, const char *text, size_t len);

/* keys of chainsByContig (chain.d:495) = source-build ("target") contigs */
int swo_num_tcontigs(const swo_ctx *ctx);
const char *swo_tcontig_name(const swo_ctx *ctx, int tcid);
/* -1 if the chain has no such contig */
int swo_tcontig_id(const swo_ctx *ctx, const char *name, size_t len);
int64_t swo_tcontig_nblocks(const swo_ctx *ctx, int tcid);
/* tSize of the chain headers naming this contig (chain.d:298; last header wins) */
int64_t swo_tcontig_size(const swo_ctx *ctx, int tcid);
int64_t swo_num_blocks(const swo_ctx *ctx);
/* 1 if no two blocks of a target contig overlap (true for UCSC over.chain) */
int swo_table_is_disjoint(const swo_ctx *ctx);

/* ChainFile.contigNames (chain.d:500): destination ("query") contig id -> name */
int swo_num_qcontigs(const swo_ctx *ctx);
const char *swo_qcontig_name(const swo_ctx *ctx, int qcid);
/* ChainFile.qContigSizes (chain.d:504,551), in first-insertion order */
int swo_num_qsizes(const swo_ctx *ctx);
const char *swo_qsize_name(const swo_ctx *ctx, int i);
int64_t swo_qsize_value(const swo_ctx *ctx, int i);

/* Copy `n` blocks of target contig `tcid`, starting at `first`, in index order
 * (tStart,tEnd,qStart) — the flattened form of the reference's per-contig tree
 * (chain.d:537-548).  Any output pointer may be NULL. */
int swo_get_blocks(const swo_ctx *ctx, int tcid, int64_t first, int64_t n, int32_t *t_start,
                   int32_t *t_end, int32_t *q_start, int32_t *qcid, int8_t *invert);

/* ---- ChainFile.lift + liftBED record body, binary  (chain.d:664-713,
 *      bed.d:129-196) ----------------------------------------------------- */

/* One lifted row.  Rows of one query are consecutive, sorted by raw qStart
 * (bed.d:156-157; ties in block-index order), rows of different queries in
 * query order. */
typedef struct swo_row {
    int32_t qcid_strand; /* bits 0-15 qcid; bits 16-23 strand byte after the
                            cumulative STRAND_TABLE flip (bed.d:173), 0 if no
                            strand input */
    int32_t start;       /* orderStartEnd(qStart,qEnd).start (bed.d:165-168) */
    int32_t end;
    uint32_t src;        /* index of the source query */
} swo_row;

typedef struct swo_thick {
    int32_t thick_start, thick_end; /* bed.d:175-196 */
} swo_thick;

typedef struct swo_lifted {
    uint64_t n_rows;
    uint64_t n_unmatched;           /* queries with zero overlapping blocks */
    const uint32_t *row_offset;     /* [n+1] exclusive prefix of per-query fan-out */
    const swo_row *rows;            /* [n_rows] */
    const swo_thick *thick;         /* [n_rows] or NULL when no thick input */
} swo_lifted;

/* Host arrays in, pinned host arrays out.  tcid = swo_tcontig_id (negative ->
 * no hit).  strand, thick_start/thick_end may be NULL (BED3 / BED6).
 * Any record order is accepted and gives the same rows; a batch sorted by
 * (contig, start) is recognised on the device and runs the faster variant. */
int swo_lift_intervals(swo_ctx *ctx, size_t n, const int32_t *tcid, const int32_t *start,
                       const int32_t *end, const uint8_t *strand, const int32_t *thick_start,
                       const int32_t *thick_end, swo_lifted *out);

/* Same kernel, everything resident on device `devices[0]` (kernel-only
 * measurement; no host copies).  d_row_offset[n+1], d_rows[rows_cap],
 * d_thick[rows_cap] (or NULL) are caller-allocated device buffers.
 * Asynchronous on `stream` (a cudaStream_t, NULL = legacy default stream);
 * totals[0]=n_rows, totals[1]=n_unmatched are written to the DEVICE array
 * d_totals[2].  If n_rows exceeds rows_cap nothing is written past the
 * capacity and d_totals[0] still holds the required count. */
int swo_lift_intervals_dev(swo_ctx *ctx, size_t n, const int32_t *d_tcid, const int32_t *d_start,
                           const int32_t *d_end, const uint8_t *d_strand,
                           const int32_t *d_thick_start, const int32_t *d_thick_end,
                           uint32_t *d_row_offset, swo_row *d_rows, swo_thick *d_thick,
                           size_t rows_cap, uint64_t *d_totals, void *stream);

/* ---- liftCoordOnly / liftDirectly, batched  (chain.d:604-654) ----------- */

/* hit[i] = 0|1; on 1 out_pos[i] = lifted coordinate and out_qcid[i] the
 * destination contig id; on 0 out_pos[i] = pos[i], out_qcid[i] = -1. */
int swo_lift_points(swo_ctx *ctx, size_t n, const int32_t *tcid, const int32_t *pos,
                    int32_t *out_qcid, int32_t *out_pos, uint8_t *hit);

/* ---- liftBED, text -> text  (bed.d:72-204) ------------------------------ */

typedef struct swo_text_out {
    const char *lifted;       /* what the reference writes to -o (bed.d:199) */
    size_t lifted_len;
    const char *unmatched;    /* what it writes to -u (bed.d:151)            */
    size_t unmatched_len;
    uint64_t n_records;       /* lines that were queried   */
    uint64_t n_rows;          /* lifted rows               */
    uint64_t n_unmatched;     /* records without any block */
} swo_text_out;

/* `in` = the whole BED file (any host memory; pinned is faster).  With several devices the
 * input is cut into byte-range shards (one per device, a shard owns the lines that start
 * inside it), every shard streams through its own device, and the outputs come back in
 * shard order = single-device order (the reference is one sequential loop,
 * bed.d:113).  A malformed line (fewer than 3 fields, non-integer coordinate; the reference
 * dies there with a range violation / ConvException, bed.d:125-126) gives SWO_E_PARSE, and
 * `out` then holds what the reference had written by that point: the outputs of every line in
 * front of it. */
int swo_lift_bed_text(swo_ctx *ctx, const char *in, size_t in_len, swo_text_out *out);

/* The result of the last swo_lift_bed_text as per-shard pieces in output order (which: 0 =
 * lifted, 1 = unmatched): the concatenation of the pieces is the output.  With
 * swo_set_option(ctx, "concat", 0) the pieces are all there is (out->lifted / out->unmatched
 * are NULL, the lengths are the totals): a caller that writes to a file, like liftBED does
 * (bed.d:151,199), writes the pieces in order and no concatenation copy is made. */
int swo_text_num_segments(const swo_ctx *ctx);
int swo_text_segment(const swo_ctx *ctx, int which, int i, const char **ptr, size_t *len);

/* Same kernels on device-resident text (kernel-only measurement; launches of one ctx that use
 * different streams are ordered one after the other on the device: they share per-device scratch).
 * d_in must be
 * 16-byte aligned.  d_sizes[0]=lifted bytes, [1]=unmatched bytes, [2]=records,
 * [3]=rows, [4]=unmatched records, [5]=error (0 = every line parsed, else
 * UINT64_MAX - byte offset of the first malformed line) — a DEVICE array of 6
 * uint64.  Asynchronous on `stream`. */
int swo_lift_bed_text_dev(swo_ctx *ctx, const char *d_in, size_t in_len, char *d_lifted,
                          size_t lifted_cap, char *d_unmatched, size_t unmatched_cap,
                          uint64_t *d_sizes, void *stream);

/* ---- liftVCF record body  (vcf.d:115-169) ------------------------------- */

/* Destination genome, 2 bit/base (A=0 C=1 G=2 T=3): base b of destination
 * contig q is bits [2*(g%4), 2*(g%4)+2) of packed[g/4], g = contig_off[q]+b.
 * Replaces IndexedFastaFile (vcf.d:38-39); indexed by qcid. */
int swo_load_genome_2bit(swo_ctx *ctx, const uint8_t *packed, size_t packed_bytes,
                         const int64_t *contig_off, const int64_t *contig_len, int n_contigs);

/* Destination genome from an uncompressed FASTA file (IndexedFastaFile(genomefile), vcf.d:38):
 * packed to 2 bit/base with side channels for lower-case (soft-masked) bases and for every
 * non-ACGT symbol, so that the REF check stays byte exact (vcf.d:144).  Sequences are keyed by
 * the chain's destination contigs: call after swo_load_chain. */
int swo_load_genome_fasta(swo_ctx *ctx, const char *path);
/* fa.hasSeq(name) / fa.seqLen(name) (vcf.d:79-83); -1 if absent */
int swo_genome_has_seq(const swo_ctx *ctx, const char *name);
int64_t swo_genome_seq_len(const swo_ctx *ctx, const char *name);

/* flags bit0 = matched (vcf.d:134), bit1 = refchg (vcf.d:144-161).  out_ref has
 * the layout of ref_bytes/ref_off and receives the destination REF
 * (fetchSequence, vcf.d:143, clipped at the contig end), out_reflen its length.
 * Host arrays in, host arrays out; blocking.  The records travel in chunks of 8 Mi through three streams
 * (H2D / kernel / D2H overlap); with the arrays in pinned memory (swo_host_alloc) the copies run at link
 * speed: 2e8 records in 0.1 s on one B200. */
int swo_lift_vcf_points(swo_ctx *ctx, size_t n, const int32_t *tcid, const int32_t *pos,
                        const int32_t *reflen, const uint64_t *ref_off, const char *ref_bytes,
                        size_t ref_bytes_len, int32_t *out_qcid, int32_t *out_pos, uint8_t *flags,
                        int32_t *out_reflen, char *out_ref);

/* Same kernel on device-resident arrays (kernel-only measurement, or callers that already hold the
 * records on the GPU): every pointer is a device pointer on devices[0]; asynchronous on `stream`. */
int swo_lift_vcf_points_dev(swo_ctx *ctx, size_t n, const int32_t *d_tcid, const int32_t *d_pos,
                            const int32_t *d_reflen, const uint64_t *d_ref_off, const char *d_ref_bytes,
                            int32_t *d_out_qcid, int32_t *d_out_pos, uint8_t *d_flags, int32_t *d_out_reflen,
                            char *d_out_ref, void *stream);

/* liftVCF, text -> text (vcf.d:29-176) for an uncompressed text VCF held in memory.  The
 * header rewrite (vcf.d:59-108) and the column split / join run on the host; the per-record
 * work of vcf.d:115-169 (liftDirectly, REF fetch and compare, refchg) runs in
 * lift_vcf_points_kernel.  Columns the reference does not touch are passed through byte for
 * byte.  chainfile / genomefile: the strings written to the ##liftoverChainfile / ##contig
 * source= / ##liftoverGenome lines; filedate: "YYYYMMDD" for ##fileDate (NULL = today).
 * out->lifted = what the reference writes to -o, out->unmatched = to -u (input header + the
 * records without a block).  Needs swo_load_genome_fasta or swo_load_genome_2bit. */
int swo_lift_vcf_text(swo_ctx *ctx, const char *in, size_t in_len, const char *chainfile,
                      const char *genomefile, const char *filedate, swo_text_out *out);

/* ---- liftMAF  (maf.d:119-319) -------------------------------------------- */

/* Record body of liftMAF (maf.d:172-310) on arrays: record i is the zero-based half-open
 * interval [start[i], end[i]) (= MAF start-1, end; maf.d:186) on source contig tcid[i].
 *   nhits[i]  0 -> unmatched (maf.d:192), 1 -> lifted, 2 -> more than one block: dropped (maf.d:298)
 *   on one hit: out_qcid, out_start/out_end (one-based closed, orderStartEnd then start+1,
 *   maf.d:208-214); flags bit0 = lifted, bit2 = '-' block (the caller reverse-complements the
 *   tumor alleles, maf.d:220-224), bit1 = Reference_Allele differs from the destination genome
 *   (maf.d:265); out_ref + out_off[i] receives fetchSequence(contig, out_start, out_end)
 *   (maf.d:264, clipped at the end of the sequence), out_reflen[i] its length.  reflen[i] = -1
 *   marks a "-" Reference_Allele: nothing is fetched or compared (maf.d:263).  The caller sizes
 *   out_ref so that record i owns at least min(end-start, longest block) bytes. */
int swo_lift_maf_records(swo_ctx *ctx, size_t n, const int32_t *tcid, const int32_t *start,
                         const int32_t *end, const int32_t *reflen, const uint64_t *ref_off,
                         const char *ref_bytes, size_t ref_bytes_len, const uint64_t *out_off,
                         size_t out_ref_cap, uint8_t *nhits, int32_t *out_qcid, int32_t *out_start,
                         int32_t *out_end, uint8_t *flags, int32_t *out_reflen, char *out_ref);

/* counters the reference logs at the end of liftMAF (maf.d:160, 309-312) */
typedef struct swo_maf_stats {
    uint64_t n_records, n_matched, n_unmatched, n_multiple, n_refchg, n_neither;
} swo_maf_stats;

/* liftMAF, text -> text (maf.d:119-319) for a MAF held in memory: the first two lines go to both
 * outputs (maf.d:166-171); the column split / join and the tumor-allele reverse complement
 * (maf.d:36-54) run on the host, lift + REF fetch/compare in lift_maf_records_kernel.
 * genomebuild replaces NCBI_Build (maf.d:294).  A record with fewer than 13 fields, a
 * non-integer start/end or start > end is SWO_E_PARSE (reference: range violation /
 * ConvException / assert).  stats may be NULL.  Needs a genome (swo_load_genome_fasta). */
int swo_lift_maf_text(swo_ctx *ctx, const char *in, size_t in_len, const char *genomebuild,
                      swo_text_out *out, swo_maf_stats *stats);

/* ---- helpers ------------------------------------------------------------ */

/* pinned host memory for inputs (cudaHostAlloc); faster H2D than pageable */
void *swo_host_alloc(size_t bytes);
void swo_host_free(void *p);

/* tuning knobs: "chunk_bytes" = bytes of text per H2D/kernel/D2H pipeline stage
 * (default 64 MiB, up to 8 in flight; the tests shrink it to exercise chunk boundaries);
 * "zero_copy_input" = 1: pinned input is read by the kernel over PCIe instead of being staged (default 0);
 * "tile_bytes" = 0 (default: 16 KiB, lowered per chunk when the output is much larger than the input),
 * 4096, 8192 or 16384: text tile size of lift_bed_text_kernel */
int swo_set_option(swo_ctx *ctx, const char *key, int64_t value);

/* number of kernel launches issued by this ctx since creation (bench.py's
 * gpu_launches claim) */
uint64_t swo_launch_count(const swo_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SWIFTOVER_CUDA_H */
