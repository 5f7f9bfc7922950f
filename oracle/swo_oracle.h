/*
 * swo_oracle.h — CPU ORACLE for the swiftover liftover hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or execute it.  The product (libswiftover_cuda.so) never calls it.
 *
 * It is a plain-C restatement of the reference's algorithm for the path
 *   chain file -> per-contig block index -> interval/point lookup -> BED/VCF rows
 * written from the behaviour of (paths relative to /root/reference):
 *   source/swiftover/chain.d:272-405   Chain ctor (block generation)
 *   source/swiftover/chain.d:509-597   ChainFile ctor (file segmentation, ids)
 *   source/swiftover/chain.d:604-713   liftDirectly / liftCoordOnly / lift
 *   source/swiftover/chain.d:728-743   intersect
 *   source/swiftover/chain.d:787-805   orderStartEnd, ContigIDorAdd
 *   source/swiftover/bed.d:34-52,113-202   STRAND_TABLE, liftBED record loop
 *   source/swiftover/vcf.d:115-169     liftVCF record loop
 *   source/swiftover/maf.d:21-54,119-319   NT_COMP_TABLE, reverse_complement, liftMAF
 *
 * Third-party pieces that are NOT in /root/reference and are restated from
 * their published behaviour (dub.json:7-11, versions floating):
 *   intervaltree ~>0.22.1  findOverlapsWith / overlaps (half-open)  -> sorted
 *       array + prefix-max-end + binary search here (same hit SET; the
 *       traversal ORDER of the splay tree is history dependent, see below)
 *   Phobos std.algorithm.sort (unstable) at bed.d:157
 *   dklib ~>0.1.1 khash (name -> tree / id / size)
 *
 * PARITY PINNING.  The reference ships no golden outputs.  The
 * oracle is pinned against the only known answers the reference holds:
 *   chain.d:204-221  (ChainLink.qEnd / size / opCmp KATs)
 *   chain.d:451-481  (19-block chrX chain -> nlinks == 19)
 * and against hand-derivable answers on resources/minimal.chain (tests/).
 * For everything else (lift/intersect/liftBED text) parity is UNPINNED BY THE
 * REFERENCE; rules cite the lines above.
 *
 * DEFINED DEVIATIONS (the reference's behaviour is undefined there):
 *   - order of equal-qStart rows inside one record: reference = splay traversal
 *     order x unstable sort (bed.d:154-157 "order is undefined").  Here: stable
 *     sort of hits taken in (tStart,tEnd,qStart) order.
 *   - first hit of a point query over target-overlapping chains (o.front(),
 *     chain.d:613,646): here the lowest block in (tStart,tEnd,qStart) order.
 *   - contig missing from the chain: reference indexes a hash with a missing
 *     key (README.md:59-62, crash).  Here: zero hits -> record is unmatched.
 *   - line with < 3 fields / non-integer coordinate: reference out-of-bounds
 *     or ConvException.  Here: error ORC_E_PARSE, nothing emitted.
 *   - blocks equal in (tStart,tEnd,qStart) on one contig (opCmp == 0,
 *     chain.d:147-164): tree insert keeps the first; so does the oracle.
 */
#ifndef SWO_ORACLE_H
#define SWO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_OK = 0, ORC_E_IO = -1, ORC_E_PARSE = -2, ORC_E_ARG = -3, ORC_E_OOM = -4 };

/* chain.d:107-118 */
typedef struct orc_link {
    int64_t tStart, tEnd, qStart;
    int32_t qcid;
    int32_t invert; /* -1 | +1 */
} orc_link;

typedef struct orc_chainfile orc_chainfile;

typedef struct orc_buf {
    char *data;
    size_t len, cap;
} orc_buf;

typedef struct orc_bed_stats {
    uint64_t n_lines;     /* physical lines seen            */
    uint64_t n_records;   /* lines that were lifted/queried */
    uint64_t n_rows;      /* lifted rows written            */
    uint64_t n_unmatched; /* records routed to unmatched    */
} orc_bed_stats;

/* chain.d:128-144 */
int64_t orc_link_qend(const orc_link *l);
int64_t orc_link_size(const orc_link *l);
/* chain.d:147-171 */
int64_t orc_link_cmp(const orc_link *a, const orc_link *b);
int orc_link_cmp_scalar(const orc_link *a, int64_t x);
/* chain.d:728-743 */
orc_link orc_intersect(orc_link link, int64_t start, int64_t end);
/* chain.d:787-793 */
void orc_order_start_end(int64_t *start, int64_t *end);
/* bed.d:34-52,173 : new strand byte for (old byte, invert) */
unsigned char orc_strand_flip(unsigned char c, int invert);

/* chain.d:272-405 on one chain's lines; returns number of links or <0.  Used
 * by the KAT for chain.d:451-481. */
int64_t orc_chain_count_links(const char *text, size_t len);

/* chain.d:509-597 */
orc_chainfile *orc_chain_parse(const char *text, size_t len, char *err, size_t errcap);
orc_chainfile *orc_chain_load(const char *path, char *err, size_t errcap);
void orc_chain_free(orc_chainfile *cf);

int orc_num_tcontigs(const orc_chainfile *cf);
const char *orc_tcontig_name(const orc_chainfile *cf, int i);
int64_t orc_tcontig_nblocks(const orc_chainfile *cf, int i);
int orc_tcontig_id(const orc_chainfile *cf, const char *name, size_t len); /* -1 if absent */
/* block j of target contig i in (tStart,tEnd,qStart) order */
orc_link orc_tcontig_block(const orc_chainfile *cf, int i, int64_t j);
int64_t orc_total_blocks(const orc_chainfile *cf);

/* ChainFile.contigNames (chain.d:500) */
int orc_num_qcontigs(const orc_chainfile *cf);
const char *orc_qcontig_name(const orc_chainfile *cf, int id);
/* ChainFile.qContigSizes (chain.d:504, 551) in first-insertion order */
int orc_num_qsizes(const orc_chainfile *cf);
const char *orc_qsize_name(const orc_chainfile *cf, int i);
int64_t orc_qsize_value(const orc_chainfile *cf, int i);

/* chain.d:664-713.  Hits in (tStart,tEnd,qStart) order, each intersected.
 * *out is malloc'd (caller frees).  tcid < 0 -> 0 hits. */
int64_t orc_lift(const orc_chainfile *cf, int tcid, int64_t start, int64_t end, orc_link **out);
/* chain.d:637-654.  returns 0|1, rewrites *coord on 1 */
int orc_lift_coord_only(const orc_chainfile *cf, int tcid, int64_t *coord);
/* chain.d:604-627.  returns 0|1, rewrites *qcid,*coord on 1 */
int orc_lift_directly(const orc_chainfile *cf, int tcid, int32_t *qcid, int64_t *coord);

/* bed.d:113-202 text -> text.  Appends to lifted / unmatched. */
int orc_lift_bed(const orc_chainfile *cf, const char *in, size_t in_len, orc_buf *lifted,
                 orc_buf *unmatched, orc_bed_stats *st, char *err, size_t errcap);
/* same, over `nthreads` line-aligned byte ranges in parallel (pthreads),
 * outputs concatenated in range order (== single-thread output). */
int orc_lift_bed_mt(const orc_chainfile *cf, const char *in, size_t in_len, int nthreads,
                    orc_buf *lifted, orc_buf *unmatched, orc_bed_stats *st, char *err,
                    size_t errcap);
void orc_buf_free(orc_buf *b);

/* Binary form of the same record body (bed.d:129-196 without the text):
 * per query i the rows sorted by (qStart, hit order); strand/thick optional
 * (NULL).  Outputs are malloc'd; caller frees with orc_free. */
typedef struct orc_rows {
    uint64_t n_rows, n_unmatched;
    uint32_t *row_offset; /* [n+1] */
    int32_t *qcid, *start, *end;
    uint32_t *src;
    uint8_t *strand;             /* [n_rows] if strand given */
    int32_t *thick_start, *thick_end; /* [n_rows] if thick given */
    uint32_t *unmatched_idx;     /* [n_unmatched] */
} orc_rows;
int orc_lift_intervals(const orc_chainfile *cf, size_t n, const int32_t *tcid,
                       const int32_t *start, const int32_t *end, const uint8_t *strand,
                       const int32_t *thick_start, const int32_t *thick_end, orc_rows *out);
void orc_rows_free(orc_rows *r);

/* vcf.d:115-169 on (contig,pos0,REF) triples with a 2-bit destination genome.
 * genome: base b of query contig q is bits [2*(g%4), 2*(g%4)+2) of
 * packed[g/4], g = contig_off[q] + b; codes A=0 C=1 G=2 T=3.
 * flags bit0 = matched, bit1 = refchg.  out_ref (same layout as ref_bytes,
 * ref_off) receives the destination REF (clipped at contig end like faidx),
 * out_reflen its length. */
int orc_lift_vcf_points(const orc_chainfile *cf, const uint8_t *packed,
                        const int64_t *contig_off, const int64_t *contig_len, int n_gcontigs,
                        size_t n, const int32_t *tcid, const int32_t *pos,
                        const int32_t *reflen, const uint64_t *ref_off, const char *ref_bytes,
                        int32_t *out_qcid, int32_t *out_pos, uint8_t *flags,
                        int32_t *out_reflen, char *out_ref);

/* liftVCF on text (vcf.d:29-176), restated for an UNCOMPRESSED text VCF and an in-memory FASTA.
 * The reference goes through htslib/dhtslib (absent here, parity unpinned): this restatement
 * keeps every column it does not touch byte for byte, which is what htslib does for records
 * whose numeric fields are already in canonical form.
 *   header (vcf.d:59-108): input "##" lines in order; then, before "#CHROM",
 *     ##INFO=<ID=refchg,Number=0,Type=Flag,Description="REF allele changed in new genome">
 *     ##contig=<ID=k,length=n,source=<chainfile>>  for every destination contig k of the chain
 *        (contigSort order, vcf.d:181-215) that the FASTA has with the same length, unless a
 *        contig line with that ID already exists (htslib does not add a duplicate ID)
 *     ##fileDate=<filedate>, ##liftoverProgram=swiftover, ##liftoverProgramURL=...,
 *     ##liftoverChainfile=<chainfile>, ##liftoverGenome=<genomefile>
 *     the unmatched file gets the input header unchanged (vcf.d:107-108)
 *   record (vcf.d:115-169): liftDirectly(CHROM, POS-1); no hit -> unmatched file, unchanged;
 *     hit -> CHROM, POS rewritten; newRef = FASTA[chrom][pos, pos+len(REF)) clipped at the end
 *     of the sequence; if REF != newRef (byte exact): REF = newRef and INFO gets the flag
 *     "refchg" ("." -> "refchg", else appended with ';').  ALT untouched, no strand handling.
 * fasta: plain FASTA text ('>' name up to the first white space; sequence lines joined). */
typedef struct orc_vcf_stats {
    uint64_t n_records, n_matched, n_unmatched, n_refchg;
} orc_vcf_stats;
int orc_lift_vcf_text(const orc_chainfile *cf, const char *fasta, size_t fasta_len, const char *in,
                      size_t in_len, const char *chainfile, const char *genomefile,
                      const char *filedate, orc_buf *lifted, orc_buf *unmatched,
                      orc_vcf_stats *st, char *err, size_t errcap);

/* liftMAF on text (maf.d:119-319), with an in-memory FASTA for the destination genome.
 *   lines 1 and 2 (version header, column header; maf.d:166-171) go to both outputs verbatim;
 *   record (maf.d:172-310): split on white-space runs; start,end = to!long of columns 6,7
 *     (1-based closed); hits = lift(contig, start-1, end)
 *     0 hits  -> unmatched file, fields tab-joined (maf.d:190-193)
 *     1 hit   -> contig = destination name; start,end = orderStartEnd(qStart,qEnd), start+1;
 *                '-' block: Tumor_Seq_Allele1/2 reverse-complemented through NT_COMP_TABLE
 *                (maf.d:21-54: only ACGTacgt and '-' map, every other byte becomes NUL, "NA" stays);
 *                newRef = REF == "-" ? "-" : FASTA[contig][start-1, end) (clipped at the end of the
 *                sequence); REF != newRef -> REF = newRef (maf.d:263-291); NCBI_Build = genomebuild;
 *                fields tab-joined to the lifted file (maf.d:294-296)
 *     >1 hits -> record is dropped from both outputs (maf.d:298-307)
 * DEFINED DEVIATIONS: a record with fewer than 13 fields, a non-integer start/end or
 * start > end is ORC_E_PARSE (reference: range violation / ConvException / assert); allele bytes >= 128 map to NUL
 * (reference indexes a 128-entry table out of bounds); a destination contig the FASTA lacks
 * yields an empty newRef.  dhtslib/htslib faidx is absent from /root/reference: parity unpinned. */
typedef struct orc_maf_stats {
    uint64_t n_records, n_matched, n_unmatched, n_multiple, n_refchg, n_neither; /* maf.d:160, 309-312 */
} orc_maf_stats;
int orc_lift_maf_text(const orc_chainfile *cf, const char *fasta, size_t fasta_len, const char *in,
                      size_t in_len, const char *genomebuild, orc_buf *lifted, orc_buf *unmatched,
                      orc_maf_stats *st, char *err, size_t errcap);
/* maf.d:36-54; out has room for n bytes; returns the output length (== n) */
size_t orc_reverse_complement(const char *allele, size_t n, char *out);

/* Workload helper for bench.py --impl reference (not part of the restatement):
 * writes "name\tstart\tend\n" rows for n records, names taken from the chain's
 * source contigs.  Returns bytes written (<= cap) or 0 if cap is too small. */
size_t orc_format_bed3(const orc_chainfile *cf, size_t n, const int32_t *tcid, const int32_t *start,
                       const int32_t *end, char *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
