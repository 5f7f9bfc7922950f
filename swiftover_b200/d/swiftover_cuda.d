/// swiftover_cuda.d — extern(C) binding of libswiftover_cuda (include/swiftover_cuda.h)
/// for the D host.  NOT BUILT in this repository's image (no ldc2/dmd/gdc/dub);
/// kept so that the reference's D host can switch its ChainFile / liftBED bodies to the
/// CUDA engine: see INTEGRATION.md.
module swiftover.cuda;

extern (C) nothrow @nogc:

struct swo_ctx;

enum : int
{
    SWO_OK = 0, SWO_E_IO = -1, SWO_E_PARSE = -2, SWO_E_CUDA = -3, SWO_E_OOM = -4,
    SWO_E_ARG = -5, SWO_E_STATE = -6, SWO_E_CAPACITY = -7
}

struct swo_row { int qcid_strand; int start; int end; uint src; }
struct swo_thick { int thick_start; int thick_end; }
struct swo_lifted
{
    ulong n_rows; ulong n_unmatched;
    const(uint)* row_offset; const(swo_row)* rows; const(swo_thick)* thick;
}
struct swo_text_out
{
    const(char)* lifted; size_t lifted_len;
    const(char)* unmatched; size_t unmatched_len;
    ulong n_records; ulong n_rows; ulong n_unmatched;
}

int swo_create(swo_ctx** ctx, const(int)* devices, int ndev);
int swo_create_inspect(swo_ctx** ctx);
void swo_destroy(swo_ctx* ctx);
const(char)* swo_last_error(const(swo_ctx)* ctx);
const(char)* swo_version();

int swo_load_chain(swo_ctx* ctx, const(char)* path);
int swo_load_chain_mem(swo_ctx* ctx, const(char)* text, size_t len);
int swo_num_tcontigs(const(swo_ctx)* ctx);
const(char)* swo_tcontig_name(const(swo_ctx)* ctx, int tcid);
int swo_tcontig_id(const(swo_ctx)* ctx, const(char)* name, size_t len);
long swo_tcontig_nblocks(const(swo_ctx)* ctx, int tcid);
long swo_tcontig_size(const(swo_ctx)* ctx, int tcid);
long swo_num_blocks(const(swo_ctx)* ctx);
int swo_table_is_disjoint(const(swo_ctx)* ctx);
int swo_num_qcontigs(const(swo_ctx)* ctx);
const(char)* swo_qcontig_name(const(swo_ctx)* ctx, int qcid);
int swo_num_qsizes(const(swo_ctx)* ctx);
const(char)* swo_qsize_name(const(swo_ctx)* ctx, int i);
long swo_qsize_value(const(swo_ctx)* ctx, int i);
int swo_get_blocks(const(swo_ctx)* ctx, int tcid, long first, long n, int* t_start, int* t_end,
                   int* q_start, int* qcid, byte* invert);

int swo_lift_intervals(swo_ctx* ctx, size_t n, const(int)* tcid, const(int)* start, const(int)* end,
                       const(ubyte)* strand, const(int)* thick_start, const(int)* thick_end, swo_lifted* o);
int swo_lift_intervals_dev(swo_ctx* ctx, size_t n, const(int)* d_tcid, const(int)* d_start, const(int)* d_end,
                           const(ubyte)* d_strand, const(int)* d_thick_start, const(int)* d_thick_end,
                           uint* d_row_offset, swo_row* d_rows, swo_thick* d_thick, size_t rows_cap,
                           ulong* d_totals, void* stream);
int swo_lift_points(swo_ctx* ctx, size_t n, const(int)* tcid, const(int)* pos, int* out_qcid, int* out_pos,
                    ubyte* hit);
int swo_lift_bed_text(swo_ctx* ctx, const(char)* inbuf, size_t in_len, swo_text_out* o);
int swo_text_num_segments(const(swo_ctx)* ctx);
int swo_text_segment(const(swo_ctx)* ctx, int which, int i, const(char)** ptr, size_t* len);
int swo_lift_bed_text_dev(swo_ctx* ctx, const(char)* d_in, size_t in_len, char* d_lifted, size_t lifted_cap,
                          char* d_unmatched, size_t unmatched_cap, ulong* d_sizes, void* stream);
int swo_load_genome_2bit(swo_ctx* ctx, const(ubyte)* packed, size_t packed_bytes, const(long)* contig_off,
                         const(long)* contig_len, int n_contigs);
int swo_lift_vcf_points(swo_ctx* ctx, size_t n, const(int)* tcid, const(int)* pos, const(int)* reflen,
                        const(ulong)* ref_off, const(char)* ref_bytes, size_t ref_bytes_len, int* out_qcid,
                        int* out_pos, ubyte* flags, int* out_reflen, char* out_ref);
int swo_lift_vcf_points_dev(swo_ctx* ctx, size_t n, const(int)* d_tcid, const(int)* d_pos, const(int)* d_reflen,
                            const(ulong)* d_ref_off, const(char)* d_ref_bytes, int* d_out_qcid, int* d_out_pos,
                            ubyte* d_flags, int* d_out_reflen, char* d_out_ref, void* stream);
int swo_load_genome_fasta(swo_ctx* ctx, const(char)* path);
int swo_genome_has_seq(const(swo_ctx)* ctx, const(char)* name);
long swo_genome_seq_len(const(swo_ctx)* ctx, const(char)* name);
int swo_lift_vcf_text(swo_ctx* ctx, const(char)* inbuf, size_t in_len, const(char)* chainfile,
                      const(char)* genomefile, const(char)* filedate, swo_text_out* o);
/// counters liftMAF logs (maf.d:160, 309-312)
struct swo_maf_stats { ulong n_records, n_matched, n_unmatched, n_multiple, n_refchg, n_neither; }
int swo_lift_maf_records(swo_ctx* ctx, size_t n, const(int)* tcid, const(int)* start, const(int)* end,
                         const(int)* reflen, const(ulong)* ref_off, const(char)* ref_bytes, size_t ref_bytes_len,
                         const(ulong)* out_off, size_t out_ref_cap, ubyte* nhits, int* out_qcid, int* out_start,
                         int* out_end, ubyte* flags, int* out_reflen, char* out_ref);
int swo_lift_maf_text(swo_ctx* ctx, const(char)* inbuf, size_t in_len, const(char)* genomebuild, swo_text_out* o,
                      swo_maf_stats* stats);
void* swo_host_alloc(size_t bytes);
void swo_host_free(void* p);
int swo_set_option(swo_ctx* ctx, const(char)* key, long value);
ulong swo_launch_count(const(swo_ctx)* ctx);
