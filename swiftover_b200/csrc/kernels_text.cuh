// kernels_text.cuh — liftBED on device: BED text in, lifted + unmatched BED text out.
//
// Replaces the reference's record loop (/root/reference/source/swiftover/bed.d:113-202):
//   skip '#', "track", "brows" lines              bed.d:115-116
//   split on white-space runs                      bed.d:118-119
//   start/end = to!int                             bed.d:125-126
//   cf.lift                                        bed.d:129 -> chain.d:664-713
//   thickStart/thickEnd point lifts                bed.d:134-148
//   unmatched: fields joined by '\t'               bed.d:150-151
//   rows sorted by qStart, strand table, thick clamp, joined by '\t'   bed.d:156-199
//
// ONE persistent kernel, every input byte read once from HBM and every output
// byte written once.  A tile is 16 KiB of input text staged in shared memory
// with 128-bit loads; a tile owns the lines that START inside it.  Each thread
// owns the lines starting in its 64-byte segment.  Pass A parses + searches and
// sums the exact output byte counts of both streams; the tile aggregates are
// chained with a decoupled look-back (two independent chains, one per output
// stream); pass B formats into a shared-memory stage that is then copied out
// with aligned 128-bit stores.  Lines with >= 2 hits are expanded cooperatively
// (warp / block), each row computing its own sorted position and byte offset.
#pragma once
#include "device_util.cuh"
#include "lift_core.cuh"

namespace swo {

constexpr int TX_TPB = 256;
constexpr int TX_NW = TX_TPB / 32;
constexpr int TX_SEG = 64;                   // bytes of input owned per thread
constexpr int TX_TILE = TX_TPB * TX_SEG;     // 16384
constexpr int TX_PRE = 16;                   // bytes staged before the tile
constexpr int TX_LOOK = 1024;                // bytes staged after the tile
constexpr int TX_WIN = TX_PRE + TX_TILE + TX_LOOK;
constexpr int TX_OUT_STAGE = 20480;          // lifted bytes staged per tile
constexpr int TX_UNM_STAGE = 4096;           // unmatched bytes staged per tile
constexpr int TX_WL_CAP = 128;
constexpr int TX_HEAVY_MIN = 256;

// open-addressing name table entry (source contig name -> tcid)
struct NameEnt {
    uint32_t hash;
    int32_t id;  // -1 = empty
    uint32_t off, len;
};

struct TextArgs {
    TableView T;
    const NameEnt *thash;
    uint32_t thash_mask;
    const char *tname_chars;
    const uint32_t *qname_off;  // [nq+1]
    const char *qname_chars;
    const char *in;
    uint64_t in_len;
    int ntiles;
    char *lifted;
    uint64_t lifted_cap;
    char *unm;
    uint64_t unm_cap;
    uint64_t *desc_l, *desc_u;  // [ntiles] each, zeroed
    unsigned *ticket;           // zeroed
    uint64_t *sizes;            // [6], zeroed: lifted, unmatched, records, rows, unmatched recs, error
};

// ---- input access ------------------------------------------------------------
struct Reader {
    const unsigned char *win;  // shared-memory window
    uint64_t ws;               // absolute offset of win[0]
    uint32_t wlen;
    const unsigned char *g;
    uint64_t in_len;
    // byte at absolute position p; the end of input reads as '\n' (byLine yields a
    // final unterminated line, bed.d:113)
    __device__ __forceinline__ int at(uint64_t p) const {
        if (p >= in_len) return '\n';
        const uint64_t r = p - ws;
        if (r < wlen) return win[r];
        return __ldg(g + p);
    }
};

__device__ __forceinline__ bool tx_space(int c) { return c == ' ' || (c >= 9 && c <= 13); }  // '\n' included

// next field at or after p on the current line: [a,b).  false at end of line.
__device__ __forceinline__ bool next_field(const Reader &R, uint64_t p, uint64_t &a, uint64_t &b) {
    int c = R.at(p);
    while (c != '\n' && tx_space(c)) c = R.at(++p);
    if (c == '\n') return false;
    a = p;
    while (!tx_space(c)) c = R.at(++p);
    b = p;
    return true;
}

// to!int on [a,b)  (bed.d:125-126,135-136)
__device__ __forceinline__ bool field_i32(const Reader &R, uint64_t a, uint64_t b, int &out) {
    int c = R.at(a);
    bool neg = false;
    if (c == '+' || c == '-') {
        neg = c == '-';
        a++;
    }
    if (a >= b) return false;
    long long v = 0;
    for (; a < b; a++) {
        const unsigned d = (unsigned)(R.at(a) - '0');
        if (d > 9u) return false;
        v = v * 10 + d;
        if (v > 2147483648ll) return false;
    }
    if (neg) v = -v;
    if (v > 2147483647ll) return false;
    out = (int)v;
    return true;
}

__device__ __forceinline__ int lookup_tcontig(const TextArgs &A, const Reader &R, uint64_t a, uint64_t b) {
    uint32_t h = 0x9E3779B9u;
    for (uint64_t p = a; p < b; p++) h = (h ^ (uint32_t)R.at(p)) * 0x01000193u;
    h ^= h >> 15;
    const uint32_t len = (uint32_t)(b - a);
    for (uint32_t s = h & A.thash_mask;; s = (s + 1) & A.thash_mask) {
        const NameEnt e = A.thash[s];
        if (e.id < 0) return -1;
        if (e.hash == h && e.len == len) {
            uint32_t i = 0;
            while (i < len && __ldg(A.tname_chars + e.off + i) == (char)R.at(a + i)) i++;
            if (i == len) return e.id;
        }
    }
}

enum { LN_RECORD = 0, LN_SKIP = 1, LN_ERROR = 2 };

struct LineInfo {
    int status;
    int numf;  // saturates at 8
    int tcid, start, end;
    int ts, te;             // thickStart, thickEnd text values (numf >= 8)
    unsigned char strand;   // first byte of field 5 (numf >= 6)
    uint64_t off3;          // where to resume tokenising for fields 3..
    uint32_t head_len;      // len(f0)+len(f1)+len(f2)
    uint32_t pre_len;       // sum over fields 3..min(numf,6)-1 of (1+len)
    uint32_t f6len, f7len;  // text lengths of fields 6,7
    uint32_t post_len;      // sum over fields 8.. of (1+len)
};

__device__ __forceinline__ LineInfo scan_line(const TextArgs &A, const Reader &R, uint64_t p0) {
    LineInfo L;
    L.status = LN_SKIP;
    L.numf = 0;
    L.tcid = -1;
    L.start = L.end = L.ts = L.te = 0;
    L.strand = 0;
    L.off3 = p0;
    L.head_len = L.pre_len = L.f6len = L.f7len = L.post_len = 0;
    {  // bed.d:115-116 (tests on the raw line)
        const int c0 = R.at(p0);
        if (c0 == '#') return L;
        if (c0 == 't' || c0 == 'b') {
            const int c1 = R.at(p0 + 1);
            if (c1 != '\n') {
                const int c2 = R.at(p0 + 2);
                if (c2 != '\n') {
                    const int c3 = R.at(p0 + 3);
                    if (c3 != '\n') {
                        const int c4 = R.at(p0 + 4);
                        if (c4 != '\n') {
                            if (c0 == 't' && c1 == 'r' && c2 == 'a' && c3 == 'c' && c4 == 'k') return L;
                            if (c0 == 'b' && c1 == 'r' && c2 == 'o' && c3 == 'w' && c4 == 's') return L;
                        }
                    }
                }
            }
        }
    }
    L.status = LN_ERROR;
    uint64_t a, b;
    if (!next_field(R, p0, a, b)) return L;  // blank line: < 3 fields
    L.tcid = lookup_tcontig(A, R, a, b);
    L.head_len = (uint32_t)(b - a);
    if (!next_field(R, b, a, b)) return L;
    if (!field_i32(R, a, b, L.start)) return L;
    L.head_len += (uint32_t)(b - a);
    if (!next_field(R, b, a, b)) return L;
    if (!field_i32(R, a, b, L.end)) return L;
    L.head_len += (uint32_t)(b - a);
    L.numf = 3;
    L.off3 = b;
    uint64_t a6 = 0, b6 = 0, a7 = 0, b7 = 0;
    while (next_field(R, b, a, b)) {
        const uint32_t len = (uint32_t)(b - a);
        const int c = L.numf;
        if (c < 6) {
            L.pre_len += 1 + len;
            if (c == 5) L.strand = (unsigned char)R.at(a);
        } else if (c == 6) {
            L.f6len = len;
            a6 = a;
            b6 = b;
        } else if (c == 7) {
            L.f7len = len;
            a7 = a;
            b7 = b;
        } else {
            L.post_len += 1 + len;
        }
        if (L.numf < 8) L.numf++;
    }
    if (L.numf >= 8) {  // bed.d:134-136
        if (!field_i32(R, a6, b6, L.ts) || !field_i32(R, a7, b7, L.te)) return L;
    }
    L.status = LN_RECORD;
    return L;
}

// ---- formatting ----------------------------------------------------------------
__device__ __forceinline__ int ndigits_i32(int v) {  // length of the decimal text of v
    unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
    int n = v < 0 ? 2 : 1;
    if (u >= 100000u) {
        if (u >= 10000000u) n += u >= 1000000000u ? 9 : (u >= 100000000u ? 8 : 7);
        else n += u >= 1000000u ? 6 : 5;
    } else {
        if (u >= 1000u) n += u >= 10000u ? 4 : 3;
        else n += u >= 100u ? 2 : (u >= 10u ? 1 : 0);
    }
    return n;
}
__device__ __forceinline__ char *put_i32(char *d, int v) {
    const int n = ndigits_i32(v);
    unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
    char *e = d + n;
    char *p = e;
    do {
        const unsigned q = u / 10u;
        *--p = (char)('0' + (u - q * 10u));
        u = q;
    } while (u);
    if (v < 0) *--p = '-';
    return e;
}

__device__ __forceinline__ uint32_t qname_len(const TextArgs &A, int qcid) {
    return __ldg(A.qname_off + qcid + 1) - __ldg(A.qname_off + qcid);
}

// bytes of one lifted row (bed.d:161-199)
__device__ __forceinline__ uint32_t row_len(const TextArgs &A, const LineInfo &L, const Thick &th, const RowCore &r) {
    uint32_t n = qname_len(A, r.qcid) + 1 + ndigits_i32(r.s) + 1 + ndigits_i32(r.e) + L.pre_len + L.post_len + 1;
    if (L.numf == 7) n += 1 + L.f6len;
    if (L.numf >= 8) {
        int c7, c8;
        thick_clamp(th, r.s, r.e, c7, c8);
        n += 2 + ndigits_i32(c7) + ndigits_i32(c8);
    }
    return n;
}
// bytes of the unmatched line (bed.d:151)
__device__ __forceinline__ uint32_t unmatched_len(const LineInfo &L) {
    uint32_t n = L.head_len + 2 + L.pre_len + L.post_len + 1;
    if (L.numf >= 7) n += 1 + L.f6len;
    if (L.numf >= 8) n += 1 + L.f7len;
    return n;
}

__device__ __forceinline__ char *copy_field(char *d, const Reader &R, uint64_t a, uint64_t b) {
    for (; a < b; a++) *d++ = (char)R.at(a);
    return d;
}

__device__ __forceinline__ char *emit_row(char *d, const TextArgs &A, const Reader &R, const LineInfo &L,
                                          const Thick &th, const RowCore &r, unsigned char strand_byte) {
    {
        const uint32_t o = __ldg(A.qname_off + r.qcid), e = __ldg(A.qname_off + r.qcid + 1);
        for (uint32_t i = o; i < e; i++) *d++ = __ldg(A.qname_chars + i);  // bed.d:163
    }
    *d++ = '\t';
    d = put_i32(d, r.s);  // bed.d:169
    *d++ = '\t';
    d = put_i32(d, r.e);  // bed.d:170
    if (L.numf > 3) {
        int c7 = 0, c8 = 0;
        if (L.numf >= 8) thick_clamp(th, r.s, r.e, c7, c8);
        uint64_t a, b = L.off3;
        int c = 3;
        while (next_field(R, b, a, b)) {
            *d++ = '\t';
            if (c == 5) {  // bed.d:173
                *d++ = (char)strand_byte;
                d = copy_field(d, R, a + 1, b);
            } else if (c == 6 && L.numf >= 8) {
                d = put_i32(d, c7);
            } else if (c == 7 && L.numf >= 8) {
                d = put_i32(d, c8);
            } else {
                d = copy_field(d, R, a, b);
            }
            c++;
        }
    }
    *d++ = '\n';
    return d;
}

__device__ __forceinline__ char *emit_unmatched(char *d, const Reader &R, uint64_t p0) {
    uint64_t a, b = p0;
    bool first = true;
    while (next_field(R, b, a, b)) {
        if (!first) *d++ = '\t';
        first = false;
        d = copy_field(d, R, a, b);
    }
    *d++ = '\n';
    return d;
}

// Expand one multi-hit line: each candidate block computes its rank, its byte
// offset inside the record's output (sum of the lengths of the rows sorted before
// it) and its strand byte, then formats its row.  (first,stride) = work split.
__device__ __forceinline__ void expand_line(const TextArgs &A, const Reader &R, uint64_t p0, int lo, int ncand,
                                            char *dst, int first, int stride) {
    const TableView &T = A.T;
    const LineInfo L = scan_line(A, R, p0);
    Thick th{0, 0, 0};
    if (L.numf >= 8) th = lift_thick(T, L.tcid, L.ts, L.te);
    for (int c = first; c < ncand; c += stride) {
        const Block bj = load_block(T, lo + c);
        if (!T.disjoint && !block_hits(bj, L.start)) continue;
        const RowCore rj = make_row(bj, L.start, L.end);
        int rank = 0, negs = rj.inv < 0 ? 1 : 0;
        uint32_t off = 0;
        int best_key = rj.key, best_k = c, best_inv = rj.inv;
        for (int k = 0; k < ncand; k++) {
            if (k == c) continue;
            const Block bk = load_block(T, lo + k);
            if (!T.disjoint && !block_hits(bk, L.start)) continue;
            const RowCore rk = make_row(bk, L.start, L.end);
            const bool before = rk.key < rj.key || (rk.key == rj.key && k < c);
            if (before) {
                rank++;
                negs += rk.inv < 0 ? 1 : 0;
                off += row_len(A, L, th, rk);
            }
            if (rk.key < best_key || (rk.key == best_key && k < best_k)) {
                best_key = rk.key;
                best_k = k;
                best_inv = rk.inv;
            }
        }
        const unsigned char sb = L.numf >= 6 ? strand_at_rank(L.strand, best_inv, rank, negs) : 0;
        emit_row(dst + off, A, R, L, th, rj, sb);
    }
}

// copy n bytes from shared memory (4-byte aligned base) to global dst with
// 16-byte aligned vector stores in the middle
__device__ __forceinline__ void copy_out(char *dst, const uint32_t *s_words, uint32_t n) {
    const unsigned char *s = reinterpret_cast<const unsigned char *>(s_words);
    uint32_t head = (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
    if (head > n) head = n;
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = (char)s[i];
    const uint32_t nvec = (n - head) >> 4;
    const uint32_t sh = (head & 3u) * 8u;
    const uint32_t w0 = head >> 2;
    for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
        const uint32_t w = w0 + v * 4u;
        const uint32_t a0 = s_words[w], a1 = s_words[w + 1], a2 = s_words[w + 2], a3 = s_words[w + 3], a4 = s_words[w + 4];
        int4 o;
        o.x = (int)__funnelshift_r(a0, a1, sh);
        o.y = (int)__funnelshift_r(a1, a2, sh);
        o.z = (int)__funnelshift_r(a2, a3, sh);
        o.w = (int)__funnelshift_r(a3, a4, sh);
        st_stream_v4(dst + head + (size_t)v * 16u, o);
    }
    const uint32_t done = head + nvec * 16u;
    for (uint32_t i = done + threadIdx.x; i < n; i += blockDim.x) dst[i] = (char)s[i];
}

struct TextSmem {
    uint32_t in_words[TX_WIN / 4 + 4];
    uint32_t out_words[TX_OUT_STAGE / 4 + 8];
    uint32_t unm_words[TX_UNM_STAGE / 4 + 8];
    uint32_t scan_a[TX_NW + 1];
    uint32_t scan_b[TX_NW + 1];
    uint64_t base_l, base_u;
    int tile;
    int wl_n;
    uint32_t wl_p0[TX_WL_CAP];   // line start relative to tile start
    int wl_lo[TX_WL_CAP];
    int wl_cnt[TX_WL_CAP];
    uint32_t wl_off[TX_WL_CAP];  // byte offset inside the tile's lifted output
};

__global__ void __launch_bounds__(TX_TPB) lift_bed_text_kernel(const TextArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TextSmem &S = *reinterpret_cast<TextSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const TableView &T = A.T;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            S.tile = (int)atomicAdd(A.ticket, 1u);
            S.wl_n = 0;
        }
        __syncthreads();
        const int tile = S.tile;
        if (tile >= A.ntiles) break;
        const uint64_t t0 = (uint64_t)tile * TX_TILE;

        // ---- stage the window [t0-16, t0+TILE+LOOK) -------------------------
        Reader R;
        R.win = reinterpret_cast<const unsigned char *>(S.in_words);
        R.ws = tile == 0 ? 0 : t0 - TX_PRE;
        {
            const uint64_t wend = min(A.in_len, t0 + TX_TILE + TX_LOOK);
            R.wlen = (uint32_t)(wend - R.ws);
        }
        R.g = reinterpret_cast<const unsigned char *>(A.in);
        R.in_len = A.in_len;
        {
            const uint32_t nvec = R.wlen >> 4;
            const int4 *src = reinterpret_cast<const int4 *>(A.in + R.ws);
            int4 *dstv = reinterpret_cast<int4 *>(S.in_words);
            for (uint32_t v = tid; v < nvec; v += TX_TPB) dstv[v] = ld_stream_v4(src + v);
            unsigned char *sb = reinterpret_cast<unsigned char *>(S.in_words);
            for (uint32_t i = (nvec << 4) + tid; i < R.wlen; i += TX_TPB) sb[i] = R.g[R.ws + i];
        }
        __syncthreads();

        // ---- pass A: parse + search + exact output byte counts -------------------
        const uint64_t seg = t0 + (uint64_t)tid * TX_SEG;
        uint64_t starts = 0;  // bit i: a line starts at seg+i
        if (seg < A.in_len) {
#pragma unroll
            for (int w = 0; w < TX_SEG / 4; w++) {
                // bytes seg-1+4w .. seg+2+4w  (a '\n' at p-1 means a line starts at p)
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t p = seg + 4 * w + k;
                    const bool st = p < A.in_len && (p == 0 || R.at(p - 1) == '\n');
                    x |= st ? (1u << k) : 0u;
                }
                starts |= (uint64_t)x << (4 * w);
            }
        }
        uint32_t a_l = 0, a_u = 0, a_rec = 0, a_rows = 0, a_unm = 0;
        for (uint64_t m = starts; m; m &= m - 1) {
            const uint64_t p0 = seg + (uint64_t)(__ffsll((long long)m) - 1);
            const LineInfo L = scan_line(A, R, p0);
            if (L.status == LN_SKIP) continue;
            if (L.status == LN_ERROR) {
                atomicMax((unsigned long long *)(A.sizes + 5), ~(unsigned long long)p0);  // max(~p0) == first bad line
                continue;
            }
            a_rec++;
            const Cand cd = find_candidates(T, L.tcid, L.start, L.end);
            const int n = count_hits(T, cd, L.start);
            if (n == 0) {
                a_u += unmatched_len(L);
                a_unm++;
                continue;
            }
            a_rows += (uint32_t)n;
            Thick th{0, 0, 0};
            if (L.numf >= 8) th = lift_thick(T, L.tcid, L.ts, L.te);
            for (int j = cd.lo; j < cd.hi; j++) {
                const Block b = load_block(T, j);
                if (!T.disjoint && !block_hits(b, L.start)) continue;
                a_l += row_len(A, L, th, make_row(b, L.start, L.end));
            }
        }
        uint32_t tot_l, tot_u;
        const uint32_t ex_l = block_exclusive_scan<uint32_t, TX_NW>(a_l, S.scan_a, &tot_l);
        const uint32_t ex_u = block_exclusive_scan<uint32_t, TX_NW>(a_u, S.scan_b, &tot_u);
        if (warp_id() == 0) {
            const uint64_t b = lookback_exclusive(A.desc_l, tile, (uint64_t)tot_l);
            if (tid == 0) S.base_l = b;
        } else if (warp_id() == 1) {
            const uint64_t b = lookback_exclusive(A.desc_u, tile, (uint64_t)tot_u);
            if (lane_id() == 0) S.base_u = b;
        }
        {  // record counters (every warp)
            const uint32_t r = warp_sum(a_rec), w = warp_sum(a_rows), u = warp_sum(a_unm);
            if (lane_id() == 0) {
                if (r) atomicAdd((unsigned long long *)(A.sizes + 2), (unsigned long long)r);
                if (w) atomicAdd((unsigned long long *)(A.sizes + 3), (unsigned long long)w);
                if (u) atomicAdd((unsigned long long *)(A.sizes + 4), (unsigned long long)u);
            }
        }
        __syncthreads();
        const uint64_t gb_l = S.base_l, gb_u = S.base_u;
        if (tile == A.ntiles - 1 && tid == 0) {
            A.sizes[0] = gb_l + tot_l;
            A.sizes[1] = gb_u + tot_u;
        }
        const bool fit_l = gb_l + tot_l <= A.lifted_cap, fit_u = gb_u + tot_u <= A.unm_cap;
        const bool stage_l = tot_l <= TX_OUT_STAGE, stage_u = tot_u <= TX_UNM_STAGE;
        char *base_l = stage_l ? reinterpret_cast<char *>(S.out_words) : A.lifted + gb_l;
        char *base_u = stage_u ? reinterpret_cast<char *>(S.unm_words) : A.unm + gb_u;

        // ---- pass B: format ---------------------------------------------------------
        uint32_t o_l = ex_l, o_u = ex_u;
        for (uint64_t m = starts; m; m &= m - 1) {
            const uint64_t p0 = seg + (uint64_t)(__ffsll((long long)m) - 1);
            const LineInfo L = scan_line(A, R, p0);
            if (L.status != LN_RECORD) continue;
            const Cand cd = find_candidates(T, L.tcid, L.start, L.end);
            const int n = count_hits(T, cd, L.start);
            if (n == 0) {
                if (fit_u) emit_unmatched(base_u + o_u, R, p0);
                o_u += unmatched_len(L);
                continue;
            }
            Thick th{0, 0, 0};
            if (L.numf >= 8) th = lift_thick(T, L.tcid, L.ts, L.te);
            if (n == 1) {
                int j = cd.lo;
                Block b = load_block(T, j);
                if (!T.disjoint)
                    while (!block_hits(b, L.start)) b = load_block(T, ++j);
                const RowCore r = make_row(b, L.start, L.end);
                if (fit_l) emit_row(base_l + o_l, A, R, L, th, r, strand_flip(L.strand, r.inv));
                o_l += row_len(A, L, th, r);
                continue;
            }
            // n >= 2: total length again (cheap relative to the expansion), queue the expansion
            uint32_t len = 0;
            for (int j = cd.lo; j < cd.hi; j++) {
                const Block b = load_block(T, j);
                if (!T.disjoint && !block_hits(b, L.start)) continue;
                len += row_len(A, L, th, make_row(b, L.start, L.end));
            }
            if (fit_l) {
                const int slot = atomicAdd(&S.wl_n, 1);
                if (slot < TX_WL_CAP) {
                    S.wl_p0[slot] = (uint32_t)(p0 - t0);
                    S.wl_lo[slot] = cd.lo;
                    S.wl_cnt[slot] = cd.hi - cd.lo;
                    S.wl_off[slot] = o_l;
                } else {
                    expand_line(A, R, p0, cd.lo, cd.hi - cd.lo, base_l + o_l, 0, 1);
                }
            }
            o_l += len;
        }
        __syncthreads();
        const int nwl = min(S.wl_n, TX_WL_CAP);
        for (int e = warp_id(); e < nwl; e += TX_NW)
            if (S.wl_cnt[e] < TX_HEAVY_MIN)
                expand_line(A, R, t0 + S.wl_p0[e], S.wl_lo[e], S.wl_cnt[e], base_l + S.wl_off[e], lane_id(), 32);
        for (int e = 0; e < nwl; e++)
            if (S.wl_cnt[e] >= TX_HEAVY_MIN)
                expand_line(A, R, t0 + S.wl_p0[e], S.wl_lo[e], S.wl_cnt[e], base_l + S.wl_off[e], tid, TX_TPB);
        __syncthreads();
        // ---- copy the stages out ---------------------------------------------------
        if (stage_l && fit_l && tot_l) copy_out(A.lifted + gb_l, S.out_words, tot_l);
        if (stage_u && fit_u && tot_u) copy_out(A.unm + gb_u, S.unm_words, tot_u);
    }
}

}  // namespace swo
