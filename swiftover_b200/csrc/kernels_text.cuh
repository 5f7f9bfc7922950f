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
// ONE persistent kernel; every input byte is read from HBM once and every output
// byte written once.  A tile is 16 KiB of text staged in shared memory by one
// 1-D bulk copy (TMA, completion on an mbarrier); a tile owns the lines that
// START inside it.
//   1. line starts: word-wide (SWAR) newline masks of the staged window, block
//      scan, table of line-start offsets in shared memory;
//   2. count phase: lines are dealt to threads round-robin (up to TX_NB = 3 per
//      thread, results cached in registers).  A thread parses its line ONCE with
//      word-wide loads (SWAR separator / digit masks and decimal conversion),
//      looks the contig name up, searches the block table (upper levels =
//      Eytzinger-ordered group maxima in shared memory, then a quaternary search
//      inside one group), and computes the exact byte length of its output;
//   3. one packed multi-scan places lifted and unmatched text of all lines of the
//      tile; the tile totals are published to the decoupled look-back chains;
//   4. emit phase: rows are formatted from the cached state into ONE shared-memory
//      stage holding the tile's whole output (lifted text from the front, unmatched
//      text at the end);
//   5. the resolve of the chained offsets and the copy-out (aligned 128-bit
//      stores) of a tile are deferred until the NEXT tile has been counted, so
//      that the look-back never waits for a predecessor's formatting.
// Fast path = canonical lines (single tabs, plain digits, inside the window).
// Anything else (other white space, signs, >9 digits, lines longer than the
// look-ahead, malformed lines) takes the generic per-line path below, which
// follows the reference's tokenizer exactly.  Records with >= 2 hits: rows already
// in qStart order (or exactly reversed) are written sequentially by the owning
// thread; the rest are expanded cooperatively (warp / block), each row computing
// its own sorted position, byte offset and cumulative strand byte.  A stream that
// does not fit its stage is written directly to global memory at the chained offset.
#pragma once
#include "device_util.cuh"
#include "lift_core.cuh"

namespace swo {

#ifndef TX_MIN_CTAS
#define TX_MIN_CTAS 2  // resident CTAs per SM the register budget is sized for
#endif
#ifndef TX_THREADS
#define TX_THREADS 256
#endif
#ifndef TX_NB
#define TX_NB 3  // batches of TX_THREADS lines whose parse results stay in registers
#endif
constexpr int TX_TPB = TX_THREADS;
constexpr int TX_NW = TX_TPB / 32;
#ifndef TX_TILE_V
#define TX_TILE_V 16384
#endif
constexpr int TX_TILE = TX_TILE_V;               // largest tile (bytes of text); TextArgs.tile_bytes picks 16, 8 or 4 KiB
constexpr int TX_MSEG = 64;                  // bytes per newline-mask word
static_assert(TX_TILE / TX_TPB == 64 || TX_TILE / TX_TPB == 32, "TX_THREADS must be 256 or 512");
constexpr int TX_PRE = 64;                   // bytes staged before the tile
constexpr int TX_LOOK = 1024;                // bytes staged after the tile
constexpr int TX_WIN = TX_PRE + TX_TILE + TX_LOOK;
constexpr int TX_PAD = 32;                   // always '\n' after the window
constexpr int TX_NSEG = TX_WIN / TX_MSEG;    // 273
constexpr int TX_OUT_STAGE = TX_TILE / 4 * 5;       // lifted bytes staged per tile
constexpr int TX_ROUND_SLACK = 4096;         // multi-round emit: a line's rows may run this far past its round's window
constexpr int TX_UNM_STAGE = TX_TILE + TX_LOOK;  // unmatched bytes staged per tile: a tile's unmatched text never exceeds its input
// One stage for both streams: the unmatched text (known size: the scan has run) sits at the end, the lifted
// text may use everything in front of it — with few unmatched records a tile's lifted text may be up to
// 2.2x its input before the tile has to go out in rounds (with two fixed stages, 20 + 17 KiB, BED8 input
// whose output is 1.15x the input sent a good part of its tiles through the rounds: configs[2] text
// 127 -> see DESIGN.md).  The rounds themselves use the first TX_OUT_STAGE bytes only.
constexpr int TX_STAGE_ALL = TX_OUT_STAGE + TX_UNM_STAGE;
constexpr int TX_LCAP = TX_TILE / 8;           // line-start table entries
constexpr int TX_QN_MAX = 256;               // destination contig names kept in shared memory ...
constexpr int TX_QCHARS = 3072;              // ... if their characters fit here
constexpr int TX_WL_CAP = 256;
#ifndef TX_HEAVY_MIN_V
#define TX_HEAVY_MIN_V 64
#endif
constexpr int TX_HEAVY_MIN = TX_HEAVY_MIN_V;            // candidates above which the whole block expands a line
constexpr int TX_XSCR = 512;                 // shared-memory scratch entries (= TX_NW warp regions of TX_HEAVY_MIN)
constexpr int TX_GSCR = 8192;                // per-CTA global scratch entries for heavier lines
#ifndef TX_CW_MIN_V
#define TX_CW_MIN_V 48
#endif
constexpr int TX_CW_MIN = TX_CW_MIN_V;                // candidates above which a line's output length is summed by the block
constexpr int TX_CW_CAP = 64;                // such lines per tile
static_assert(TX_NW * TX_HEAVY_MIN <= TX_XSCR, "one scratch region per warp");
constexpr int TX_EY_CAP = ChainTable::EY_CAP;

// open-addressing name table entry (source contig name -> tcid): two 128-bit loads, the
// first 16 bytes of the name inline so that the usual verification needs no further load
struct NameEnt {
    uint32_t hash;
    int32_t id;  // -1 = empty
    uint32_t off, len;
    uint32_t w[4];  // name bytes 0..15, little endian, zero padded
};

struct TextArgs {
    TableView T;
    unsigned full_mask;  // 0xffffffff as a run-time value (converge_warp, device_util.cuh)
    const NameEnt *thash;
    uint32_t thash_mask;
    const char *tname_chars;
    const uint32_t *qname_off;  // [nq+1]
    const char *qname_chars;
    int nq;
    const char *in;   // 16-byte aligned
    uint64_t in_len;  // includes `skip`
    uint32_t skip;    // leading bytes (< 16) that belong to the previous chunk; the text starts at in + skip
    uint32_t tile_bytes;  // TX_TILE, or a smaller power of two (>= 16 * TX_THREADS) for inputs whose output is
                          // much larger than the input, so that a tile's lifted text still fits the stage
    int ntiles;
    char *lifted;
    uint64_t lifted_cap;
    char *unm;
    uint64_t unm_cap;
    unsigned long long *gscr;   // per-CTA scratch for very heavy lines: grid * TX_GSCR * 12 bytes (or null)
    uint64_t *desc_l, *desc_u;  // [ntiles] each, zeroed
    unsigned *ticket;           // zeroed
    uint64_t *sizes;            // [6], zeroed: lifted, unmatched, records, rows, unmatched recs, error
};

// =====================================================================================
// Generic per-line path (exact tokenizer; any white space; lines of any length)
// =====================================================================================
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

// name table probe: hash h, length len, first 16 bytes w0..w3; `tail_eq(off)` compares
// the bytes from 16 on against the stored name at tname_chars + off (names longer than 16)
template <typename TailEq>
__device__ __forceinline__ int probe_tcontig(const TextArgs &A, uint32_t h, uint32_t len, uint32_t w0, uint32_t w1,
                                             uint32_t w2, uint32_t w3, TailEq tail_eq) {
    const int4 *tab = reinterpret_cast<const int4 *>(A.thash);
    for (uint32_t s = h & A.thash_mask;; s = (s + 1) & A.thash_mask) {
        const int4 e0 = __ldg(tab + 2 * s), e1 = __ldg(tab + 2 * s + 1);
        if (e0.y < 0) return -1;
        if ((uint32_t)e0.x == h && (uint32_t)e0.w == len && (uint32_t)e1.x == w0 && (uint32_t)e1.y == w1 &&
            (uint32_t)e1.z == w2 && (uint32_t)e1.w == w3 && (len <= 16u || tail_eq((uint32_t)e0.z)))
            return e0.y;
    }
}

__device__ __forceinline__ int lookup_tcontig(const TextArgs &A, const Reader &R, uint64_t a, uint64_t b) {
    const uint32_t len = (uint32_t)(b - a);
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; i++)
        if ((uint32_t)i < len) w[i >> 2] |= (uint32_t)R.at(a + i) << (8 * (i & 3));
    uint32_t h = name_hash_head(w[0], w[1], w[2], w[3], len);
    for (uint64_t p = a + 16; p < b; p++) h = name_hash_tail(h, (uint8_t)R.at(p));
    h = name_hash_fin(h);
    return probe_tcontig(A, h, len, w[0], w[1], w[2], w[3], [&](uint32_t off) {
        uint32_t i = 16;
        while (i < len && __ldg(A.tname_chars + off + i) == (char)R.at(a + i)) i++;
        return i == len;
    });
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

__device__ __noinline__ void scan_line(const TextArgs &A, const Reader &R, uint64_t p0, LineInfo &L) {
    L.status = LN_SKIP;
    L.numf = 0;
    L.tcid = -1;
    L.start = L.end = L.ts = L.te = 0;
    L.strand = 0;
    L.off3 = p0;
    L.head_len = L.pre_len = L.f6len = L.f7len = L.post_len = 0;
    {  // bed.d:115-116 (tests on the raw line)
        const int c0 = R.at(p0);
        if (c0 == '#') return;
        if (c0 == 't' || c0 == 'b') {
            const int c1 = R.at(p0 + 1);
            if (c1 != '\n') {
                const int c2 = R.at(p0 + 2);
                if (c2 != '\n') {
                    const int c3 = R.at(p0 + 3);
                    if (c3 != '\n') {
                        const int c4 = R.at(p0 + 4);
                        if (c4 != '\n') {
                            if (c0 == 't' && c1 == 'r' && c2 == 'a' && c3 == 'c' && c4 == 'k') return;
                            if (c0 == 'b' && c1 == 'r' && c2 == 'o' && c3 == 'w' && c4 == 's') return;
                        }
                    }
                }
            }
        }
    }
    L.status = LN_ERROR;
    uint64_t a, b;
    if (!next_field(R, p0, a, b)) return;  // blank line: < 3 fields
    L.tcid = lookup_tcontig(A, R, a, b);
    L.head_len = (uint32_t)(b - a);
    if (!next_field(R, b, a, b)) return;
    if (!field_i32(R, a, b, L.start)) return;
    L.head_len += (uint32_t)(b - a);
    if (!next_field(R, b, a, b)) return;
    if (!field_i32(R, a, b, L.end)) return;
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
        if (!field_i32(R, a6, b6, L.ts) || !field_i32(R, a7, b7, L.te)) return;
    }
    L.status = LN_RECORD;
}

// ---- formatting ----------------------------------------------------------------
__device__ __forceinline__ int ndigits_u32(unsigned u) {
    int n = 1;
    if (u >= 100000u) {
        if (u >= 10000000u) n += u >= 1000000000u ? 9 : (u >= 100000000u ? 8 : 7);
        else n += u >= 1000000u ? 6 : 5;
    } else {
        if (u >= 1000u) n += u >= 10000u ? 4 : 3;
        else n += u >= 100u ? 2 : (u >= 10u ? 1 : 0);
    }
    return n;
}
__device__ __forceinline__ int ndigits_i32(int v) {  // length of the decimal text of v
    return ndigits_u32(v < 0 ? 0u - (unsigned)v : (unsigned)v) + (v < 0 ? 1 : 0);
}
__device__ __forceinline__ char *put_i32(char *d, int v) {
    const int n = ndigits_i32(v);
    unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
    char *e = d + n;
    char *p = e;
    while (u >= 100u) {  // two digits per step
        const unsigned q = u / 100u, r = u - q * 100u;
        const unsigned t = (r * 103u) >> 10;  // r / 10 for r < 100
        p -= 2;
        p[0] = (char)('0' + t);
        p[1] = (char)('0' + (r - t * 10u));
        u = q;
    }
    if (u >= 10u) {
        const unsigned t = (u * 103u) >> 10;
        p -= 2;
        p[0] = (char)('0' + t);
        p[1] = (char)('0' + (u - t * 10u));
    } else {
        *--p = (char)('0' + u);
    }
    if (v < 0) *--p = '-';
    return e;
}

// Destination-name table: usually the CTA's shared-memory copy, else the global one; the two
// pointers live in shared memory (TextSmem::qoff_p / qchars_p, set once per CTA).  The kernel's
// argument struct itself is never copied or modified: it stays in the constant bank.
__device__ __forceinline__ const uint32_t *tx_qoff();
__device__ __forceinline__ const char *tx_qchars();
__device__ __forceinline__ uint32_t qname_len(const TextArgs &, int qcid) {
    const uint32_t *o = tx_qoff();
    return o[qcid + 1] - o[qcid];
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

__device__ __forceinline__ char *emit_qname(char *d, const TextArgs &, int qcid) {
    const uint32_t *off = tx_qoff();
    const char *ch = tx_qchars();
    const uint32_t o = off[qcid], e = off[qcid + 1];
    for (uint32_t i = o; i < e; i++) *d++ = ch[i];  // bed.d:163
    return d;
}

__device__ __forceinline__ char *emit_row(char *d, const TextArgs &A, const Reader &R, const LineInfo &L,
                                          const Thick &th, const RowCore &r, unsigned char strand_byte) {
    d = emit_qname(d, A, r.qcid);
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

// Generic line, count side: output byte counts of both streams.
struct SlowCount {
    int status;
    uint32_t len_l, len_u;
    int lo, ncand, nhits;
};
__device__ __noinline__ void slow_count(const TextArgs &A, const int2 *ey, const Reader &R, uint64_t p0, SlowCount &C) {
    const TableView &T = A.T;
    LineInfo L;
    scan_line(A, R, p0, L);
    C.status = L.status;
    C.len_l = C.len_u = 0;
    C.lo = C.ncand = C.nhits = 0;
    if (L.status != LN_RECORD) return;
    const Cand cd = find_candidates(T, ey, L.tcid, L.start, L.end);
    const int n = count_hits(T, cd, L.start);
    C.lo = cd.lo;
    C.ncand = cd.hi - cd.lo;
    C.nhits = n;
    if (n == 0) {
        C.len_u = unmatched_len(L);
        return;
    }
    Thick th{0, 0, 0};
    if (L.numf >= 8) th = lift_thick(T, ey, L.tcid, L.ts, L.te);
    for (int j = cd.lo; j < cd.hi; j++) {
        const Block b = load_block(T, j);
        if (!T.disjoint && !block_hits(b, L.start)) continue;
        C.len_l += row_len(A, L, th, make_row(b, L.start, L.end));
    }
}
// Generic line, emit side for 0 or 1 hit (>= 2 hits go through expand_line).
__device__ __noinline__ void slow_emit(const TextArgs &A, const int2 *ey, const Reader &R, uint64_t p0, int lo, int nhits,
                                       char *dst) {
    const TableView &T = A.T;
    if (nhits == 0) {
        emit_unmatched(dst, R, p0);
        return;
    }
    LineInfo L;
    scan_line(A, R, p0, L);
    Thick th{0, 0, 0};
    if (L.numf >= 8) th = lift_thick(T, ey, L.tcid, L.ts, L.te);
    int j = lo;
    Block b = load_block(T, j);
    if (!T.disjoint)
        while (!block_hits(b, L.start)) b = load_block(T, ++j);
    const RowCore r = make_row(b, L.start, L.end);
    emit_row(dst, A, R, L, th, r, strand_flip(L.strand, r.inv));
}

// Expand one multi-hit line: each candidate block computes its rank, its byte
// offset inside the record's output (sum of the lengths of the rows sorted before
// it) and its strand byte, then formats its row.  (first,stride) = work split.
__device__ __noinline__ void expand_line(const TextArgs &A, const int2 *ey, const Reader &R, uint64_t p0, int lo,
                                         int ncand, char *dst, int first, int stride) {
    const TableView &T = A.T;
    LineInfo L;
    scan_line(A, R, p0, L);
    Thick th{0, 0, 0};
    if (L.numf >= 8) th = lift_thick(T, ey, L.tcid, L.ts, L.te);
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

// Cooperative form used for the queued multi-hit lines (one warp, or the whole block for
// heavy fan-out).  Linear or n log^2 n, never quadratic:
//   1. one 64-bit sort key (qStart, candidate index) per candidate goes to `scr`, its row
//      length and invert to `aux` (shared memory, or a per-CTA global scratch for very heavy lines);
//   2. if the keys are not already ascending (mixed chains, '-' chains) they are sorted with a
//      bitonic network;
//   3. rows are taken in sorted order; a running scan of the lengths gives every row its byte
//      offset, a running count of inverted rows its cumulative strand byte (bed.d:173).
// BLOCK: all threads of the CTA call it (scan_scr: TX_NW+1 words for the block scan).
template <bool BLOCK>
__device__ __noinline__ void expand_line_coop(const TextArgs &A, const int2 *ey, const Reader &R, uint64_t p0, int lo,
                                              int ncand, char *dst, unsigned long long *scr, uint32_t *aux,
                                              unsigned long long *scan_scr, int first, int stride) {
    const TableView &T = A.T;
    LineInfo L;
    scan_line(A, R, p0, L);
    Thick th{0, 0, 0};
    if (L.numf >= 8) th = lift_thick(T, ey, L.tcid, L.ts, L.te);
    int P = 32;
    while (P < ncand) P <<= 1;
    for (int c = first; c < P; c += stride) {
        unsigned long long key = ~0ull;  // padding and non-hits sort last
        if (c < ncand) {
            const Block b = load_block(T, lo + c);
            const RowCore r = make_row(b, L.start, L.end);
            if (T.disjoint || block_hits(b, L.start)) {
                key = ((unsigned long long)((uint32_t)r.key ^ 0x80000000u) << 32) | (uint32_t)c;
                aux[c] = (row_len(A, L, th, r) << 1) | (r.inv < 0 ? 1u : 0u);
            }
        }
        scr[c] = key;
    }
    converge_warp(A.full_mask);  // (a plain __syncwarp() may be dropped by ptxas: device_util.cuh)
    if (BLOCK) __syncthreads();
    bool mine = true;
    for (int c = first; c < P; c += stride)
        if (c > 0 && scr[c - 1] > scr[c]) mine = false;
    const bool sorted = BLOCK ? (__syncthreads_and(mine) != 0) : (__all_sync(FULL, mine) != 0);
    if (!sorted) {
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = first; i < P; i += stride) {
                    const int x = i ^ j;
                    if (x > i) {
                        const unsigned long long va = scr[i], vb = scr[x];
                        if ((va > vb) == ((i & k) == 0)) {
                            scr[i] = vb;
                            scr[x] = va;
                        }
                    }
                }
                if (BLOCK) __syncthreads();
                else converge_warp(A.full_mask);
            }
        }
    }
    // rows in output order
    unsigned long long carry = 0;  // (bytes << 32) | inverted rows, of all earlier positions
    const int inv_first = (aux[(uint32_t)scr[0]] & 1u) ? -1 : 1;  // position 0 is a hit: the line has >= 2 hits
    for (int base = 0; base < P; base += stride) {
        const int j = base + first;
        const unsigned long long e = j < P ? scr[j] : ~0ull;
        const bool hit = e != ~0ull;
        const uint32_t c = (uint32_t)e;
        const uint32_t av = hit ? aux[c] : 0u;
        const unsigned long long v = ((unsigned long long)(av >> 1) << 32) | (av & 1u);
        unsigned long long ex, tot;
        if (BLOCK) {
            ex = block_exclusive_scan<unsigned long long, TX_NW>(v, scan_scr, &tot);
        } else {
            const unsigned long long inc = warp_inclusive_scan(v);
            ex = inc - v;
            tot = __shfl_sync(FULL, inc, 31);
        }
        if (hit) {
            const unsigned long long at = carry + ex;
            const int negs_le = (int)(uint32_t)at + (int)(av & 1u);
            const RowCore rj = make_row(load_block(T, lo + (int)c), L.start, L.end);
            const unsigned char sb = L.numf >= 6 ? strand_at_rank(L.strand, inv_first, j, negs_le) : 0;
            emit_row(dst + (uint32_t)(at >> 32), A, R, L, th, rj, sb);
        }
        carry += tot;
    }
    if (BLOCK) __syncthreads();  // scratch is reused by the next heavy line
}

// =====================================================================================
// Fast path: canonical line, parsed once, kept in registers between count and emit
// =====================================================================================
enum { FL_NONE = 0, FL_FAST = 1, FL_SLOW = 2 };

struct FastLine {
    int numf;             // 3..8 (saturated)
    int tcid, start, end;
    uint32_t len0;        // length of the contig name
    uint32_t q;           // window offset of the terminating '\n'
    uint32_t e2;          // window offset of the separator after field 2 (== q when numf == 3)
    uint32_t a5, a6, b7;  // first byte of field 5, first byte of field 6, end of field 7
    int v6, v7;           // thickStart, thickEnd
    unsigned char strand;
};

// ---- word-wide access to the staged window -----------------------------------------
// 16 bytes starting at window offset i (any alignment): five aligned word loads and four
// funnel shifts, issued together (one shared-memory latency for the whole chunk)
__device__ __forceinline__ void load16(const uint32_t *W, uint32_t i, uint32_t &x0, uint32_t &x1, uint32_t &x2,
                                       uint32_t &x3) {
    const uint32_t wi = i >> 2, sh = (i & 3u) * 8u;
    const uint32_t a = W[wi], b = W[wi + 1], c = W[wi + 2], d = W[wi + 3], e = W[wi + 4];
    x0 = __funnelshift_r(a, b, sh);
    x1 = __funnelshift_r(b, c, sh);
    x2 = __funnelshift_r(c, d, sh);
    x3 = __funnelshift_r(d, e, sh);
}
__device__ __forceinline__ void load12(const uint32_t *W, uint32_t i, uint32_t &x0, uint32_t &x1, uint32_t &x2) {
    const uint32_t wi = i >> 2, sh = (i & 3u) * 8u;
    const uint32_t a = W[wi], b = W[wi + 1], c = W[wi + 2], d = W[wi + 3];
    x0 = __funnelshift_r(a, b, sh);
    x1 = __funnelshift_r(b, c, sh);
    x2 = __funnelshift_r(c, d, sh);
}
// bit 7 of every byte of x that is <= 0x20 (field separators and other control bytes)
__device__ __forceinline__ uint32_t le20_marks(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) | x) & 0x80808080u;
}
// bit 7 of every byte of x that is not an ASCII digit
__device__ __forceinline__ uint32_t nondigit_marks(uint32_t x) {
    const uint32_t t = x ^ 0x30303030u;
    return (((t & 0x7F7F7F7Fu) + 0x76767676u) | t) & 0x80808080u;
}
// byte k (0..11) of the 12-byte chunk x0,x1,x2
__device__ __forceinline__ uint32_t chunk_byte(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t k) {
    const uint32_t w = k < 4u ? x0 : (k < 8u ? x1 : x2);
    return (w >> ((k & 3u) * 8u)) & 0xFFu;
}
// Unsigned decimal of 1..9 digits at window offset i: value, digit count and the byte that
// follows.  false: no digit, or more than 9 digits (generic path decides).
__device__ __forceinline__ bool parse_uint_swar(const uint32_t *W, uint32_t i, uint32_t &val, uint32_t &n,
                                                uint32_t &term) {
    uint32_t x0, x1, x2;
    load12(W, i, x0, x1, x2);
    const uint32_t m0 = nondigit_marks(x0), m1 = nondigit_marks(x1), m2 = nondigit_marks(x2);
    if (m0) n = (uint32_t)(__ffs((int)m0) - 1) >> 3;
    else if (m1) n = 4u + ((uint32_t)(__ffs((int)m1) - 1) >> 3);
    else if (m2) n = 8u + ((uint32_t)(__ffs((int)m2) - 1) >> 3);
    else return false;
    if (n == 0u || n > 9u) return false;
    term = chunk_byte(x0, x1, x2, n);
    const uint32_t k = n < 8u ? n : 8u;
    // the first k digits, as 8 text positions with leading zeros
    unsigned long long tt = ((unsigned long long)(x1 ^ 0x30303030u) << 32) | (x0 ^ 0x30303030u);
    tt <<= 8u * (8u - k);
    const uint32_t lo = (uint32_t)tt, hi = (uint32_t)(tt >> 32);
    const uint32_t ul = (lo * 10u + (lo >> 8)) & 0x00FF00FFu, uh = (hi * 10u + (hi >> 8)) & 0x00FF00FFu;
    const uint32_t vl = (ul & 0xFFFFu) * 100u + (ul >> 16), vh = (uh & 0xFFFFu) * 100u + (uh >> 16);
    val = vl * 10000u + vh;
    if (n == 9u) val = val * 10u + ((x2 ^ 0x30u) & 0xFFu);
    return true;
}

// Parse the line starting at window offset p.  `limit`: a terminator at or beyond
// this offset is not a real line end (the line runs out of the staged window).
// On FL_FAST the first 16 name bytes are left in nw[4] (zero padded) for the lookup.
__device__ __forceinline__ int fast_parse(const TextArgs &A, const unsigned char *win, uint32_t p, uint32_t limit,
                                          FastLine &F, uint32_t (&nw)[4]) {
    const uint32_t *W = reinterpret_cast<const uint32_t *>(win);
    load16(W, p, nw[0], nw[1], nw[2], nw[3]);
    {
        const uint32_t c = nw[0] & 0xFFu;
        if (c == '#') return FL_NONE;  // bed.d:115
        if ((nw[0] == 0x63617274u && (nw[1] & 0xFFu) == 'k') || (nw[0] == 0x776F7262u && (nw[1] & 0xFFu) == 's'))
            return FL_NONE;  // bed.d:116 "track" / "brows"
    }
    // contig name: up to the first byte <= 0x20, which must be a tab
    uint32_t len0;
    {
        const uint32_t m0 = le20_marks(nw[0]), m1 = le20_marks(nw[1]), m2 = le20_marks(nw[2]), m3 = le20_marks(nw[3]);
        if (m0) len0 = (uint32_t)(__ffs((int)m0) - 1) >> 3;
        else if (m1) len0 = 4u + ((uint32_t)(__ffs((int)m1) - 1) >> 3);
        else if (m2) len0 = 8u + ((uint32_t)(__ffs((int)m2) - 1) >> 3);
        else if (m3) len0 = 12u + ((uint32_t)(__ffs((int)m3) - 1) >> 3);
        else {  // 16 bytes or more
            len0 = 16;
            while (win[p + len0] > 0x20u) len0++;
        }
    }
    if (len0 == 0u || win[p + len0] != '\t') return FL_SLOW;
    if (len0 < 16u) {  // zero the bytes after the name
        const uint32_t keep = len0 & 3u ? 0xFFFFFFFFu >> (32u - 8u * (len0 & 3u)) : 0u;
        const uint32_t wi = len0 >> 2;
        nw[0] = wi == 0 ? nw[0] & keep : nw[0];
        nw[1] = wi == 1 ? nw[1] & keep : (wi < 1 ? 0u : nw[1]);
        nw[2] = wi == 2 ? nw[2] & keep : (wi < 2 ? 0u : nw[2]);
        nw[3] = wi == 3 ? nw[3] & keep : (wi < 3 ? 0u : nw[3]);
    }
    F.len0 = len0;
    uint32_t v, n, term;
    uint32_t i = p + len0 + 1;
    if (!parse_uint_swar(W, i, v, n, term) || term != '\t') return FL_SLOW;
    F.start = (int)v;
    i += n + 1;
    if (!parse_uint_swar(W, i, v, n, term)) return FL_SLOW;
    F.end = (int)v;
    uint32_t j = i + n;
    F.e2 = j;
    F.a5 = F.a6 = F.b7 = 0;
    F.v6 = F.v7 = 0;
    F.strand = 0;
    uint32_t c = term;
    int f = 3;
    bool ok67 = true;
    while (c == '\t') {
        const uint32_t a = j + 1;
        if (f == 6 || f == 7) {
            if (parse_uint_swar(W, a, v, n, term) && term <= 0x20u) {
                j = a + n;
                c = term;
            } else {  // sign, junk or too long: generic path decides
                ok67 = false;
                j = a;
                c = win[j];
                while (c > 0x20u) c = win[++j];
            }
            if (f == 6) {
                F.a6 = a;
                F.v6 = (int)v;
            } else {
                F.b7 = j;
                F.v7 = (int)v;
            }
        } else {
            j = a;
            c = win[j];
            while (c > 0x20u) c = win[++j];
            if (f == 5) {
                F.a5 = a;
                F.strand = win[a];
            }
        }
        if (j == a) return FL_SLOW;  // empty field: white-space run, generic tokenizer
        f++;
    }
    if (c != '\n' || j >= limit) return FL_SLOW;
    F.q = j;
    F.numf = f < 8 ? f : 8;
    if (F.numf >= 8 && !ok67) return FL_SLOW;
    return FL_FAST;
}

// contig lookup for a fast-path line: nw = first 16 name bytes (zero padded)
__device__ __forceinline__ int lookup_tcontig_win(const TextArgs &A, const unsigned char *win, uint32_t p, uint32_t len,
                                                  const uint32_t (&nw)[4]) {
    uint32_t h = name_hash_head(nw[0], nw[1], nw[2], nw[3], len);
    for (uint32_t i = 16; i < len; i++) h = name_hash_tail(h, win[p + i]);
    h = name_hash_fin(h);
    return probe_tcontig(A, h, len, nw[0], nw[1], nw[2], nw[3], [&](uint32_t off) {
        uint32_t i = 16;
        while (i < len && (unsigned char)__ldg(A.tname_chars + off + i) == win[p + i]) i++;
        return i == len;
    });
}

// bytes of a fast-path row that do not depend on the block
__device__ __forceinline__ uint32_t fast_fixed_len(const FastLine &F) {
    uint32_t n = 3 + (F.q - F.e2);  // two tabs, '\n', pass-through tail (with its leading tab)
    if (F.numf >= 8) n -= F.b7 - F.a6;
    return n;
}
__device__ __forceinline__ uint32_t fast_row_len(const TextArgs &A, const FastLine &F, uint32_t fixed, const Thick &th,
                                                 const RowCore &r) {
    uint32_t n = fixed + qname_len(A, r.qcid) + ndigits_i32(r.s) + ndigits_i32(r.e);
    if (F.numf >= 8) {
        int c7, c8;
        thick_clamp(th, r.s, r.e, c7, c8);
        n += 1 + ndigits_i32(c7) + ndigits_i32(c8);
    }
    return n;
}
__device__ __forceinline__ char *copy_span(char *d, const unsigned char *win, uint32_t a, uint32_t b) {
    for (; a < b; a++) *d++ = (char)win[a];
    return d;
}
__device__ __forceinline__ char *fast_emit_row(char *d, const TextArgs &A, const unsigned char *win, const FastLine &F,
                                               const Thick &th, const RowCore &r, unsigned char strand_byte) {
    d = emit_qname(d, A, r.qcid);
    *d++ = '\t';
    d = put_i32(d, r.s);
    *d++ = '\t';
    d = put_i32(d, r.e);
    if (F.numf > 3) {
        if (F.numf < 6) {
            d = copy_span(d, win, F.e2, F.q);
        } else {
            d = copy_span(d, win, F.e2, F.a5);
            *d++ = (char)strand_byte;  // bed.d:173
            if (F.numf < 8) {
                d = copy_span(d, win, F.a5 + 1, F.q);
            } else {
                int c7, c8;
                thick_clamp(th, r.s, r.e, c7, c8);  // bed.d:175-196
                d = copy_span(d, win, F.a5 + 1, F.a6);
                d = put_i32(d, c7);
                *d++ = '\t';
                d = put_i32(d, c8);
                d = copy_span(d, win, F.b7, F.q);
            }
        }
    }
    *d = '\n';
    return d + 1;
}

// copy n bytes, starting at byte offset so of the shared-memory word array s_words, to global
// dst with 16-byte aligned vector stores in the middle
__device__ __forceinline__ void copy_out(char *dst, const uint32_t *s_words, uint32_t n, uint32_t so = 0) {
    const unsigned char *s = reinterpret_cast<const unsigned char *>(s_words) + so;
    uint32_t head = (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
    if (head > n) head = n;
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = (char)s[i];
    const uint32_t nvec = (n - head) >> 4;
    const uint32_t q = so + head;  // byte offset of the first vector's source
    const uint32_t sh = (q & 3u) * 8u;
    const uint32_t w0 = q >> 2;
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

// 4-bit mask of the bytes of x equal to '\n' (bit b = byte b)
__device__ __forceinline__ uint32_t nl_mask4(uint32_t x) {
    x ^= 0x0A0A0A0Au;
    uint32_t t = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;  // bit 7 of a byte clear <=> byte == 0
    t = ~t & 0x80808080u;
    return ((t >> 7) * 0x00204081u) >> 21 & 0xFu;
}


// per-tile constants shared by the helpers; one copy per CTA in shared memory (TextSmem::X),
// written by thread 0 while the tile is staged
struct TileCtx {
    Reader R;
    uint64_t g0;     // absolute offset of window offset s0
    uint32_t s0;
    uint32_t limit;  // see fast_parse
};

// pending tile (formatted, not yet copied out): block-uniform, kept in shared memory
struct PendTile {
    int pend, have_base, direct_l, direct_u, tile;
    uint32_t tot_l, tot_u;
    uint32_t u_off;  // byte offset of the unmatched text in the stage
    uint64_t gb_l, gb_u;
};

struct TextSmem {
    uint32_t in_words[(TX_WIN + TX_PAD) / 4];
    // the tile's output: lifted text from the front, unmatched text at the very end (see TX_STAGE_ALL)
    uint32_t stage_words[TX_STAGE_ALL / 4 + 8];
    unsigned long long nlm[TX_NSEG + 1];
    int2 ey[TX_EY_CAP];
    union {  // the line table is dead once the emit phase is over, when the expansions run
        uint16_t lstart[TX_LCAP];
        uint32_t xaux[TX_XSCR];
    };
    unsigned long long scan64[TX_NB * TX_NW];
    uint32_t scan32[TX_NW + 1];
    uint64_t base_l, base_u;
    uint64_t in_bar;  // mbarrier of the window's bulk copy
    TileCtx X;
    PendTile P;
    unsigned long long lb_part[TX_NW][2];
    int lb_found[TX_NW][2];
    int tile;
    int wl_n;
    uint32_t rlo, rhi;  // byte range staged in the current round (multi-round emit of a large tile)
    // destination contig names (copy, when they fit) and the pointers every formatter uses
    const uint32_t *qoff_p;
    const char *qchars_p;
    uint32_t qoff[TX_QN_MAX + 1];
    char qchars[TX_QCHARS];
    unsigned long long xscr[TX_XSCR];  // expand_line_coop scratch: TX_NW warp regions / one block region
    // lines whose fan-out is too large to be summed by one thread (count phase)
    int cw_n;
    uint32_t cw_p[TX_CW_CAP];
    int cw_lo[TX_CW_CAP], cw_cnt[TX_CW_CAP], cw_tcid[TX_CW_CAP];
    uint32_t cw_len[TX_CW_CAP];
    uint32_t red32[TX_NW];
    uint32_t wl_p0[TX_WL_CAP];   // line start, window offset
    int wl_lo[TX_WL_CAP];
    int wl_cnt[TX_WL_CAP];
    uint32_t wl_off[TX_WL_CAP];  // byte offset inside the tile's lifted output
};

extern __shared__ __align__(16) unsigned char tx_smem_raw[];
__device__ __forceinline__ const uint32_t *tx_qoff() { return reinterpret_cast<const TextSmem *>(tx_smem_raw)->qoff_p; }
__device__ __forceinline__ const char *tx_qchars() { return reinterpret_cast<const TextSmem *>(tx_smem_raw)->qchars_p; }
// fixed shared-memory addresses: no register or load is spent on them
__device__ __forceinline__ TextSmem &tx_smem() { return *reinterpret_cast<TextSmem *>(tx_smem_raw); }
__device__ __forceinline__ const unsigned char *tx_win() { return reinterpret_cast<const unsigned char *>(tx_smem().in_words); }
__device__ __forceinline__ const int2 *tx_ey() { return tx_smem().ey; }

// What a thread keeps about one line between the count and the emit phase (8 words).
//   meta: bits 0-1 kind, 2-3 hit class (0, 1, 2 = two or more), 4-6 numf-3, 7-14 strand byte,
//         one hit:  15 invert<0, 16 thick.has, 17-31 qcid
//         >= 2 hits: 15 rows already ascending in qStart, 16 strictly descending, 17-31 tcid
struct LineState {
    uint32_t pq;    // p | q << 16 : window offsets of the line start and of its '\n'
    uint32_t e2a5;  // FastLine.e2 | a5 << 16
    uint32_t a6b7;  // FastLine.a6 | b7 << 16
    uint32_t meta;
    int a, b;       // fast line, 1 hit: row start,end ; fast line, >= 2 hits: query start,end
    int x, y;       // fast line, 1 hit: lifted thickStart,thickEnd ; otherwise: first candidate block, candidates
};
__device__ __forceinline__ int st_kind(const LineState &st) { return (int)(st.meta & 3u); }
__device__ __forceinline__ int st_hits(const LineState &st) { return (int)((st.meta >> 2) & 3u); }


// Count side of one line: parse, look up, search, exact output lengths.  Returns the
// number of overlapping blocks, -1 for a line that is not a record, -2 for a malformed one.
// What a thread remembers from its previous line (256 lines earlier in the tile).  In sorted input
// the next line names the same contig and starts no earlier, so the contig lookup is a register
// compare and the block search a short gallop from the previous lower bound.
struct LineHint {
    uint32_t nw[4];  // first 16 name bytes
    uint32_t len0;   // name length; 0 = no hint
    int tcid, lo, start, contig_end;
};

// allow_cw: a heavy-fan-out line may be handed to the block (count work list); such a line is marked
// with both order bits (15 and 16) set and keeps its work-list slot in st.e2a5.
__device__ __forceinline__ int line_count(const TextArgs &A, const TileCtx &X, uint32_t p, LineState &st,
                                          uint32_t &len_l, uint32_t &len_u, const bool allow_cw, LineHint &H) {
    const TableView &T = A.T;
    len_l = len_u = 0;
    st.pq = p;
    st.e2a5 = st.a6b7 = 0;
    st.a = st.b = st.x = st.y = 0;
    st.meta = FL_NONE;
    FastLine F;
    uint32_t nw[4];
    const int kind = fast_parse(A, tx_win(), p, X.limit, F, nw);
    if (kind == FL_NONE) return -1;
    if (kind == FL_FAST) {
        TextSmem &S = tx_smem();
        Cand cd;
        Block b0{0, 0, 0, 0};
        {
            const bool same = H.len0 == F.len0 && F.len0 <= 16u && H.nw[0] == nw[0] && H.nw[1] == nw[1] &&
                              H.nw[2] == nw[2] && H.nw[3] == nw[3];
            F.tcid = same ? H.tcid : lookup_tcontig_win(A, tx_win(), p, F.len0, nw);
            if (same && F.tcid >= 0 && F.start >= H.start) {
                cd = find_candidates_from(T, H.contig_end, H.lo, F.start, F.end, b0);
            } else {
                cd = find_candidates_b0(T, tx_ey(), F.tcid, F.start, F.end, b0);
                H.contig_end = F.tcid >= 0 && F.tcid < T.ntc ? load_cmeta(T, F.tcid).y : 0;
            }
            H.nw[0] = nw[0], H.nw[1] = nw[1], H.nw[2] = nw[2], H.nw[3] = nw[3];
            H.len0 = F.len0;
            H.tcid = F.tcid;
            H.lo = cd.lo;
            H.start = F.start;
        }
        const int nh = count_hits(T, cd, F.start);
        st.pq = p | (F.q << 16);
        st.e2a5 = F.e2 | (F.a5 << 16);
        st.a6b7 = F.a6 | (F.b7 << 16);
        uint32_t meta = FL_FAST | ((uint32_t)(nh < 2 ? nh : 2) << 2) | ((uint32_t)(F.numf - 3) << 4) |
                        ((uint32_t)F.strand << 7);
        int slot = TX_CW_CAP;
        if (allow_cw && nh > 0 && cd.hi - cd.lo > TX_CW_MIN) slot = atomicAdd(&S.cw_n, 1);
        if (nh == 0) {
            len_u = F.q - p + 1;
        } else if (slot < TX_CW_CAP) {
            // heavy fan-out: the block sums this line's output length after the per-thread phase
            S.cw_p[slot] = p;
            S.cw_lo[slot] = cd.lo;
            S.cw_cnt[slot] = cd.hi - cd.lo;
            S.cw_tcid[slot] = F.tcid;
            st.e2a5 = (uint32_t)slot;
            meta |= 3u << 15;
            st.a = F.start;
            st.b = F.end;
            st.x = cd.lo;
            st.y = cd.hi - cd.lo;
            meta |= (uint32_t)(F.tcid < 0x7FFF ? F.tcid : 0x7FFF) << 17;
        } else {
            const uint32_t fixed = fast_fixed_len(F);
            Thick th{0, 0, 0};
            // thickStart / thickEnd normally lie inside the interval: gallop from its lower bound
            if (F.numf >= 8) th = lift_thick_from(T, tx_ey(), F.tcid, F.v6, F.v7, F.start, cd.lo, H.contig_end);
            RowCore r;
            r.key = r.s = r.e = r.qcid = 0;
            r.inv = 1;
            bool asc = true, desc = true, first = true;
            int prev_key = 0;
            for (int j = cd.lo; j < cd.hi; j++) {
                const Block b = j > cd.lo ? load_block(T, j) : b0;
                if (!T.disjoint && !block_hits(b, F.start)) continue;
                r = make_row(b, F.start, F.end);
                len_l += fast_row_len(A, F, fixed, th, r);
                if (!first) {  // are the rows already in output order (bed.d:156-157), or exactly reversed?
                    asc = asc && r.key >= prev_key;
                    desc = desc && r.key < prev_key;
                }
                first = false;
                prev_key = r.key;
            }
            if (nh == 1) {
                st.a = r.s;
                st.b = r.e;
                st.x = th.ts;
                st.y = th.te;
                meta |= (r.inv < 0 ? (1u << 15) : 0u) | (th.has ? (1u << 16) : 0u) | ((uint32_t)r.qcid << 17);
            } else {
                st.a = F.start;
                st.b = F.end;
                st.x = cd.lo;
                st.y = cd.hi - cd.lo;
                // bits 15/16: candidate order is ascending / strictly descending in qStart; 17-31: tcid
                meta |= (asc ? (1u << 15) : 0u) | (!asc && desc ? (1u << 16) : 0u) |
                        ((uint32_t)(F.tcid < 0x7FFF ? F.tcid : 0x7FFF) << 17);
            }
        }
        st.meta = meta;
        return nh;
    }
    SlowCount C;
    slow_count(A, tx_ey(), X.R, X.g0 + (p - X.s0), C);
    if (C.status == LN_ERROR) return -2;
    if (C.status == LN_SKIP) return -1;
    len_l = C.len_l;
    len_u = C.len_u;
    st.x = C.lo;
    st.y = C.ncand;
    st.meta = FL_SLOW | ((uint32_t)(C.nhits < 2 ? C.nhits : 2) << 2);
    return C.nhits;
}

__device__ __forceinline__ void state_to_fast(const LineState &st, FastLine &F) {
    F.numf = (int)((st.meta >> 4) & 7u) + 3;
    F.strand = (unsigned char)((st.meta >> 7) & 255u);
    F.q = st.pq >> 16;
    F.e2 = st.e2a5 & 0xFFFFu;
    F.a5 = st.e2a5 >> 16;
    F.a6 = st.a6b7 & 0xFFFFu;
    F.b7 = st.a6b7 >> 16;
    F.tcid = 0;
    F.start = st.a;
    F.end = st.b;
    F.v6 = F.v7 = 0;
}

// A fast line with a handful of hits in no particular order, expanded by its own thread:
// every row's sorted position (qStart, then block order: bed.d:156-157), byte offset and
// cumulative strand byte (bed.d:173) are computed directly.
__device__ __noinline__ void fast_expand_small(const TextArgs &A, const unsigned char *win, const FastLine &F, int tcid,
                                                  int lo, int ncand, char *dst) {
    const TableView &T = A.T;
    const uint32_t fixed = fast_fixed_len(F);
    Thick th{0, 0, 0};
    if (F.numf >= 8) {  // thickStart / thickEnd are re-read from the line (not kept between the phases)
        const uint32_t *W = reinterpret_cast<const uint32_t *>(win);
        uint32_t v6 = 0, v7 = 0, n6 = 0, n7 = 0, term;
        parse_uint_swar(W, F.a6, v6, n6, term);
        parse_uint_swar(W, F.a6 + n6 + 1, v7, n7, term);
        th = lift_thick_from(T, tx_ey(), tcid, (int)v6, (int)v7, F.start, lo, load_cmeta(T, tcid).y);
    }
    for (int c = 0; c < ncand; c++) {
        const Block bj = load_block(T, lo + c);
        if (!T.disjoint && !block_hits(bj, F.start)) continue;
        const RowCore rj = make_row(bj, F.start, F.end);
        int rank = 0, negs = rj.inv < 0 ? 1 : 0;
        uint32_t off = 0;
        int best_key = rj.key, best_k = c, best_inv = rj.inv;
        for (int k = 0; k < ncand; k++) {
            if (k == c) continue;
            const Block bk = load_block(T, lo + k);
            if (!T.disjoint && !block_hits(bk, F.start)) continue;
            const RowCore rk = make_row(bk, F.start, F.end);
            const bool before = rk.key < rj.key || (rk.key == rj.key && k < c);
            if (before) {
                rank++;
                negs += rk.inv < 0 ? 1 : 0;
                off += fast_row_len(A, F, fixed, th, rk);
            }
            if (rk.key < best_key || (rk.key == best_key && k < best_k)) {
                best_key = rk.key;
                best_k = k;
                best_inv = rk.inv;
            }
        }
        const unsigned char sb = F.numf >= 6 ? strand_at_rank(F.strand, best_inv, rank, negs) : 0;
        fast_emit_row(dst + off, A, win, F, th, rj, sb);
    }
}

// A fast line whose hits are already in output order (or exactly reversed, '-' chains):
// one sequential pass, the strand byte flipping cumulatively as the reference's in-place
// update does (bed.d:173).
__device__ __noinline__ void fast_expand_sorted(const TextArgs &A, const TileCtx &X, const FastLine &F, int tcid,
                                                   int lo, int ncand, bool reversed, char *dst) {
    const TableView &T = A.T;
    Thick th{0, 0, 0};
    if (F.numf >= 8) {  // thickStart / thickEnd are re-read from the line (not kept between the phases)
        const uint32_t *W = reinterpret_cast<const uint32_t *>(tx_win());
        uint32_t v6 = 0, v7 = 0, n6 = 0, n7 = 0, term;
        parse_uint_swar(W, F.a6, v6, n6, term);
        parse_uint_swar(W, F.a6 + n6 + 1, v7, n7, term);
        th = lift_thick_from(T, tx_ey(), tcid, (int)v6, (int)v7, F.start, lo, load_cmeta(T, tcid).y);
    }
    unsigned char sb = F.strand;
    char *d = dst;
    for (int c = 0; c < ncand; c++) {
        const Block b = load_block(T, reversed ? lo + ncand - 1 - c : lo + c);
        if (!T.disjoint && !block_hits(b, F.start)) continue;
        const RowCore r = make_row(b, F.start, F.end);
        if (F.numf >= 6) sb = strand_flip(sb, r.inv);
        d = fast_emit_row(d, A, tx_win(), F, th, r, sb);
    }
}

constexpr int TX_SMALL_FANOUT = 6;
constexpr int TX_SEQ_FANOUT = 48;  // sorted multi-hit lines up to this many candidates stay with their thread

// Emit side of one line.  dst_l / dst_u = where this line's lifted / unmatched text goes
// (nullptr: that stream is not written, e.g. the caller's buffer is too small).
__device__ __forceinline__ void line_emit(const TextArgs &A, const TileCtx &X, TextSmem &S, const LineState &st,
                                          char *dst_l, char *dst_u, uint32_t o_l) {
    const int kind = st_kind(st);
    if (kind == FL_NONE) return;
    const int nhc = st_hits(st);
    const uint32_t p = st.pq & 0xFFFFu;
    if (nhc == 0) {
        if (!dst_u) return;
        if (kind == FL_FAST) copy_span(dst_u, tx_win(), p, (st.pq >> 16) + 1);
        else slow_emit(A, tx_ey(), X.R, X.g0 + (p - X.s0), st.x, 0, dst_u);
        return;
    }
    if (!dst_l) return;
    if (kind == FL_FAST && nhc == 1) {
        FastLine F;
        state_to_fast(st, F);
        Thick th;
        th.has = (int)((st.meta >> 16) & 1u);
        th.ts = st.x;
        th.te = st.y;
        RowCore r;
        r.s = st.a;
        r.e = st.b;
        r.qcid = (int)(st.meta >> 17);
        r.inv = (st.meta >> 15) & 1u ? -1 : 1;
        r.key = 0;
        fast_emit_row(dst_l, A, tx_win(), F, th, r, strand_flip(F.strand, r.inv));
    } else if (kind == FL_FAST && (st.meta & (3u << 15)) && st.y <= TX_SEQ_FANOUT && (st.meta >> 17) != 0x7FFFu) {
        FastLine F;
        state_to_fast(st, F);
        fast_expand_sorted(A, X, F, (int)(st.meta >> 17), st.x, st.y, (st.meta >> 16) & 1u, dst_l);
    } else if (kind == FL_FAST && st.y <= TX_SMALL_FANOUT && (st.meta >> 17) != 0x7FFFu) {
        FastLine F;
        state_to_fast(st, F);
        fast_expand_small(A, tx_win(), F, (int)(st.meta >> 17), st.x, st.y, dst_l);
    } else if (nhc == 1) {
        slow_emit(A, tx_ey(), X.R, X.g0 + (p - X.s0), st.x, 1, dst_l);
    } else {
        const int slot = atomicAdd(&S.wl_n, 1);
        if (slot < TX_WL_CAP) {
            S.wl_p0[slot] = p;
            S.wl_lo[slot] = st.x;
            S.wl_cnt[slot] = st.y;
            S.wl_off[slot] = o_l;
        } else {
            expand_line(A, tx_ey(), X.R, X.g0 + (p - X.s0), st.x, st.y, dst_l, 0, 1);
        }
    }
}

// Out-of-line copies for the rare call sites (dense tiles beyond the cached batches, multi-round
// and direct emits): the hot loop keeps one inlined copy per cached batch and the kernel stays
// small enough for the instruction cache.
__device__ __noinline__ void line_emit_cold(const TextArgs &A, const TileCtx &X, TextSmem &S, const LineState st,
                                            char *dst_l, char *dst_u, uint32_t o_l) {
    line_emit(A, X, S, st, dst_l, dst_u, o_l);
}
__device__ __noinline__ int line_count_cold(const TextArgs &A, const TileCtx &X, uint32_t p, LineState &st,
                                            uint32_t &len_l, uint32_t &len_u) {
    LineHint H;
    H.len0 = 0;
    return line_count(A, X, p, st, len_l, len_u, false, H);
}

// Block-cooperative sum of the output lengths of the tile's heavy-fan-out lines (count work
// list): results in S.cw_len.  Every thread of the CTA calls it.
__device__ __noinline__ void coop_count_lines(const TextArgs &A, const TileCtx &X, TextSmem &S) {
    const TableView &T = A.T;
    const int tid = threadIdx.x;
    const int ncw = min(S.cw_n, TX_CW_CAP);
    for (int e = 0; e < ncw; e++) {
        FastLine F;
        uint32_t nw[4];
        fast_parse(A, tx_win(), S.cw_p[e], X.limit, F, nw);
        const uint32_t fixed = fast_fixed_len(F);
        Thick th{0, 0, 0};
        if (F.numf >= 8) th = lift_thick(T, tx_ey(), S.cw_tcid[e], F.v6, F.v7);
        uint32_t part = 0;
        const int lo = S.cw_lo[e], cnt = S.cw_cnt[e];
        for (int c = tid; c < cnt; c += TX_TPB) {
            const Block b = load_block(T, lo + c);
            if (!T.disjoint && !block_hits(b, F.start)) continue;
            part += fast_row_len(A, F, fixed, th, make_row(b, F.start, F.end));
        }
        converge_warp(A.full_mask);  // lanes leave the loop above at different times
        part = warp_sum(part);
        if (lane_id() == 0) S.red32[warp_id()] = part;
        __syncthreads();
        if (tid == 0) {
            uint32_t sum = 0;
            for (int w = 0; w < TX_NW; w++) sum += S.red32[w];
            S.cw_len[e] = sum;
        }
        __syncthreads();
    }
}

// NB independent exclusive block scans with one pair of barriers.
template <int NB>
__device__ __forceinline__ void block_scan_multi(const unsigned long long (&v)[NB], unsigned long long (&ex)[NB],
                                                 unsigned long long (&tot)[NB], unsigned long long *scratch) {
    unsigned long long inc[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) inc[b] = warp_inclusive_scan(v[b]);
    __syncthreads();
    if (lane_id() == 31) {
#pragma unroll
        for (int b = 0; b < NB; b++) scratch[b * TX_NW + warp_id()] = inc[b];
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; b++) {
        unsigned long long wbase = 0, t = 0;
#pragma unroll
        for (int w = 0; w < TX_NW; w++) {
            const unsigned long long s = scratch[b * TX_NW + w];
            if (w < warp_id()) wbase += s;
            t += s;
        }
        ex[b] = wbase + inc[b] - v[b];
        tot[b] = t;
    }
}

// Block-wide resolve of both look-back chains: 256 predecessors per round instead of 32.
// Every thread of the block calls it; returns the exclusive prefixes in S.base_l / S.base_u.
__device__ __forceinline__ void lookback_resolve2_block(TextSmem &S, uint64_t *desc_l, uint64_t *desc_u, int tile,
                                                        uint64_t agg_l, uint64_t agg_u) {
    uint64_t ex_l = 0, ex_u = 0;
    bool done_l = tile == 0, done_u = tile == 0;
    int look = tile - 1;
    while (!(done_l && done_u)) {
        const int idx = look - (int)threadIdx.x;
        uint64_t vl = LB_INC, vu = LB_INC;  // virtual tile before tile 0: inclusive prefix 0
        if (idx >= 0) {
            if (!done_l) {
                vl = ld_relaxed_u64(desc_l + idx);
                while ((vl & LB_STATE) == LB_EMPTY) {
                    __nanosleep(20);
                    vl = ld_relaxed_u64(desc_l + idx);
                }
            }
            if (!done_u) {
                vu = ld_relaxed_u64(desc_u + idx);
                while ((vu & LB_STATE) == LB_EMPTY) {
                    __nanosleep(20);
                    vu = ld_relaxed_u64(desc_u + idx);
                }
            }
        }
        {
            const unsigned ml = __ballot_sync(FULL, (vl & LB_STATE) == LB_INC);
            const int fl = ml ? (__ffs(ml) - 1) : 32;
            const uint64_t sl = warp_sum((lane_id() <= fl) ? (vl & LB_VALUE) : 0ull);
            const unsigned mu = __ballot_sync(FULL, (vu & LB_STATE) == LB_INC);
            const int fu = mu ? (__ffs(mu) - 1) : 32;
            const uint64_t su = warp_sum((lane_id() <= fu) ? (vu & LB_VALUE) : 0ull);
            __syncthreads();  // previous round's readers are done
            if (lane_id() == 0) {
                S.lb_part[warp_id()][0] = sl;
                S.lb_part[warp_id()][1] = su;
                S.lb_found[warp_id()][0] = ml != 0;
                S.lb_found[warp_id()][1] = mu != 0;
            }
            __syncthreads();
        }
        if (!done_l) {
            for (int w = 0; w < TX_NW; w++) {  // warps are ordered by distance from the tile
                ex_l += S.lb_part[w][0];
                if (S.lb_found[w][0]) {
                    done_l = true;
                    break;
                }
            }
        }
        if (!done_u) {
            for (int w = 0; w < TX_NW; w++) {
                ex_u += S.lb_part[w][1];
                if (S.lb_found[w][1]) {
                    done_u = true;
                    break;
                }
            }
        }
        look -= TX_TPB;
    }
    if (threadIdx.x == 0) {
        if (tile > 0) {
            st_relaxed_u64(desc_l + tile, LB_INC | (ex_l + agg_l));
            st_relaxed_u64(desc_u + tile, LB_INC | (ex_u + agg_u));
        }
        S.base_l = ex_l;
        S.base_u = ex_u;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TX_TPB, TX_MIN_CTAS) lift_bed_text_kernel(const __grid_constant__ TextArgs A) {
    TextSmem &S = *reinterpret_cast<TextSmem *>(tx_smem_raw);
    const int tid = threadIdx.x;
    {  // destination contig names -> shared memory when they fit (once per persistent CTA)
        const int nq = A.nq;
        const bool fits = nq <= TX_QN_MAX && A.qname_off[nq] <= (uint32_t)TX_QCHARS;
        if (fits) {
            for (int i = tid; i <= nq; i += TX_TPB) S.qoff[i] = A.qname_off[i];
            const uint32_t nc = A.qname_off[nq];
            for (uint32_t i = tid; i < nc; i += TX_TPB) S.qchars[i] = A.qname_chars[i];
        }
        if (tid == 0) {
            S.qoff_p = fits ? S.qoff : A.qname_off;
            S.qchars_p = fits ? S.qchars : A.qname_chars;
        }
    }
    const TableView &T = A.T;
    unsigned char *const win = reinterpret_cast<unsigned char *>(S.in_words);
    char *const stage_l = reinterpret_cast<char *>(S.stage_words);

    // upper search levels -> shared memory (once per persistent CTA)
    for (int i = tid; i < T.ey_n; i += TX_TPB) S.ey[i] = __ldg(T.ey + i);

    // The formatted tile stays in the stages while the NEXT tile is staged, parsed and
    // counted (none of which touches the stages); only then is its chained offset resolved
    // and the stage copied out.  Predecessor tiles therefore have a whole extra count phase
    // to publish their aggregates and the look-back practically never waits.
    if (tid == 0) {
        S.P.pend = 0;
        mbar_init(&S.in_bar, 1);
        mbar_init_fence();
    }
    unsigned in_uses = 0;  // completed phases of S.in_bar (block-uniform)
    auto flush = [&]() {  // block-uniform
        PendTile &P = S.P;
        if (!P.pend) return;
        uint64_t gb_l = P.gb_l, gb_u = P.gb_u;
        if (!P.have_base) {
            lookback_resolve2_block(S, A.desc_l, A.desc_u, P.tile, (uint64_t)P.tot_l, (uint64_t)P.tot_u);
            gb_l = S.base_l;
            gb_u = S.base_u;
        }
        if (P.tile == A.ntiles - 1 && tid == 0) {
            A.sizes[0] = gb_l + P.tot_l;
            A.sizes[1] = gb_u + P.tot_u;
        }
        if (!P.direct_l && P.tot_l && gb_l + P.tot_l <= A.lifted_cap) copy_out(A.lifted + gb_l, S.stage_words, P.tot_l);
        if (!P.direct_u && P.tot_u && gb_u + P.tot_u <= A.unm_cap) copy_out(A.unm + gb_u, S.stage_words, P.tot_u, P.u_off);
        __syncthreads();  // the stages may be overwritten now
        if (tid == 0) P.pend = 0;
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            S.tile = (int)atomicAdd(A.ticket, 1u);
            S.wl_n = 0;
            S.cw_n = 0;
        }
        __syncthreads();
        const int tile = S.tile;
        if (tile >= A.ntiles) break;
        const uint32_t tile_b = A.tile_bytes, seg = tile_b / TX_TPB;  // bytes a thread enumerates line starts of
        const uint64_t t0 = (uint64_t)tile * tile_b;

        // ---- stage the window [t0-PRE, t0+TILE+LOOK); everything else reads as '\n' ----
        TileCtx &X = S.X;
        const uint64_t g0 = tile == 0 ? 0 : t0 - TX_PRE;  // absolute offset of the first staged byte
        const uint32_t s0 = tile == 0 ? TX_PRE : 0;       // its window offset
        const uint64_t gend = min(A.in_len, t0 + tile_b + TX_LOOK);
        const uint32_t nst = (uint32_t)(gend - g0);
        const uint32_t wv = s0 + nst;  // window offsets < wv hold input bytes
        {
            // one 1-D bulk copy (TMA) by thread 0 brings the window in; the other threads write the padding and
            // ask for the window of the tile a grid's width ahead — about the tile some CTA draws when this one
            // is done — to be brought into L2, so that its bulk copy finds the text there
            const uint32_t nbulk = nst & ~15u;
            if (tid == 0) {
                fence_proxy_async();  // the previous tile's (generic-proxy) accesses to the window come first
                mbar_expect_tx(&S.in_bar, nbulk);
                if (nbulk) bulk_g2s(win + s0, A.in + g0, nbulk, &S.in_bar);
            }
            for (uint32_t i = nbulk + tid; i < nst; i += TX_TPB) win[s0 + i] = (unsigned char)A.in[g0 + i];
            for (uint32_t i = tid; i < s0; i += TX_TPB) win[i] = '\n';
            for (uint32_t i = wv + tid; i < TX_WIN + TX_PAD; i += TX_TPB) win[i] = '\n';
            const uint64_t pf = t0 + (uint64_t)gridDim.x * tile_b + (uint64_t)tid * 128u;
            if ((uint32_t)tid * 128u < tile_b + TX_LOOK && pf < A.in_len) prefetch_l2(A.in + pf);
        }
        if (tid == 0) {
            X.g0 = g0;
            X.s0 = s0;
            X.R.win = win + s0;
            X.R.ws = g0;
            X.R.wlen = nst;
            X.R.g = reinterpret_cast<const unsigned char *>(A.in);
            X.R.in_len = A.in_len;
            // a terminator found at or beyond this window offset is not the line's real end
            X.limit = gend == A.in_len ? wv + 1 : wv;
        }
        mbar_wait(&S.in_bar, in_uses & 1u);
        in_uses++;
        __syncthreads();

        // ---- newline masks of the window, 64 bytes per word -----------------------------
        for (int sg = tid; sg < (int)((TX_PRE + tile_b + TX_LOOK) / TX_MSEG); sg += TX_TPB) {
            const uint4 *w = reinterpret_cast<const uint4 *>(win + sg * TX_MSEG);
            unsigned long long m = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint4 x = w[k];
                const uint32_t m16 = nl_mask4(x.x) | (nl_mask4(x.y) << 4) | (nl_mask4(x.z) << 8) | (nl_mask4(x.w) << 12);
                m |= (unsigned long long)m16 << (16 * k);
            }
            S.nlm[sg] = m;
        }
        __syncthreads();
        // line starts inside this thread's 64 bytes of the tile: the byte before is '\n'
        unsigned long long starts;
        {
            const int w = (int)((tid * seg) / TX_MSEG);  // mask word of this thread's bytes (window word w + 1)
            starts = (S.nlm[w + 1] << 1) | (S.nlm[w] >> 63);
            if (seg < TX_MSEG) starts = (starts >> ((tid * seg) % TX_MSEG)) & ((1ull << seg) - 1ull);
        }
        if (tile == 0 && tid == 0)  // the text starts (with a line) at offset `skip`
            starts = (starts & ~((1ull << A.skip) - 1ull)) | (1ull << A.skip);
        unsigned long long valid = starts;
        {
            const long long rem = (long long)A.in_len - (long long)(t0 + (uint64_t)tid * seg);
            if (rem <= 0) valid = 0;
            else if (rem < (long long)seg) valid &= (1ull << rem) - 1ull;
        }
        // line index = prefix over ALL starts (the valid ones are a prefix of them)
        uint32_t packed_tot;
        const uint32_t packed_ex = block_exclusive_scan<uint32_t, TX_NW>(
            (uint32_t)__popcll(starts) | ((uint32_t)__popcll(valid) << 16), S.scan32, &packed_tot);
        const int my_first = (int)(packed_ex & 0xFFFFu);
        const int nlines = (int)(packed_tot >> 16);  // lines owned by this tile
        auto build_table = [&](int wbase) {
            __syncthreads();  // previous users of the table are done
            int j = my_first;
            for (unsigned long long m = starts; m; m &= m - 1, j++) {
                const int slot = j - wbase;
                if (slot >= 0 && slot < TX_LCAP)
                    S.lstart[slot] = (uint16_t)(TX_PRE + tid * seg + __ffsll((long long)m) - 1);
            }
            __syncthreads();
        };
        build_table(0);

        // ---- phase C: parse + search + exact output lengths; no barrier inside ---------
        uint32_t c_rec = 0, c_rows = 0, c_unm = 0;
        LineHint H;
        H.len0 = 0;
        LineState st[TX_NB];
        unsigned long long v[TX_NB], ex[TX_NB], bt[TX_NB];
        // one copy of the (large) per-line code: the batch loop is not unrolled, the cached states
        // rotate through st[] so that every index stays static (registers)
#pragma unroll
        for (int b = 0; b < TX_NB; b++) st[b] = LineState{};
#pragma unroll 1
        for (int b = 0; b < TX_NB; b++) {
            const int li = b * TX_TPB + tid;
            LineState cur = LineState{};
            cur.meta = FL_NONE;
            unsigned long long cv = 0;
            if (li < nlines) {
                uint32_t ll, lu;
                const int nh = line_count(A, X, S.lstart[li], cur, ll, lu, true, H);
                cv = ((unsigned long long)ll << 32) | lu;
                if (nh == -2)  // max(~p0) == first bad line
                    atomicMax((unsigned long long *)(A.sizes + 5),
                              ~(unsigned long long)(X.g0 + ((cur.pq & 0xFFFFu) - X.s0)));
                if (nh >= 0) {
                    c_rec++;
                    c_rows += (uint32_t)nh;
                    c_unm += nh == 0 ? 1u : 0u;
                }
            }
#pragma unroll
            for (int k = 0; k + 1 < TX_NB; k++) {
                st[k] = st[k + 1];
                v[k] = v[k + 1];
            }
            st[TX_NB - 1] = cur;
            v[TX_NB - 1] = cv;
        }
        converge_warp(A.full_mask);  // the count phase is long divergent per-thread work
        block_scan_multi<TX_NB>(v, ex, bt, S.scan64);
        // lines with heavy fan-out: the block sums their output lengths (a single thread would
        // hold the whole tile, and every successor's look-back, for hundreds of microseconds),
        // then the tile is scanned again.  S.cw_n is stable here: the scan above had barriers.
        if (S.cw_n > 0) {
            coop_count_lines(A, X, S);
#pragma unroll
            for (int b = 0; b < TX_NB; b++)
                if ((st[b].meta & ((3u << 15) | 15u)) == ((3u << 15) | (2u << 2) | FL_FAST))
                    v[b] += (unsigned long long)S.cw_len[st[b].e2a5] << 32;
            block_scan_multi<TX_NB>(v, ex, bt, S.scan64);
        }
        uint32_t run_l = 0, run_u = 0;
        uint32_t len_l[TX_NB];  // lifted bytes of each cached line (the multi-round emit places lines by it)
#pragma unroll
        for (int b = 0; b < TX_NB; b++) {
            ex[b] += ((unsigned long long)run_l << 32) | run_u;
            run_l += (uint32_t)(bt[b] >> 32);
            run_u += (uint32_t)bt[b];
            len_l[b] = (uint32_t)(v[b] >> 32);
        }
        const uint32_t cached_l = run_l, cached_u = run_u;  // output bytes of the cached batches
        // tiles with more than TX_NB*256 lines (very short lines): count the rest now, redo them when emitting
        if (nlines > TX_NB * TX_TPB) {
            int wbase = 0;
            for (int lb = TX_NB * TX_TPB; lb < nlines; lb += TX_TPB) {
                if (lb >= wbase + TX_LCAP) {
                    wbase += TX_LCAP;
                    build_table(wbase);
                }
                const int li = lb + tid;
                uint32_t ll = 0, lu = 0;
                if (li < nlines) {
                    LineState t;
                    const int nh = line_count_cold(A, X, S.lstart[li - wbase], t, ll, lu);
                    if (nh == -2)
                        atomicMax((unsigned long long *)(A.sizes + 5),
                                  ~(unsigned long long)(X.g0 + ((t.pq & 0xFFFFu) - X.s0)));
                    if (nh >= 0) {
                        c_rec++;
                        c_rows += (uint32_t)nh;
                        c_unm += nh == 0 ? 1u : 0u;
                    }
                }
                unsigned long long t1;
                block_exclusive_scan<unsigned long long, TX_NW>(((unsigned long long)ll << 32) | lu, S.scan64, &t1);
                run_l += (uint32_t)(t1 >> 32);
                run_u += (uint32_t)t1;
            }
            if (wbase) build_table(0);
        }
        const uint32_t tot_l = run_l, tot_u = run_u;

        // ---- publish the tile aggregates now; successors never wait on our formatting ---
        if (warp_id() == 0) lookback_publish(A.desc_l, tile, (uint64_t)tot_l);
        else if (warp_id() == 1) lookback_publish(A.desc_u, tile, (uint64_t)tot_u);
        {  // record counters
            const uint32_t rr = warp_sum(c_rec), w = warp_sum(c_rows), u = warp_sum(c_unm);
            if (lane_id() == 0) {
                if (rr) atomicAdd((unsigned long long *)(A.sizes + 2), (unsigned long long)rr);
                if (w) atomicAdd((unsigned long long *)(A.sizes + 3), (unsigned long long)w);
                if (u) atomicAdd((unsigned long long *)(A.sizes + 4), (unsigned long long)u);
            }
        }
        // a stream that does not fit its stage is written straight to global memory,
        // which needs the chained offset before formatting
        const bool direct_u = tot_u > TX_UNM_STAGE;
        const uint32_t u_off = direct_u ? (uint32_t)TX_STAGE_ALL : (uint32_t)TX_STAGE_ALL - ((tot_u + 15u) & ~15u);  // >= TX_OUT_STAGE
        const bool direct_l = tot_l > u_off;
        char *const stage_u = stage_l + u_off;
        uint64_t gb_l = 0, gb_u = 0;
        bool have_base = false;
        flush();  // previous tile: resolve + copy out, frees the stages
        auto resolve_now = [&]() {  // block-uniform
            lookback_resolve2_block(S, A.desc_l, A.desc_u, tile, (uint64_t)tot_l, (uint64_t)tot_u);
            gb_l = S.base_l;
            gb_u = S.base_u;
            have_base = true;
        };
        // an oversized unmatched stream (lines longer than the look-ahead only) needs its offset now; an
        // oversized lifted stream resolves after its first round has been formatted (see below)
        if (direct_u) resolve_now();
        char *const base_u = direct_u ? (gb_u + tot_u <= A.unm_cap ? A.unm + gb_u : nullptr) : stage_u;

        auto run_expansions = [&](char *base) {  // cooperative expansion of the queued multi-hit lines (block-uniform)
            converge_warp(A.full_mask);  // behind the emit phase
            __syncthreads();
            const int nwl = min(S.wl_n, TX_WL_CAP);
            if (nwl == 0) return;
            for (int e = warp_id(); e < nwl; e += TX_NW)
                if (S.wl_cnt[e] <= TX_HEAVY_MIN)
                    expand_line_coop<false>(A, tx_ey(), X.R, X.g0 + (S.wl_p0[e] - X.s0), S.wl_lo[e], S.wl_cnt[e],
                                            base + S.wl_off[e], S.xscr + warp_id() * TX_HEAVY_MIN,
                                            S.xaux + warp_id() * TX_HEAVY_MIN, nullptr, lane_id(), 32);
            __syncthreads();
            for (int e = 0; e < nwl; e++) {
                const int cnt = S.wl_cnt[e];
                if (cnt <= TX_HEAVY_MIN) continue;
                const uint64_t p0 = X.g0 + (S.wl_p0[e] - X.s0);
                if (cnt <= TX_XSCR) {
                    expand_line_coop<true>(A, tx_ey(), X.R, p0, S.wl_lo[e], cnt, base + S.wl_off[e], S.xscr, S.xaux, S.scan64,
                                           tid, TX_TPB);
                } else if (cnt <= TX_GSCR && A.gscr) {
                    unsigned long long *g = A.gscr + (size_t)blockIdx.x * (TX_GSCR + TX_GSCR / 2);
                    expand_line_coop<true>(A, tx_ey(), X.R, p0, S.wl_lo[e], cnt, base + S.wl_off[e], g,
                                           reinterpret_cast<uint32_t *>(g + TX_GSCR), S.scan64, tid, TX_TPB);
                } else {
                    expand_line(A, tx_ey(), X.R, p0, S.wl_lo[e], cnt, base + S.wl_off[e], tid, TX_TPB);
                }
            }
            __syncthreads();
        };
        // lines beyond the cached batches (very dense tiles): re-counted and formatted batch by batch
        auto emit_uncached = [&](char *base_l) {
            if (nlines <= TX_NB * TX_TPB) return;
            int wbase = 0;
            build_table(0);  // the expansions of a multi-round emit reuse the table's memory
            uint32_t r_l = cached_l, r_u = cached_u;
            for (int lb = TX_NB * TX_TPB; lb < nlines; lb += TX_TPB) {
                if (lb >= wbase + TX_LCAP) {
                    wbase += TX_LCAP;
                    build_table(wbase);
                }
                const int li = lb + tid;
                uint32_t ll = 0, lu = 0;
                LineState t;
                t.meta = FL_NONE;
                if (li < nlines) line_count_cold(A, X, S.lstart[li - wbase], t, ll, lu);
                unsigned long long t1;
                const unsigned long long e1 =
                    block_exclusive_scan<unsigned long long, TX_NW>(((unsigned long long)ll << 32) | lu, S.scan64, &t1);
                const uint32_t o_l = r_l + (uint32_t)(e1 >> 32), o_u = r_u + (uint32_t)e1;
                r_l += (uint32_t)(t1 >> 32);
                r_u += (uint32_t)t1;
                line_emit_cold(A, X, S, t, base_l ? base_l + o_l : nullptr, base_u ? base_u + o_u : nullptr, o_l);
            }
        };

        if (!direct_l) {
            // ---- phase E: format from the cached state into the stages; no barrier inside ---
#pragma unroll 1
            for (int b = 0; b < TX_NB; b++) {  // same rotation: one copy of line_emit
                const uint32_t o_l = (uint32_t)(ex[0] >> 32), o_u = (uint32_t)ex[0];
                line_emit(A, X, S, st[0], stage_l + o_l, base_u ? base_u + o_u : nullptr, o_l);
                const LineState t = st[0];
                const unsigned long long te = ex[0];
#pragma unroll
                for (int k = 0; k + 1 < TX_NB; k++) {
                    st[k] = st[k + 1];
                    ex[k] = ex[k + 1];
                }
                st[TX_NB - 1] = t;
                ex[TX_NB - 1] = te;
            }
            emit_uncached(stage_l);
            run_expansions(stage_l);
        } else {
            // ---- large tile (fan-out): the lifted text goes out in rounds of TX_OUT_STAGE -
            // TX_ROUND_SLACK bytes through the stage.  A line belongs to the round its first row
            // starts in; lines longer than the slack, and the uncached ones, are written to
            // global memory directly AFTER the rounds (a round's copy-out may cover their bytes).
#pragma unroll
            for (int b = 0; b < TX_NB; b++)  // unmatched stream first
                line_emit_cold(A, X, S, st[b], nullptr, base_u ? base_u + (uint32_t)ex[b] : nullptr, 0);
            {
                constexpr uint32_t W = TX_OUT_STAGE - TX_ROUND_SLACK;
                const uint32_t cached_end = cached_l;
                for (uint32_t r0 = 0; r0 < cached_end; r0 += W) {
                    bool any = false;
#pragma unroll
                    for (int b = 0; b < TX_NB; b++) {
                        const uint32_t o_l = (uint32_t)(ex[b] >> 32), len = len_l[b];
                        any = any || (len > 0 && len <= TX_ROUND_SLACK && o_l >= r0 && o_l - r0 < W);
                    }
                    if (tid == 0) {
                        S.rlo = 0xFFFFFFFFu;
                        S.rhi = 0;
                        S.wl_n = 0;
                    }
                    converge_warp(A.full_mask);  // behind line_emit_cold / the previous round
                    if (!__syncthreads_or(any)) continue;
#pragma unroll
                    for (int b = 0; b < TX_NB; b++) {
                        const uint32_t o_l = (uint32_t)(ex[b] >> 32), len = len_l[b];
                        if (len > 0 && len <= TX_ROUND_SLACK && o_l >= r0 && o_l - r0 < W) {
                            line_emit_cold(A, X, S, st[b], stage_l + (o_l - r0), nullptr, o_l - r0);
                            atomicMin(&S.rlo, o_l);
                            atomicMax(&S.rhi, o_l + len);
                        }
                    }
                    run_expansions(stage_l);
                    // the chained offset is needed only now: the formatting of the first round gave the
                    // predecessors time to publish their aggregates
                    if (!have_base) resolve_now();
                    if (gb_l + tot_l <= A.lifted_cap)
                        copy_out(A.lifted + gb_l + S.rlo, S.stage_words, S.rhi - S.rlo, S.rlo - r0);
                    __syncthreads();
                }
                if (!have_base) resolve_now();
                char *const glob_l = gb_l + tot_l <= A.lifted_cap ? A.lifted + gb_l : nullptr;
                if (tid == 0) S.wl_n = 0;
                converge_warp(A.full_mask);
                __syncthreads();
                if (glob_l) {
#pragma unroll
                    for (int b = 0; b < TX_NB; b++) {
                        const uint32_t o_l = (uint32_t)(ex[b] >> 32), len = len_l[b];
                        if (len > TX_ROUND_SLACK) line_emit_cold(A, X, S, st[b], glob_l + o_l, nullptr, o_l);
                    }
                }
                emit_uncached(glob_l);
                if (glob_l) run_expansions(glob_l);
            }
        }
        // resolve + copy-out of this tile are deferred until the next tile has been counted
        converge_warp(A.full_mask);
        __syncthreads();  // (also orders this tile's last reads of S.P before the update)
        if (tid == 0) {
            PendTile &P = S.P;
            P.pend = 1;
            P.have_base = have_base;
            P.direct_l = direct_l;
            P.direct_u = direct_u;
            P.tile = tile;
            P.tot_l = tot_l;
            P.tot_u = tot_u;
            P.u_off = u_off;
            P.gb_l = gb_l;
            P.gb_u = gb_u;
        }
    }
    __syncthreads();
    flush();
}

// sizes of a finished chunk -> pinned host memory (stream-ordered after the lift kernel)
__global__ void copy_sizes_kernel(const uint64_t *__restrict__ d_sizes, uint64_t *__restrict__ h_sizes) {
    if (threadIdx.x < 6) h_sizes[threadIdx.x] = d_sizes[threadIdx.x];
}

}  // namespace swo
