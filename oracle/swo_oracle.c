/*
 * swo_oracle.c — CPU ORACLE (test infrastructure; see swo_oracle.h header).
 * Plain C99 + pthreads.  Every function cites the reference lines it follows
 * (paths relative to /root/reference).
 */
#define _GNU_SOURCE
#include "swo_oracle.h"

#include <limits.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ utils */

static void set_err(char *err, size_t cap, const char *fmt, ...) {
    if (!err || !cap) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, cap, fmt, ap);
    va_end(ap);
}

/* ASCII subset of std.uni.isWhite used by the no-arg splitter (bed.d:119,
 * chain.d:347).  Non-ASCII white space is out of scope (BED/chain are ASCII). */
static inline int is_white(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0D); }

/* std.conv.to!int on a token: [+-]?[0-9]+, whole token, must fit int32
 * (bed.d:125-126,135-136; chain.d:298-327,347). */
static int parse_int32(const char *s, size_t n, int32_t *out) {
    size_t i = 0;
    int neg = 0;
    if (n == 0) return -1;
    if (s[0] == '+' || s[0] == '-') {
        neg = s[0] == '-';
        i = 1;
    }
    if (i == n) return -1;
    int64_t v = 0;
    for (; i < n; i++) {
        unsigned d = (unsigned char)s[i] - '0';
        if (d > 9) return -1;
        v = v * 10 + d;
        if (v > (int64_t)INT32_MAX + 1) return -1;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return -1;
    *out = (int32_t)v;
    return 0;
}

static int buf_reserve(orc_buf *b, size_t extra) {
    if (b->len + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 4096;
    while (nc < b->len + extra) nc *= 2;
    char *p = (char *)realloc(b->data, nc);
    if (!p) return -1;
    b->data = p;
    b->cap = nc;
    return 0;
}
static inline int buf_put(orc_buf *b, const char *s, size_t n) {
    if (buf_reserve(b, n)) return -1;
    memcpy(b->data + b->len, s, n);
    b->len += n;
    return 0;
}
static inline int buf_putc(orc_buf *b, char c) { return buf_put(b, &c, 1); }
/* decimal text of a long (toChars, bed.d:169-170) */
static inline int buf_put_i64(orc_buf *b, int64_t v) {
    char tmp[24];
    int n = 0;
    uint64_t u = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    do {
        tmp[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) tmp[n++] = '-';
    if (buf_reserve(b, (size_t)n)) return -1;
    while (n) b->data[b->len++] = tmp[--n];
    return 0;
}
void orc_buf_free(orc_buf *b) {
    free(b->data);
    b->data = NULL;
    b->len = b->cap = 0;
}

/* ------------------------------------------------------- ChainLink helpers */

/* chain.d:128-134 */
int64_t orc_link_qend(const orc_link *l) { return l->qStart + l->invert * (l->tEnd - l->tStart); }
/* chain.d:141-144 — note: the reference returns invert*(tEnd-tStart); its own
 * unittest (chain.d:219-220) expects 1000 for an invert=-1 link, which the
 * code as written does not give (-1000).  The unittest is stale;
 * the oracle follows the CODE and the KAT test checks |size|. */
int64_t orc_link_size(const orc_link *l) { return l->invert * (l->tEnd - l->tStart); }

/* chain.d:147-164 */
int64_t orc_link_cmp(const orc_link *a, const orc_link *b) {
    if (a->tStart < b->tStart) return -1;
    if (a->tStart > b->tStart) return 1;
    if (a->tEnd < b->tEnd) return -1;
    if (a->tEnd > b->tEnd) return 1;
    return a->qStart - b->qStart;
}
/* chain.d:166-171 */
int orc_link_cmp_scalar(const orc_link *a, int64_t x) {
    if (a->tStart < x) return -1;
    if (a->tStart > x) return 1;
    return 0;
}

/* chain.d:728-743 */
orc_link orc_intersect(orc_link link, int64_t start, int64_t end) {
    int64_t ts = link.tStart > start ? link.tStart : start;
    int64_t te = link.tEnd < end ? link.tEnd : end;
    int64_t delta = ts - link.tStart;
    link.qStart += link.invert * delta;
    link.tStart = ts;
    link.tEnd = te;
    return link;
}

/* chain.d:787-793 */
void orc_order_start_end(int64_t *start, int64_t *end) {
    int64_t s = *start, e = *end;
    *start = s < e ? s : e;
    *end = s > e ? s : e;
}

/* bed.d:34-52,173: STRAND_TABLE[c + invert]; entries 0x2A->'-', 0x2C->'+',
 * 0x2E->'-', everything else 33 ('!').  Index 256 (c=0xFF,+1) is past the
 * table in the reference; defined here as '!'. */
unsigned char orc_strand_flip(unsigned char c, int invert) {
    int idx = (int)c + invert;
    if (idx == 0x2A || idx == 0x2E) return '-';
    if (idx == 0x2C) return '+';
    return '!';
}

/* ------------------------------------------------------------- chain file */

typedef struct {
    orc_link l;
    uint64_t ord; /* insertion ordinal (tie-break for stable sort / dedup) */
} link_ord;

typedef struct {
    char *name;
    size_t name_len;
    link_ord *raw; /* insertion order, during build */
    size_t n_raw, cap_raw;
    /* final, sorted by (tStart,tEnd,qStart), deduplicated */
    size_t n;
    int64_t *tStart, *tEnd, *qStart, *pmax;
    int32_t *qcid;
    int8_t *invert;
} tcontig;

typedef struct {
    char *name;
    int64_t size;
} qsize_ent;

struct orc_chainfile {
    tcontig *tc;
    int ntc, captc;
    char **qnames;
    int nq, capq;
    qsize_ent *qsizes;
    int nqs, capqs;
    /* open-addressing name -> tcid */
    int32_t *hash;
    uint32_t hmask;
    uint64_t ord;
};

static uint32_t fnv1a(const char *s, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) {
        h ^= (unsigned char)s[i];
        h *= 16777619u;
    }
    return h;
}

static char *dup_str(const char *s, size_t n) {
    char *p = (char *)malloc(n + 1);
    if (p) {
        memcpy(p, s, n);
        p[n] = 0;
    }
    return p;
}

static int tc_find_linear(orc_chainfile *cf, const char *s, size_t n) {
    for (int i = 0; i < cf->ntc; i++)
        if (cf->tc[i].name_len == n && memcmp(cf->tc[i].name, s, n) == 0) return i;
    return -1;
}
/* chainsByContig.require(tName, new tree)  chain.d:537 */
static int tc_require(orc_chainfile *cf, const char *s, size_t n) {
    int i = tc_find_linear(cf, s, n);
    if (i >= 0) return i;
    if (cf->ntc == cf->captc) {
        cf->captc = cf->captc ? cf->captc * 2 : 64;
        cf->tc = (tcontig *)realloc(cf->tc, sizeof(tcontig) * (size_t)cf->captc);
    }
    tcontig *t = &cf->tc[cf->ntc];
    memset(t, 0, sizeof *t);
    t->name = dup_str(s, n);
    t->name_len = n;
    return cf->ntc++;
}
/* ContigIDorAdd  chain.d:799-805 */
static int q_id_or_add(orc_chainfile *cf, const char *s, size_t n) {
    for (int i = 0; i < cf->nq; i++)
        if (strlen(cf->qnames[i]) == n && memcmp(cf->qnames[i], s, n) == 0) return i;
    if (cf->nq == cf->capq) {
        cf->capq = cf->capq ? cf->capq * 2 : 64;
        cf->qnames = (char **)realloc(cf->qnames, sizeof(char *) * (size_t)cf->capq);
    }
    cf->qnames[cf->nq] = dup_str(s, n);
    return cf->nq++;
}
/* qContigSizes[qName] = qSize  chain.d:551,579 (last write wins) */
static void qsize_set(orc_chainfile *cf, const char *s, size_t n, int64_t size) {
    for (int i = 0; i < cf->nqs; i++)
        if (strlen(cf->qsizes[i].name) == n && memcmp(cf->qsizes[i].name, s, n) == 0) {
            cf->qsizes[i].size = size;
            return;
        }
    if (cf->nqs == cf->capqs) {
        cf->capqs = cf->capqs ? cf->capqs * 2 : 64;
        cf->qsizes = (qsize_ent *)realloc(cf->qsizes, sizeof(qsize_ent) * (size_t)cf->capqs);
    }
    cf->qsizes[cf->nqs].name = dup_str(s, n);
    cf->qsizes[cf->nqs].size = size;
    cf->nqs++;
}

typedef struct {
    const char *p;
    size_t n;
} slice;

/* One chain = header line + data lines (chain.d:272-405).  `emit` receives
 * every link in file order.  Returns number of links or <0. */
typedef struct {
    slice tName, qName;
    int32_t tSize, qSize;
    int invert;
} chain_hdr;

typedef int (*link_cb)(void *ctx, const chain_hdr *h, const orc_link *l);

static int64_t parse_one_chain(const slice *lines, size_t nlines, link_cb cb, void *ctx,
                               chain_hdr *hdr_out, char *err, size_t errcap) {
    if (nlines == 0) {
        set_err(err, errcap, "empty chain");
        return ORC_E_PARSE;
    }
    /* header: split on single spaces  chain.d:293 */
    slice f[16];
    int nf = 0;
    {
        const char *p = lines[0].p;
        size_t n = lines[0].n;
        if (n && p[n - 1] == '\r') n--; /* tolerated; the reference would throw */
        size_t a = 0;
        for (size_t i = 0; i <= n; i++) {
            if (i == n || p[i] == ' ') {
                if (nf < 16) {
                    f[nf].p = p + a;
                    f[nf].n = i - a;
                }
                nf++;
                a = i + 1;
            }
        }
    }
    if (nf < 13 || f[0].n != 5 || memcmp(f[0].p, "chain", 5) != 0) {
        set_err(err, errcap, "malformed chain header: %.*s", (int)(lines[0].n > 80 ? 80 : lines[0].n),
                lines[0].p);
        return ORC_E_PARSE;
    }
    chain_hdr h;
    int32_t tS, tE, qS, qE, id;
    h.tName = f[2];
    h.qName = f[7];
    if (parse_int32(f[3].p, f[3].n, &h.tSize) || parse_int32(f[5].p, f[5].n, &tS) ||
        parse_int32(f[6].p, f[6].n, &tE) || parse_int32(f[8].p, f[8].n, &h.qSize) ||
        parse_int32(f[10].p, f[10].n, &qS) || parse_int32(f[11].p, f[11].n, &qE) ||
        parse_int32(f[12].p, f[12].n, &id) || f[4].n < 1 || f[9].n < 1) {
        set_err(err, errcap, "malformed chain header numbers: %.*s",
                (int)(lines[0].n > 80 ? 80 : lines[0].n), lines[0].p);
        return ORC_E_PARSE;
    }
    char tStrand = f[4].p[0], qStrand = f[9].p[0];
    /* chain.d:300-325 */
    int32_t targetStart = (tStrand == '+') ? tS : h.tSize - tS;
    int32_t queryStart = (qStrand == '+') ? qS : h.qSize - qS;
    (void)tE;
    (void)qE;
    h.invert = (tStrand == qStrand) ? 1 : -1; /* chain.d:330 */
    if (hdr_out) *hdr_out = h;

    int32_t tFrom = targetStart, qFrom = queryStart; /* chain.d:339-340 (int) */
    int64_t nlinks = 0;
    for (size_t li = 1; li < nlines; li++) {
        const char *p = lines[li].p;
        size_t n = lines[li].n;
        if (n == 0) continue; /* chain.d:344 */
        int32_t d[3];
        int nd = 0;
        size_t i = 0;
        int over = 0;
        while (i < n) {
            while (i < n && is_white((unsigned char)p[i])) i++;
            if (i >= n) break;
            size_t a = i;
            while (i < n && !is_white((unsigned char)p[i])) i++;
            if (nd < 3) {
                if (parse_int32(p + a, i - a, &d[nd])) {
                    set_err(err, errcap, "bad integer in chain data line: %.*s", (int)(n > 60 ? 60 : n), p);
                    return ORC_E_PARSE;
                }
                nd++;
            } else
                over = 1;
        }
        if (nd == 0) continue; /* white-space-only line: defined as blank */
        if (over || nd == 2) { /* chain.d:403 assert(0) */
            set_err(err, errcap, "unexpected length of alignment data line: %.*s", (int)(n > 60 ? 60 : n), p);
            return ORC_E_PARSE;
        }
        orc_link l;
        l.tStart = tFrom; /* chain.d:358-369 */
        l.tEnd = (int64_t)tFrom + d[0];
        l.qStart = qFrom;
        l.qcid = -1;
        l.invert = h.invert;
        if (cb) {
            int rc = cb(ctx, &h, &l);
            if (rc) return rc;
        }
        nlinks++;
        if (nd == 3) { /* chain.d:397-401 (int arithmetic) */
            tFrom += d[0] + d[1];
            qFrom += h.invert * (d[0] + d[2]);
        }
    }
    return nlinks;
}

static size_t split_lines(const char *text, size_t len, slice **out) {
    /* File.byLineCopy: '\n' terminated, terminator dropped, a final
     * unterminated line is yielded, no empty line after a final '\n'. */
    size_t cap = 1024, n = 0;
    slice *v = (slice *)malloc(cap * sizeof(slice));
    size_t a = 0;
    while (a < len) {
        const char *nl = (const char *)memchr(text + a, '\n', len - a);
        size_t e = nl ? (size_t)(nl - text) : len;
        if (n == cap) {
            cap *= 2;
            v = (slice *)realloc(v, cap * sizeof(slice));
        }
        v[n].p = text + a;
        v[n].n = e - a;
        n++;
        a = e + 1;
    }
    *out = v;
    return n;
}

int64_t orc_chain_count_links(const char *text, size_t len) {
    slice *lines;
    size_t n = split_lines(text, len, &lines);
    char err[128];
    int64_t r = parse_one_chain(lines, n, NULL, NULL, NULL, err, sizeof err);
    free(lines);
    return r;
}

static int add_link_cb(void *ctx, const chain_hdr *h, const orc_link *l) {
    orc_chainfile *cf = (orc_chainfile *)ctx;
    /* ContigIDorAdd per link  chain.d:364 */
    int qcid = q_id_or_add(cf, h->qName.p, h->qName.n);
    /* tree.insert(*link) into the target contig's tree  chain.d:537-548 */
    int t = tc_require(cf, h->tName.p, h->tName.n);
    tcontig *tc = &cf->tc[t];
    if (tc->n_raw == tc->cap_raw) {
        tc->cap_raw = tc->cap_raw ? tc->cap_raw * 2 : 256;
        tc->raw = (link_ord *)realloc(tc->raw, tc->cap_raw * sizeof(link_ord));
        if (!tc->raw) return ORC_E_OOM;
    }
    tc->raw[tc->n_raw].l = *l;
    tc->raw[tc->n_raw].l.qcid = qcid;
    tc->raw[tc->n_raw].ord = cf->ord++;
    tc->n_raw++;
    return 0;
}

static int cmp_link_ord(const void *a, const void *b) {
    const link_ord *x = (const link_ord *)a, *y = (const link_ord *)b;
    int64_t c = orc_link_cmp(&x->l, &y->l);
    if (c < 0) return -1;
    if (c > 0) return 1;
    return x->ord < y->ord ? -1 : (x->ord > y->ord ? 1 : 0);
}

static int finalize(orc_chainfile *cf) {
    for (int t = 0; t < cf->ntc; t++) {
        tcontig *tc = &cf->tc[t];
        qsort(tc->raw, tc->n_raw, sizeof(link_ord), cmp_link_ord);
        /* insert() of an opCmp-equal link returns the existing node: keep first */
        size_t m = 0;
        for (size_t i = 0; i < tc->n_raw; i++) {
            if (m && orc_link_cmp(&tc->raw[m - 1].l, &tc->raw[i].l) == 0) continue;
            tc->raw[m++] = tc->raw[i];
        }
        tc->n = m;
        size_t a = m ? m : 1;
        tc->tStart = (int64_t *)malloc(a * sizeof(int64_t));
        tc->tEnd = (int64_t *)malloc(a * sizeof(int64_t));
        tc->qStart = (int64_t *)malloc(a * sizeof(int64_t));
        tc->pmax = (int64_t *)malloc(a * sizeof(int64_t));
        tc->qcid = (int32_t *)malloc(a * sizeof(int32_t));
        tc->invert = (int8_t *)malloc(a);
        int64_t pm = INT64_MIN;
        for (size_t i = 0; i < m; i++) {
            const orc_link *l = &tc->raw[i].l;
            tc->tStart[i] = l->tStart;
            tc->tEnd[i] = l->tEnd;
            tc->qStart[i] = l->qStart;
            tc->qcid[i] = l->qcid;
            tc->invert[i] = (int8_t)l->invert;
            if (l->tEnd > pm) pm = l->tEnd;
            tc->pmax[i] = pm;
        }
        free(tc->raw);
        tc->raw = NULL;
        tc->n_raw = tc->cap_raw = 0;
    }
    uint32_t hs = 16;
    while (hs < (uint32_t)cf->ntc * 4u) hs <<= 1;
    cf->hash = (int32_t *)malloc(hs * sizeof(int32_t));
    for (uint32_t i = 0; i < hs; i++) cf->hash[i] = -1;
    cf->hmask = hs - 1;
    for (int t = 0; t < cf->ntc; t++) {
        uint32_t h = fnv1a(cf->tc[t].name, cf->tc[t].name_len) & cf->hmask;
        while (cf->hash[h] >= 0) h = (h + 1) & cf->hmask;
        cf->hash[h] = t;
    }
    return 0;
}

/* chain.d:509-597 */
orc_chainfile *orc_chain_parse(const char *text, size_t len, char *err, size_t errcap) {
    slice *lines;
    size_t nl = split_lines(text, len, &lines);
    orc_chainfile *cf = (orc_chainfile *)calloc(1, sizeof *cf);
    if (nl == 0) {
        set_err(err, errcap, "empty chain file");
        goto fail;
    }
    size_t chainStart = 0;
    for (size_t i = 0; i < nl; i++) {
        if (lines[i].n > 5 && memcmp(lines[i].p, "chain", 5) == 0) { /* chain.d:526 */
            int64_t chainEnd = (int64_t)i - 1;
            if (chainEnd > 0) { /* chain.d:530 (quirk: a chain that ends at line 0 is dropped) */
                chain_hdr h;
                int64_t r = parse_one_chain(lines + chainStart, i - chainStart, add_link_cb, cf, &h, err, errcap);
                if (r < 0) goto fail;
                qsize_set(cf, h.qName.p, h.qName.n, h.qSize); /* chain.d:551 */
            }
            chainStart = i;
        }
    }
    { /* last chain  chain.d:556-579 */
        chain_hdr h;
        int64_t r = parse_one_chain(lines + chainStart, nl - chainStart, add_link_cb, cf, &h, err, errcap);
        if (r < 0) goto fail;
        qsize_set(cf, h.qName.p, h.qName.n, h.qSize);
    }
    free(lines);
    finalize(cf);
    return cf;
fail:
    free(lines);
    orc_chain_free(cf);
    return NULL;
}

static char *slurp(const char *path, size_t *len) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *p = (char *)malloc((size_t)n + 1);
    if (p && fread(p, 1, (size_t)n, f) != (size_t)n) {
        free(p);
        p = NULL;
    }
    fclose(f);
    if (p) *len = (size_t)n;
    return p;
}

orc_chainfile *orc_chain_load(const char *path, char *err, size_t errcap) {
    size_t n;
    char *t = slurp(path, &n);
    if (!t) {
        set_err(err, errcap, "File does not exist: %s", path); /* chain.d:513-514 */
        return NULL;
    }
    orc_chainfile *cf = orc_chain_parse(t, n, err, errcap);
    free(t);
    return cf;
}

void orc_chain_free(orc_chainfile *cf) {
    if (!cf) return;
    for (int t = 0; t < cf->ntc; t++) {
        tcontig *tc = &cf->tc[t];
        free(tc->name);
        free(tc->raw);
        free(tc->tStart);
        free(tc->tEnd);
        free(tc->qStart);
        free(tc->pmax);
        free(tc->qcid);
        free(tc->invert);
    }
    free(cf->tc);
    for (int i = 0; i < cf->nq; i++) free(cf->qnames[i]);
    free(cf->qnames);
    for (int i = 0; i < cf->nqs; i++) free(cf->qsizes[i].name);
    free(cf->qsizes);
    free(cf->hash);
    free(cf);
}

int orc_num_tcontigs(const orc_chainfile *cf) { return cf->ntc; }
const char *orc_tcontig_name(const orc_chainfile *cf, int i) { return cf->tc[i].name; }
int64_t orc_tcontig_nblocks(const orc_chainfile *cf, int i) { return (int64_t)cf->tc[i].n; }
int64_t orc_total_blocks(const orc_chainfile *cf) {
    int64_t s = 0;
    for (int i = 0; i < cf->ntc; i++) s += (int64_t)cf->tc[i].n;
    return s;
}
int orc_tcontig_id(const orc_chainfile *cf, const char *name, size_t len) {
    uint32_t h = fnv1a(name, len) & cf->hmask;
    while (cf->hash[h] >= 0) {
        const tcontig *t = &cf->tc[cf->hash[h]];
        if (t->name_len == len && memcmp(t->name, name, len) == 0) return cf->hash[h];
        h = (h + 1) & cf->hmask;
    }
    return -1;
}
orc_link orc_tcontig_block(const orc_chainfile *cf, int i, int64_t j) {
    const tcontig *t = &cf->tc[i];
    orc_link l = {t->tStart[j], t->tEnd[j], t->qStart[j], t->qcid[j], t->invert[j]};
    return l;
}
int orc_num_qcontigs(const orc_chainfile *cf) { return cf->nq; }
const char *orc_qcontig_name(const orc_chainfile *cf, int id) { return cf->qnames[id]; }
int orc_num_qsizes(const orc_chainfile *cf) { return cf->nqs; }
const char *orc_qsize_name(const orc_chainfile *cf, int i) { return cf->qsizes[i].name; }
int64_t orc_qsize_value(const orc_chainfile *cf, int i) { return cf->qsizes[i].size; }

/* ------------------------------------------------------------- the search
 * findOverlapsWith (intervaltree, half-open): hit set = blocks with
 * tStart < end && tEnd > start.  Sorted array: [lo,hi) with
 *   lo = first j with prefix-max(tEnd)[j] > start,  hi = first j with tStart[j] >= end
 * then filter tEnd[j] > start (a no-op for target-disjoint chains). */
static inline size_t lower_gt(const int64_t *a, size_t n, int64_t x) { /* first a[j] > x */
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t m = (lo + hi) >> 1;
        if (a[m] > x) hi = m;
        else lo = m + 1;
    }
    return lo;
}
static inline size_t lower_ge(const int64_t *a, size_t n, int64_t x) { /* first a[j] >= x */
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t m = (lo + hi) >> 1;
        if (a[m] >= x) hi = m;
        else lo = m + 1;
    }
    return lo;
}

typedef struct {
    orc_link *v;
    size_t n, cap;
} linkvec;

static int64_t lift_into(const orc_chainfile *cf, int tcid, int64_t start, int64_t end, linkvec *o) {
    o->n = 0;
    if (tcid < 0 || tcid >= cf->ntc) return 0;
    const tcontig *t = &cf->tc[tcid];
    size_t lo = lower_gt(t->pmax, t->n, start);
    size_t hi = lower_ge(t->tStart, t->n, end);
    for (size_t j = lo; j < hi; j++) {
        if (t->tEnd[j] <= start) continue;
        if (o->n == o->cap) {
            o->cap = o->cap ? o->cap * 2 : 16;
            o->v = (orc_link *)realloc(o->v, o->cap * sizeof(orc_link));
        }
        orc_link l = {t->tStart[j], t->tEnd[j], t->qStart[j], t->qcid[j], t->invert[j]};
        o->v[o->n++] = orc_intersect(l, start, end); /* chain.d:690,707 */
    }
    return (int64_t)o->n;
}

int64_t orc_lift(const orc_chainfile *cf, int tcid, int64_t start, int64_t end, orc_link **out) {
    linkvec v = {0, 0, 0};
    int64_t n = lift_into(cf, tcid, start, end, &v);
    *out = v.v;
    return n;
}

/* first hit of [c,c+1) in (tStart,tEnd,qStart) order, or -1 */
static inline int64_t point_hit(const tcontig *t, int64_t c) {
    size_t lo = lower_gt(t->pmax, t->n, c);
    size_t hi = lower_ge(t->tStart, t->n, c + 1);
    for (size_t j = lo; j < hi; j++)
        if (t->tEnd[j] > c) return (int64_t)j;
    return -1;
}

/* chain.d:637-654 */
int orc_lift_coord_only(const orc_chainfile *cf, int tcid, int64_t *coord) {
    if (tcid < 0 || tcid >= cf->ntc) return 0;
    const tcontig *t = &cf->tc[tcid];
    int64_t j = point_hit(t, *coord);
    if (j < 0) return 0;
    *coord = t->qStart[j] + t->invert[j] * (*coord - t->tStart[j]); /* intersect().qStart */
    return 1;
}
/* chain.d:604-627 */
int orc_lift_directly(const orc_chainfile *cf, int tcid, int32_t *qcid, int64_t *coord) {
    if (tcid < 0 || tcid >= cf->ntc) return 0;
    const tcontig *t = &cf->tc[tcid];
    int64_t j = point_hit(t, *coord);
    if (j < 0) return 0;
    *coord = t->qStart[j] + t->invert[j] * (*coord - t->tStart[j]);
    *qcid = t->qcid[j];
    return 1;
}

/* stable sort by qStart (bed.d:156-157; tie order = hit order, see header) */
static void sort_links_by_qstart(orc_link *v, size_t n, orc_link *tmp) {
    if (n < 2) return;
    if (n <= 16) {
        for (size_t i = 1; i < n; i++) {
            orc_link x = v[i];
            size_t j = i;
            while (j > 0 && v[j - 1].qStart > x.qStart) {
                v[j] = v[j - 1];
                j--;
            }
            v[j] = x;
        }
        return;
    }
    size_t h = n / 2;
    sort_links_by_qstart(v, h, tmp);
    sort_links_by_qstart(v + h, n - h, tmp);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (v[j].qStart < v[i].qStart) ? v[j++] : v[i++];
    while (i < h) tmp[k++] = v[i++];
    while (j < n) tmp[k++] = v[j++];
    memcpy(v, tmp, n * sizeof(orc_link));
}

/* ------------------------------------------------------------ BED driver */

typedef struct {
    slice *f;
    size_t n, cap;
} fieldvec;

/* bed.d:113-202 */
static int lift_bed_range(const orc_chainfile *cf, const char *in, size_t a, size_t b, orc_buf *lifted,
                          orc_buf *unmatched, orc_bed_stats *st, char *err, size_t errcap) {
    fieldvec fv = {0, 0, 0};
    linkvec lv = {0, 0, 0};
    orc_link *tmp = NULL;
    size_t tmpcap = 0;
    int rc = ORC_OK;
    size_t p = a;
    while (p < b) {
        const char *nlp = (const char *)memchr(in + p, '\n', b - p);
        size_t e = nlp ? (size_t)(nlp - in) : b;
        const char *line = in + p;
        size_t n = e - p;
        size_t line_off = p;
        p = e + 1;
        st->n_lines++;
        if (n > 0 && line[0] == '#') continue; /* bed.d:115 */
        if (n > 4 && (memcmp(line, "track", 5) == 0 || memcmp(line, "brows", 5) == 0)) continue; /* :116 */

        fv.n = 0; /* bed.d:118-119 */
        size_t i = 0;
        while (i < n) {
            while (i < n && is_white((unsigned char)line[i])) i++;
            if (i >= n) break;
            size_t s0 = i;
            while (i < n && !is_white((unsigned char)line[i])) i++;
            if (fv.n == fv.cap) {
                fv.cap = fv.cap ? fv.cap * 2 : 32;
                fv.f = (slice *)realloc(fv.f, fv.cap * sizeof(slice));
            }
            fv.f[fv.n].p = line + s0;
            fv.f[fv.n].n = i - s0;
            fv.n++;
        }
        size_t numf = fv.n;
        if (numf < 3) {
            set_err(err, errcap, "BED line at byte %zu has %zu fields (< 3)", line_off, numf);
            rc = ORC_E_PARSE;
            goto done;
        }
        int32_t s32, e32;
        if (parse_int32(fv.f[1].p, fv.f[1].n, &s32) || parse_int32(fv.f[2].p, fv.f[2].n, &e32)) {
            set_err(err, errcap, "BED line at byte %zu: bad start/end", line_off);
            rc = ORC_E_PARSE;
            goto done;
        }
        int64_t start = s32, end = e32; /* bed.d:125-126 */
        st->n_records++;
        int tcid = orc_tcontig_id(cf, fv.f[0].p, fv.f[0].n);
        lift_into(cf, tcid, start, end, &lv); /* bed.d:129 */

        int64_t thickStart = 0, thickEnd = 0;
        int hasThick = 0;
        if (numf >= 8) { /* bed.d:134-148 */
            int32_t ts32, te32;
            if (parse_int32(fv.f[6].p, fv.f[6].n, &ts32) || parse_int32(fv.f[7].p, fv.f[7].n, &te32)) {
                set_err(err, errcap, "BED line at byte %zu: bad thickStart/thickEnd", line_off);
                rc = ORC_E_PARSE;
                goto done;
            }
            thickStart = ts32;
            thickEnd = te32;
            if (thickStart != thickEnd) {
                hasThick = 1;
                if (!orc_lift_coord_only(cf, tcid, &thickStart) || !orc_lift_coord_only(cf, tcid, &thickEnd))
                    hasThick = 0;
                else
                    orc_order_start_end(&thickStart, &thickEnd);
            }
        }

        if (lv.n == 0) { /* bed.d:150-151 */
            st->n_unmatched++;
            for (size_t k = 0; k < numf; k++) {
                if (k && buf_putc(unmatched, '\t')) goto oom;
                if (buf_put(unmatched, fv.f[k].p, fv.f[k].n)) goto oom;
            }
            if (buf_putc(unmatched, '\n')) goto oom;
            continue;
        }
        if (lv.n > tmpcap) {
            tmpcap = lv.n * 2;
            tmp = (orc_link *)realloc(tmp, tmpcap * sizeof(orc_link));
        }
        sort_links_by_qstart(lv.v, lv.n, tmp); /* bed.d:156-157 */
        unsigned char strand0 = 0;
        if (numf >= 6) strand0 = (unsigned char)fv.f[5].p[0];
        for (size_t k = 0; k < lv.n; k++) {
            const orc_link *l = &lv.v[k];
            const char *qn = cf->qnames[l->qcid]; /* bed.d:163 */
            int64_t s = l->qStart, en = orc_link_qend(l);
            orc_order_start_end(&s, &en); /* bed.d:165-168 */
            if (numf >= 6) strand0 = orc_strand_flip(strand0, l->invert); /* bed.d:173, cumulative */
            int64_t c7 = s, c8 = s;
            if (numf >= 8) { /* bed.d:175-196 */
                if (!(!hasThick || thickStart >= en || thickEnd <= s)) {
                    c7 = thickStart < s ? s : thickStart;
                    c8 = thickEnd > en ? en : thickEnd;
                }
            }
            if (buf_put(lifted, qn, strlen(qn)) || buf_putc(lifted, '\t') || buf_put_i64(lifted, s) ||
                buf_putc(lifted, '\t') || buf_put_i64(lifted, en))
                goto oom;
            for (size_t c = 3; c < numf; c++) {
                if (buf_putc(lifted, '\t')) goto oom;
                if (c == 5) {
                    if (buf_putc(lifted, (char)strand0)) goto oom;
                    if (fv.f[5].n > 1 && buf_put(lifted, fv.f[5].p + 1, fv.f[5].n - 1)) goto oom;
                } else if (c == 6 && numf >= 8) {
                    if (buf_put_i64(lifted, c7)) goto oom;
                } else if (c == 7 && numf >= 8) {
                    if (buf_put_i64(lifted, c8)) goto oom;
                } else if (buf_put(lifted, fv.f[c].p, fv.f[c].n))
                    goto oom;
            }
            if (buf_putc(lifted, '\n')) goto oom; /* bed.d:199 */
            st->n_rows++;
        }
    }
    goto done;
oom:
    set_err(err, errcap, "out of memory");
    rc = ORC_E_OOM;
done:
    free(fv.f);
    free(lv.v);
    free(tmp);
    return rc;
}

int orc_lift_bed(const orc_chainfile *cf, const char *in, size_t in_len, orc_buf *lifted, orc_buf *unmatched,
                 orc_bed_stats *st, char *err, size_t errcap) {
    orc_bed_stats local = {0, 0, 0, 0};
    int rc = lift_bed_range(cf, in, 0, in_len, lifted, unmatched, &local, err, errcap);
    if (st) *st = local;
    return rc;
}

typedef struct {
    const orc_chainfile *cf;
    const char *in;
    size_t a, b;
    orc_buf lifted, unmatched;
    orc_bed_stats st;
    char err[256];
    int rc;
} mt_job;

static void *mt_run(void *arg) {
    mt_job *j = (mt_job *)arg;
    j->rc = lift_bed_range(j->cf, j->in, j->a, j->b, &j->lifted, &j->unmatched, &j->st, j->err, sizeof j->err);
    return NULL;
}

int orc_lift_bed_mt(const orc_chainfile *cf, const char *in, size_t in_len, int nthreads, orc_buf *lifted,
                    orc_buf *unmatched, orc_bed_stats *st, char *err, size_t errcap) {
    if (nthreads < 1) nthreads = 1;
    mt_job *jobs = (mt_job *)calloc((size_t)nthreads, sizeof(mt_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    size_t prev = 0;
    for (int t = 0; t < nthreads; t++) {
        size_t cut = (t == nthreads - 1) ? in_len : (size_t)((__uint128_t)in_len * (unsigned)(t + 1) / (unsigned)nthreads);
        if (cut < prev) cut = prev;
        /* a range owns the lines that START inside it */
        while (cut < in_len && cut > 0 && in[cut - 1] != '\n') cut++;
        jobs[t].cf = cf;
        jobs[t].in = in;
        jobs[t].a = prev;
        jobs[t].b = cut;
        prev = cut;
    }
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, mt_run, &jobs[t]);
    int rc = ORC_OK;
    orc_bed_stats tot = {0, 0, 0, 0};
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (rc == ORC_OK && jobs[t].rc != ORC_OK) {
            rc = jobs[t].rc;
            set_err(err, errcap, "%s", jobs[t].err);
        }
    }
    for (int t = 0; t < nthreads; t++) {
        if (rc == ORC_OK) {
            if (buf_put(lifted, jobs[t].lifted.data, jobs[t].lifted.len) ||
                buf_put(unmatched, jobs[t].unmatched.data, jobs[t].unmatched.len))
                rc = ORC_E_OOM;
            tot.n_lines += jobs[t].st.n_lines;
            tot.n_records += jobs[t].st.n_records;
            tot.n_rows += jobs[t].st.n_rows;
            tot.n_unmatched += jobs[t].st.n_unmatched;
        }
        orc_buf_free(&jobs[t].lifted);
        orc_buf_free(&jobs[t].unmatched);
    }
    if (st) *st = tot;
    free(jobs);
    free(th);
    return rc;
}

/* ----------------------------------------------------- binary record body */

void orc_rows_free(orc_rows *r) {
    free(r->row_offset);
    free(r->qcid);
    free(r->start);
    free(r->end);
    free(r->src);
    free(r->strand);
    free(r->thick_start);
    free(r->thick_end);
    free(r->unmatched_idx);
    memset(r, 0, sizeof *r);
}

/* bed.d:129-196 on integers.  Pass 1 counts, pass 2 fills. */
int orc_lift_intervals(const orc_chainfile *cf, size_t n, const int32_t *tcid, const int32_t *start,
                       const int32_t *end, const uint8_t *strand, const int32_t *thick_start,
                       const int32_t *thick_end, orc_rows *out) {
    memset(out, 0, sizeof *out);
    linkvec lv = {0, 0, 0};
    orc_link *tmp = NULL;
    size_t tmpcap = 0;
    out->row_offset = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    uint64_t rows = 0, unm = 0;
    for (size_t i = 0; i < n; i++) {
        out->row_offset[i] = (uint32_t)rows;
        int64_t h = lift_into(cf, tcid[i], start[i], end[i], &lv);
        rows += (uint64_t)h;
        if (!h) unm++;
    }
    out->row_offset[n] = (uint32_t)rows;
    if (rows > UINT32_MAX) {
        free(lv.v);
        orc_rows_free(out);
        return ORC_E_ARG;
    }
    out->n_rows = rows;
    out->n_unmatched = unm;
    size_t ra = rows ? rows : 1, ua = unm ? unm : 1;
    out->qcid = (int32_t *)malloc(ra * sizeof(int32_t));
    out->start = (int32_t *)malloc(ra * sizeof(int32_t));
    out->end = (int32_t *)malloc(ra * sizeof(int32_t));
    out->src = (uint32_t *)malloc(ra * sizeof(uint32_t));
    if (strand) out->strand = (uint8_t *)malloc(ra);
    if (thick_start && thick_end) {
        out->thick_start = (int32_t *)malloc(ra * sizeof(int32_t));
        out->thick_end = (int32_t *)malloc(ra * sizeof(int32_t));
    }
    out->unmatched_idx = (uint32_t *)malloc(ua * sizeof(uint32_t));
    uint64_t r = 0, u = 0;
    for (size_t i = 0; i < n; i++) {
        lift_into(cf, tcid[i], start[i], end[i], &lv);
        if (lv.n == 0) {
            out->unmatched_idx[u++] = (uint32_t)i;
            continue;
        }
        int64_t ts = 0, te = 0;
        int hasThick = 0;
        if (out->thick_start) {
            ts = thick_start[i];
            te = thick_end[i];
            if (ts != te) {
                hasThick = 1;
                if (!orc_lift_coord_only(cf, tcid[i], &ts) || !orc_lift_coord_only(cf, tcid[i], &te)) hasThick = 0;
                else orc_order_start_end(&ts, &te);
            }
        }
        if (lv.n > tmpcap) {
            tmpcap = lv.n * 2;
            tmp = (orc_link *)realloc(tmp, tmpcap * sizeof(orc_link));
        }
        sort_links_by_qstart(lv.v, lv.n, tmp);
        unsigned char sb = strand ? strand[i] : 0;
        for (size_t k = 0; k < lv.n; k++, r++) {
            const orc_link *l = &lv.v[k];
            int64_t s = l->qStart, en = orc_link_qend(l);
            orc_order_start_end(&s, &en);
            out->qcid[r] = l->qcid;
            out->start[r] = (int32_t)s;
            out->end[r] = (int32_t)en;
            out->src[r] = (uint32_t)i;
            if (strand) {
                sb = orc_strand_flip(sb, l->invert);
                out->strand[r] = sb;
            }
            if (out->thick_start) {
                int64_t c7 = s, c8 = s;
                if (!(!hasThick || ts >= en || te <= s)) {
                    c7 = ts < s ? s : ts;
                    c8 = te > en ? en : te;
                }
                out->thick_start[r] = (int32_t)c7;
                out->thick_end[r] = (int32_t)c8;
            }
        }
    }
    free(lv.v);
    free(tmp);
    return ORC_OK;
}

/* ------------------------------------------------------------ VCF points */

static inline char base_at(const uint8_t *packed, int64_t g) {
    return "ACGT"[(packed[g >> 2] >> ((g & 3) * 2)) & 3];
}

/* vcf.d:115-169 */
int orc_lift_vcf_points(const orc_chainfile *cf, const uint8_t *packed, const int64_t *contig_off,
                        const int64_t *contig_len, int n_gcontigs, size_t n, const int32_t *tcid,
                        const int32_t *pos, const int32_t *reflen, const uint64_t *ref_off,
                        const char *ref_bytes, int32_t *out_qcid, int32_t *out_pos, uint8_t *flags,
                        int32_t *out_reflen, char *out_ref) {
    for (size_t i = 0; i < n; i++) {
        int32_t q = -1;
        int64_t c = pos[i];
        int hit = orc_lift_directly(cf, tcid[i], &q, &c); /* vcf.d:125 */
        if (!hit) {                                      /* vcf.d:134-137: record unchanged */
            out_qcid[i] = -1;
            out_pos[i] = pos[i];
            flags[i] = 0;
            out_reflen[i] = reflen[i];
            if (out_ref && reflen[i] > 0) memcpy(out_ref + ref_off[i], ref_bytes + ref_off[i], (size_t)reflen[i]);
            continue;
        }
        out_qcid[i] = q;
        out_pos[i] = (int32_t)c;
        uint8_t fl = 1;
        /* fetchSequence(chrom, pos, pos+refLen), clipped to the contig (vcf.d:143) */
        int64_t L = (q >= 0 && q < n_gcontigs) ? contig_len[q] : 0;
        int64_t m = 0;
        if (c >= 0 && c < L) {
            m = reflen[i];
            if (c + m > L) m = L - c;
        }
        int diff = (m != reflen[i]);
        for (int64_t k = 0; k < m; k++) {
            char b = base_at(packed, contig_off[q] + c + k);
            if (out_ref) out_ref[ref_off[i] + (uint64_t)k] = b;
            if (b != ref_bytes[ref_off[i] + (uint64_t)k]) diff = 1; /* vcf.d:144 byte-exact */
        }
        out_reflen[i] = (int32_t)m;
        if (diff) fl |= 2; /* vcf.d:156-161 */
        flags[i] = fl;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------ liftVCF on text */
static int is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
static int buf_puts(orc_buf *b, const char *s) { return buf_put(b, s, strlen(s)); }
typedef struct {
    const char *name;
    size_t name_len;
    char *seq;
    int64_t len;
} fa_seq;

static int fa_parse(const char *fa, size_t n, fa_seq **out, int *nout) {
    int cap = 16, cnt = 0;
    fa_seq *v = (fa_seq *)malloc(sizeof(fa_seq) * (size_t)cap);
    size_t i = 0;
    while (i < n) {
        size_t e = i;
        while (e < n && fa[e] != '\n') e++;
        if (fa[i] == '>') {
            if (cnt == cap) {
                cap *= 2;
                v = (fa_seq *)realloc(v, sizeof(fa_seq) * (size_t)cap);
            }
            size_t a = i + 1, b = a;
            while (b < e && !is_space((unsigned char)fa[b])) b++;
            v[cnt].name = fa + a;
            v[cnt].name_len = b - a;
            v[cnt].seq = (char *)malloc(n - e + 1);
            v[cnt].len = 0;
            cnt++;
        } else if (cnt > 0) {
            for (size_t k = i; k < e; k++)
                if (!is_space((unsigned char)fa[k])) v[cnt - 1].seq[v[cnt - 1].len++] = fa[k];
        }
        i = e + 1;
    }
    *out = v;
    *nout = cnt;
    return ORC_OK;
}
static const fa_seq *fa_find(const fa_seq *v, int n, const char *name, size_t len) {
    for (int i = 0; i < n; i++)
        if (v[i].name_len == len && memcmp(v[i].name, name, len) == 0) return &v[i];
    return NULL;
}

/* vcf.d:200-215 */
static int contig_numeric(const char *z, size_t n) {
    if (n >= 3 && memcmp(z, "chr", 3) == 0) {
        z += 3;
        n -= 3;
    }
    if (n == 0) return 200; /* the reference would index out of bounds */
    if (z[0] == 'X') return 100;
    if (z[0] == 'Y') return 101;
    if (z[0] == 'M') return 102;
    size_t k = 0;
    long v = 0;
    while (k < n && z[k] >= '0' && z[k] <= '9') v = v * 10 + (z[k++] - '0');
    if (k == 0) return 200;
    return (int)v;
}
/* vcf.d:181-196: true if x sorts before y */
static int contig_before(const char *x, const char *y) {
    const size_t nx = strlen(x), ny = strlen(y);
    const int ux = memchr(x, '_', nx) != NULL, uy = memchr(y, '_', ny) != NULL;
    if (ux && !uy) return 0;
    if (!ux && uy) return 1;
    const int a = contig_numeric(x, nx), b = contig_numeric(y, ny);
    if (a != b) return a < b;
    return strcmp(x, y) < 0;
}

int orc_lift_vcf_text(const orc_chainfile *cf, const char *fasta, size_t fasta_len, const char *in,
                      size_t in_len, const char *chainfile, const char *genomefile,
                      const char *filedate, orc_buf *lifted, orc_buf *unmatched,
                      orc_vcf_stats *st, char *err, size_t errcap) {
    fa_seq *fa = NULL;
    int nfa = 0;
    fa_parse(fasta, fasta_len, &fa, &nfa);
    orc_vcf_stats s = {0, 0, 0, 0};
    int rc = ORC_OK;
    int header_done = 0;
    size_t i = 0;
    /* IDs of the contig lines of the input header */
    while (i < in_len && rc == ORC_OK) {
        size_t e = i;
        while (e < in_len && in[e] != '\n') e++;
        const char *ln = in + i;
        size_t n = e - i;
        if (n > 0 && ln[n - 1] == '\r') n--;
        i = e + 1;
        if (n == 0) continue;
        if (ln[0] == '#') {
            if (n >= 2 && ln[1] == '#') { /* meta line: both files */
                buf_put(lifted, ln, n);
                buf_putc(lifted, '\n');
                buf_put(unmatched, ln, n);
                buf_putc(unmatched, '\n');
                continue;
            }
            /* "#CHROM" line: the additions go in front of it (vcf.d:62-103) */
            if (!header_done) {
                static const char *refchg =
                    "##INFO=<ID=refchg,Number=0,Type=Flag,Description=\"REF allele changed in new genome\">\n";
                buf_put(lifted, refchg, strlen(refchg));
                const int nq = orc_num_qsizes(cf);
                int *ord = (int *)malloc(sizeof(int) * (size_t)(nq > 0 ? nq : 1));
                for (int k = 0; k < nq; k++) ord[k] = k;
                for (int a = 1; a < nq; a++) { /* insertion sort by contigSort */
                    int v = ord[a], b = a - 1;
                    while (b >= 0 && contig_before(orc_qsize_name(cf, v), orc_qsize_name(cf, ord[b]))) {
                        ord[b + 1] = ord[b];
                        b--;
                    }
                    ord[b + 1] = v;
                }
                for (int a = 0; a < nq; a++) {
                    const char *k = orc_qsize_name(cf, ord[a]);
                    const fa_seq *q = fa_find(fa, nfa, k, strlen(k));
                    if (!q || q->len != orc_qsize_value(cf, ord[a])) continue; /* vcf.d:79-87 */
                    /* an ID the header already has is not added again */
                    char pat[512];
                    snprintf(pat, sizeof pat, "##contig=<ID=%s", k);
                    const size_t pl = strlen(pat);
                    int dup = 0;
                    for (size_t h = 0; h + pl < lifted->len && !dup; h++)
                        if ((h == 0 || lifted->data[h - 1] == '\n') && memcmp(lifted->data + h, pat, pl) == 0 &&
                            (lifted->data[h + pl] == ',' || lifted->data[h + pl] == '>'))
                            dup = 1;
                    if (dup) continue;
                    buf_put(lifted, pat, pl);
                    buf_puts(lifted, ",length=");
                    buf_put_i64(lifted, orc_qsize_value(cf, ord[a]));
                    buf_puts(lifted, ",source=");
                    buf_puts(lifted, chainfile);
                    buf_puts(lifted, ">\n");
                }
                free(ord);
                buf_puts(lifted, "##fileDate=");
                buf_puts(lifted, filedate);
                buf_puts(lifted, "\n##liftoverProgram=swiftover\n");
                buf_puts(lifted, "##liftoverProgramURL=https://github.com/blachlylab/swiftover\n");
                buf_puts(lifted, "##liftoverChainfile=");
                buf_puts(lifted, chainfile);
                buf_puts(lifted, "\n##liftoverGenome=");
                buf_puts(lifted, genomefile);
                buf_putc(lifted, '\n');
                header_done = 1;
            }
            buf_put(lifted, ln, n);
            buf_putc(lifted, '\n');
            buf_put(unmatched, ln, n);
            buf_putc(unmatched, '\n');
            continue;
        }
        /* record: columns are tab separated */
        size_t c0[9], c1[9];
        int nc = 0;
        size_t a = 0;
        for (size_t k = 0; k <= n && nc < 9; k++) {
            if (k == n || ln[k] == '\t') {
                c0[nc] = a;
                c1[nc] = (nc == 8) ? n : k;
                nc++;
                a = k + 1;
                if (nc == 8 && k < n) { /* everything after INFO passes through */
                    c0[8] = a;
                    c1[8] = n;
                    nc = 9;
                }
            }
        }
        if (nc < 8) {
            snprintf(err, errcap, "VCF record with fewer than 8 columns at byte %zu", (size_t)(ln - in));
            rc = ORC_E_PARSE;
            break;
        }
        int64_t pos1 = 0;
        int ok = c1[1] > c0[1];
        for (size_t k = c0[1]; k < c1[1]; k++) {
            if (ln[k] < '0' || ln[k] > '9') ok = 0;
            else pos1 = pos1 * 10 + (ln[k] - '0');
            if (pos1 > 2147483647LL) ok = 0;
        }
        if (!ok) {
            snprintf(err, errcap, "bad POS at byte %zu", (size_t)(ln - in));
            rc = ORC_E_PARSE;
            break;
        }
        s.n_records++;
        const int tcid = orc_tcontig_id(cf, ln + c0[0], c1[0] - c0[0]);
        int32_t q = -1;
        int64_t c = pos1 - 1; /* rec.pos is 0-based */
        const int hit = orc_lift_directly(cf, tcid, &q, &c); /* vcf.d:125 */
        if (!hit) {                                        /* vcf.d:134-137 */
            s.n_unmatched++;
            buf_put(unmatched, ln, n);
            buf_putc(unmatched, '\n');
            continue;
        }
        s.n_matched++;
        const char *qn = orc_qcontig_name(cf, q);
        const size_t reflen = c1[3] - c0[3];
        const fa_seq *fs = fa_find(fa, nfa, qn, strlen(qn));
        int64_t m = 0;
        if (fs && c >= 0 && c < fs->len) { /* fetchSequence, clipped (vcf.d:143) */
            m = (int64_t)reflen;
            if (c + m > fs->len) m = fs->len - c;
        }
        const int diff = ((size_t)m != reflen) || (m > 0 && memcmp(fs->seq + c, ln + c0[3], (size_t)m) != 0);
        buf_put(lifted, qn, strlen(qn));
        buf_putc(lifted, '\t');
        buf_put_i64(lifted, c + 1);
        buf_putc(lifted, '\t');
        buf_put(lifted, ln + c0[2], c1[2] - c0[2]);
        buf_putc(lifted, '\t');
        if (diff) {
            s.n_refchg++;
            if (m > 0) buf_put(lifted, fs->seq + c, (size_t)m); /* vcf.d:156-157 */
        } else {
            buf_put(lifted, ln + c0[3], reflen);
        }
        buf_putc(lifted, '\t');
        buf_put(lifted, ln + c0[4], c1[6] - c0[4]); /* ALT QUAL FILTER */
        buf_putc(lifted, '\t');
        if (diff) { /* vcf.d:161 */
            if (c1[7] - c0[7] == 1 && ln[c0[7]] == '.') buf_put(lifted, "refchg", 6);
            else {
                buf_put(lifted, ln + c0[7], c1[7] - c0[7]);
                buf_put(lifted, ";refchg", 7);
            }
        } else {
            buf_put(lifted, ln + c0[7], c1[7] - c0[7]);
        }
        if (nc == 9) {
            buf_putc(lifted, '\t');
            buf_put(lifted, ln + c0[8], c1[8] - c0[8]);
        }
        buf_putc(lifted, '\n');
    }
    for (int k = 0; k < nfa; k++) free(fa[k].seq);
    free(fa);
    if (st) *st = s;
    return rc;
}

/* ------------------------------------------------------------ liftMAF text */
/* maf.d:21-32 */
static unsigned char nt_comp(unsigned char c) {
    switch (c) {
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
/* maf.d:36-54 */
size_t orc_reverse_complement(const char *allele, size_t n, char *out) {
    if (n == 2 && allele[0] == 'N' && allele[1] == 'A') { /* maf.d:41-42 */
        out[0] = 'N';
        out[1] = 'A';
        return 2;
    }
    for (size_t i = 0; i < n; i++) out[n - 1 - i] = (char)nt_comp((unsigned char)allele[i]);
    return n;
}

/* std.conv.to!long: optional sign, then digits only */
static int parse_int64(const char *s, size_t n, int64_t *out) {
    size_t i = 0;
    int neg = 0;
    if (n && (s[0] == '-' || s[0] == '+')) {
        neg = s[0] == '-';
        i = 1;
    }
    if (i == n) return 0;
    uint64_t v = 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return 0;
        const uint64_t d = (uint64_t)(s[i] - '0');
        if (v > (UINT64_MAX - d) / 10) return 0;
        v = v * 10 + d;
    }
    if (neg ? v > (uint64_t)INT64_MAX + 1 : v > (uint64_t)INT64_MAX) return 0;
    *out = neg ? (int64_t)(0 - v) : (int64_t)v;
    return 1;
}

int orc_lift_maf_text(const orc_chainfile *cf, const char *fasta, size_t fasta_len, const char *in,
                      size_t in_len, const char *genomebuild, orc_buf *lifted, orc_buf *unmatched,
                      orc_maf_stats *st, char *err, size_t errcap) {
    enum { M_BUILD = 3, M_CONTIG = 4, M_START = 5, M_END = 6, M_VTYPE = 9, M_REF = 10, M_T1 = 11, M_T2 = 12 }; /* maf.d:57-72 */
    fa_seq *fa = NULL;
    int nfa = 0;
    fa_parse(fasta, fasta_len, &fa, &nfa);
    orc_maf_stats s = {0, 0, 0, 0, 0, 0};
    int rc = ORC_OK;
    size_t i = 0;
    for (int h = 0; h < 2 && i < in_len; h++) { /* readln x2, terminator kept (maf.d:166-171) */
        size_t e = i;
        while (e < in_len && in[e] != '\n') e++;
        if (e < in_len) e++;
        buf_put(lifted, in + i, e - i);
        buf_put(unmatched, in + i, e - i);
        i = e;
    }
    size_t fcap = 64, *f0 = (size_t *)malloc(sizeof(size_t) * fcap), *f1 = (size_t *)malloc(sizeof(size_t) * fcap);
    char *tmp = NULL;
    size_t tmpcap = 0;
    while (i < in_len && rc == ORC_OK) {
        size_t e = i;
        while (e < in_len && in[e] != '\n') e++;
        const char *ln = in + i;
        const size_t n = e - i;
        const size_t at = i;
        i = e + 1;
        size_t nf = 0, k = 0; /* splitter(): runs of white space (maf.d:175) */
        while (k < n) {
            while (k < n && is_space((unsigned char)ln[k])) k++;
            if (k == n) break;
            if (nf == fcap) {
                fcap *= 2;
                f0 = (size_t *)realloc(f0, sizeof(size_t) * fcap);
                f1 = (size_t *)realloc(f1, sizeof(size_t) * fcap);
            }
            f0[nf] = k;
            while (k < n && !is_space((unsigned char)ln[k])) k++;
            f1[nf++] = k;
        }
        int64_t start, end;
        if (nf < 13) {
            snprintf(err, errcap, "MAF record with fewer than 13 fields at byte %zu", at);
            rc = ORC_E_PARSE;
            break;
        }
        if (!parse_int64(ln + f0[M_START], f1[M_START] - f0[M_START], &start) ||
            !parse_int64(ln + f0[M_END], f1[M_END] - f0[M_END], &end)) {
            snprintf(err, errcap, "bad start/end at byte %zu", at);
            rc = ORC_E_PARSE;
            break;
        }
        if (start > end) { /* assert(start <= end), maf.d:183 */
            snprintf(err, errcap, "start > end at byte %zu", at);
            rc = ORC_E_PARSE;
            break;
        }
        s.n_records++;
        start -= 1; /* maf.d:186 */
        const int tcid = orc_tcontig_id(cf, ln + f0[M_CONTIG], f1[M_CONTIG] - f0[M_CONTIG]);
        orc_link *hits = NULL;
        const int64_t nh = orc_lift(cf, tcid, start, end, &hits); /* maf.d:189 */
        if (nh == 0) { /* maf.d:192-195 */
            s.n_unmatched++;
            for (size_t j = 0; j < nf; j++) {
                if (j) buf_putc(unmatched, '\t');
                buf_put(unmatched, ln + f0[j], f1[j] - f0[j]);
            }
            buf_putc(unmatched, '\n');
        } else if (nh > 1) { /* maf.d:298-307 */
            s.n_matched++;
            s.n_multiple++;
        } else {
            s.n_matched++;
            const orc_link l = hits[0];
            int64_t qs = l.qStart, qe = orc_link_qend(&l);
            orc_order_start_end(&qs, &qe); /* maf.d:211 */
            qs++;                          /* maf.d:212 */
            const char *qn = orc_qcontig_name(cf, l.qcid);
            /* new REF (maf.d:263-264) */
            const char *ref = ln + f0[M_REF];
            const size_t reflen = f1[M_REF] - f0[M_REF];
            const char *newref;
            size_t newlen;
            if (reflen == 1 && ref[0] == '-') {
                newref = "-";
                newlen = 1;
            } else {
                const fa_seq *fs = fa_find(fa, nfa, qn, strlen(qn));
                const int64_t a = qs - 1, b = qe; /* 1-based closed -> [a,b) */
                newref = "";
                newlen = 0;
                if (fs && a >= 0 && a < fs->len && b > a) {
                    newref = fs->seq + a;
                    newlen = (size_t)((b > fs->len ? fs->len : b) - a);
                }
            }
            /* tumor alleles after the strand flip (maf.d:220-224) */
            const size_t l1 = f1[M_T1] - f0[M_T1], l2 = f1[M_T2] - f0[M_T2];
            if (l1 + l2 + 2 > tmpcap) {
                tmpcap = (l1 + l2 + 2) * 2;
                tmp = (char *)realloc(tmp, tmpcap);
            }
            const char *t1 = ln + f0[M_T1], *t2 = ln + f0[M_T2];
            if (l.invert < 0) {
                orc_reverse_complement(t1, l1, tmp);
                orc_reverse_complement(t2, l2, tmp + l1);
                t1 = tmp;
                t2 = tmp + l1;
            }
            const int chg = reflen != newlen || memcmp(ref, newref, reflen) != 0; /* maf.d:265 */
            if (chg) {
                s.n_refchg++;
                const size_t vl = f1[M_VTYPE] - f0[M_VTYPE];
                const int is_del = vl == 3 && memcmp(ln + f0[M_VTYPE], "DEL", 3) == 0;
                const int eq1 = newlen == l1 && memcmp(newref, t1, l1) == 0;
                const int eq2 = newlen == l2 && memcmp(newref, t2, l2) == 0;
                if (!is_del && !eq1 && !eq2) s.n_neither++; /* maf.d:273-277 */
            }
            for (size_t j = 0; j < nf; j++) { /* maf.d:296 */
                if (j) buf_putc(lifted, '\t');
                if (j == M_BUILD) buf_puts(lifted, genomebuild);
                else if (j == M_CONTIG) buf_puts(lifted, qn);
                else if (j == M_START) buf_put_i64(lifted, qs);
                else if (j == M_END) buf_put_i64(lifted, qe);
                else if (j == M_REF && chg) buf_put(lifted, newref, newlen);
                else if (j == M_T1) buf_put(lifted, t1, l1);
                else if (j == M_T2) buf_put(lifted, t2, l2);
                else buf_put(lifted, ln + f0[j], f1[j] - f0[j]);
            }
            buf_putc(lifted, '\n');
        }
        free(hits);
    }
    free(f0);
    free(f1);
    free(tmp);
    for (int k = 0; k < nfa; k++) free(fa[k].seq);
    free(fa);
    if (st) *st = s;
    return rc;
}

/* ---------------------------------------------------- bench workload helper */
size_t orc_format_bed3(const orc_chainfile *cf, size_t n, const int32_t *tcid, const int32_t *start,
                       const int32_t *end, char *out, size_t cap) {
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        const tcontig *t = &cf->tc[tcid[i]];
        if (o + t->name_len + 26 > cap) return 0;
        memcpy(out + o, t->name, t->name_len);
        o += t->name_len;
        o += (size_t)sprintf(out + o, "\t%d\t%d\n", start[i], end[i]);
    }
    return o;
}
