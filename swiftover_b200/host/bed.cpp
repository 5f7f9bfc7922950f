// bed.cpp — liftBED host driver: the reference's record loop (bed.d:113-202) becomes
// read-all -> swo_lift_bed_text (one batched call into the CUDA engine) -> write-all.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <iterator>

#include "chainfile.hpp"

namespace swiftover {

static bool slurp(const std::string &path, char *&buf, size_t &len) {
    FILE *f = (path == "-" || path.empty()) ? stdin : fopen(path.c_str(), "rb");  // bed.d:89-92
    if (!f) return false;
    size_t cap = 1 << 22;
    len = 0;
    buf = (char *)swo_host_alloc(cap);  // pinned: full-rate H2D
    if (!buf) throw Error(SWO_E_OOM, "out of pinned host memory");
    for (;;) {
        size_t r = fread(buf + len, 1, cap - len, f);
        len += r;
        if (r == 0) break;
        if (len == cap) {
            char *nb = (char *)swo_host_alloc(cap * 2);
            if (!nb) {
                swo_host_free(buf);
                throw Error(SWO_E_OOM, "out of pinned host memory");
            }
            memcpy(nb, buf, len);
            swo_host_free(buf);
            buf = nb;
            cap *= 2;
        }
    }
    const bool bad = ferror(f) != 0;
    if (f != stdin) fclose(f);
    if (bad) {
        swo_host_free(buf);
        throw Error(SWO_E_IO, "read error on " + path);
    }
    return true;
}

// the shards' pieces in order are the file (no concatenation copy: swo_set_option "concat" 0)
static void write_pieces(swo_ctx *ctx, int which, FILE *f, const std::string &name) {
    for (int i = 0; i < swo_text_num_segments(ctx); i++) {
        const char *p = nullptr;
        size_t n = 0;
        if (swo_text_segment(ctx, which, i, &p, &n) != SWO_OK || (n && fwrite(p, 1, n, f) != n))
            throw Error(SWO_E_IO, "write error on " + name);
    }
}

void liftBED(const std::string &chainfile, const std::string &infile, const std::string &outfile,
             const std::string &unmatched) {
    ChainFile cf(chainfile);  // bed.d:76
    char *in = nullptr;
    size_t n = 0;
    if (!slurp(infile, in, n)) throw Error(SWO_E_IO, "Input file does not exist");
    swo_set_option(cf.handle(), "concat", 0);
    swo_text_out out;
    const int rc = swo_lift_bed_text(cf.handle(), in, n, &out);
    swo_host_free(in);
    // on a malformed line the records in front of it are still written, like the reference's stream (bed.d:113-199)
    if (rc != SWO_OK && rc != SWO_E_PARSE) cf.check(rc);
    FILE *fo = (outfile == "-" || outfile.empty()) ? stdout : fopen(outfile.c_str(), "wb");  // bed.d:94-97
    if (!fo) throw Error(SWO_E_IO, "cannot open " + outfile);
    write_pieces(cf.handle(), 0, fo, outfile);
    if ((fo != stdout ? fclose(fo) : fflush(stdout)) != 0) throw Error(SWO_E_IO, "write error on " + outfile);
    if (!unmatched.empty()) {  // bed.d:99-100
        FILE *fu = fopen(unmatched.c_str(), "wb");
        if (!fu) throw Error(SWO_E_IO, "cannot open " + unmatched);
        write_pieces(cf.handle(), 1, fu, unmatched);
        if (fclose(fu) != 0) throw Error(SWO_E_IO, "write error on " + unmatched);
    }
    cf.check(rc);
}

}  // namespace swiftover
