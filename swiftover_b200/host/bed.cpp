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
    for (;;) {
        size_t r = fread(buf + len, 1, cap - len, f);
        len += r;
        if (r == 0) break;
        if (len == cap) {
            char *nb = (char *)swo_host_alloc(cap * 2);
            memcpy(nb, buf, len);
            swo_host_free(buf);
            buf = nb;
            cap *= 2;
        }
    }
    if (f != stdin) fclose(f);
    return true;
}

void liftBED(const std::string &chainfile, const std::string &infile, const std::string &outfile,
             const std::string &unmatched) {
    ChainFile cf(chainfile);  // bed.d:76
    char *in = nullptr;
    size_t n = 0;
    if (!slurp(infile, in, n)) throw Error(SWO_E_IO, "Input file does not exist");
    swo_text_out out;
    int rc = swo_lift_bed_text(cf.handle(), in, n, &out);
    swo_host_free(in);
    cf.check(rc);
    FILE *fo = (outfile == "-" || outfile.empty()) ? stdout : fopen(outfile.c_str(), "wb");  // bed.d:94-97
    if (!fo) throw Error(SWO_E_IO, "cannot open " + outfile);
    fwrite(out.lifted, 1, out.lifted_len, fo);
    if (fo != stdout) fclose(fo);
    else fflush(stdout);
    if (!unmatched.empty()) {  // bed.d:99-100
        FILE *fu = fopen(unmatched.c_str(), "wb");
        if (!fu) throw Error(SWO_E_IO, "cannot open " + unmatched);
        fwrite(out.unmatched, 1, out.unmatched_len, fu);
        fclose(fu);
    }
}

}  // namespace swiftover
