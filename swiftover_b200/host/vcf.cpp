// vcf.cpp — liftVCF host driver: the reference's htslib record loop (vcf.d:36-176) becomes
// read-all -> swo_lift_vcf_text (header rewrite + column split on the host, liftDirectly +
// REF fetch/compare on the GPU) -> write-all.  Uncompressed text VCF and FASTA only.
#include <cstdio>
#include <cstring>
#include <string>

#include "chainfile.hpp"

namespace swiftover {

void liftVCF(const std::string &chainfile, const std::string &genomefile, const std::string &infile,
             const std::string &outfile, const std::string &unmatched) {
    ChainFile cf(chainfile);                                         // vcf.d:36
    cf.check(swo_load_genome_fasta(cf.handle(), genomefile.c_str())); // vcf.d:38
    FILE *f = fopen(infile.c_str(), "rb");
    if (!f) throw Error(SWO_E_IO, "Input file does not exist");
    std::string data;
    char buf[1 << 16];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, r);
    fclose(f);
    swo_text_out out;
    cf.check(swo_lift_vcf_text(cf.handle(), data.data(), data.size(), chainfile.c_str(), genomefile.c_str(), nullptr, &out));
    FILE *fo = fopen(outfile.c_str(), "wb");
    if (!fo) throw Error(SWO_E_IO, "cannot open " + outfile);
    fwrite(out.lifted, 1, out.lifted_len, fo);
    fclose(fo);
    FILE *fu = fopen(unmatched.c_str(), "wb");
    if (!fu) throw Error(SWO_E_IO, "cannot open " + unmatched);
    fwrite(out.unmatched, 1, out.unmatched_len, fu);
    fclose(fu);
    // vcf.d:171-172
    fprintf(stderr, "[I::liftVCF] Matched %llu records (%llu unmatched)\n", (unsigned long long)out.n_rows,
            (unsigned long long)out.n_unmatched);
}

}  // namespace swiftover
