// maf.cpp — liftMAF host driver: the reference's record loop (maf.d:119-319) becomes
// read-all -> swo_lift_maf_text (column split/join + tumor-allele reverse complement on the host,
// lift + REF fetch/compare on the GPU) -> write-all.  stdin/stdout as in maf.d:139-147.
#include <cstdio>
#include <cstring>
#include <string>

#include "chainfile.hpp"

namespace swiftover {

void liftMAF(const std::string &chainfile, const std::string &genomefile, const std::string &genomebuild,
             const std::string &infile, const std::string &outfile, const std::string &unmatched) {
    fprintf(stderr, "[W::liftMAF] Only fields build, chr, start, end, ref/tumor1/tumor2 are updated.\n");  // maf.d:123-124
    fprintf(stderr, "[W::liftMAF] A reference allele change might invalidate other fields.\n");
    ChainFile cf(chainfile);                                           // maf.d:132
    cf.check(swo_load_genome_fasta(cf.handle(), genomefile.c_str()));  // maf.d:126
    FILE *f = (infile == "-" || infile.empty()) ? stdin : fopen(infile.c_str(), "rb");
    if (!f) throw Error(SWO_E_IO, "Input file does not exist");
    std::string data;
    char buf[1 << 16];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, r);
    if (f != stdin) fclose(f);
    swo_text_out out;
    swo_maf_stats st;
    cf.check(swo_lift_maf_text(cf.handle(), data.data(), data.size(), genomebuild.c_str(), &out, &st));
    FILE *fo = (outfile == "-" || outfile.empty()) ? stdout : fopen(outfile.c_str(), "wb");
    if (!fo) throw Error(SWO_E_IO, "cannot open " + outfile);
    fwrite(out.lifted, 1, out.lifted_len, fo);
    if (fo != stdout) fclose(fo);
    else fflush(fo);
    FILE *fu = fopen(unmatched.c_str(), "wb");
    if (!fu) throw Error(SWO_E_IO, "cannot open " + unmatched);
    fwrite(out.unmatched, 1, out.unmatched_len, fu);
    fclose(fu);
    // maf.d:309-316
    fprintf(stderr, "[I::liftMAF] Matched %llu records (%llu multiply, skipped) and %llu unmatched\n",
            (unsigned long long)st.n_matched, (unsigned long long)st.n_multiple, (unsigned long long)st.n_unmatched);
    fprintf(stderr, "[I::liftMAF] %llu REF allele changes (%llu not matching either tumor allele)\n",
            (unsigned long long)st.n_refchg, (unsigned long long)st.n_neither);
    if (st.n_matched == 0)
        fprintf(stderr, "[W::liftMAF] Could you have a chromosome name mismatch between chain file and MAF?\n");
}

}  // namespace swiftover
