// main.cpp — the `swiftover` command line over the CUDA engine: same flags, checks and
// exit codes as the reference (/root/reference/source/swiftover/main.d:28-195).
//   -t|--type bed|maf|vcf   -c|--chainfile F   -g|--genome F   -b|--build S
//   -i|--infile F (- or omitted: stdin)    -o|--outfile F (- or omitted: stdout)   -u|--unmatched F
#include <sys/stat.h>

#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "chainfile.hpp"

using namespace swiftover;

static void usage() {  // defaultGetoptPrinter, main.d:74-75
    fprintf(stdout,
            "swift liftover (version: B200 block table)\n"
            "-t      --type Required: File type: bed|maf|vcf\n"
            "-c --chainfile Required: UCSC-format chain file\n"
            "-g    --genome           Genome FASTA (destination build; req. for MAF/VCF)\n"
            "-b     --build           Genome build name (destination; req. for MAF)\n"
            "-i    --infile           Input file; - or omit for stdin\n"
            "-o   --outfile           Output file; - or omit for stdout\n"
            "-u --unmatched           Unmatched output file\n"
            "-h      --help           This help information.\n");
}
static bool exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
static bool ends_with(const std::string &s, const char *suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

int main(int argc, char **argv) {
    std::string type, chain, genome, build, in, out, unm;
    bool help = false, bad = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], v;
        auto eq = a.find('=');
        bool has_v = false;
        if (a.rfind("--", 0) == 0 && eq != std::string::npos) {
            v = a.substr(eq + 1);
            a = a.substr(0, eq);
            has_v = true;
        }
        auto val = [&]() -> std::string {
            if (has_v) return v;
            if (i + 1 < argc) return argv[++i];
            bad = true;
            return "";
        };
        if (a == "-t" || a == "--type") type = val();
        else if (a == "-c" || a == "--chainfile") chain = val();
        else if (a == "-g" || a == "--genome") genome = val();
        else if (a == "-b" || a == "--build") build = val();
        else if (a == "-i" || a == "--infile") in = val();
        else if (a == "-o" || a == "--outfile") out = val();
        else if (a == "-u" || a == "--unmatched") unm = val();
        else if (a == "-h" || a == "--help") help = true;
        else bad = true;
    }
    if (bad || help || type.empty() || chain.empty()) {  // main.d:54-85: help, return 1
        usage();
        return 1;
    }
    try {
        if (!exists(chain)) throw Error(SWO_E_IO, "Chainfile does not exist");  // main.d:91-92
        if (ends_with(chain, ".gz") || ends_with(chain, ".bz2") || ends_with(chain, ".Z") || ends_with(chain, ".zip")) {
            fprintf(stderr, "Only uncompressed chainfiles are supported (for now)\n");  // main.d:94-99
            return -1;
        }
        if (in != "-" && !in.empty() && !exists(in)) throw Error(SWO_E_IO, "Input file does not exist");  // :100
        if (out != "-" && !out.empty()) {  // test write, main.d:103-107
            FILE *f = fopen(out.c_str(), "w");
            if (!f) throw Error(SWO_E_IO, "cannot write " + out);
            fclose(f);
        }
        if (!unm.empty()) {
            FILE *f = fopen(unm.c_str(), "w");
            if (!f) throw Error(SWO_E_IO, "cannot write " + unm);
            fclose(f);
        }
        if (unm.empty()) throw Error(SWO_E_ARG, "<unmatched> output file required");  // main.d:114-115
        if (type == "bed") {
            liftBED(chain, in, out, unm);  // main.d:121
        } else if (type == "vcf") {
            if (genome.empty()) throw Error(SWO_E_ARG, "Genome FASTA (for the destination build) required.");
            if (!exists(genome)) throw Error(SWO_E_IO, "Genome FASTA does not exist");
            if (in.empty() || in == "-" || out.empty() || out == "-")  // vcf.d:42 "need dhtslib to support stdin/stdout"
                throw Error(SWO_E_ARG, "VCF liftover needs -i and -o file names");
            liftVCF(chain, genome, in, out, unm);  // main.d:130-134
        } else if (type == "maf") {
            if (genome.empty()) throw Error(SWO_E_ARG, "Genome FASTA (for the destination build) required.");  // main.d:124-127
            if (build.empty()) throw Error(SWO_E_ARG, "Genome build name (for the destination) required.");
            if (!exists(genome)) throw Error(SWO_E_IO, "Genome FASTA does not exist");
            liftMAF(chain, genome, build, in, out, unm);  // main.d:128
        } else {
            throw Error(SWO_E_ARG, "Unknown file type. Use \"bed\" or \"vcf\".");  // main.d:136
        }
    } catch (const Error &e) {  // an uncaught D exception: message on stderr, exit status 1
        fprintf(stderr, "swiftover: %s\n", e.what());
        return 1;
    }
    return 0;
}
