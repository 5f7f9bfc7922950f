// chainfile.hpp — C++ host mirror of the reference's `struct ChainFile`
// (/root/reference/source/swiftover/chain.d:485-714) on top of the C ABI
// (include/swiftover_cuda.h).  Same member names and meaning; the D host would
// bind the same C functions with extern(C) (see swiftover_b200/d/ and
// INTEGRATION.md).  The reference is compiled D; no D toolchain exists in this
// image, so this C++ host is the executable artefact.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "swiftover_cuda.h"

namespace swiftover {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// one lifted piece of an interval (what liftBED prints per ChainLink, bed.d:161-170)
struct LiftedInterval {
    std::string contig;
    long start, end;
    char strand;  // after the cumulative STRAND_TABLE flip (bed.d:173); 0 if none given
};

class ChainFile {
  public:
    std::vector<std::string> contigNames;         // chain.d:500
    std::map<std::string, long> qContigSizes;     // chain.d:504

    explicit ChainFile(const std::string &fn, const std::vector<int> &devices = {}) {  // chain.d:509
        int rc = swo_create(&ctx_, devices.empty() ? nullptr : devices.data(), (int)devices.size());
        if (rc) throw Error(rc, "no CUDA device (libswiftover_cuda has no CPU fallback)");
        rc = swo_load_chain(ctx_, fn.c_str());
        if (rc) {
            std::string m = swo_last_error(ctx_);
            swo_destroy(ctx_);
            ctx_ = nullptr;
            throw Error(rc, m);
        }
        for (int i = 0; i < swo_num_qcontigs(ctx_); i++) contigNames.emplace_back(swo_qcontig_name(ctx_, i));
        for (int i = 0; i < swo_num_qsizes(ctx_); i++) qContigSizes[swo_qsize_name(ctx_, i)] = (long)swo_qsize_value(ctx_, i);
    }
    ~ChainFile() {
        if (ctx_) swo_destroy(ctx_);
    }
    ChainFile(const ChainFile &) = delete;
    ChainFile &operator=(const ChainFile &) = delete;

    swo_ctx *handle() const { return ctx_; }

    // chain.d:664-713 (+ the sort / orderStartEnd / strand of bed.d:156-173)
    std::vector<LiftedInterval> lift(const std::string &contig, long start, long end, char strand = 0) {
        const int32_t tc = swo_tcontig_id(ctx_, contig.data(), contig.size());
        const int32_t s = (int32_t)start, e = (int32_t)end;
        const uint8_t sb = (uint8_t)strand;
        swo_lifted out;
        check(swo_lift_intervals(ctx_, 1, &tc, &s, &e, strand ? &sb : nullptr, nullptr, nullptr, &out));
        std::vector<LiftedInterval> v;
        for (uint64_t k = 0; k < out.n_rows; k++)
            v.push_back({contigNames[out.rows[k].qcid_strand & 0xFFFF], out.rows[k].start, out.rows[k].end,
                         (char)((out.rows[k].qcid_strand >> 16) & 0xFF)});
        return v;
    }
    // chain.d:637-654
    int liftCoordOnly(const std::string &contig, long &coord) {
        std::string c = contig;
        return liftDirectly(c, coord);
    }
    // chain.d:604-627
    int liftDirectly(std::string &contig, long &coord) {
        const int32_t tc = swo_tcontig_id(ctx_, contig.data(), contig.size());
        const int32_t p = (int32_t)coord;
        int32_t q, op;
        uint8_t hit;
        check(swo_lift_points(ctx_, 1, &tc, &p, &q, &op, &hit));
        if (!hit) return 0;
        contig = contigNames[q];
        coord = op;
        return 1;
    }
    void check(int rc) const {
        if (rc) throw Error(rc, swo_last_error(ctx_));
    }

  private:
    swo_ctx *ctx_ = nullptr;
};

// bed.d:72
void liftBED(const std::string &chainfile, const std::string &infile, const std::string &outfile,
             const std::string &unmatched);

// vcf.d:29-32 (uncompressed text VCF, uncompressed FASTA)
void liftVCF(const std::string &chainfile, const std::string &genomefile, const std::string &infile,
             const std::string &outfile, const std::string &unmatched);
// maf.d:119-121
void liftMAF(const std::string &chainfile, const std::string &genomefile, const std::string &genomebuild,
             const std::string &infile, const std::string &outfile, const std::string &unmatched);

}  // namespace swiftover
