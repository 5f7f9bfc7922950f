// genome.hpp — destination genome for the VCF REF check: FASTA -> 2 bit/base + side channels.
//
// Replaces dhtslib's IndexedFastaFile (/root/reference/source/swiftover/vcf.d:38-39: hasSeq,
// seqLen, fetchSequence).  The reference compares REF with the fetched bases byte for byte
// (vcf.d:144) and writes whatever the FASTA holds, so case and non-ACGT symbols must survive
// the packing:
//   packed : 2 bit/base (A=0 C=1 G=2 T=3), base b of destination contig q at g = off[q] + b
//   lower  : 1 bit/base, set for soft-masked (lower-case) bases            (absent if none)
//   exc    : runs {start g, length, byte} of every other symbol (N, IUPAC codes), sorted,
//            plus a bitmap of the 4096-base chunks that contain one        (absent if none)
// Sequences are keyed by the chain's destination contig id; FASTA records the chain does not
// name are ignored, chain contigs the FASTA lacks have length 0 (hasSeq == false).
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace swo {

struct GenomeStore {
    std::vector<int64_t> off, len;  // [nq]
    std::vector<uint8_t> packed, lower, exc_chunk;
    std::vector<int64_t> exc_start;
    std::vector<int32_t> exc_len;
    std::vector<uint8_t> exc_byte;
    bool any_lower = false;
    std::unordered_map<std::string, int64_t> fasta_len;  // every FASTA record: name -> length

    bool has_seq(std::string_view name) const { return fasta_len.count(std::string(name)) != 0; }
    int64_t seq_len(std::string_view name) const {
        auto it = fasta_len.find(std::string(name));
        return it == fasta_len.end() ? -1 : it->second;
    }
};

// qnames: destination contigs of the chain (id order).  The FASTA is read twice: once to measure its
// records, once to pack them — a contig's slot is as long as its FASTA record, whatever size the chain
// header states (the reference only warns on a mismatch, vcf.d:83-86).  qsizes is not used any more.
bool genome_from_fasta(const char *path, const std::vector<std::string> &qnames, const std::vector<int64_t> &qsizes,
                       GenomeStore &out, std::string &err);
bool genome_from_fasta_text(std::string_view text, const std::vector<std::string> &qnames,
                            const std::vector<int64_t> &qsizes, GenomeStore &out, std::string &err);

}  // namespace swo
