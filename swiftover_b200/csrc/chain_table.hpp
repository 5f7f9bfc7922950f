// chain_table.hpp — host-side chain-file parser and flattened block table.
//
// Replaces the build phase of the reference's ChainFile constructor
// (/root/reference/source/swiftover/chain.d:509-597) and Chain constructor
// (chain.d:272-405).  Instead of one splay/AVL/IITree per source contig
// (chain.d:537-548) the blocks of every contig are laid out as ONE flat table,
// sorted per contig by the reference's opCmp order (tStart, tEnd, qStart)
// (chain.d:147-164), with a prefix-max-of-tEnd augmentation so that a half-open
// overlap query is two monotone searches plus a forward scan.
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

namespace swo {

// One alignment block as the device sees it: a single 128-bit load.
// meta = (qcid << 1) | (invert == -1).   (ChainLink, chain.d:107-118)
struct alignas(16) Block {
    int32_t tStart, tEnd, qStart, meta;
};

struct ChainTable {
    // source-build ("target") contigs, in order of first appearance
    std::vector<std::string> tnames;
    std::unordered_map<std::string, int> tindex;
    std::vector<int32_t> toff;  // [ntc+1] block range of each contig
    std::vector<int64_t> tsizes;  // tSize from the chain headers (chain.d:298)

    std::vector<Block> blocks;      // [B] payload, contig-major, opCmp order
    std::vector<int32_t> tStartKey; // [B] = blocks[i].tStart (search key)
    std::vector<int32_t> pmaxKey;   // [B] prefix max of tEnd inside the contig
    bool disjoint = true;           // no two blocks of a contig overlap

    // Upper search levels (staged in shared memory by the kernels): for every
    // contig the maxima pmaxKey[a + i*S + S-1] of its full groups of S = 1<<eyShift
    // consecutive blocks, stored in Eytzinger (BFS) order as {key, block index}
    // pairs at ey[eyoff[c] ...].  S is the smallest power of two for which all
    // samples fit in EY_CAP entries.
    static constexpr int EY_CAP = 2048;
    int eyShift = 0;
    std::vector<int32_t> eyKey, eyIdx;  // [nsamples] each
    std::vector<int32_t> eyoff;         // [ntc+1]

    // destination ("query") contigs: ChainFile.contigNames (chain.d:500)
    std::vector<std::string> qnames;
    // ChainFile.qContigSizes (chain.d:504), first-insertion order
    std::vector<std::pair<std::string, int64_t>> qsizes;

    int ntc() const { return (int)tnames.size(); }
    int64_t nblocks() const { return (int64_t)blocks.size(); }
    int tcontig_id(std::string_view name) const;
};

// Returns true on success; on failure `err` holds the message.
bool parse_chain_text(std::string_view text, ChainTable &out, std::string &err);

// Name hash shared by host table construction and the device text parser.
inline uint32_t name_hash(const char *s, size_t n) {
    uint32_t h = 0x9E3779B9u;
    for (size_t i = 0; i < n; i++) h = (h ^ (uint8_t)s[i]) * 0x01000193u;
    return h ^ (h >> 15);
}

}  // namespace swo
