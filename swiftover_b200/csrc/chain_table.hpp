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

#ifndef SWO_EY_CAP
#define SWO_EY_CAP 2048  // Eytzinger samples staged in shared memory (16 KiB)
#endif

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
    static constexpr int EY_CAP = SWO_EY_CAP;
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

// Name hash shared by host table construction and the device text parser: the first 16
// bytes enter as four little-endian words (zero padded), longer names add an FNV tail.
#if defined(__CUDACC__)
#define SWO_HASH_HD __host__ __device__ inline
#else
#define SWO_HASH_HD inline
#endif
SWO_HASH_HD uint32_t name_hash_head(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t len) {
    return (w0 * 0x9E3779B1u) ^ (w1 * 0x85EBCA77u) ^ (w2 * 0xC2B2AE3Du) ^ (w3 * 0x27D4EB2Fu) ^ (len * 0x165667B1u);
}
SWO_HASH_HD uint32_t name_hash_tail(uint32_t h, uint8_t c) { return (h ^ c) * 0x01000193u; }
SWO_HASH_HD uint32_t name_hash_fin(uint32_t h) {
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    return h ^ (h >> 12);
}
inline void name_words(const char *s, size_t n, uint32_t w[4]) {
    w[0] = w[1] = w[2] = w[3] = 0;
    for (size_t i = 0; i < n && i < 16; i++) w[i >> 2] |= (uint32_t)(uint8_t)s[i] << (8 * (i & 3));
}
inline uint32_t name_hash(const char *s, size_t n) {
    uint32_t w[4];
    name_words(s, n, w);
    uint32_t h = name_hash_head(w[0], w[1], w[2], w[3], (uint32_t)n);
    for (size_t i = 16; i < n; i++) h = name_hash_tail(h, (uint8_t)s[i]);
    return name_hash_fin(h);
}

}  // namespace swo
