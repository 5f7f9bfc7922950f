// chain_table.cpp — see chain_table.hpp.  Host C++ only (no CUDA).
#include "chain_table.hpp"

#include <algorithm>
#include <charconv>
#include <climits>
#include <numeric>

namespace swo {

namespace {

inline bool ascii_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// std.conv.to!int on one token: optional sign, digits, nothing else, int32.
bool token_to_i32(std::string_view t, int32_t &v) {
    if (t.empty()) return false;
    const char *b = t.data(), *e = b + t.size();
    if (*b == '+') {
        ++b;
        if (b == e || *b == '-' || *b == '+') return false;
    }
    auto r = std::from_chars(b, e, v, 10);
    return r.ec == std::errc() && r.ptr == e;
}

struct RawBlock {
    int64_t tStart, tEnd, qStart;
    int32_t qcid;
    int8_t invert;
    uint32_t ord;  // insertion order inside the contig
};

struct Builder {
    ChainTable &tab;
    std::vector<std::vector<RawBlock>> per_contig;
    std::unordered_map<std::string, int> qindex;
    std::unordered_map<std::string, size_t> qsize_index;

    int target(std::string_view name) {  // chainsByContig.require  chain.d:537
        auto it = tab.tindex.find(std::string(name));
        if (it != tab.tindex.end()) return it->second;
        int id = (int)tab.tnames.size();
        tab.tnames.emplace_back(name);
        tab.tindex.emplace(std::string(name), id);
        tab.tsizes.push_back(0);
        per_contig.emplace_back();
        return id;
    }
    int query(std::string_view name) {  // ContigIDorAdd  chain.d:799-805
        auto it = qindex.find(std::string(name));
        if (it != qindex.end()) return it->second;
        int id = (int)tab.qnames.size();
        tab.qnames.emplace_back(name);
        qindex.emplace(std::string(name), id);
        return id;
    }
    void qsize(std::string_view name, int64_t size) {  // chain.d:551,579
        auto it = qsize_index.find(std::string(name));
        if (it != qsize_index.end()) {
            tab.qsizes[it->second].second = size;
            return;
        }
        qsize_index.emplace(std::string(name), tab.qsizes.size());
        tab.qsizes.emplace_back(std::string(name), size);
    }
};

std::string clip(std::string_view s) { return std::string(s.substr(0, 80)); }

// One chain: lines[0] is the header.  chain.d:272-405
bool add_chain(Builder &b, const std::vector<std::string_view> &lines, size_t first, size_t last,
               std::string &err) {
    std::string_view hdr = lines[first];
    if (!hdr.empty() && hdr.back() == '\r') hdr.remove_suffix(1);
    std::string_view f[13];
    size_t nf = 0, pos = 0;
    while (nf < 13) {  // split on single spaces  chain.d:293
        size_t sp = hdr.find(' ', pos);
        f[nf++] = hdr.substr(pos, sp == std::string_view::npos ? std::string_view::npos : sp - pos);
        if (sp == std::string_view::npos) break;
        pos = sp + 1;
    }
    int32_t tSize, tS, tE, qSize, qS, qE, id;
    if (nf < 13 || f[0] != "chain" || f[4].empty() || f[9].empty() || !token_to_i32(f[3], tSize) ||
        !token_to_i32(f[5], tS) || !token_to_i32(f[6], tE) || !token_to_i32(f[8], qSize) ||
        !token_to_i32(f[10], qS) || !token_to_i32(f[11], qE) || !token_to_i32(f[12], id)) {
        err = "malformed chain header: " + clip(hdr);
        return false;
    }
    // the 13th field ends at the next space, if any (extra fields are ignored by the reference)
    const char tStrand = f[4][0], qStrand = f[9][0];
    int32_t tFrom = tStrand == '+' ? tS : tSize - tS;  // chain.d:300-310
    int32_t qFrom = qStrand == '+' ? qS : qSize - qS;  // chain.d:315-325
    const int invert = tStrand == qStrand ? 1 : -1;    // chain.d:330
    const std::string_view tName = f[2], qName = f[7];

    for (size_t li = first + 1; li < last; li++) {
        std::string_view ln = lines[li];
        if (ln.empty()) continue;  // chain.d:344
        int32_t d[3];
        int nd = 0;
        bool extra = false;
        size_t i = 0;
        while (i < ln.size()) {
            while (i < ln.size() && ascii_space(ln[i])) i++;
            if (i == ln.size()) break;
            size_t a = i;
            while (i < ln.size() && !ascii_space(ln[i])) i++;
            if (nd == 3) {
                extra = true;
                break;
            }
            if (!token_to_i32(ln.substr(a, i - a), d[nd++])) {
                err = "bad integer in chain data line: " + clip(ln);
                return false;
            }
        }
        if (nd == 0) continue;
        if (extra || nd == 2) {  // chain.d:403
            err = "unexpected length of alignment data line: " + clip(ln);
            return false;
        }
        const int qcid = b.query(qName);  // chain.d:364
        const int t = b.target(tName);
        b.tab.tsizes[t] = tSize;
        RawBlock rb;
        rb.tStart = tFrom;
        rb.tEnd = (int64_t)tFrom + d[0];
        rb.qStart = qFrom;
        rb.qcid = qcid;
        rb.invert = (int8_t)invert;
        rb.ord = (uint32_t)b.per_contig[t].size();
        const int64_t qEnd = rb.qStart + (int64_t)invert * d[0];
        if (rb.tEnd > INT32_MAX || qEnd > INT32_MAX || qEnd < INT32_MIN || rb.tEnd < INT32_MIN) {
            err = "chain block outside the 32-bit coordinate range: " + clip(ln);
            return false;
        }
        b.per_contig[t].push_back(rb);
        if (nd == 3) {  // chain.d:397-401
            const int64_t nt = (int64_t)tFrom + d[0] + d[1];
            const int64_t nq = (int64_t)qFrom + (int64_t)invert * ((int64_t)d[0] + d[2]);
            if (nt > INT32_MAX || nt < INT32_MIN || nq > INT32_MAX || nq < INT32_MIN) {
                err = "chain coordinates overflow 32 bits: " + clip(ln);
                return false;
            }
            tFrom = (int32_t)nt;
            qFrom = (int32_t)nq;
        }
    }
    b.qsize(qName, qSize);
    return true;
}

}  // namespace

int ChainTable::tcontig_id(std::string_view name) const {
    auto it = tindex.find(std::string(name));
    return it == tindex.end() ? -1 : it->second;
}

// In-order fill of the implicit tree rooted at k (1-based) over m sorted samples.
static void ey_fill(const ChainTable &tab, int a, int S, int m, int k, int &next, int32_t *key, int32_t *idx) {
    if (k > m) return;
    ey_fill(tab, a, S, m, 2 * k, next, key, idx);
    const int j = a + next * S + S - 1;
    key[k - 1] = tab.pmaxKey[j];
    idx[k - 1] = j;
    next++;
    ey_fill(tab, a, S, m, 2 * k + 1, next, key, idx);
}

static void build_eytzinger(ChainTable &tab) {
    int sh = 0;
    for (;; sh++) {
        size_t tot = 0;
        for (int c = 0; c < tab.ntc(); c++) tot += (size_t)(tab.toff[c + 1] - tab.toff[c]) >> sh;
        if (tot <= (size_t)ChainTable::EY_CAP) break;
    }
    tab.eyShift = sh;
    tab.eyKey.clear();
    tab.eyIdx.clear();
    tab.eyoff.assign(1, 0);
    for (int c = 0; c < tab.ntc(); c++) {
        const int a = tab.toff[c], m = (tab.toff[c + 1] - a) >> sh;
        const size_t base = tab.eyKey.size();
        tab.eyKey.resize(base + m);
        tab.eyIdx.resize(base + m);
        int next = 0;
        ey_fill(tab, a, 1 << sh, m, 1, next, tab.eyKey.data() + base, tab.eyIdx.data() + base);
        tab.eyoff.push_back((int32_t)tab.eyKey.size());
    }
}

bool parse_chain_text(std::string_view text, ChainTable &tab, std::string &err) {
    tab = ChainTable();
    // byLineCopy (chain.d:520): '\n'-terminated, no empty line after a final '\n'
    std::vector<std::string_view> lines;
    for (size_t a = 0; a < text.size();) {
        size_t nl = text.find('\n', a);
        if (nl == std::string_view::npos) nl = text.size();
        lines.push_back(text.substr(a, nl - a));
        a = nl + 1;
    }
    if (lines.empty()) {
        err = "empty chain file";
        return false;
    }
    Builder b{tab, {}, {}, {}};
    size_t chainStart = 0;
    for (size_t i = 0; i < lines.size(); i++) {
        if (lines[i].size() > 5 && lines[i].substr(0, 5) == "chain") {  // chain.d:526
            // chain.d:528-530: the slice [chainStart, i) is a chain only when i-1 > 0
            if (i >= 2 && !add_chain(b, lines, chainStart, i, err)) return false;
            chainStart = i;
        }
    }
    if (!add_chain(b, lines, chainStart, lines.size(), err)) return false;  // chain.d:556-579

    // flatten: per contig sort by (tStart,tEnd,qStart); an opCmp-equal block is
    // the same tree node, the first one inserted survives.
    tab.toff.assign(1, 0);
    tab.disjoint = true;
    for (auto &v : b.per_contig) {
        std::sort(v.begin(), v.end(), [](const RawBlock &x, const RawBlock &y) {
            if (x.tStart != y.tStart) return x.tStart < y.tStart;
            if (x.tEnd != y.tEnd) return x.tEnd < y.tEnd;
            if (x.qStart != y.qStart) return x.qStart < y.qStart;
            return x.ord < y.ord;
        });
        int64_t pm = INT64_MIN;
        const RawBlock *prev = nullptr;
        for (const RawBlock &r : v) {
            if (prev && prev->tStart == r.tStart && prev->tEnd == r.tEnd && prev->qStart == r.qStart) continue;
            if (r.qcid > 0x7FFF) {
                err = "more than 32768 destination contigs";
                return false;
            }
            if (prev && r.tStart < pm) tab.disjoint = false;
            pm = std::max(pm, r.tEnd);
            tab.blocks.push_back(Block{(int32_t)r.tStart, (int32_t)r.tEnd, (int32_t)r.qStart,
                                       (int32_t)((r.qcid << 1) | (r.invert < 0 ? 1 : 0))});
            tab.tStartKey.push_back((int32_t)r.tStart);
            tab.pmaxKey.push_back((int32_t)pm);
            prev = &r;
        }
        if (tab.blocks.size() > (size_t)INT32_MAX) {
            err = "too many blocks";
            return false;
        }
        tab.toff.push_back((int32_t)tab.blocks.size());
    }
    build_eytzinger(tab);
    return true;
}

}  // namespace swo
