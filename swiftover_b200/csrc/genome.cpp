// genome.cpp — see genome.hpp.  Host C++ only.
#include "genome.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace swo {

namespace {

struct Packer {
    const std::vector<std::string> &qnames;
    GenomeStore &g;
    std::unordered_map<std::string, int> qid;
    int cur = -1;          // destination contig id of the record being read, -1 = ignored record
    std::string cur_name;  // FASTA record name
    int64_t cur_len = 0;
    int64_t cap = 0;       // bases reserved for cur
    // open exception run
    int64_t run_start = -1;
    int32_t run_len = 0;
    uint8_t run_byte = 0;

    void flush_run() {
        if (run_start >= 0) {
            g.exc_start.push_back(run_start);
            g.exc_len.push_back(run_len);
            g.exc_byte.push_back(run_byte);
        }
        run_start = -1;
        run_len = 0;
    }
    void end_record() {
        flush_run();
        if (!cur_name.empty()) g.fasta_len[cur_name] = cur_len;
        if (cur >= 0) g.len[cur] = cur_len;
        cur = -1;
        cur_name.clear();
        cur_len = 0;
    }
    void begin_record(std::string_view name) {
        end_record();
        cur_name = std::string(name);
        auto it = qid.find(cur_name);
        cur = it == qid.end() ? -1 : it->second;
        cap = cur >= 0 ? ((cur + 1 < (int)g.off.size() ? g.off[cur + 1] : (int64_t)g.packed.size() * 4) - g.off[cur]) : 0;
    }
    bool add_bases(const char *s, size_t n, std::string &err) {
        if (cur < 0) {
            for (size_t i = 0; i < n; i++) cur_len += !(s[i] == ' ' || (s[i] >= 9 && s[i] <= 13));
            return true;
        }
        for (size_t i = 0; i < n; i++) {
            const unsigned char c = (unsigned char)s[i];
            if (c == ' ' || (c >= 9 && c <= 13)) continue;
            if (cur_len >= cap) {
                err = "FASTA sequence " + cur_name + " changed between the two passes over the file";
                return false;
            }
            const int64_t p = g.off[cur] + cur_len++;
            int code = -1;
            switch (c | 0x20) {
                case 'a': code = 0; break;
                case 'c': code = 1; break;
                case 'g': code = 2; break;
                case 't': code = 3; break;
            }
            if (code >= 0) {
                g.packed[p >> 2] |= (uint8_t)(code << ((p & 3) * 2));
                if (c & 0x20) {
                    g.lower[p >> 3] |= (uint8_t)(1u << (p & 7));
                    g.any_lower = true;
                }
                if (run_start >= 0) flush_run();
            } else {
                if (run_start >= 0 && run_byte == c && run_start + run_len == p && run_len < 0x7fffffff) {
                    run_len++;
                } else {
                    flush_run();
                    run_start = p;
                    run_len = 1;
                    run_byte = c;
                }
                g.exc_chunk[p >> 15] |= (uint8_t)(1u << ((p >> 12) & 7));
            }
        }
        return true;
    }
};

void layout(const std::vector<std::string> &qnames, const std::vector<int64_t> &qsizes, GenomeStore &g) {
    const size_t nq = qnames.size();
    g = GenomeStore();
    g.off.assign(nq, 0);
    g.len.assign(nq, 0);
    int64_t tot = 0;
    for (size_t i = 0; i < nq; i++) {
        g.off[i] = tot;
        tot += (std::max<int64_t>(qsizes[i], 0) + 31) / 32 * 32;
    }
    g.packed.assign((size_t)(tot / 4) + 16, 0);
    g.lower.assign((size_t)(tot / 8) + 16, 0);
    g.exc_chunk.assign((size_t)(tot >> 15) + 16, 0);
}

// FASTA records come in file order, the layout is in destination-contig-id order: sort the runs
void sort_runs(GenomeStore &g) {
    const size_t n = g.exc_start.size();
    if (std::is_sorted(g.exc_start.begin(), g.exc_start.end())) return;
    std::vector<size_t> ord(n);
    for (size_t i = 0; i < n; i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return g.exc_start[a] < g.exc_start[b]; });
    std::vector<int64_t> s(n);
    std::vector<int32_t> l(n);
    std::vector<uint8_t> b(n);
    for (size_t i = 0; i < n; i++) {
        s[i] = g.exc_start[ord[i]];
        l[i] = g.exc_len[ord[i]];
        b[i] = g.exc_byte[ord[i]];
    }
    g.exc_start.swap(s);
    g.exc_len.swap(l);
    g.exc_byte.swap(b);
}

}  // namespace

namespace {

// stream the FASTA text through a packer (header lines start records, everything else is bases)
bool feed_text(std::string_view text, Packer &pk, std::string &err) {
    size_t i = 0;
    while (i < text.size()) {
        size_t e = text.find('\n', i);
        if (e == std::string_view::npos) e = text.size();
        if (text[i] == '>') {
            size_t b = i + 1;
            while (b < e && !(text[b] == ' ' || (text[b] >= 9 && text[b] <= 13))) b++;
            pk.begin_record(text.substr(i + 1, b - i - 1));
        } else if (!pk.add_bases(text.data() + i, e - i, err)) {
            return false;
        }
        i = e + 1;
    }
    pk.end_record();
    return true;
}

bool feed_file(const char *path, Packer &pk, std::string &err) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        err = std::string("Genome FASTA does not exist: ") + path;
        return false;
    }
    std::vector<char> buf(1 << 22);
    std::string header;  // a '>' line may straddle two reads
    bool in_header = false, at_line_start = true, ok = true;
    size_t r;
    while (ok && (r = fread(buf.data(), 1, buf.size(), f)) > 0) {
        size_t i = 0;
        while (i < r && ok) {
            if (in_header) {
                const char *nl = (const char *)memchr(buf.data() + i, '\n', r - i);
                const size_t e = nl ? (size_t)(nl - buf.data()) : r;
                header.append(buf.data() + i, e - i);
                i = e;
                if (nl) {
                    size_t b = 0;
                    while (b < header.size() && !(header[b] == ' ' || (header[b] >= 9 && header[b] <= 13))) b++;
                    pk.begin_record(std::string_view(header).substr(0, b));
                    header.clear();
                    in_header = false;
                    at_line_start = true;
                    i++;
                }
                continue;
            }
            if (at_line_start && buf[i] == '>') {
                in_header = true;
                i++;
                continue;
            }
            const char *nl = (const char *)memchr(buf.data() + i, '\n', r - i);
            const size_t e = nl ? (size_t)(nl - buf.data()) : r;
            ok = pk.add_bases(buf.data() + i, e - i, err);
            at_line_start = nl != nullptr;
            i = nl ? e + 1 : e;
        }
    }
    fclose(f);
    if (!ok) return false;
    if (in_header) {
        size_t b = 0;
        while (b < header.size() && !(header[b] == ' ' || (header[b] >= 9 && header[b] <= 13))) b++;
        pk.begin_record(std::string_view(header).substr(0, b));
    }
    pk.end_record();
    return true;
}

// Two passes over the source: the first only measures every record (all records "ignored"), so that
// each destination contig's slot is sized from the FASTA itself — the reference fetches by FASTA
// coordinates whatever the chain header says about the contig's size, and only warns (and leaves the
// ##contig line out) when the two lengths differ (vcf.d:83-86); the second packs.
template <typename Feed>
bool build(Feed feed, const std::vector<std::string> &qnames, GenomeStore &out, std::string &err) {
    GenomeStore probe;
    Packer measure{qnames, probe, {}, -1, {}, 0, 0, -1, 0, 0};
    if (!feed(measure, err)) return false;
    std::vector<int64_t> sizes(qnames.size(), 0);
    for (size_t i = 0; i < qnames.size(); i++) {
        auto it = probe.fasta_len.find(qnames[i]);
        if (it != probe.fasta_len.end()) sizes[i] = it->second;
    }
    layout(qnames, sizes, out);
    Packer pk{qnames, out, {}, -1, {}, 0, 0, -1, 0, 0};
    for (size_t i = 0; i < qnames.size(); i++) pk.qid.emplace(qnames[i], (int)i);
    if (!feed(pk, err)) return false;
    sort_runs(out);
    return true;
}

}  // namespace

bool genome_from_fasta_text(std::string_view text, const std::vector<std::string> &qnames, const std::vector<int64_t> &,
                            GenomeStore &out, std::string &err) {
    return build([&](Packer &pk, std::string &e) { return feed_text(text, pk, e); }, qnames, out, err);
}

bool genome_from_fasta(const char *path, const std::vector<std::string> &qnames, const std::vector<int64_t> &,
                       GenomeStore &out, std::string &err) {
    return build([&](Packer &pk, std::string &e) { return feed_file(path, pk, e); }, qnames, out, err);
}

}  // namespace swo
