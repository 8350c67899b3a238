// g2u_rgfa.hpp — host-side build of the tables gaf2unstable needs from a minigraph rGFA.
//
// Runs once per process on the host (table build, not the per-record
// path) and produces flat arrays that are uploaded for the kernels:
//
//   * stable contig -> its nodes sorted by SO offset             (get_unstable_mapping,
//     reference gaf2unstable_main.cpp:34-68; node length = length of the sequence column,
//     not the LN tag; nodes with an already seen offset on the same contig are dropped,
//     as std::set::insert does)
//   * node -> reference contig                                   (rgfa2contig,
//     reference rgfa-split.cpp:35-161: rank-0 nodes take their SN with an "id=…|" prefix
//     stripped, rank>0 nodes inherit the unique contig of already assigned neighbours,
//     in rank order)
//
// Line scanning follows the vendored gfakluge semantics the reference relies on
// (gfakluge.hpp:757-824 S lines, :826-967 L/E lines): a line is recognised by its first
// byte, tokens split on TAB *or* space, every delimiter starts a new token, optional
// fields are key:type:value with the value allowed to contain ':'.
#pragma once
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace g2p {

struct RgfaNode {
    std::string name;
    int64_t offset;
    int64_t length;
};
struct RgfaNodeLess {
    bool operator()(const RgfaNode& a, const RgfaNode& b) const { return a.offset < b.offset; }
};

struct RgfaTables {
    // The same container type and insertion sequence as the reference, so that iterating it
    // (the -o node-lengths file, gaf2unstable_main.cpp:280-284) yields the same row order.
    std::unordered_map<std::string, std::set<RgfaNode, RgfaNodeLess>> mapping;
    std::unordered_map<int64_t, int64_t> node_to_contig;   // node id -> reference contig id
    std::vector<std::string> ref_contigs;                  // reference contig id -> name
    std::string error;       // message for exit(1) conditions
    int exit_code = 0;       // 0 ok, 1 = reference exit(1), 134 = reference abort
};

namespace rgfa_detail {

struct Tok {
    const char* p;
    size_t n;
    std::string str() const { return std::string(p, n); }
};

// tokens of the line starting at buf[i] (first byte is the record type); returns index of the
// terminating '\n' (or size).  Matches the scanner loops at gfakluge.hpp:772-782 / :835-848:
// the type byte itself is not part of token 0, a NUL byte ends a token like a delimiter.
inline size_t line_tokens(const char* buf, size_t size, size_t i, std::vector<Tok>& toks) {
    toks.clear();
    size_t a = i + 1;
    size_t j = a;
    for (;;) {
        if (j >= size) { toks.push_back(Tok{buf + a, j - a}); return size; }
        char c = buf[j];
        if (c == '\n') { toks.push_back(Tok{buf + a, j - a}); return j; }
        if (c == '\t' || c == ' ' || c == 0) { toks.push_back(Tok{buf + a, j - a}); a = j + 1; }
        ++j;
    }
}

inline bool all_digits(const Tok& t) {
    if (t.n == 0) return false;
    for (size_t i = 0; i < t.n; ++i) if (t.p[i] < '0' || t.p[i] > '9') return false;
    return true;
}

// key:type:value ; false when the token has fewer than two ':' separated parts
inline bool split_tag(const Tok& t, std::string& key, std::string& val) {
    const char* c1 = static_cast<const char*>(memchr(t.p, ':', t.n));
    if (!c1) return false;
    const char* c2 = static_cast<const char*>(memchr(c1 + 1, ':', t.n - (c1 + 1 - t.p)));
    key.assign(t.p, c1 - t.p);
    if (!c2) { val.clear(); return true; }   // "k:t" -> value is the empty join
    val.assign(c2 + 1, t.p + t.n - (c2 + 1));
    return true;
}

inline bool stol_ok(const std::string& s, int64_t& v) {
    try { v = std::stol(s); return true; } catch (...) { return false; }
}

// rgfa-split.hpp:79-83 node_id
inline bool node_id_of(const std::string& name, int64_t& id) {
    size_t off = name.find('s') + 1;   // npos + 1 == 0
    return stol_ok(name.substr(off), id);
}

}  // namespace rgfa_detail

// Builds everything from the rGFA text.  On a condition where the reference dies, exit_code
// is set (134 for asserts / uncaught exceptions, 1 for its explicit exit(1) paths).
inline void build_rgfa_tables(const char* buf, size_t size, RgfaTables& T) {
    using namespace rgfa_detail;
    std::vector<Tok> toks;
    std::map<int64_t, std::list<int64_t>> rank_to_nodes;
    std::unordered_map<int64_t, int64_t> node_to_rank;
    std::unordered_multimap<int64_t, int64_t> edges;
    std::unordered_map<std::string, int64_t> contig_ids;
    auto die = [&](int code, const std::string& msg) { if (!T.exit_code) { T.exit_code = code; T.error = msg; } };

    // ---- pass 1: S lines -> stable mapping (gaf2unstable_main.cpp:34-68)
    // ---- pass 2: S lines -> ranks and rank-0 contigs (rgfa-split.cpp:54-91)
    // The reference scans the file twice; both visitors see the same lines in the same
    // order, so one scan feeds both, with pass-1 failures taking precedence.
    struct Seg { std::string name, sn_raw; int64_t so, sr; bool has_sr; };
    std::vector<Seg> segs;
    for (size_t i = 0; i < size; ++i) {
        if (buf[i] != 'S' || !(i == 0 || buf[i - 1] == '\n')) continue;
        size_t eol = line_tokens(buf, size, i, toks);
        if (toks.size() < 3) { die(134, "short S line"); return; }
        Seg s;
        s.name = toks[1].str();
        size_t tag_index = 3;
        int64_t seq_len;
        if (all_digits(toks[2])) {   // gfakluge treats this as a GFA2 "S name len seq" line (:788-792)
            if (toks.size() < 4) { die(134, "short GFA2 S line"); return; }
            seq_len = (int64_t)toks[3].n;
            tag_index = 4;
        } else {
            seq_len = (int64_t)toks[2].n;
        }
        bool has_sn = false, has_so = false;
        s.has_sr = false; s.so = 0; s.sr = 0;
        bool dup_sr = false;
        if (toks.size() > 3) {
            for (size_t j = tag_index; j < toks.size(); ++j) {
                std::string key, val;
                if (!split_tag(toks[j], key, val)) { die(134, "malformed optional field on S line"); return; }
                if (key == "SN") {
                    if (has_sn) { die(134, "assert(found_SN == false)"); return; }
                    s.sn_raw = val; has_sn = true;
                } else if (key == "SO") {
                    if (has_so) { die(134, "assert(found_SO == false)"); return; }
                    if (!stol_ok(val, s.so)) { die(134, "stol(SO)"); return; }
                    if (s.so < 0) { die(134, "assert(offset >= 0)"); return; }
                    has_so = true;
                } else if (key == "SR") {
                    if (s.has_sr) dup_sr = true;
                    else if (!stol_ok(val, s.sr)) { die(134, "stol(SR)"); return; }
                    s.has_sr = true;
                }
            }
        }
        if (!has_sn || !has_so) { die(134, "assert(found_SN && found_SO)"); return; }
        T.mapping[s.sn_raw].insert(RgfaNode{s.name, s.so, seq_len});
        if (dup_sr) s.sr = INT64_MIN;   // pass 2 asserts on it
        segs.push_back(std::move(s));
        i = eol;
    }
    for (const Seg& s : segs) {
        int64_t id;
        if (!node_id_of(s.name, id)) { die(134, "stol in node_id"); return; }
        if (s.sr == INT64_MIN) { die(134, "assert(found_SR == false)"); return; }
        if (!s.has_sr) { die(134, "assert(found_SR)"); return; }
        if (s.sr < 0) { die(134, "assert(rank >= 0)"); return; }
        // strip_prefix (rgfa-split.cpp:12-19)
        std::string contig = s.sn_raw;
        if (contig.compare(0, 3, "id=") == 0) {
            size_t p = contig.find('|', 3);
            if (p == std::string::npos) { die(134, "assert(p != npos) in strip_prefix"); return; }
            contig = contig.substr(p + 1);
        }
        rank_to_nodes[s.sr].push_back(id);
        node_to_rank[id] = s.sr;
        if (s.sr == 0) {
            auto it = contig_ids.find(contig);
            int64_t cid;
            if (it != contig_ids.end()) cid = it->second;
            else { cid = (int64_t)contig_ids.size(); contig_ids[contig] = cid; T.ref_contigs.push_back(contig); }
            T.node_to_contig[id] = cid;
        }
    }
    // ---- pass 3: L (and E) lines -> undirected adjacency (rgfa-split.cpp:94-99)
    for (size_t i = 0; i < size; ++i) {
        if (!(i == 0 || buf[i - 1] == '\n')) continue;
        const char t = buf[i];
        if (t != 'L' && t != 'E' && t != 'C') continue;
        size_t eol = line_tokens(buf, size, i, toks);
        std::string a, b;
        if (t == 'L') {
            if (toks.size() < 5) { die(134, "short L line"); return; }
            a = toks[1].str(); b = toks[3].str();
        } else if (t == 'E') {
            if (toks.size() < 9 || toks[2].n == 0 || toks[3].n == 0) { die(134, "short E line"); return; }
            a.assign(toks[2].p, toks[2].n - 1); b.assign(toks[3].p, toks[3].n - 1);
        }   // 'C': empty names -> node_id throws in the reference
        int64_t ia, ib;
        if (!node_id_of(a, ia) || !node_id_of(b, ib)) { die(134, "stol in node_id (edge)"); return; }
        edges.insert(std::make_pair(ia, ib));
        edges.insert(std::make_pair(ib, ia));
        i = eol;
    }
    // ---- contigs of rank>0 nodes, in rank order (rgfa-split.cpp:108-158)
    for (auto& rn : rank_to_nodes) {
        const int64_t rank = rn.first;
        if (rank <= 0) continue;
        std::list<int64_t>& todo = rn.second;
        int64_t pushes = 0;
        while (!todo.empty()) {
            const int64_t node = todo.back();
            todo.pop_back();
            std::unordered_map<int64_t, int64_t> counts;
            auto range = edges.equal_range(node);
            for (auto e = range.first; e != range.second; ++e) {
                const int64_t other = e->second;
                const int64_t orank = node_to_rank[other];
                if (orank < rank || (orank == rank && T.node_to_contig.count(other))) ++counts[T.node_to_contig[other]];
            }
            if (counts.empty()) {
                todo.push_front(node);
                ++pushes;
                if (pushes > (int64_t)todo.size()) {
                    std::string m = "[error] Unable to assign contigs for the following nodes at rank " + std::to_string(rank) + ":\n";
                    for (int64_t x : todo) m += std::to_string(x) + "\n";
                    die(1, m);
                    return;
                }
            } else if (counts.size() > 1) {
                std::string m = "[error] Conflict found for node \"" + std::to_string(node) + "\" with rank \"" + std::to_string(rank) + ":\n";
                for (auto& c : counts) m += "\tcontig=" + T.ref_contigs[c.first] + " count=" + std::to_string(c.second) + "\n";
                die(1, m);
                return;
            } else {
                T.node_to_contig[node] = counts.begin()->first;
                pushes = 0;
            }
        }
    }
}

}  // namespace g2p
