// gaf2unstable_oracle — CPU restatement of the reference's stable -> unstable GAF rewrite.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
// product (libg2p.so, gaf2paf, gaf2unstable); only tests/ and __graft_entry__.smoke() may run
// it, as the checker.
//
// Parity status: PINNED against the reference binary (oracle/_ref/gaf2unstable, compiled from
// /root/reference by oracle/build_ref.sh): the committed known-answer vectors
// (tests/golden/gaf2unstable_kat.json, generated from that binary) and seeded differential runs
// on synthetic rGFA (tests/test_oracle.py) must match byte for byte, including the -o file.
//
// It follows the reference's own structure -- std containers, records materialised into step
// vectors and a tag map -- and shares no code with the device implementation (g2u_core.cuh,
// g2u_rgfa.hpp).  Inputs the reference would die on (asserts, uncaught exceptions) are reported
// as exit code 134 without trying to reproduce the message.
//
// usage: gaf2unstable_oracle -g graph.gfa [-o node-lengths.tsv] <gaf|->
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Abort : std::runtime_error {
    using std::runtime_error::runtime_error;
};
void require(bool ok, const char* what) { if (!ok) throw Abort(what); }

long to_long(const std::string& s) {   // std::stol; throws like the reference would
    size_t pos = 0;
    return std::stol(s, &pos);
}

// ---- rGFA scan (gfakluge.hpp:757-824 S lines, :826-967 L lines): tokens split on TAB or space
std::vector<std::string> tokens_of(const std::string& buf, size_t& i) {
    std::vector<std::string> toks(1);
    while (i < buf.size()) {
        char c = 0;
        while (i + 1 <= buf.size()) {
            ++i;
            if (i >= buf.size()) { c = '\n'; break; }
            c = buf[i];
            if (c == 0 || c == '\t' || c == ' ' || c == '\n') break;
            toks.back().push_back(c);
        }
        if (c == '\n') break;
        toks.emplace_back();
    }
    return toks;
}

struct Tag { std::string key, val; };
Tag split_tag(const std::string& t) {   // key:type:value, value may contain ':' (gfakluge.hpp:806-812)
    size_t c1 = t.find(':');
    require(c1 != std::string::npos, "optional field without ':'");
    size_t c2 = t.find(':', c1 + 1);
    Tag g;
    g.key = t.substr(0, c1);
    g.val = c2 == std::string::npos ? "" : t.substr(c2 + 1);
    return g;
}

struct MGSeq {   // gaf2unstable_main.cpp:18-30
    std::string name;
    int64_t offset = 0, length = 0;
    bool operator<(const MGSeq& o) const { return offset < o.offset; }
};

int64_t node_id(const std::string& name) {   // rgfa-split.hpp:79-83
    return to_long(name.substr(name.find('s') + 1));
}

std::string strip_prefix(const std::string& sn) {   // rgfa-split.cpp:12-19
    if (sn.compare(0, 3, "id=") == 0) {
        size_t p = sn.find('|', 3);
        require(p != std::string::npos, "strip_prefix");
        return sn.substr(p + 1);
    }
    return sn;
}

struct Tables {
    std::unordered_map<std::string, std::set<MGSeq>> lookup;   // get_unstable_mapping, gaf2unstable_main.cpp:34-68
    std::unordered_map<int64_t, int64_t> node_to_contig;        // rgfa2contig, rgfa-split.cpp:35-161
    std::vector<std::string> contigs;
};

int load_rgfa(const std::string& buf, Tables& T) {
    std::map<int64_t, std::list<int64_t>> rank_to_nodes;
    std::unordered_map<int64_t, int64_t> node_to_rank;
    std::unordered_multimap<int64_t, int64_t> edges;
    std::unordered_map<std::string, int64_t> contig_map;
    struct Seg { std::string name; std::vector<Tag> tags; };
    std::vector<Seg> segs;
    // pass 1 (gaf2unstable_main.cpp:43-66): sequence length, SN, SO
    for (size_t i = 0; i < buf.size(); ++i) {
        if (buf[i] != 'S' || !(i == 0 || buf[i - 1] == '\n')) continue;
        std::vector<std::string> tk = tokens_of(buf, i);
        require(tk.size() >= 3, "short S line");
        size_t tag_index = 3;
        std::string seq = tk[2];
        bool numeric = !tk[2].empty();
        for (char c : tk[2]) numeric = numeric && c >= '0' && c <= '9';
        if (numeric) { require(tk.size() >= 4, "short GFA2 S line"); seq = tk[3]; tag_index = 4; }   // gfakluge.hpp:788-792
        Seg sg;
        sg.name = tk[1];
        if (tk.size() > 3)
            for (size_t j = tag_index; j < tk.size(); ++j) sg.tags.push_back(split_tag(tk[j]));
        MGSeq mg;
        mg.name = sg.name;
        mg.length = (int64_t)seq.size();
        std::string contig;
        bool sn = false, so = false;
        for (const Tag& g : sg.tags) {
            if (g.key == "SN") { require(!sn, "two SN"); contig = g.val; sn = true; }
            else if (g.key == "SO") { require(!so, "two SO"); mg.offset = to_long(g.val); require(mg.offset >= 0, "SO<0"); so = true; }
        }
        require(sn && so, "S line without SN/SO");
        T.lookup[contig].insert(mg);
        segs.push_back(sg);
    }
    // pass 2 (rgfa-split.cpp:55-91): ranks and rank-0 contigs
    for (const Seg& sg : segs) {
        int64_t id = node_id(sg.name), rank = 0;
        std::string contig;
        bool sn = false, sr = false;
        for (const Tag& g : sg.tags) {
            if (g.key == "SN") { require(!sn, "two SN"); contig = strip_prefix(g.val); sn = true; }
            else if (g.key == "SR") { require(!sr, "two SR"); rank = to_long(g.val); require(rank >= 0, "SR<0"); sr = true; }
        }
        require(sn && sr, "S line without SN/SR");
        rank_to_nodes[rank].push_back(id);
        node_to_rank[id] = rank;
        if (rank == 0) {
            int64_t cid;
            if (contig_map.count(contig)) cid = contig_map[contig];
            else { cid = (int64_t)contig_map.size(); contig_map[contig] = cid; T.contigs.push_back(contig); }
            T.node_to_contig[id] = cid;
        }
    }
    // pass 3 (rgfa-split.cpp:94-99): L lines -> undirected adjacency
    for (size_t i = 0; i < buf.size(); ++i) {
        if (buf[i] != 'L' || !(i == 0 || buf[i - 1] == '\n')) continue;
        std::vector<std::string> tk = tokens_of(buf, i);
        require(tk.size() >= 5, "short L line");
        int64_t a = node_id(tk[1]), b = node_id(tk[3]);
        edges.insert({a, b});
        edges.insert({b, a});
    }
    // contigs of rank>0 nodes, in rank order (rgfa-split.cpp:108-158)
    for (auto& rn : rank_to_nodes) {
        if (rn.first <= 0) continue;
        const int64_t rank = rn.first;
        std::list<int64_t>& todo = rn.second;
        int64_t pushes = 0;
        while (!todo.empty()) {
            int64_t node = todo.back();
            todo.pop_back();
            std::unordered_map<int64_t, int64_t> counts;
            auto er = edges.equal_range(node);
            for (auto e = er.first; e != er.second; ++e) {
                int64_t other = e->second, orank = node_to_rank[other];
                if (orank < rank || (orank == rank && T.node_to_contig.count(other))) ++counts[T.node_to_contig[other]];
            }
            if (counts.empty()) {
                todo.push_front(node);
                if (++pushes > (int64_t)todo.size()) {
                    std::cerr << "[error] Unable to assign contigs for the following nodes at rank " << rank << ":\n";
                    for (int64_t x : todo) std::cerr << x << std::endl;
                    return 1;
                }
            } else if (counts.size() > 1) {
                std::cerr << "[error] Conflict found for node \"" << node << "\" with rank \"" << rank << ":\n";
                for (auto& c : counts) std::cerr << "\tcontig=" << T.contigs[c.first] << " count=" << c.second << std::endl;
                return 1;
            } else {
                T.node_to_contig[node] = counts.begin()->first;
                pushes = 0;
            }
        }
    }
    return 0;
}

// ---- GAF record (gafkluge.hpp:43-79, parse :84-204, print :274-323)
struct Step { std::string name; bool rev = false, stable = false, interval = false; int64_t start = 0, end = 0; };
struct Rec {
    std::string qname;
    int64_t qlen = 0, qs = 0, qe = 0, plen = 0, ps = 0, pe = 0, m = 0, b = 0;
    int32_t mapq = 0;
    char strand = '+';
    std::vector<Step> path;
    std::map<std::string, std::pair<std::string, std::string>> tags;
};

int64_t gaf_int(const std::string& s) { return s == "*" ? -1 : to_long(s); }   // gafkluge.hpp:30-34
std::string int_str(int64_t v) { return v == -1 ? "*" : std::to_string(v); }    // gafkluge.hpp:25-29

void parse_record(const std::string& line, Rec& r) {
    std::vector<std::string> col;
    {
        std::istringstream in(line);
        std::string tok;
        while (std::getline(in, tok, '\t')) col.push_back(tok);
    }
    require(col.size() >= 12, "fewer than 12 columns");
    for (int k = 0; k < 12; ++k) require(!col[k].empty(), "empty column");
    r.qname = col[0];
    r.qlen = gaf_int(col[1]); r.qs = gaf_int(col[2]); r.qe = gaf_int(col[3]);
    require(col[4].size() == 1 && (col[4][0] == '+' || col[4][0] == '-' || col[4][0] == '*'), "strand");
    r.strand = col[4][0];
    r.path.clear();
    const std::string& p = col[5];
    if (p[0] == '<' || p[0] == '>') {   // gafkluge.hpp:118-147
        size_t i = 0;
        while (i < p.size()) {
            size_t j = i + 1;
            while (j < p.size() && p[j] != '<' && p[j] != '>') ++j;
            Step s;
            s.rev = p[i] == '<';
            std::string tok = p.substr(i + 1, j - i - 1);
            size_t c = tok.find(':');
            if (c == std::string::npos) s.name = tok;
            else {
                s.name = tok.substr(0, c);
                s.stable = s.interval = true;
                size_t d = tok.find('-', c + 1);
                require(d != std::string::npos, "range without '-'");
                s.start = to_long(tok.substr(c + 1, d - c));
                s.end = to_long(tok.substr(d + 1));
            }
            r.path.push_back(s);
            i = j;
        }
    } else if (p != "*") {              // a bare stable contig: one step (gafkluge.hpp:148-157)
        Step s;
        s.name = p;
        s.stable = true;
        r.path.push_back(s);
    }
    r.plen = gaf_int(col[6]); r.ps = gaf_int(col[7]); r.pe = gaf_int(col[8]); r.m = gaf_int(col[9]); r.b = gaf_int(col[10]);
    if (col[11] == "*") r.mapq = -1;
    else { int v = std::stoi(col[11]); r.mapq = v >= 255 ? -1 : v; }   // gafkluge.hpp:176-183
    r.tags.clear();
    for (size_t k = 12; k < col.size(); ++k) {   // gafkluge.hpp:185-202
        const std::string& f = col[k];
        if (f.empty()) continue;
        size_t c1 = f.find(':'), c2 = c1 == std::string::npos ? c1 : f.find(':', c1 + 1);
        require(f.size() >= 5 && c1 != std::string::npos && c2 != std::string::npos, "optional tag");
        std::string key = f.substr(0, c1);
        require(!r.tags.count(key), "duplicate tag");
        r.tags[key] = {f.substr(c1 + 1, c2 - c1 - 1), f.substr(c2 + 1)};
    }
}

void print_record(const Rec& r, std::string& out) {
    out += r.qname + "\t" + int_str(r.qlen) + "\t" + int_str(r.qs) + "\t" + int_str(r.qe) + "\t";
    out.push_back(r.strand);
    out += "\t";
    if (r.path.empty()) out += "*\t*\t*\t*\t*\t*\t";
    else {
        for (const Step& s : r.path) {   // gafkluge.hpp:274-283
            if (!s.stable || s.interval) out.push_back(s.rev ? '<' : '>');
            out += s.name;
            if (s.interval) out += ":" + std::to_string(s.start) + "-" + std::to_string(s.end);
        }
        out += "\t" + int_str(r.plen) + "\t" + int_str(r.ps) + "\t" + int_str(r.pe) + "\t" + int_str(r.m) + "\t" + int_str(r.b) + "\t";
    }
    out += std::to_string(r.mapq == -1 ? 255 : r.mapq);
    for (const auto& t : r.tags) out += "\t" + t.first + ":" + t.second.first + ":" + t.second.second;
    out += "\n";
}

// get_unstable_interval (gaf2unstable_main.cpp:70-107)
std::vector<MGSeq> unstable_interval(const Tables& T, const std::string& contig, int64_t start, int64_t end) {
    auto it = T.lookup.find(contig);
    require(it != T.lookup.end(), "contig not in rGFA");
    const std::set<MGSeq>& nodes = it->second;
    MGSeq q;
    q.offset = start;
    auto i = nodes.upper_bound(q);
    require(i != nodes.begin(), "interval starts before the first node");
    --i;
    q.offset = end;
    auto j = nodes.lower_bound(q);
    require(j != nodes.begin(), "interval ends before the first node");
    std::vector<MGSeq> v;
    int64_t total = 0;
    for (auto k = i; k != j; ++k) { v.push_back(*k); total += k->length; }
    require(!v.empty(), "empty interval");
    total -= start - v.front().offset;
    if (total > end - start) {
        require(v.back().length - (total - (end - start)) > 0, "end clip");
        total = end - start;
    }
    require(total == end - start, "interval length");
    return v;
}

// gaf2unstable (gaf2unstable_main.cpp:109-175); returns the reference's warning text or ""
std::string to_unstable(const Tables& T, Rec& r) {
    std::vector<Step> out;
    for (const Step& s : r.path) {
        std::vector<MGSeq> nodes;
        if (!s.interval) {
            require(r.path.size() == 1, "bare contig in a multi-step path");
            nodes = unstable_interval(T, s.name, r.ps, r.pe);
            int64_t len = r.pe - r.ps;
            r.ps -= nodes.front().offset;
            r.pe = r.ps + len;
            r.plen = 0;
            for (const MGSeq& n : nodes) r.plen += n.length;
        } else {
            nodes = unstable_interval(T, s.name, s.start, s.end);
        }
        if (s.rev) nodes = std::vector<MGSeq>(nodes.rbegin(), nodes.rend());
        for (const MGSeq& n : nodes) {
            Step u;
            u.name = n.name;
            u.rev = s.rev;
            out.push_back(u);
        }
    }
    r.path = out;
    std::set<int64_t> refs;
    for (const Step& s : r.path) {
        auto it = T.node_to_contig.find(node_id(s.name));
        require(it != T.node_to_contig.end(), "node without reference contig");
        refs.insert(it->second);
    }
    std::string warn;
    if (refs.size() > 1) {
        warn = "[gaf2unstable] warning: Target path spans multiple reference contigs ";
        for (int64_t id : refs) warn += T.contigs.at(id) + ", ";
        warn += "\nthe (unstable) record is\n";
        print_record(r, warn);
    }
    if (refs.size() == 1) r.tags["rc"] = {"Z", T.contigs.at(*refs.begin())};
    return warn;
}

bool slurp(const char* path, std::string& out) {
    FILE* f = std::strcmp(path, "-") == 0 ? stdin : std::fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t k;
    while ((k = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    if (f != stdin) std::fclose(f);
    return true;
}

}  // namespace

int main(int argc, char** argv) {
    const char *gfa = nullptr, *olen = nullptr, *in = nullptr;
    for (int i = 1; i < argc; ++i) {
        if ((!std::strcmp(argv[i], "-g") || !std::strcmp(argv[i], "--rgfa")) && i + 1 < argc) gfa = argv[++i];
        else if (!std::strcmp(argv[i], "-o") && i + 1 < argc) olen = argv[++i];
        else in = argv[i];
    }
    if (!gfa || !in) { std::fprintf(stderr, "usage: gaf2unstable_oracle -g graph.gfa [-o lengths.tsv] <gaf|->\n"); return 1; }
    std::string rg, gaf, out;
    if (!slurp(in, gaf)) { std::fprintf(stderr, "[gaf2unstable] error: unable to open input: %s\n", in); return 1; }
    if (!slurp(gfa, rg)) { std::fprintf(stderr, "[gaf2unstable] error: Could not open %s\n", gfa); return 1; }
    Tables T;
    try {
        int rc = load_rgfa(rg, T);
        if (rc) return rc;
        if (olen) {   // gaf2unstable_main.cpp:274-285: iteration order of the unordered_map, nodes by offset
            std::ofstream o(olen);
            if (!o) { std::fprintf(stderr, "[gaf2unstable] error: unable to open output: %s\n", olen); return 1; }
            for (const auto& cs : T.lookup)
                for (const MGSeq& s : cs.second) o << s.name << "\t" << s.length << "\n";
        }
        std::istringstream lines(gaf);
        std::string line;
        Rec r;
        while (std::getline(lines, line)) {   // gaf2unstable_main.cpp:288-297
            if (line[0] == '*') continue;
            try {
                parse_record(line, r);
                std::string warn = to_unstable(T, r);
                std::fputs(warn.c_str(), stderr);
                print_record(r, out);
            } catch (const std::exception& e) {
                std::fwrite(out.data(), 1, out.size(), stdout);
                std::fflush(stdout);
                std::fprintf(stderr, "abort: %s\n", e.what());
                return 134;
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "abort: %s\n", e.what());
        return 134;
    }
    std::fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
