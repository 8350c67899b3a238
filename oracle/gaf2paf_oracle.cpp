// gaf2paf_oracle — CPU restatement of the reference's GAF -> PAF path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by
// the product (libg2p.so, gaf2paf, gaf2unstable); only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may run it, as the checker.
//
// Parity status: PINNED.  The reference has no golden vectors for this path
// (SURVEY.md §4), so this restatement is pinned against the reference binary itself,
// compiled from /root/reference by oracle/build_ref.sh into oracle/_ref/: the
// committed known-answer vectors (tests/golden/*.json, produced by
// tests/golden/make_golden.py from that binary) and seeded differential runs
// (tests/test_oracle.py) must match byte for byte.
//
// It follows the reference's own structure (materialised step vector and CIGAR
// vector, physical flip for '-' records) and is deliberately unrelated to the
// streaming two-cursor formulation the CUDA kernels use, so that the two can check
// each other.  Each function cites the reference lines it restates.
//
// usage: gaf2paf_oracle -l lengths.tsv <gaf|-> [gaf2 ...]  > out.paf
//        (exit 0 / 1 / 134 exactly as the reference; aborts are reported as 134)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Abort : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct Exit1 : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// std::stol as the reference uses it; failures are uncaught there -> SIGABRT.
long stol_or_abort(const std::string& s) {
    try {
        return std::stol(s);
    } catch (const std::exception&) {
        throw Abort("stol: " + s);
    }
}

// paf.hpp:31-47 split_delims (TAB only here): empty tokens are dropped
std::vector<std::string> split_tabs_nonempty(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && s[i] == '\t') ++i;
        size_t a = i;
        while (i < s.size() && s[i] != '\t') ++i;
        if (i > a) out.push_back(s.substr(a, i - a));
    }
    return out;
}

// gaf2paf_main.cpp:22-45 get_len_map
std::unordered_map<std::string, int64_t> load_lengths(const std::string& text) {
    std::unordered_map<std::string, int64_t> m;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::vector<std::string> toks = split_tabs_nonempty(text.substr(pos, eol - pos));
        if (toks.size() > 1) m[toks[0]] = stol_or_abort(toks[1]);
        pos = eol + 1;
    }
    return m;
}

struct Step {   // gafkluge.hpp:43-51
    std::string name;
    bool is_reverse = false, is_interval = false;
    int64_t start = 0, end = 0;
};
struct Op {     // gaf2paf_main.cpp:47-48
    char c;
    int64_t len;
};
struct Record { // gafkluge.hpp:56-79
    std::string qname;
    int64_t qlen = -1, qs = -1, qe = -1, plen = -1, ps = -1, pe = -1, matches = -1, blen = -1;
    int32_t mapq = 255;
    char strand = '*';
    std::vector<Step> path;
    std::map<std::string, std::pair<std::string, std::string>> tags;
};

// gafkluge.hpp:32-34
int64_t string_to_int(const std::string& s) { return s == "*" ? -1 : stol_or_abort(s); }

// gafkluge.hpp:84-204 parse_gaf_record
void parse_record(const std::string& line, Record& r) {
    size_t pos = 0;
    bool fail = false;   // stream failbit: set when a read starts at end-of-stream
    int col = 1;
    auto next = [&](std::string& buf) -> bool {   // getline(in, buf, '\t'); returns stream state
        if (fail || pos > line.size()) { fail = true; buf.clear(); return false; }
        if (pos == line.size()) { pos = line.size() + 1; fail = true; buf.clear(); return false; }
        size_t t = line.find('\t', pos);
        if (t == std::string::npos) { buf = line.substr(pos); pos = line.size() + 1; }
        else { buf = line.substr(pos, t - pos); pos = t + 1; if (pos == line.size()) { /* trailing tab: next read hits EOF */ } }
        return true;
    };
    std::string buf;
    auto scan = [&]() {   // :90-96
        bool ok = next(buf);
        if (!ok || buf.empty()) throw Abort("Error parsing GAF column " + std::to_string(col));
        ++col;
    };
    scan(); r.qname = buf;
    scan(); r.qlen = string_to_int(buf);
    scan(); r.qs = string_to_int(buf);
    scan(); r.qe = string_to_int(buf);
    scan();
    if (buf == "-" || buf == "*" || buf == "+") r.strand = buf[0];
    else throw Abort("Error parsing GAF strand: " + buf);
    r.path.clear();
    scan();
    if (buf[0] == '<' || buf[0] == '>') {   // :119-148
        size_t p = 0, nx;
        do {
            Step st;
            p = buf.find_first_of("><", p);
            nx = buf.find_first_of("><", p + 1);
            std::string tok = buf.substr(p, nx != std::string::npos ? nx - p : std::string::npos);
            size_t colon = tok.find_first_of(':');
            st.is_reverse = tok[0] == '<';
            if (colon == std::string::npos) {
                st.name = tok.substr(1);
            } else {
                st.name = tok.substr(1, colon - 1);
                st.is_interval = true;
                size_t dash = tok.find_first_of('-', colon);
                if (dash == std::string::npos) throw Abort("Error parsing GAF range of " + tok);
                st.start = stol_or_abort(tok.substr(colon + 1, dash - colon));
                st.end = stol_or_abort(tok.substr(dash + 1));
            }
            r.path.push_back(st);
            p = nx;
        } while (nx != std::string::npos);
    } else if (buf != "*") {   // :149-156
        Step st;
        st.name = buf;
        r.path.push_back(st);
    }
    scan(); r.plen = string_to_int(buf);
    scan(); r.ps = string_to_int(buf);
    scan(); r.pe = string_to_int(buf);
    scan(); r.matches = string_to_int(buf);
    scan(); r.blen = string_to_int(buf);
    scan();
    if (buf == "*") {
        r.mapq = -1;
    } else {   // :179-183 std::stoi
        long v = stol_or_abort(buf);
        if (v > 2147483647L || v < -2147483648L) throw Abort("stoi range");
        r.mapq = v >= 255 ? -1 : (int32_t)v;
    }
    r.tags.clear();
    for (;;) {   // :186-202
        bool ok = next(buf);
        if (ok && !buf.empty()) {
            size_t c1 = buf.find_first_of(':');
            size_t c2 = c1 == std::string::npos ? std::string::npos : buf.find_first_of(':', c1 + 1);
            if (buf.length() < 5 || c1 == std::string::npos || c2 == std::string::npos) throw Abort("Unable to parse optional tag " + buf);
            std::string tag = buf.substr(0, c1);
            if (r.tags.count(tag)) throw Abort("Duplicate optional field found: " + tag);
            r.tags[tag] = std::make_pair(buf.substr(c1 + 1, c2 - c1 - 1), buf.substr(c2 + 1));
        }
        if (!ok) break;
    }
}

// gafkluge.hpp:226-239 for_each_cg: a token runs up to the next of "MIDNSHPX=", its length is
// std::stol of the text before the letter (blanks, sign and trailing junk accepted as stol does).
std::vector<Op> parse_cg(const Record& r) {
    std::vector<Op> ops;
    auto it = r.tags.find("cg");
    if (it == r.tags.end()) return ops;
    const std::string& cg = it->second.second;
    size_t co = 0;
    while (co < cg.size()) {
        size_t nx = cg.find_first_of("MIDNSHPX=", co);
        if (nx == std::string::npos) throw Abort("for_each_cg assert");
        std::string num = cg.substr(co, nx - co);
        ops.push_back(Op{cg[nx], (int64_t)stol_or_abort(num)});
        co = nx + 1;
    }
    return ops;
}

bool consumes_query(char c) { return c == 'M' || c == 'I' || c == 'S' || c == '=' || c == 'X'; }    // gaf2paf_main.cpp:50-52
bool consumes_target(char c) { return c == 'M' || c == 'D' || c == 'N' || c == '=' || c == 'X'; }   // :54-56

int64_t lookup(const std::unordered_map<std::string, int64_t>& len_map, const std::string& name) {
    auto it = len_map.find(name);
    if (it == len_map.end()) throw Exit1("[gaf2paf] error: unable to find " + name + " in lengths map");
    return it->second;
}

// gaf2paf_main.cpp:92-131 flip_gaf
void flip(Record& r, std::vector<Op>& ops, const std::unordered_map<std::string, int64_t>& len_map) {
    r.strand = r.strand == '+' ? '-' : '+';
    if (ops.empty()) throw Abort("assert(!cigar.empty())");
    std::vector<Op> rev(ops.rbegin(), ops.rend());
    ops.swap(rev);
    std::vector<Step> rp(r.path.rbegin(), r.path.rend());
    r.path.swap(rp);
    int64_t total = 0;
    for (Step& s : r.path) {
        s.is_reverse = !s.is_reverse;
        total += s.is_interval ? s.end - s.start : lookup(len_map, s.name);
    }
    int64_t rs = total - r.pe, re = total - r.ps;
    r.ps = rs;
    r.pe = re;
}

std::string format_g(double v) {   // ostream << double at default precision == %g
    char b[64];
    snprintf(b, sizeof b, "%g", v);
    return b;
}

// gaf2paf_main.cpp:134-264 gaf2paf (with cigar_cut :59-68 and cigar_next_by_target :71-90
// done on an index + "remaining length of the head op" instead of list surgery)
void convert(const Record& r, std::vector<Op> ops, const std::unordered_map<std::string, int64_t>& len_map, std::string& out) {
    if (r.strand != '+') throw Abort("assert(strand == '+')");
    const int64_t path_len = r.pe - r.ps;
    size_t pos = 0;   // ops[pos] is the head of the unconsumed CIGAR (its len may have been reduced by a cut)
    int64_t qcount = 0, tcount = 0;
    for (size_t i = 0; i < r.path.size(); ++i) {
        Step st = r.path[i];
        const int64_t tlen = lookup(len_map, st.name);
        if (!st.is_interval) { st.start = 0; st.end = tlen; }
        int64_t so = i == 0 ? r.ps : 0;
        int64_t eo = i == r.path.size() - 1 ? tcount + (st.end - st.start) - path_len - so : 0;
        if (so < 0 || eo < 0) throw Abort("assert(start_offset >= 0 && end_offset >= 0)");
        const int64_t want = (st.end - eo) - (st.start + so);
        if (want < 0) throw Abort("negative quota (reference: undefined behaviour)");
        // cigar_next_by_target
        std::vector<Op> piece;
        int64_t cur = 0;
        while (pos < ops.size() && cur < want) {
            Op o = ops[pos];
            if (consumes_target(o.c) && cur + o.len > want) {
                int64_t keep = want - cur;       // first part stays in this step …
                ops[pos].len = o.len - keep;     // … the rest heads the next step (cigar_cut)
                o.len = keep;
                cur += keep;
                piece.push_back(o);
                break;
            }
            if (consumes_target(o.c)) cur += o.len;
            piece.push_back(o);
            ++pos;
        }
        if (cur != want) throw Abort("assert(cur_len > target_len)");
        char strand = '+';
        if (st.is_reverse) {
            std::swap(so, eo);
            std::vector<Op> rp(piece.rbegin(), piece.rend());
            piece.swap(rp);
            strand = '-';
        }
        int64_t q = 0, t = 0, nm = 0, nb = 0;
        std::string cig;
        for (const Op& o : piece) {
            if (consumes_query(o.c)) q += o.len;
            if (consumes_target(o.c)) t += o.len;
            if (o.c == 'M' || o.c == '=') nm += o.len;
            nb += o.len;
            cig += std::to_string(o.len);
            cig.push_back(o.c);
        }
        if (nm > 0) {   // :225
            const int64_t q0 = r.qs + qcount;
            out += r.qname; out.push_back('\t');
            out += std::to_string(r.qlen); out.push_back('\t');
            out += std::to_string(q0); out.push_back('\t');
            out += std::to_string(q0 + q); out.push_back('\t');
            out.push_back(strand); out.push_back('\t');
            out += st.name; out.push_back('\t');
            out += std::to_string(tlen); out.push_back('\t');
            out += std::to_string(st.start + so); out.push_back('\t');
            out += std::to_string(st.end - eo); out.push_back('\t');
            out += std::to_string(nm); out.push_back('\t');
            out += std::to_string(nb); out.push_back('\t');
            out += std::to_string((int64_t)r.mapq);
            auto tp = r.tags.find("tp");
            if (tp != r.tags.end()) out += "\ttp:" + tp->second.first + ":" + tp->second.second;
            auto rc = r.tags.find("rc");
            if (rc != r.tags.end()) out += "\trc:" + rc->second.first + ":" + rc->second.second;
            out += "\tgm:i:" + std::to_string(r.matches);
            out += "\tgl:i:" + std::to_string(r.blen);
            volatile double identity = 0;
            if (r.blen > 0) {
                identity = (double)r.matches / (double)r.blen;
                volatile double t1 = identity * 1000;
                volatile double t2 = t1 + 0.5;
                identity = std::floor(t2) / 1000;
            }
            out += "\tgi:f:" + format_g(identity);
            out += "\tcg:Z:" + cig + "\n";
        }
        qcount += q;
        tcount += t;
    }
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

// gaf2paf_main.cpp:274-377 main (argument handling reduced to what the tests need)
int main(int argc, char** argv) {
    const char* lengths = nullptr;
    std::vector<const char*> inputs;
    for (int i = 1; i < argc; ++i) {
        if ((!std::strcmp(argv[i], "-l") || !std::strcmp(argv[i], "--lengths")) && i + 1 < argc) lengths = argv[++i];
        else inputs.push_back(argv[i]);
    }
    if (!lengths || inputs.empty()) { std::fprintf(stderr, "usage: gaf2paf_oracle -l lengths.tsv <gaf> ...\n"); return 1; }
    std::string tsv;
    if (!slurp(lengths, tsv)) { std::fprintf(stderr, "[gaf2paf] error: unable to open %s\n", lengths); return 1; }
    std::unordered_map<std::string, int64_t> len_map;
    try {
        len_map = load_lengths(tsv);
    } catch (const Abort& e) {
        std::fprintf(stderr, "abort: %s\n", e.what());
        return 134;
    }
    std::string out;
    for (const char* path : inputs) {
        std::string gaf;
        if (!slurp(path, gaf)) {
            std::fwrite(out.data(), 1, out.size(), stdout);
            std::fprintf(stderr, "[gaf2paf] error: unable to open input: %s\n", path);
            return 1;
        }
        size_t pos = 0;
        while (pos < gaf.size()) {
            size_t eol = gaf.find('\n', pos);
            if (eol == std::string::npos) eol = gaf.size();
            std::string line = gaf.substr(pos, eol - pos);
            pos = eol + 1;
            if (line[0] == '*') continue;   // :360-363 (an empty line has line[0] == '\0')
            const size_t keep = out.size();
            try {
                Record r;
                parse_record(line, r);
                if (!r.tags.count("cg")) throw Exit1("[gaf2paf] error: cg cigar not found. This tool only works on output of minigraph -c");
                std::vector<Op> ops = parse_cg(r);
                if (r.strand == '-') flip(r, ops, len_map);
                convert(r, ops, len_map, out);
            } catch (const Exit1& e) {
                std::fwrite(out.data(), 1, out.size(), stdout);   // exit(1) flushes what was written
                std::fprintf(stderr, "%s\n", e.what());
                return 1;
            } catch (const Abort& e) {
                out.resize(keep);
                std::fwrite(out.data(), 1, out.size(), stdout);
                std::fprintf(stderr, "abort: %s\n", e.what());
                return 134;
            }
            if (out.size() > (1u << 24)) { std::fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
        }
    }
    std::fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
