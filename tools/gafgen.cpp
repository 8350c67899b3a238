// gafgen — deterministic synthetic minigraph-style GAF + lengths-table generator.
//
// Test / benchmark infrastructure (not part of the converter).  Produces the
// workload shapes of SURVEY.md §8(d): short-read node-space records (config 3),
// assembly-scale records with ~10^4 steps (config 4), stable-interval records
// (config 1 shape) and mixes of them.  Every record is generated from
// hash(seed, record index), so any sub-range can be produced independently (per
// shard, per thread) and is reproducible on the GPU box without shipping data.
//
// Invariants kept so that the reference converts every record without asserting:
//   * path_start < len(step 0), end clip < len(last step)        (gaf2paf_main.cpp:176-178)
//   * sum of CIGAR target bases == path_end - path_start          (:80)
//
// Build:  g++ -O2 -std=c++17 -shared -fPIC -pthread tools/gafgen.cpp -o build/libgafgen.so
//         g++ -O2 -std=c++17 -DGAFGEN_MAIN -pthread tools/gafgen.cpp -o build/gafgen
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

extern "C" {
struct gafgen_params {
    uint64_t seed;
    uint32_t n_nodes;        // node table s1..sN (node mode) or number of contigs (stable mode)
    uint32_t node_len_lo, node_len_hi;
    uint32_t steps_lo, steps_hi;
    uint32_t mrun_lo, mrun_hi;     // length of M (or =) runs
    uint32_t indel_lo, indel_hi;   // length of I / D ops between runs
    uint32_t max_runs;       // cap on M runs per record (0 = no cap): short reads use ~5
    uint32_t pct_rev;        // percent of steps written '<'
    uint32_t pct_minus;      // percent of records on '-' strand
    uint32_t use_eqx;        // 1: '=' runs with 1 bp 'X' instead of plain 'M'
    uint32_t stable;         // 1: steps are >ctgK:a-b intervals, table holds contigs
    uint32_t qlen_min;       // query length is max(qlen_min, query end)
    uint32_t pct_star;       // percent of extra "*\t..." -S style lines interleaved
    uint32_t mix_every;      // K > 0: every K-th record has the assembly shape (5k-15k steps) over the same node table
    uint32_t qname_len;      // > 0: query names are padded to this length (instrument-style read names)
    uint32_t extra_tag_len;  // > 0: an extra "zd:Z:<text>" tag of this many value bytes before cg
};
}

namespace {

typedef uint64_t u64;
typedef uint32_t u32;

inline u64 mix(u64 x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

struct Rng {
    u64 s;
    explicit Rng(u64 seed) : s(seed) {}
    u64 next() { s += 0x9e3779b97f4a7c15ULL; return mix(s); }
    // uniform in [lo, hi]
    u64 range(u64 lo, u64 hi) { return hi <= lo ? lo : lo + next() % (hi - lo + 1); }
    bool pct(u32 p) { return next() % 100 < p; }
};

inline u32 node_len(const gafgen_params& P, u64 id) {
    u64 span = (u64)P.node_len_hi - P.node_len_lo + 1;
    return P.node_len_lo + (u32)(mix(P.seed * 0x51ed27 + id) % span);
}
inline u64 contig_len(const gafgen_params& P, u64 id) {
    return 1000000ULL + mix(P.seed * 0x7777 + id) % 200000000ULL;
}

inline void put_u64(std::string& o, u64 v) {
    char b[24];
    int n = 0;
    do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) o.push_back(b[--n]);
}

void gen_record(const gafgen_params& P0, u64 idx, std::string& o) {
    gafgen_params P = P0;
    if (P.mix_every && idx % P.mix_every == P.mix_every - 1) {   // the assembly-shaped record of a mixed file (config 5)
        P.steps_lo = 5000; P.steps_hi = 15000; P.mrun_lo = 20; P.mrun_hi = 800; P.indel_lo = 1; P.indel_hi = 50;
        P.max_runs = 0; P.qlen_min = 1000000; P.qname_len = 0; P.extra_tag_len = 0;
    }
    Rng R(mix(P.seed ^ mix(idx + 1)));
    if (P.pct_star && R.pct(P.pct_star)) {
        o += "*\t>s"; put_u64(o, R.range(1, P.n_nodes)); o += "\t97\t12\t0\t6\t92\n";
    }
    const u32 n = (u32)R.range(P.steps_lo, P.steps_hi);
    // steps
    static thread_local std::vector<u64> ids, sa;
    static thread_local std::vector<u32> ln;
    static thread_local std::vector<char> rv;
    ids.resize(n); ln.resize(n); rv.resize(n); sa.resize(n);
    u64 total = 0;
    for (u32 i = 0; i < n; ++i) {
        ids[i] = R.range(1, P.n_nodes);
        if (P.stable) {
            u64 cl = contig_len(P, ids[i]);
            u32 l = (u32)R.range(P.node_len_lo, P.node_len_hi);
            sa[i] = R.range(0, cl - l);
            ln[i] = l;
        } else {
            ln[i] = node_len(P, ids[i]);
        }
        rv[i] = R.pct(P.pct_rev);
        total += ln[i];
    }
    // clips: keep the alignment close to the inner end of the first/last step
    u64 ps, pe;
    if (n == 1) {
        ps = R.range(0, ln[0] - 1);
        u64 maxw = ln[0] - ps;
        u64 w = R.range(1, maxw);
        pe = ps + w;
    } else {
        ps = R.range(0, ln[0] - 1);
        u64 endclip = R.range(0, ln[n - 1] - 1);
        pe = total - endclip;
    }
    const u64 W = pe - ps;
    // CIGAR with W target bases
    static thread_local std::string cg;
    cg.clear();
    u64 tleft = W, qcons = 0, nmatch = 0, blen = 0, runs = 0;
    while (tleft > 0) {
        u64 m = R.range(P.mrun_lo, P.mrun_hi);
        ++runs;
        if (m > tleft || (P.max_runs && runs >= P.max_runs)) m = tleft;
        if (P.use_eqx && m > 2 && R.pct(50)) {
            u64 a = R.range(1, m - 2);
            put_u64(cg, a); cg.push_back('=');
            cg += "1X";
            put_u64(cg, m - a - 1); cg.push_back('=');
            nmatch += m - 1;
        } else {
            put_u64(cg, m); cg.push_back(P.use_eqx ? '=' : 'M');
            nmatch += m;
        }
        qcons += m; blen += m; tleft -= m;
        if (tleft == 0) break;
        u64 g = R.range(P.indel_lo, P.indel_hi);
        if (R.pct(50)) {
            put_u64(cg, g); cg.push_back('I');
            qcons += g; blen += g;
        } else {
            if (g >= tleft) g = tleft > 1 ? tleft - 1 : 0;
            if (g) { put_u64(cg, g); cg.push_back('D'); tleft -= g; blen += g; }
        }
    }
    const u64 qs = R.range(0, 40);
    const u64 qe = qs + qcons;
    const u64 qlen = std::max<u64>(P.qlen_min, qe + R.range(0, 10));
    const bool minus = R.pct(P.pct_minus);
    static const int mapqs[5] = {0, 1, 30, 60, 255};
    // line
    {
        const size_t n0 = o.size();
        o += "read"; put_u64(o, idx);
        for (u32 k = 0; o.size() - n0 < P.qname_len; ++k) o.push_back(k % 9 == 0 ? ':' : (char)('0' + mix(idx * 131 + k) % 10));
        o.push_back('\t');
    }
    put_u64(o, qlen); o.push_back('\t');
    put_u64(o, qs); o.push_back('\t');
    put_u64(o, qe); o.push_back('\t');
    o.push_back(minus ? '-' : '+'); o.push_back('\t');
    for (u32 i = 0; i < n; ++i) {
        o.push_back(rv[i] ? '<' : '>');
        if (P.stable) {
            o += "ctg"; put_u64(o, ids[i]); o.push_back(':');
            put_u64(o, sa[i]); o.push_back('-'); put_u64(o, sa[i] + ln[i]);
        } else {
            o.push_back('s'); put_u64(o, ids[i]);
        }
    }
    o.push_back('\t');
    put_u64(o, total); o.push_back('\t');
    put_u64(o, ps); o.push_back('\t');
    put_u64(o, pe); o.push_back('\t');
    put_u64(o, nmatch); o.push_back('\t');
    put_u64(o, blen); o.push_back('\t');
    put_u64(o, mapqs[R.next() % 5]);
    o += R.pct(80) ? "\ttp:A:P" : "\ttp:A:S";
    o += "\tcm:i:"; put_u64(o, R.range(1, 40));
    o += "\ts1:i:"; put_u64(o, R.range(10, 150));
    o += "\ts2:i:"; put_u64(o, R.range(0, 100));
    o += "\tdv:f:0.0"; put_u64(o, R.range(100, 999));
    if (P.extra_tag_len) {
        o += "\tzd:Z:";
        for (u32 k = 0; k < P.extra_tag_len; ++k) o.push_back((char)('A' + mix(idx * 977 + k) % 26));
    }
    o += "\tcg:Z:"; o += cg;
    o.push_back('\n');
}

}  // namespace

extern "C" {

// "name\tlength\n" rows of the table the records refer to.  Returns the byte count;
// writes only if it fits in cap.
size_t gafgen_lengths(const gafgen_params* P, char* buf, size_t cap) {
    std::string o;
    o.reserve((size_t)P->n_nodes * 14);
    for (u64 id = 1; id <= P->n_nodes; ++id) {
        if (P->stable) { o += "ctg"; put_u64(o, id); o.push_back('\t'); put_u64(o, contig_len(*P, id)); }
        else { o.push_back('s'); put_u64(o, id); o.push_back('\t'); put_u64(o, node_len(*P, id)); }
        o.push_back('\n');
    }
    if (o.size() <= cap && buf) std::memcpy(buf, o.data(), o.size());
    return o.size();
}

// Records [first, first+count).  Returns the byte count; writes only if it fits.
size_t gafgen_records(const gafgen_params* P, uint64_t first, uint64_t count, char* buf, size_t cap, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > count) threads = count ? (int)count : 1;
    std::vector<std::string> parts(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) {
        th.emplace_back([&, t]() {
            u64 a = first + count * t / threads, b = first + count * (t + 1) / threads;
            std::string& o = parts[t];
            for (u64 i = a; i < b; ++i) gen_record(*P, i, o);
        });
    }
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    if (total <= cap && buf) {
        size_t off = 0;
        std::vector<std::thread> cp;
        for (int t = 0; t < threads; ++t) {
            char* dst = buf + off;
            off += parts[t].size();
            cp.emplace_back([&, t, dst]() { std::memcpy(dst, parts[t].data(), parts[t].size()); });
        }
        for (auto& x : cp) x.join();
    }
    return total;
}

// Generates records [first, first+count) once and returns a malloc'd buffer (free with
// gafgen_free); *size receives the byte count.
char* gafgen_records_alloc(const gafgen_params* P, uint64_t first, uint64_t count, int threads, size_t* size) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > count) threads = count ? (int)count : 1;
    std::vector<std::string> parts(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) {
        th.emplace_back([&, t]() {
            u64 a = first + count * t / threads, b = first + count * (t + 1) / threads;
            std::string& o = parts[t];
            o.reserve((size_t)((b - a) * 160));
            for (u64 i = a; i < b; ++i) gen_record(*P, i, o);
        });
    }
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    char* buf = static_cast<char*>(std::malloc(total ? total : 1));
    if (!buf) { *size = 0; return nullptr; }
    size_t off = 0;
    std::vector<std::thread> cp;
    for (int t = 0; t < threads; ++t) {
        char* dst = buf + off;
        off += parts[t].size();
        cp.emplace_back([&, t, dst]() { std::memcpy(dst, parts[t].data(), parts[t].size()); });
    }
    for (auto& x : cp) x.join();
    *size = total;
    return buf;
}
void gafgen_free(char* p) { std::free(p); }

void gafgen_preset(const char* name, gafgen_params* P) {
    std::memset(P, 0, sizeof *P);
    P->seed = 1;
    P->pct_rev = 30; P->pct_minus = 50; P->qlen_min = 150;
    if (!std::strcmp(name, "short")) {
        P->n_nodes = 200000; P->node_len_lo = 20; P->node_len_hi = 400;
        P->steps_lo = 1; P->steps_hi = 5; P->mrun_lo = 10; P->mrun_hi = 120; P->indel_lo = 1; P->indel_hi = 3;
        P->max_runs = 5;
    } else if (!std::strcmp(name, "short_eqx")) {
        gafgen_preset("short", P); P->use_eqx = 1;
    } else if (!std::strcmp(name, "asm")) {     // config 4
        P->n_nodes = 2000000; P->node_len_lo = 50; P->node_len_hi = 2000;
        P->steps_lo = 5000; P->steps_hi = 15000; P->mrun_lo = 20; P->mrun_hi = 800; P->indel_lo = 1; P->indel_hi = 50;
        P->qlen_min = 1000000;
    } else if (!std::strcmp(name, "stable")) {  // config 1 shape
        P->n_nodes = 24; P->node_len_lo = 200; P->node_len_hi = 20000; P->stable = 1;
        P->steps_lo = 1; P->steps_hi = 50; P->mrun_lo = 20; P->mrun_hi = 3000; P->indel_lo = 1; P->indel_hi = 50;
        P->qlen_min = 10000;
    } else if (!std::strcmp(name, "tagged")) {  // short reads with instrument-style names and a long extra tag: 250-500 B records
        gafgen_preset("short", P); P->qname_len = 40; P->extra_tag_len = 120;
    } else if (!std::strcmp(name, "mixed")) {   // config 5 shape: ~90 % of the bytes short-read records, ~10 % assembly-scale records, one node table
        gafgen_preset("short", P);
        P->n_nodes = 2000000; P->node_len_lo = 50; P->node_len_hi = 2000; P->mix_every = 16400;
    } else if (!std::strcmp(name, "medium")) {  // long-read-like, used for skew tests
        P->n_nodes = 200000; P->node_len_lo = 20; P->node_len_hi = 400;
        P->steps_lo = 20; P->steps_hi = 200; P->mrun_lo = 5; P->mrun_hi = 300; P->indel_lo = 1; P->indel_hi = 10;
        P->qlen_min = 5000;
    }
}

}  // extern "C"

#ifdef GAFGEN_MAIN
// gafgen <preset> <n_records> <out.gaf> <out.lengths.tsv> [seed] [first]
int main(int argc, char** argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: gafgen <short|short_eqx|tagged|mixed|asm|stable|medium> <n_records> <out.gaf> <lengths.tsv> [seed] [first] [k=v ...]\n");
        return 1;
    }
    gafgen_params P;
    gafgen_preset(argv[1], &P);
    if (P.n_nodes == 0) { std::fprintf(stderr, "unknown preset %s\n", argv[1]); return 1; }
    u64 n = std::strtoull(argv[2], 0, 10);
    if (argc > 5) P.seed = std::strtoull(argv[5], 0, 10);
    u64 first = argc > 6 ? std::strtoull(argv[6], 0, 10) : 0;
    for (int i = 7; i < argc; ++i) {
        const char* eq = std::strchr(argv[i], '=');
        if (!eq) continue;
        std::string k(argv[i], eq - argv[i]);
        u32 v = (u32)std::strtoul(eq + 1, 0, 10);
        if (k == "n_nodes") P.n_nodes = v; else if (k == "steps_lo") P.steps_lo = v; else if (k == "steps_hi") P.steps_hi = v;
        else if (k == "pct_star") P.pct_star = v; else if (k == "pct_minus") P.pct_minus = v; else if (k == "pct_rev") P.pct_rev = v;
        else if (k == "use_eqx") P.use_eqx = v; else if (k == "node_len_lo") P.node_len_lo = v; else if (k == "node_len_hi") P.node_len_hi = v;
        else if (k == "mix_every") P.mix_every = v; else if (k == "qname_len") P.qname_len = v; else if (k == "extra_tag_len") P.extra_tag_len = v;
        else if (k == "mrun_lo") P.mrun_lo = v; else if (k == "mrun_hi") P.mrun_hi = v; else if (k == "max_runs") P.max_runs = v;
    }
    unsigned hw = std::thread::hardware_concurrency();
    size_t need = gafgen_records(&P, first, n, nullptr, 0, hw ? hw : 1);
    std::vector<char> buf(need);
    gafgen_records(&P, first, n, buf.data(), buf.size(), hw ? hw : 1);
    FILE* f = std::fopen(argv[3], "wb");
    std::fwrite(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    size_t ln = gafgen_lengths(&P, nullptr, 0);
    std::vector<char> lb(ln);
    gafgen_lengths(&P, lb.data(), lb.size());
    f = std::fopen(argv[4], "wb");
    std::fwrite(lb.data(), 1, lb.size(), f);
    std::fclose(f);
    std::fprintf(stderr, "gafgen: %llu records, %zu bytes; table %zu bytes\n", (unsigned long long)n, need, ln);
    return 0;
}
#endif
