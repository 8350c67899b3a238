// g2u_hostsim — CPU instantiation of the gaf2unstable per-record code (g2u_core.cuh) and of the
// host-side rGFA table builder (g2u_rgfa.hpp).
//
// TEST INFRASTRUCTURE ONLY: never linked into libg2p.so or the executables.  Mirrors
// g2p_load_rgfa + run_unstable (g2p_capi.cu) with plain host arrays so that CPU-only CI can check
// the rewrite against the reference binary and the golden vectors.
//
// usage: g2u_hostsim -g graph.gfa [-o node-lengths.tsv] <gaf|->      (stdout / exit code as gaf2unstable)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../cactus-gfa-tools_b200/csrc/g2u_core.cuh"
#include "../../cactus-gfa-tools_b200/csrc/g2p_table.hpp"
#include "../../cactus-gfa-tools_b200/csrc/g2u_rgfa.hpp"

using namespace g2p;

static bool slurp(const char* path, std::string& out) {
    FILE* f = std::strcmp(path, "-") == 0 ? stdin : std::fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t k;
    while ((k = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    if (f != stdin) std::fclose(f);
    return true;
}

int main(int argc, char** argv) {
    const char *gfa = nullptr, *olen = nullptr, *in = nullptr;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "-g") && i + 1 < argc) gfa = argv[++i];
        else if (!std::strcmp(argv[i], "-o") && i + 1 < argc) olen = argv[++i];
        else in = argv[i];
    }
    if (!gfa || !in) { std::fprintf(stderr, "usage: g2u_hostsim -g graph.gfa [-o lengths.tsv] <gaf>\n"); return 1; }
    std::string rg, gaf;
    if (!slurp(gfa, rg)) { std::fprintf(stderr, "[gaf2unstable] error: Could not open %s\n", gfa); return 1; }
    if (!slurp(in, gaf)) { std::fprintf(stderr, "[gaf2unstable] error: unable to open input: %s\n", in); return 1; }
    RgfaTables T;
    build_rgfa_tables(rg.data(), rg.size(), T);
    if (T.exit_code) { std::fputs(T.error.c_str(), stderr); std::fputc('\n', stderr); return T.exit_code; }
    HostLenTable ct;
    ct.reserve_for(T.mapping.size());
    std::vector<u32> begin, refoff{0};
    std::vector<UNode> nodes;
    std::vector<u8> names, refnames;
    std::string node_lengths;
    u32 ci = 0;
    for (const auto& cs : T.mapping) {
        ct.put(reinterpret_cast<const u8*>(cs.first.data()), (u32)cs.first.size(), (i64)ci++);
        begin.push_back((u32)nodes.size());
        i64 cum = 0;
        for (const RgfaNode& nd : cs.second) {
            UNode u;
            u.offset = nd.offset; u.cum = cum; u.length = (u32)nd.length;
            u.name_off = (u32)names.size(); u.name_len = (u32)nd.name.size();
            names.insert(names.end(), nd.name.begin(), nd.name.end());
            int64_t id;
            u.ref = -1;
            if (rgfa_detail::node_id_of(nd.name, id)) {
                auto it = T.node_to_contig.find(id);
                if (it != T.node_to_contig.end()) u.ref = (i32)it->second;
            }
            cum += nd.length;
            nodes.push_back(u);
            node_lengths += nd.name + "\t" + std::to_string(nd.length) + "\n";
        }
    }
    begin.push_back((u32)nodes.size());
    if (ct.arena.empty()) ct.arena.push_back(0);
    for (const std::string& c : T.ref_contigs) { refnames.insert(refnames.end(), c.begin(), c.end()); refoff.push_back((u32)refnames.size()); }
    if (olen) {
        FILE* o = std::fopen(olen, "wb");
        if (!o) { std::fprintf(stderr, "[gaf2unstable] error: unable to open output: %s\n", olen); return 1; }
        std::fwrite(node_lengths.data(), 1, node_lengths.size(), o);
        std::fclose(o);
    }
    UnstableView V;
    V.contigs = ct.view();
    V.contig_begin = begin.data(); V.nodes = nodes.data(); V.node_names = names.data();
    V.ref_off = refoff.data(); V.ref_names = refnames.data();

    if (!gaf.empty() && gaf.back() != '\n') gaf.push_back('\n');
    const u8* base = reinterpret_cast<const u8*>(gaf.data());
    std::string out;
    for (size_t i = 0; i < gaf.size();) {
        const char* nl = static_cast<const char*>(std::memchr(gaf.data() + i, '\n', gaf.size() - i));
        const u32 len = (u32)(nl - (gaf.data() + i));
        CountSink cs;
        u32 ea, eb;
        const u32 st = unstable_record(base + i, len, V, cs, ea, eb);
        if (st_is_abort(st)) {
            std::fwrite(out.data(), 1, out.size(), stdout);
            std::fflush(stdout);
            std::fprintf(stderr, "abort: status %u\n", st & 0xff);
            return 134;
        }
        if ((st & 0xff) != ST_SKIP) {
            const size_t o0 = out.size();
            out.resize(o0 + cs.n);
            StoreSink ss(reinterpret_cast<u8*>(&out[o0]));
            unstable_record(base + i, len, V, ss, ea, eb);
            if ((size_t)(ss.p - reinterpret_cast<u8*>(&out[o0])) != cs.n) { std::fprintf(stderr, "hostsim: pass mismatch\n"); return 99; }
            if ((st & 0xff) == ST_WARN_MULTIREF) std::fprintf(stderr, "[gaf2unstable] warning: Target path spans multiple reference contigs\n");
        }
        i = (size_t)(nl - gaf.data()) + 1;
    }
    std::fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
