// g2u_hostsim — CPU instantiation of the gaf2unstable per-record code (g2u_core.cuh) and of the
// host-side rGFA table builder (g2u_rgfa.hpp).
//
// TEST INFRASTRUCTURE ONLY: never linked into libg2p.so or the executables.  Mirrors
// g2p_load_rgfa + run_unstable (g2p_capi.cu) with plain host arrays so that CPU-only CI can check
// the rewrite against the reference binary and the golden vectors.
//
// usage: g2u_hostsim -g graph.gfa [-o node-lengths.tsv] <gaf|->      (stdout / exit code as gaf2unstable)
// Built with -DG2U_SIMT (build/g2u_simt): the staged kernels k_unstable_staged<false/true> themselves, run by the SIMT
// emulator (cuda_shim.hpp) over the whole input -- size pass, scan, emit pass, like run_unstable.
#if defined(G2U_SIMT)
#define HS_IMPLEMENTATION
#include "cuda_shim.hpp"
#include "../../cactus-gfa-tools_b200/csrc/g2p_kernels.cuh"
#endif
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

#if defined(G2U_SIMT)
    {
        const u64 n = gaf.size();
        std::vector<uint4> buf((n + 15) / 16 + 4);
        u8* text = reinterpret_cast<u8*>(buf.data());
        std::memcpy(text, gaf.data(), n);
        std::vector<u32> rec{0};
        for (u64 i = 0; i < n; ++i) if (text[i] == '\n') rec.push_back((u32)i + 1);
        if (n && text[n - 1] != '\n') rec.push_back((u32)n + 1);   // unterminated last line: virtual newline, as the line index does
        const u32 nrec = (u32)rec.size() - 1;
        if (nrec == 0) return 0;
        rec.push_back(0);
        std::vector<u64> off(nrec + 1, 0);
        std::vector<u32> status(nrec), wl(nrec);
        PipelineMeta meta;
        std::memset(&meta, 0, sizeof meta);
        meta.first_err = 0xFFFFFFFFu;
        const u32 ncta = (nrec + kUThreads - 1) / kUThreads;
        hs::launch(dim3(ncta), dim3(kUThreads), unstable_smem<false>(), [&] { k_unstable_staged<false>(text, n, rec.data(), nrec, V, off.data(), status.data(), nullptr, &meta, wl.data()); });
        u64 run = 0;
        for (u32 r = 0; r <= nrec; ++r) { const u64 c = r < nrec ? off[r] : 0; off[r] = run; run += c; }
        std::vector<u8> outb(run + 64, 0xEE);
        hs::launch(dim3(ncta), dim3(kUThreads), unstable_smem<true>(), [&] { k_unstable_staged<true>(text, n, rec.data(), nrec, V, off.data(), status.data(), outb.data(), &meta, wl.data()); });
        const u32 stop = meta.first_err == 0xFFFFFFFFu ? nrec : meta.first_err;
        std::fwrite(outb.data(), 1, off[stop], stdout);
        std::fflush(stdout);
        for (u32 r = 0; r < stop; ++r) if ((status[r] & 0xff) == ST_WARN_MULTIREF) std::fprintf(stderr, "[gaf2unstable] warning: Target path spans multiple reference contigs\n");
        if (stop != nrec) { std::fprintf(stderr, "abort: status %u\n", status[stop] & 0xff); return 134; }
        return 0;
    }
#endif
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
