// g2p_hostsim — CPU instantiation of the device per-record converter (g2p_core.cuh).
//
// TEST INFRASTRUCTURE ONLY.  The product (libg2p.so and the gaf2paf / gaf2unstable
// executables) never links or calls this file: it exists so that the exact code the
// sm_100a kernels execute can be fuzzed against the reference binary in a container
// that has no GPU.  It runs the same two passes as the device emitter: CountSink to
// size every record, an exclusive scan, StoreSink to write; and checks that both
// passes agree.
//
// usage: g2p_hostsim -l lengths.tsv <gaf|-> [gaf2 ...]     (exit codes as gaf2paf)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../cactus-gfa-tools_b200/csrc/g2p_core.cuh"
#include "../../cactus-gfa-tools_b200/csrc/g2p_table.hpp"

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
    const char* lengths = nullptr;
    std::vector<const char*> inputs;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "-l") && i + 1 < argc) lengths = argv[++i];
        else inputs.push_back(argv[i]);
    }
    if (!lengths || inputs.empty()) { std::fprintf(stderr, "usage: g2p_hostsim -l lengths.tsv <gaf> ...\n"); return 1; }
    std::string tsv;
    if (!slurp(lengths, tsv)) { std::fprintf(stderr, "[gaf2paf] error: unable to open %s\n", lengths); return 1; }
    HostLenTable table;
    u32 tst = build_len_table(tsv.data(), tsv.size(), table);
    if (tst != ST_OK) { std::fprintf(stderr, "abort: lengths table status %u\n", tst); return 134; }
    LenTableView T = table.view();

    std::string gaf;
    for (const char* p : inputs) {
        if (!slurp(p, gaf)) { std::fprintf(stderr, "[gaf2paf] error: unable to open input: %s\n", p); return 1; }
        if (!gaf.empty() && gaf.back() != '\n') gaf.push_back('\n');
    }
    // record index
    std::vector<size_t> starts;
    for (size_t i = 0; i < gaf.size();) {
        starts.push_back(i);
        const void* nl = std::memchr(gaf.data() + i, '\n', gaf.size() - i);
        i = (const char*)nl - gaf.data() + 1;
    }
    starts.push_back(gaf.size());
    const size_t nrec = starts.size() - 1;
    const u8* base = reinterpret_cast<const u8*>(gaf.data());

    // pass 1
    std::vector<u64> off(nrec + 1, 0);
    std::vector<u32> status(nrec, 0);
    size_t first_err = nrec;
    u32 err_a = 0, err_b = 0;
    for (size_t r = 0; r < nrec; ++r) {
        CountSink cs;
        u32 ea, eb;
        u32 len = (u32)(starts[r + 1] - starts[r] - 1);
        u32 st = convert_record(base + starts[r], len, T, cs, ea, eb);
        status[r] = st;
        u64 n = st_is_abort(st) ? 0 : cs.n;
        off[r + 1] = off[r] + n;
        if (st_is_error(st) && first_err == nrec) { first_err = r; err_a = ea; err_b = eb; }
    }
    // pass 2
    size_t upto = first_err == nrec ? nrec : first_err + 1;
    std::vector<u8> out(off[upto] + 64);
    for (size_t r = 0; r < upto; ++r) {
        if (st_is_abort(status[r])) continue;
        StoreSink ss(out.data() + off[r]);
        u32 ea, eb;
        u32 len = (u32)(starts[r + 1] - starts[r] - 1);
        u32 st = convert_record(base + starts[r], len, T, ss, ea, eb);
        if (st != status[r] || (u64)(ss.p - (out.data() + off[r])) != off[r + 1] - off[r]) {
            std::fprintf(stderr, "hostsim: pass mismatch on record %zu (st %u/%u, len %llu/%llu)\n", r, status[r], st,
                         (unsigned long long)(ss.p - (out.data() + off[r])), (unsigned long long)(off[r + 1] - off[r]));
            return 99;
        }
    }
    std::fwrite(out.data(), 1, off[upto], stdout);
    std::fflush(stdout);
    if (first_err != nrec) {
        u32 st = status[first_err] & 0xff;
        if (st == ST_ERR_NAME) {
            std::string nm((const char*)base + starts[first_err] + err_a, err_b - err_a);
            std::fprintf(stderr, "[gaf2paf] error: unable to find %s in lengths map\n", nm.c_str());
            return 1;
        }
        if (st == ST_ERR_NOCG) {
            std::fprintf(stderr, "[gaf2paf] error: cg cigar not found. This tool only works on output of minigraph -c\n");
            return 1;
        }
        std::fprintf(stderr, "abort: record %zu status %u aux %u\n", first_err, st, (status[first_err] >> 8) & 0xff);
        return 134;
    }
    return 0;
}
