// g2p_stub_capi — a CPU stand-in for the gaf2paf half of the C-ABI (include/g2p.h).
//
// TEST INFRASTRUCTURE ONLY: never part of libg2p.so or of the shipped executables.  CPU-only CI links
// csrc/gaf2paf_main.cpp against this file (build/gaf2paf_stub) so that the HOST logic of the drop-in
// executable -- the reader / converter / writer pipeline of cli_pipeline.hpp, chunk cutting, multi-file and
// stdin handling, G2P_GPUS round-robin, ordered output, error and exit-code paths -- runs without a GPU.
// The per-record conversion is the product's own scalar device code (g2p_core.cuh convert_record)
// instantiated for the host, exactly like build/g2p_hostsim; results are double-buffered per context as
// the real library's are.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../cactus-gfa-tools_b200/csrc/g2p_core.cuh"
#include "../../cactus-gfa-tools_b200/csrc/g2p_table.hpp"
#include "../../cactus-gfa-tools_b200/csrc/g2p_errfmt.hpp"

using namespace g2p;

struct g2p_ctx {
    HostLenTable table;
    bool have_table = false;
    std::vector<u8> out[2];
    int cur = 0;
    std::string err;
};

extern "C" {

int g2p_create(int device, g2p_ctx** out) {
    if (!out || device < 0 || device >= 64) return G2P_E_NO_DEVICE;
    *out = new g2p_ctx();
    return G2P_OK;
}
void g2p_destroy(g2p_ctx* ctx) { delete ctx; }
const char* g2p_last_error(const g2p_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
void* g2p_host_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void g2p_host_free(void* p) { std::free(p); }

int g2p_load_lengths(g2p_ctx* ctx, const char* tsv, size_t n) {
    if (!ctx) return G2P_E_ARG;
    ctx->table = HostLenTable();
    if (build_len_table(tsv, n, ctx->table) != ST_OK) return G2P_E_TABLE;
    ctx->have_table = true;
    return G2P_OK;
}

int g2p_convert_host(g2p_ctx* ctx, const char* gaf, size_t n, const char** out, g2p_result* res) {
    if (!ctx || !res || !out) return G2P_E_ARG;
    if (!ctx->have_table) return G2P_E_NOTABLE;
    std::memset(res, 0, sizeof *res);
    std::vector<u8>& o = ctx->out[ctx->cur ^= 1];
    o.clear();
    const LenTableView T = ctx->table.view();
    const u8* base = reinterpret_cast<const u8*>(gaf);
    size_t i = 0;
    u64 r = 0;
    while (i < n) {
        const void* nl = std::memchr(gaf + i, '\n', n - i);
        const size_t e = nl ? (size_t)(static_cast<const char*>(nl) - gaf) : n;
        CountSink cs;
        u32 ea = 0, eb = 0;
        const u32 st = convert_record(base + i, (u32)(e - i), T, cs, ea, eb);
        if (!st_is_abort(st) && (st & 0xff) != ST_SKIP && cs.n) {
            const size_t at = o.size();
            o.resize(at + cs.n + 64);
            StoreSink ss(o.data() + at);
            convert_record(base + i, (u32)(e - i), T, ss, ea, eb);
            o.resize(at + cs.n);
        }
        if (st_is_error(st)) {
            res->rec_status = st & 0xff;
            res->rec_aux = (st >> 8) & 0xff;
            res->err_record = r;
            res->err_name_off = i + ea;
            res->err_name_len = eb - ea;
            ++r;
            break;
        }
        ++r;
        i = e + 1;
    }
    res->n_records = r;
    res->out_bytes = o.size();
    o.push_back(0);
    *out = reinterpret_cast<const char*>(o.data());
    return G2P_OK;
}

int g2p_format_error(const g2p_result* res, const char* gaf, size_t n, char* buf, size_t cap) {
    return g2p_errfmt::format_error(res, gaf, n, buf, cap);
}

}  // extern "C"
