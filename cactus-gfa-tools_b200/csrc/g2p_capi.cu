// g2p_capi.cu — C-ABI (include/g2p.h) over the sm_100a kernels.
//
// Owns all CUDA state: the device copy of the name->length table, grow-only work
// buffers (record index, status, output offsets, PAF output) and the pinned host
// staging used by the host-buffer entry point.  No CPU conversion path exists here:
// every byte of PAF is produced by k_convert<true>.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/g2p.h"
#include "g2p_kernels.cuh"
#include "g2p_table.hpp"

using namespace g2p;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct g2p_ctx {
    int device = 0;
    std::string err;
    // table
    DevBuf d_slots, d_arena;
    LenTableView table{nullptr, nullptr, 0};
    uint64_t table_entries = 0;
    bool have_table = false;
    // work buffers
    DevBuf d_in, d_tiles, d_rec, d_status, d_off, d_blocks, d_out, d_meta, d_list, d_list2;
    int n_sm = 148;
    PinBuf h_out, h_meta;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[8] = {};
};

#define G2P_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            return G2P_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

extern "C" {

int g2p_create(int device, g2p_ctx** out) {
    if (!out) return G2P_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return G2P_E_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return G2P_E_NO_DEVICE;
    g2p_ctx* ctx = new (std::nothrow) g2p_ctx();
    if (!ctx) return G2P_E_ARG;
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return G2P_E_NO_DEVICE; }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    cudaFuncSetAttribute(k_short<kSG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShortSmem);
    cudaFuncSetAttribute(k_short<kSG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShortSmem);
    cudaFuncSetAttribute(k_long<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmem);
    cudaFuncSetAttribute(k_long<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmem);
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (ctx->d_meta.ensure(sizeof(PipelineMeta)) != cudaSuccess || ctx->h_meta.ensure(sizeof(PipelineMeta)) != cudaSuccess) {
        delete ctx;
        return G2P_E_NO_DEVICE;
    }
    *out = ctx;
    return G2P_OK;
}

void g2p_destroy(g2p_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&ctx->d_slots, &ctx->d_arena, &ctx->d_in, &ctx->d_tiles, &ctx->d_rec, &ctx->d_status, &ctx->d_off, &ctx->d_blocks,
                      &ctx->d_out, &ctx->d_meta, &ctx->d_list, &ctx->d_list2})
        b->release();
    ctx->h_out.release();
    ctx->h_meta.release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* g2p_last_error(const g2p_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

void* g2p_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void g2p_host_free(void* p) { if (p) cudaFreeHost(p); }

int g2p_copy_to_device(void* d_dst, const void* h_src, size_t bytes) {
    return cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? G2P_OK : G2P_E_CUDA;
}
int g2p_copy_to_host(void* h_dst, const void* d_src, size_t bytes) {
    return cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? G2P_OK : G2P_E_CUDA;
}

int g2p_load_lengths(g2p_ctx* ctx, const char* tsv, size_t n) {
    if (!ctx || (!tsv && n)) return G2P_E_ARG;
    G2P_CUDA(cudaSetDevice(ctx->device));
    HostLenTable t;
    u32 st = build_len_table(tsv, n, t);
    if (st != ST_OK) { ctx->err = "lengths table: std::stol would throw"; return G2P_E_TABLE; }
    G2P_CUDA(ctx->d_slots.ensure(t.slots.size() * sizeof(LenSlot)));
    G2P_CUDA(ctx->d_arena.ensure(t.arena.size()));
    G2P_CUDA(cudaMemcpy(ctx->d_slots.p, t.slots.data(), t.slots.size() * sizeof(LenSlot), cudaMemcpyHostToDevice));
    G2P_CUDA(cudaMemcpy(ctx->d_arena.p, t.arena.data(), t.arena.size(), cudaMemcpyHostToDevice));
    ctx->table.slots = static_cast<const LenSlot*>(ctx->d_slots.p);
    ctx->table.arena = static_cast<const u8*>(ctx->d_arena.p);
    ctx->table.nslots = (u32)t.slots.size();
    ctx->table_entries = t.n_entries;
    ctx->have_table = true;
    return G2P_OK;
}

uint64_t g2p_table_entries(const g2p_ctx* ctx) { return ctx ? ctx->table_entries : 0; }

// Line index into ctx->d_rec; leaves meta (n_lines, n_records) in ctx->h_meta.
static int run_index(g2p_ctx* ctx, const u8* d_text, size_t n, cudaStream_t st, uint32_t* launches) {
    const u32 ntiles = (u32)((n + kIdxTile - 1) / kIdxTile);
    G2P_CUDA(ctx->d_tiles.ensure(((size_t)ntiles + 1) * sizeof(u32)));
    PipelineMeta* d_meta = static_cast<PipelineMeta*>(ctx->d_meta.p);
    u32* d_tiles = static_cast<u32*>(ctx->d_tiles.p);
    if (ntiles) { k_count_lines<<<ntiles, kIdxThreads, 0, st>>>(d_text, n, d_tiles); ++*launches; }
    k_scan_tiles<<<1, 1024, 0, st>>>(d_tiles, ntiles, d_text, n, d_meta);
    ++*launches;
    G2P_CUDA(cudaMemcpyAsync(ctx->h_meta.p, d_meta, sizeof(PipelineMeta), cudaMemcpyDeviceToHost, st));
    G2P_CUDA(cudaStreamSynchronize(st));
    const PipelineMeta* hm = static_cast<const PipelineMeta*>(ctx->h_meta.p);
    G2P_CUDA(ctx->d_rec.ensure(((size_t)hm->n_records + 2) * sizeof(u32)));
    if (ntiles) {
        k_fill_lines<<<ntiles, kIdxThreads, 0, st>>>(d_text, n, d_tiles, static_cast<u32*>(ctx->d_rec.p), d_meta);
        ++*launches;
    }
    G2P_CUDA(cudaGetLastError());
    return G2P_OK;
}

int g2p_index_lines(g2p_ctx* ctx, const void* d_text, size_t n, const uint32_t** d_starts, uint64_t* n_lines, void* stream) {
    if (!ctx || !d_starts || !n_lines) return G2P_E_ARG;
    if (n >= 0xFFFFFFF0ULL) return G2P_E_TOOBIG;
    G2P_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->own_stream;
    uint32_t launches = 0;
    int rc = run_index(ctx, static_cast<const u8*>(d_text), n, st, &launches);
    if (rc) return rc;
    G2P_CUDA(cudaStreamSynchronize(st));
    *d_starts = static_cast<const uint32_t*>(ctx->d_rec.p);
    *n_lines = static_cast<const PipelineMeta*>(ctx->h_meta.p)->n_records;
    return G2P_OK;
}

int g2p_convert_device(g2p_ctx* ctx, const void* d_gaf_v, size_t n, void** d_out, g2p_result* res, void* stream) {
    if (!ctx || !res || !d_out) return G2P_E_ARG;
    if (!ctx->have_table) return G2P_E_NOTABLE;
    if (n >= 0xFFFFFFF0ULL) return G2P_E_TOOBIG;
    if ((reinterpret_cast<uintptr_t>(d_gaf_v) & 15) != 0) { ctx->err = "d_gaf must be 16-byte aligned"; return G2P_E_ARG; }
    std::memset(res, 0, sizeof *res);
    *d_out = nullptr;
    G2P_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->own_stream;
    const u8* d_gaf = static_cast<const u8*>(d_gaf_v);
    uint32_t launches = 0;

    G2P_CUDA(cudaEventRecord(ctx->ev[0], st));
    int rc = run_index(ctx, d_gaf, n, st, &launches);
    if (rc) return rc;
    G2P_CUDA(cudaEventRecord(ctx->ev[1], st));
    PipelineMeta* hm = static_cast<PipelineMeta*>(ctx->h_meta.p);
    PipelineMeta* d_meta = static_cast<PipelineMeta*>(ctx->d_meta.p);
    const u32 nrec = hm->n_records;
    res->n_records = nrec;
    if (nrec == 0) {
        G2P_CUDA(ctx->d_out.ensure(256));
        *d_out = ctx->d_out.p;
        G2P_CUDA(cudaStreamSynchronize(st));
        res->gpu_launches = launches;
        return G2P_OK;
    }
    G2P_CUDA(ctx->d_status.ensure((size_t)nrec * sizeof(u32)));
    G2P_CUDA(ctx->d_off.ensure(((size_t)nrec + 1) * sizeof(u64)));
    const u32 nscan = (nrec + kScanTile - 1) / kScanTile;
    G2P_CUDA(ctx->d_blocks.ensure((size_t)nscan * sizeof(u64)));
    u32* d_rec = static_cast<u32*>(ctx->d_rec.p);
    u32* d_status = static_cast<u32*>(ctx->d_status.p);
    u64* d_off = static_cast<u64*>(ctx->d_off.p);
    u64* d_blocks = static_cast<u64*>(ctx->d_blocks.p);
    G2P_CUDA(ctx->d_list.ensure((size_t)nrec * sizeof(u32)));
    G2P_CUDA(ctx->d_list2.ensure((size_t)nrec * sizeof(u32)));
    u32* d_list = static_cast<u32*>(ctx->d_list.p);
    u32* d_list2 = static_cast<u32*>(ctx->d_list2.p);
    const u32 ncta = (nrec + kShortRecsPerCta - 1) / kShortRecsPerCta;
    const u32 nlong = (u32)ctx->n_sm * 4u;
    const u32 nlist = std::min<u32>((nrec + kListThreads - 1) / kListThreads, (u32)ctx->n_sm * 16u);
    ShortArgs sa{d_gaf, (u64)n, d_rec, nrec, ctx->table, d_off, d_status, nullptr, d_list, &d_meta->n_deleg};
    LongArgs la{d_gaf, (u64)n, d_rec, ctx->table, d_off, d_status, nullptr, d_list, &d_meta->n_deleg, d_list2, &d_meta->n_deleg2};

    // pass 1: sizes + status.  k_short takes the short canonical records, k_long what it left,
    // the general kernel what neither converts (non-canonical or erroneous records).
    k_short<kSG, false><<<ncta, kSThreads, kShortSmem, st>>>(sa);
    k_long<false><<<nlong, kLThreads, kLongSmem, st>>>(la);
    k_convert_list<false><<<nlist, kListThreads, 0, st>>>(d_gaf, d_rec, ctx->table, d_off, d_status, nullptr, d_meta, d_list2, &d_meta->n_deleg2);
    launches += 3;
    G2P_CUDA(cudaEventRecord(ctx->ev[2], st));
    // exclusive scan -> offsets
    k_scan_reduce<<<nscan, kScanThreads, 0, st>>>(d_off, nrec, d_blocks);
    k_scan_blocks<<<1, 1024, 0, st>>>(d_blocks, nscan, d_meta);
    k_scan_apply<<<nscan, kScanThreads, 0, st>>>(d_off, nrec, d_blocks, d_meta);
    launches += 3;
    G2P_CUDA(cudaMemcpyAsync(hm, d_meta, sizeof(PipelineMeta), cudaMemcpyDeviceToHost, st));
    G2P_CUDA(cudaStreamSynchronize(st));
    const u64 out_total = hm->out_total;
    res->n_long = hm->n_deleg;
    res->n_delegated = hm->n_deleg2;
    G2P_CUDA(ctx->d_out.ensure(out_total + 256));
    u8* d_o = static_cast<u8*>(ctx->d_out.p);
    // pass 2: emit
    G2P_CUDA(cudaEventRecord(ctx->ev[3], st));
    sa.out = d_o;
    la.out = d_o;
    k_short<kSG, true><<<ncta, kSThreads, kShortSmem, st>>>(sa);
    ++launches;
    if (hm->n_deleg) {
        k_long<true><<<nlong, kLThreads, kLongSmem, st>>>(la);
        ++launches;
    }
    if (hm->n_deleg2) {
        k_convert_list<true><<<nlist, kListThreads, 0, st>>>(d_gaf, d_rec, ctx->table, d_off, d_status, d_o, d_meta, d_list2, &d_meta->n_deleg2);
        ++launches;
    }
    G2P_CUDA(cudaEventRecord(ctx->ev[4], st));
    res->out_bytes = out_total;
    if (hm->first_err != 0xFFFFFFFFu) {
        k_diagnose<<<1, 1, 0, st>>>(d_gaf, d_rec, ctx->table, d_off, d_meta);
        ++launches;
        G2P_CUDA(cudaMemcpyAsync(hm, d_meta, sizeof(PipelineMeta), cudaMemcpyDeviceToHost, st));
    }
    G2P_CUDA(cudaStreamSynchronize(st));
    G2P_CUDA(cudaGetLastError());
    if (hm->first_err != 0xFFFFFFFFu) {
        res->rec_status = hm->err_status & 0xff;
        res->rec_aux = (hm->err_status >> 8) & 0xff;
        res->err_record = hm->first_err;
        res->err_name_off = (u64)hm->err_rec_start + hm->err_a;
        res->err_name_len = hm->err_b - hm->err_a;
        res->out_bytes = hm->err_out_end;   // what the reference has written before it stops
    }
    cudaEventElapsedTime(&res->index_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&res->size_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&res->emit_ms, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&res->device_ms, ctx->ev[0], ctx->ev[4]);
    res->gpu_launches = launches;
    *d_out = d_o;
    return G2P_OK;
}

int g2p_convert_host(g2p_ctx* ctx, const char* gaf, size_t n, const char** out, g2p_result* res) {
    if (!ctx || !res || !out || (!gaf && n)) return G2P_E_ARG;
    if (n >= 0xFFFFFFF0ULL) return G2P_E_TOOBIG;
    *out = nullptr;
    G2P_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    G2P_CUDA(ctx->d_in.ensure(n + 256));
    if (n) G2P_CUDA(cudaMemcpyAsync(ctx->d_in.p, gaf, n, cudaMemcpyHostToDevice, st));
    void* d_o = nullptr;
    int rc = g2p_convert_device(ctx, ctx->d_in.p, n, &d_o, res, st);
    if (rc) return rc;
    G2P_CUDA(ctx->h_out.ensure(res->out_bytes + 1));
    if (res->out_bytes) G2P_CUDA(cudaMemcpyAsync(ctx->h_out.p, d_o, res->out_bytes, cudaMemcpyDeviceToHost, st));
    G2P_CUDA(cudaStreamSynchronize(st));
    *out = static_cast<const char*>(ctx->h_out.p);
    return G2P_OK;
}

int g2p_format_error(const g2p_result* res, const char* gaf, size_t n, char* buf, size_t cap) {
    if (!res || !buf || cap == 0) return G2P_E_ARG;
    buf[0] = 0;
    if (res->rec_status == G2P_REC_ERR_NAME) {
        std::string name;
        if (gaf && res->err_name_off + res->err_name_len <= n) name.assign(gaf + res->err_name_off, res->err_name_len);
        std::snprintf(buf, cap, "[gaf2paf] error: unable to find %s in lengths map\n", name.c_str());
    } else if (res->rec_status == G2P_REC_ERR_NOCG) {
        std::snprintf(buf, cap, "[gaf2paf] error: cg cigar not found. This tool only works on output of minigraph -c\n");
    } else if (res->rec_status >= G2P_REC_ABORT) {
        static const char* what[] = {"Error parsing GAF column", "Error parsing GAF strand", "Error parsing GAF range", "stol (invalid argument)",
                                     "stol (out of range)", "Unable to parse optional tag", "Duplicate optional field found",
                                     "malformed cg cigar", "assertion failed"};
        unsigned k = res->rec_status - G2P_REC_ABORT;
        if (res->rec_status == G2P_REC_ABORT)
            std::snprintf(buf, cap, "terminate: %s %u (record %llu)\n", what[0], res->rec_aux, (unsigned long long)res->err_record);
        else
            std::snprintf(buf, cap, "terminate: %s (record %llu)\n", k < 9 ? what[k] : "abort", (unsigned long long)res->err_record);
    }
    return G2P_OK;
}

}  // extern "C"
