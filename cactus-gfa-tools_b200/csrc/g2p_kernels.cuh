// g2p_kernels.cuh — sm_100a kernels of the GAF -> PAF pipeline.
//
//   k_fuse<Cfg> (g2p_fuse.cuh)                     the whole conversion in one pass, for blocks of canonical records of <= 1000 bytes
//   k_count_lines / k_scan_tiles / k_fill_lines   newline index (default); k_index1: the same in one pass (G2P_ONE_PASS_INDEX=1)
//   k_rec (or k_short) -> k_par_* (g2p_par.cuh) -> k_long<false> -> k_convert_list<false>
//                                                  pass 1: per-record PAF byte length, status, line descriptors; every kernel
//                                                  takes what the one before it left (<= 240 bytes / canonical / any record)
//   k_scan_*                                       exclusive scans: byte lengths -> output offsets, line counts -> line slots + line map
//   k_emit_lines<DENSE>                            pass 2: one thread per PAF line writes the bytes
//   k_long<true> / k_convert_list<true>            pass 2 for records without descriptors
//   k_diagnose                                     details of the first failing record
//   k_unstable_staged<EMIT> (k_unstable)           gaf2unstable; k_filter_* (g2p_filter.cuh): gaffilter
//
// All work is byte / integer; the pipeline is bound by HBM traffic and by the
// latency of serial per-record parsing, so the kernels stage contiguous record
// batches through shared memory with 128-bit coalesced accesses in both directions.
#pragma once
#if defined(G2P_HOSTSIM)
#include "cuda_shim.hpp"
#else
#include <cuda_runtime.h>
#define G2P_NOINLINE __noinline__
#endif

#include "g2p_core.cuh"
#include "g2p_short.cuh"
#include "g2p_rec.cuh"
#include "g2p_long.cuh"
#include "g2u_core.cuh"

namespace g2p {

struct PipelineMeta {
    u32 n_lines;       // '\n' count
    u32 n_records;     // lines incl. an unterminated last one
    u32 first_err;     // smallest failing record index (0xFFFFFFFF = none)
    u32 long_cursor;   // next entry of the delegate list a warp of k_long<false> takes
    u64 out_total;     // bytes of PAF for all records
    // k_diagnose
    u32 err_status;
    u32 err_a, err_b;  // name span relative to the record start
    u32 err_rec_start;
    u64 err_out_end;   // output offset just after the failing record's (partial) output
    u32 n_deleg;       // records k_short left to k_long
    u32 n_deleg2;      // records k_long left to the general kernel
    u32 n_desc;        // line-descriptor slots reserved by k_long for full batches (32 per batch, lower half of the array)
    u32 n_desc2;       // ... and for batches of at most 8 lines (as many as needed, rounded up to four; upper half)
    u32 n_reject;      // records k_par left to k_long
    u32 legacy_long;   // some k_long record is not described: run k_long<true>
    u64 lines_total;   // PAF lines of the records k_short converted
};

// ------------------------------------------------------------------------------
// Line index.  A tile is 256 threads x 4 x 16 B = 16 KiB, loaded as coalesced
// 128-bit vectors (vector v = k*256 + t).  Newlines are found with per-byte SIMD
// compares; ranks come from one block scan over four packed 16-bit counters.
// ------------------------------------------------------------------------------
constexpr int kIdxThreads = 256;
constexpr int kIdxVec = 4;
constexpr u32 kIdxTile = kIdxThreads * kIdxVec * 16;

__device__ __forceinline__ uint4 load_vec_guarded(const u8* base, u64 off, u64 n) { return ldg_vec_guarded(base, off, n); }

__device__ __forceinline__ u32 nl_bits(u32 w) { return zero_bytes(w ^ 0x0A0A0A0Au); }

__global__ void __launch_bounds__(kIdxThreads) k_count_lines(const u8* __restrict__ text, u64 n, u32* __restrict__ tile_count) {
    const u64 base = (u64)blockIdx.x * kIdxTile;
    u32 cnt = 0;
#pragma unroll
    for (int k = 0; k < kIdxVec; ++k) {
        u64 off = base + ((u64)k * kIdxThreads + threadIdx.x) * 16;
        if (off < n) {
            uint4 v = load_vec_guarded(text, off, n);
            cnt += __popc(nl_bits(v.x)) + __popc(nl_bits(v.y)) + __popc(nl_bits(v.z)) + __popc(nl_bits(v.w));
        }
    }
    // block reduce
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    __shared__ u32 wsum[kIdxThreads / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 s = 0;
        for (int i = 0; i < kIdxThreads / 32; ++i) s += wsum[i];
        tile_count[blockIdx.x] = s;
    }
}

// One CTA: exclusive scan of the per-tile counts in place; fills the meta record.
__global__ void __launch_bounds__(1024) k_scan_tiles(u32* __restrict__ tile_count, u32 ntiles, const u8* __restrict__ text, u64 n,
                                                     PipelineMeta* __restrict__ meta) {
    __shared__ u32 part[1024];
    const u32 per = (ntiles + 1023) / 1024;
    const u32 a = threadIdx.x * per;
    const u32 b = min(a + per, ntiles);
    u32 s = 0;
    for (u32 i = a; i < b; ++i) s += tile_count[i];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele over 1024 partials
    for (u32 o = 1; o < 1024; o <<= 1) {
        u32 v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    u32 run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    for (u32 i = a; i < b; ++i) { u32 c = tile_count[i]; tile_count[i] = run; run += c; }
    if (threadIdx.x == 1023) {
        u32 lines = part[1023];
        meta->n_lines = lines;
        meta->n_records = lines + ((n > 0 && text[n - 1] != '\n') ? 1u : 0u);
        meta->first_err = 0xFFFFFFFFu;
        meta->out_total = 0;
        meta->err_status = 0;
        meta->n_deleg = 0;
        meta->long_cursor = 0;
        meta->n_deleg2 = 0;
        meta->n_desc = 0;
        meta->n_desc2 = 0;
        meta->n_reject = 0;
        meta->legacy_long = 0;
        meta->lines_total = 0;
    }
}

__global__ void __launch_bounds__(kIdxThreads) k_fill_lines(const u8* __restrict__ text, u64 n, const u32* __restrict__ tile_off,
                                                            u32* __restrict__ rec_start, const PipelineMeta* __restrict__ meta) {
    const u64 base = (u64)blockIdx.x * kIdxTile;
    uint4 v[kIdxVec];
    u64 packed = 0;   // four 16-bit counters, one per k
#pragma unroll
    for (int k = 0; k < kIdxVec; ++k) {
        u64 off = base + ((u64)k * kIdxThreads + threadIdx.x) * 16;
        v[k] = make_uint4(0, 0, 0, 0);
        if (off < n) v[k] = load_vec_guarded(text, off, n);
        u32 c = __popc(nl_bits(v[k].x)) + __popc(nl_bits(v[k].y)) + __popc(nl_bits(v[k].z)) + __popc(nl_bits(v[k].w));
        packed |= (u64)c << (16 * k);
    }
    // block exclusive scan of `packed`
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 incl = packed;
    for (int o = 1; o < 32; o <<= 1) {
        u64 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += up;
    }
    __shared__ u64 wtot[kIdxThreads / 32];
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    u64 wpre = 0, total = 0;
    for (int i = 0; i < kIdxThreads / 32; ++i) { if (i < (int)warp) wpre += wtot[i]; total += wtot[i]; }
    const u64 excl = wpre + incl - packed;
    u32 kbase = tile_off[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (n > 0) rec_start[0] = 0;
        if (meta->n_records != meta->n_lines) rec_start[meta->n_records] = (u32)n + 1;   // unterminated last line
    }
#pragma unroll
    for (int k = 0; k < kIdxVec; ++k) {
        u32 rank = kbase + (u32)((excl >> (16 * k)) & 0xffff);
        const u64 off = base + ((u64)k * kIdxThreads + threadIdx.x) * 16;
        const u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u32 m = nl_bits(w[j]);
            while (m) {
                u32 bit = __ffs(m) - 1;
                m &= m - 1;
                u64 p = off + j * 4 + (bit >> 3);
                rec_start[rank + 1] = (u32)(p + 1);
                ++rank;
            }
        }
        kbase += (u32)((total >> (16 * k)) & 0xffff);
    }
}

// ------------------------------------------------------------------------------
// Single-pass line index.  Each CTA takes the next 16 KiB tile (ticket counter), brings it into
// shared memory with one TMA bulk copy (cp.async.bulk + mbarrier), counts its newlines with SWAR
// compares, ranks them with a block scan, obtains the number of newlines before the tile with a
// decoupled look-back over the per-tile status words (flag in the top two bits, value below:
// 1 = the tile's own count, 2 = inclusive prefix), and writes the record starts.  The text is
// read from HBM once.  rec_start has room for `cap` entries; when the text holds more lines than
// that (blank-line floods), nothing past the capacity is written and the caller falls back to
// the counting kernels above.
// ------------------------------------------------------------------------------
constexpr u64 kIdxFlagAgg = 1ull << 62, kIdxFlagPre = 2ull << 62, kIdxValMask = (1ull << 62) - 1;

__global__ void __launch_bounds__(kIdxThreads) k_index1(const u8* __restrict__ text, u64 n, u32 ntiles, u64* tile_status, u32* ticket,
                                                        u32* __restrict__ rec_start, u32 cap, PipelineMeta* __restrict__ meta) {
    __shared__ __align__(16) u8 s_text[kIdxTile];
    __shared__ u64 wtot[kIdxThreads / 32];
    __shared__ u32 s_tile, s_prefix;
#if !defined(G2P_HOSTSIM)
    __shared__ __align__(8) u64 s_bar;
    if (threadIdx.x == 0) { s_tile = atomicAdd(ticket, 1u); mbar_init(&s_bar, 1); fence_mbar_init(); }
#else
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
#endif
    __syncthreads();
    const u32 tile = s_tile;
    const u64 base = (u64)tile * kIdxTile;
    const u32 bytes = (u32)(n - base < (u64)kIdxTile ? n - base : (u64)kIdxTile);
    const u32 full = bytes & ~15u;
#if !defined(G2P_HOSTSIM)
    if (threadIdx.x == 0 && full) { mbar_expect_tx(&s_bar, full); bulk_g2s(s_text, text + base, full, &s_bar); }
    if (threadIdx.x < bytes - full) s_text[full + threadIdx.x] = text[base + full + threadIdx.x];   // last partial vector
    if (full) { u32 spins = 0; while (!mbar_try_wait(&s_bar, 0)) { if (++spins > (1u << 24)) __trap(); } }
#else
    (void)full;
    for (u32 i = threadIdx.x; i < bytes; i += kIdxThreads) s_text[i] = text[base + i];
#endif
    __syncthreads();
    uint4 v[kIdxVec];
    u64 packed = 0;   // four 16-bit counters, one per k
#pragma unroll
    for (int k = 0; k < kIdxVec; ++k) {
        const u32 off = ((u32)k * kIdxThreads + threadIdx.x) * 16;
        v[k] = make_uint4(0, 0, 0, 0);
        if (off < bytes) {
            v[k] = *reinterpret_cast<const uint4*>(s_text + off);
            if (off + 16 > bytes) {   // bytes past the end of the text are not part of it
                u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
                for (u32 i = bytes - off; i < 16; ++i) w[i >> 2] &= ~(0xffu << (8 * (i & 3)));
                v[k] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        const u32 c = __popc(nl_bits(v[k].x)) + __popc(nl_bits(v[k].y)) + __popc(nl_bits(v[k].z)) + __popc(nl_bits(v[k].w));
        packed |= (u64)c << (16 * k);
    }
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 incl = packed;
    for (int o = 1; o < 32; o <<= 1) {
        const u64 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += up;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    u64 wpre = 0, total = 0;
    for (int i = 0; i < kIdxThreads / 32; ++i) { if (i < (int)warp) wpre += wtot[i]; total += wtot[i]; }
    const u64 excl = wpre + incl - packed;
    const u32 tile_count = (u32)((total & 0xffff) + ((total >> 16) & 0xffff) + ((total >> 32) & 0xffff) + ((total >> 48) & 0xffff));
    // ---- decoupled look-back (warp 0)
    if (warp == 0) {
        volatile u64* st = tile_status;
        u64 prefix = 0;
        if (tile > 0) {
            if (lane == 0) { st[tile] = kIdxFlagAgg | tile_count; }
            __syncwarp();
            int j = (int)tile - 1;
            for (;;) {
                const int idx = j - (int)lane;
                u64 w = kIdxFlagPre;   // tiles before the first: prefix 0
                if (idx >= 0) {
                    u32 spins = 0;
                    while (((w = st[idx]) >> 62) == 0) { if (++spins > (1u << 26)) __trap(); }
                }
                const u32 pre = __ballot_sync(0xffffffffu, (w >> 62) == 2);
                const u32 upto = pre ? (u32)__ffs((int)pre) - 1u : 31u;   // lanes 0..upto contribute
                u64 val = lane <= upto ? (w & kIdxValMask) : 0;
                for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
                prefix += __shfl_sync(0xffffffffu, val, 0);
                if (pre) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            st[tile] = kIdxFlagPre | (prefix + tile_count);
            s_prefix = (u32)prefix;
            if (tile == ntiles - 1) {   // totals + the fields the later kernels expect initialised
                const u32 lines = (u32)prefix + tile_count;
                const u32 recs = lines + ((n > 0 && text[n - 1] != '\n') ? 1u : 0u);
                meta->n_lines = lines;
                meta->n_records = recs;
                meta->first_err = 0xFFFFFFFFu;
                meta->out_total = 0;
                meta->err_status = 0;
                meta->n_deleg = 0; meta->long_cursor = 0; meta->n_deleg2 = 0; meta->n_desc = 0; meta->n_desc2 = 0; meta->n_reject = 0; meta->legacy_long = 0; meta->lines_total = 0;
                if (recs != lines && recs < cap) rec_start[recs] = (u32)n + 1;   // unterminated last line
            }
            if (tile == 0 && cap) rec_start[0] = 0;
        }
    }
    __syncthreads();
    u32 kbase = s_prefix;
#pragma unroll
    for (int k = 0; k < kIdxVec; ++k) {
        u32 rank = kbase + (u32)((excl >> (16 * k)) & 0xffff);
        const u64 off = base + ((u64)k * kIdxThreads + threadIdx.x) * 16;
        const u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u32 m = nl_bits(w[j]);
            while (m) {
                const u32 bit = __ffs(m) - 1;
                m &= m - 1;
                const u64 p = off + j * 4 + (bit >> 3);
                if (rank + 1 < cap) rec_start[rank + 1] = (u32)(p + 1);
                ++rank;
            }
        }
        kbase += (u32)((total >> (16 * k)) & 0xffff);
    }
}

// ------------------------------------------------------------------------------
// Exclusive scan of u64 values in place (three small kernels).
// ------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr u32 kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ u64 block_excl_scan_u64(u64 v, u64& total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        u64 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += up;
    }
    __shared__ u64 wt[kScanThreads / 32];
    __syncthreads();
    if (lane == 31) wt[warp] = incl;
    __syncthreads();
    u64 pre = 0, tot = 0;
    for (int i = 0; i < kScanThreads / 32; ++i) { if (i < (int)warp) pre += wt[i]; tot += wt[i]; }
    total = tot;
    return pre + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const u64* __restrict__ x, u32 n, u64* __restrict__ block_sum) {
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < n) s += x[base + i];
    u64 total;
    block_excl_scan_u64(s, total);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_blocks(u64* __restrict__ block_sum, u32 nblocks, u64* __restrict__ total_out) {
    __shared__ u64 part[1024];
    const u32 per = (nblocks + 1023) / 1024;
    const u32 a = threadIdx.x * per, b = min(a + per, nblocks);
    u64 s = 0;
    for (u32 i = a; i < b; ++i) s += block_sum[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (u32 o = 1; o < 1024; o <<= 1) {
        u64 v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    u64 run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    for (u32 i = a; i < b; ++i) { u64 c = block_sum[i]; block_sum[i] = run; run += c; }
    if (threadIdx.x == 1023) *total_out = part[1023];
}

// x[0..n) lengths -> exclusive offsets; x[n] = total
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(u64* __restrict__ x, u32 n, const u64* __restrict__ block_off,
                                                             const u64* __restrict__ total_in) {
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    u64 v[kScanItems];
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { v[i] = base + i < n ? x[base + i] : 0; s += v[i]; }
    u64 total;
    u64 run = block_off[blockIdx.x] + block_excl_scan_u64(s, total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) x[base + i] = run;
        run += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) x[n] = *total_in;
}

// The two per-record arrays of gaf2paf (PAF bytes -> output offsets, PAF lines -> line slots) share
// the three launches; the apply pass also writes the line map k_emit_lines reads (it has the
// record's output offset, its first line slot and its line count in registers at that point).
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce2(const u64* __restrict__ x, const u64* __restrict__ y, u32 n,
                                                               u64* __restrict__ bsx, u64* __restrict__ bsy) {
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    u64 sx = 0, sy = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < n) { sx += x[base + i]; sy += y[base + i]; }
    u64 tx, ty;
    block_excl_scan_u64(sx, tx);
    block_excl_scan_u64(sy, ty);
    if (threadIdx.x == 0) { bsx[blockIdx.x] = tx; bsy[blockIdx.x] = ty; }
}

__global__ void __launch_bounds__(1024) k_scan_blocks2(u64* __restrict__ bsx, u64* __restrict__ bsy, u32 nblocks, u64* __restrict__ total_x,
                                                       u64* __restrict__ total_y) {
    __shared__ u64 part[1024];
    const u32 per = (nblocks + 1023) / 1024;
    const u32 a = threadIdx.x * per, b = min(a + per, nblocks);
    for (int which = 0; which < 2; ++which) {
        u64* bs = which ? bsy : bsx;
        u64 s = 0;
        for (u32 i = a; i < b; ++i) s += bs[i];
        __syncthreads();
        part[threadIdx.x] = s;
        __syncthreads();
        for (u32 o = 1; o < 1024; o <<= 1) {
            u64 v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
            __syncthreads();
            part[threadIdx.x] += v;
            __syncthreads();
        }
        u64 run = threadIdx.x ? part[threadIdx.x - 1] : 0;
        for (u32 i = a; i < b; ++i) { u64 c = bs[i]; bs[i] = run; run += c; }
        if (threadIdx.x == 1023) *(which ? total_y : total_x) = part[1023];
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply2(u64* __restrict__ x, u64* __restrict__ y, u32 n, const u64* __restrict__ bsx,
                                                              const u64* __restrict__ bsy, const u64* __restrict__ total_x,
                                                              const u64* __restrict__ total_y, const u32* __restrict__ rec_start,
                                                              LineMapEnt* __restrict__ map) {
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    u64 vx[kScanItems], vy[kScanItems];
    u64 sx = 0, sy = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        vx[i] = base + i < n ? x[base + i] : 0; sx += vx[i];
        vy[i] = base + i < n ? y[base + i] : 0; sy += vy[i];
    }
    u64 tx, ty;
    u64 runx = bsx[blockIdx.x] + block_excl_scan_u64(sx, tx);
    u64 runy = bsy[blockIdx.x] + block_excl_scan_u64(sy, ty);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) {
            x[base + i] = runx;
            y[base + i] = runy;
            if (vy[i]) {   // lines of a k_short record: descriptors (base+i) * kSMaxLines + j
                LineMapEnt m;
                m.rec_start = rec_start[base + i];
                m.out_off = runx;
                for (u32 j = 0; j < (u32)vy[i]; ++j) { m.desc_idx = (base + i) * kSMaxLines + j; map[runy + j] = m; }
            }
        }
        runx += vx[i];
        runy += vy[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { x[n] = *total_x; y[n] = *total_y; }
}

// ------------------------------------------------------------------------------
// General per-record conversion: one thread walks one record with the streaming state
// machine of g2p_core.cuh (any record length, every error path of the reference).  It
// runs over the list of records that k_short / k_long (g2p_short.cuh, g2p_long.cuh) delegated.
// ------------------------------------------------------------------------------
constexpr int kListThreads = 64;

template <class Sink>
__device__ G2P_NOINLINE u32 convert_record_global(const u8* r, u32 len, const LenTableView& T, Sink& S, u32& ea, u32& eb) {
    return convert_record(r, len, T, S, ea, eb);
}

template <bool EMIT>
__global__ void __launch_bounds__(kListThreads) k_convert_list(const u8* __restrict__ gaf, const u32* __restrict__ rec_start, LenTableView T,
                                                               u64* __restrict__ out_off, u32* __restrict__ status, u8* __restrict__ out,
                                                               PipelineMeta* __restrict__ meta, const u32* __restrict__ list, const u32* __restrict__ n_list) {
    const u32 nd = *n_list;
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += gridDim.x * blockDim.x) {
        const u32 r = list[k];
        const u32 s = rec_start[r], len = rec_start[r + 1] - s - 1;
        u32 ea = 0, eb = 0;
        if (!EMIT) {
            CountSink cs;
            const u32 st = convert_record_global(gaf + s, len, T, cs, ea, eb);
            out_off[r] = (st_is_abort(st) || (st & 0xff) == ST_SKIP) ? 0 : cs.n;
            status[r] = st;
            if (st_is_error(st)) atomicMin(&meta->first_err, r);
        } else {
            const u32 st = status[r];
            if (st_is_abort(st) || (st & 0xff) == ST_SKIP) continue;
            StoreSink ss(out + out_off[r]);
            convert_record_global(gaf + s, len, T, ss, ea, eb);
        }
    }
}

// ------------------------------------------------------------------------------
// gaf2unstable: one thread rewrites one record with
// unstable_record (g2u_core.cuh): per step a hash probe for the stable contig and two
// binary searches in its node array, then the record is re-serialised with its tags in
// name order.  Same two passes as gaf2paf (size -> scan -> emit).  Records whose path
// spans several reference contigs are listed for the host, which prints the reference's
// warning for them.
// ------------------------------------------------------------------------------
// unstable_record's phases (g2u_core.cuh) for the 32 records of a warp at once: a warp barrier after every phase of the
// parse, the two data-dependent loops (path steps, optional fields) as warp-uniform loops.  Every lane of the warp must
// call it; `act` says whether the lane has a record.
template <class Sink>
__device__ __forceinline__ u32 unstable_record_warp(bool act, const u8* r, u32 len, const UnstableView& V, Sink& S) {
    const u32 FULL = 0xffffffffu;
    URecHdr h;
    UParse c;
    URun u;
    u32 st = act ? (u32)ST_OK : (u32)ST_SKIP;
    if (st == ST_OK) st = u_hdr_a(r, len, h, c);
    __syncwarp();
    if (st == ST_OK) st = u_hdr_b(r, len, h, c);
    __syncwarp();
    if (st == ST_OK) st = u_hdr_c(r, len, h, c);
    __syncwarp();
    if (st == ST_OK) st = u_hdr_d(r, len, h, c);
    __syncwarp();
    const bool go = st == ST_OK;
    u.steps_left = u.tags_left = false;
    if (go) u_out_head(r, h, u, S);
    while (__any_sync(FULL, go && u.steps_left)) {
        if (go && u.steps_left) u_out_step(r, V, h, u, S);
    }
    if (go) u_out_mid(h, u, S);
    while (__any_sync(FULL, go && u.tags_left)) {
        if (go && u.tags_left) u_out_tag(r, len, V, h, u, S);
    }
    if (go) st = u_out_end(u, S);
    return st;
}

template <bool EMIT>
__global__ void __launch_bounds__(kListThreads) k_unstable(const u8* __restrict__ gaf, const u32* __restrict__ rec_start, u32 nrec, UnstableView V,
                                                           u64* __restrict__ out_off, u32* __restrict__ status, u8* __restrict__ out,
                                                           PipelineMeta* __restrict__ meta, u32* __restrict__ warn_list) {
    for (u32 base = blockIdx.x * blockDim.x; base < nrec; base += gridDim.x * blockDim.x) {   // (uniform: every lane of a warp takes part)
        const u32 r = base + threadIdx.x;
        const bool in = r < nrec;
        const u32 s = in ? rec_start[r] : 0u, len = in ? rec_start[r + 1] - s - 1 : 0u;
        if (!EMIT) {
            CountSink cs;
            const u32 st = unstable_record_warp(in, gaf + s, len, V, cs);
            if (in) {
                out_off[r] = (st_is_abort(st) || (st & 0xff) == ST_SKIP) ? 0 : cs.n;
                status[r] = st;
                if (st_is_error(st)) atomicMin(&meta->first_err, r);
                else if ((st & 0xff) == ST_WARN_MULTIREF) warn_list[atomicAdd(&meta->n_deleg, 1u)] = r;
            }
        } else {
            const u32 st = in ? status[r] : (u32)ST_SKIP;
            const bool act = in && !(st_is_abort(st) || (st & 0xff) == ST_SKIP);
            StoreSink ss(out + (act ? out_off[r] : 0));
            unstable_record_warp(act, gaf + s, len, V, ss);
        }
    }
}

// The same with the CTA's records staged: 128 consecutive records are one contiguous piece of the input and produce one
// contiguous piece of the output.  Both go through shared memory (coalesced 128-bit loads / stores), so the per-record
// code -- a byte-at-a-time state machine that scans a record several times (columns, path, tags in name order) -- runs
// at shared-memory latency instead of one L2 round trip per byte.  A CTA whose records (or output) exceed the buffers
// reads (writes) global memory directly.
constexpr u32 kUThreads = 128;
#ifndef G2U_IN_CAP
#define G2U_IN_CAP (24u << 10)
#endif
#ifndef G2U_OUT_CAP
#define G2U_OUT_CAP (40u << 10)
#endif
constexpr u32 kUInCap = G2U_IN_CAP, kUOutCap = G2U_OUT_CAP;   // (tests build a variant with tiny buffers to reach the direct paths)
template <bool EMIT> constexpr size_t unstable_smem() { return kUInCap + 32 + (EMIT ? kUOutCap + 32 : 0); }

template <bool EMIT>
__global__ void __launch_bounds__(kUThreads) k_unstable_staged(const u8* __restrict__ gaf, u64 n, const u32* __restrict__ rec_start, u32 nrec, UnstableView V,
                                                               u64* __restrict__ out_off, u32* __restrict__ status, u8* __restrict__ out,
                                                               PipelineMeta* __restrict__ meta, u32* __restrict__ warn_list) {
    G2P_DYN_SMEM(smem);
    u8* sm_in = smem;
    u8* sm_out = smem + kUInCap + 32;
    const u32 r0 = blockIdx.x * kUThreads, r1 = min(r0 + kUThreads, nrec);
    const u32 s0 = rec_start[r0], s1 = rec_start[r1];
    const u32 e1 = (u64)s1 < n ? s1 : (u32)n;   // (an unterminated last line ends at n)
    const u32 A = s0 & ~15u;
    const bool in_staged = e1 - A <= kUInCap;
    if (in_staged) {
        const u32 nvec = (e1 - A + 15u) >> 4;
        for (u32 v = threadIdx.x; v < nvec; v += kUThreads) reinterpret_cast<uint4*>(sm_in)[v] = ldg_vec_guarded(gaf, (u64)A + 16u * v, n);
    }
    u64 o0 = 0, o1 = 0;
    u32 opad = 0;
    bool out_staged = false;
    if (EMIT) {
        o0 = out_off[r0]; o1 = out_off[r1];
        opad = (u32)(o0 & 15u);
        out_staged = (o1 - o0) + opad <= kUOutCap;
    }
    __syncthreads();
    const u32 r = r0 + threadIdx.x;
    const bool in = r < r1;
    {
        const u32 s = in ? rec_start[r] : s0, len = in ? rec_start[r + 1] - s - 1 : 0u;
        const u8* text = in_staged ? sm_in + (s - A) : gaf + s;
        if (!EMIT) {
            CountSink cs;
            const u32 st = unstable_record_warp(in, text, len, V, cs);
            if (in) {
                out_off[r] = (st_is_abort(st) || (st & 0xff) == ST_SKIP) ? 0 : cs.n;
                status[r] = st;
                if (st_is_error(st)) atomicMin(&meta->first_err, r);
                else if ((st & 0xff) == ST_WARN_MULTIREF) warn_list[atomicAdd(&meta->n_deleg, 1u)] = r;
            }
        } else {
            const u32 st = in ? status[r] : (u32)ST_SKIP;
            const bool act = in && !(st_is_abort(st) || (st & 0xff) == ST_SKIP);
            const u64 o = act ? out_off[r] : o0;
            StoreSink ss(out_staged ? sm_out + opad + (u32)(o - o0) : out + o);
            unstable_record_warp(act, text, len, V, ss);
        }
    }
    if (EMIT && out_staged) {
        __syncthreads();
        const u32 total = opad + (u32)(o1 - o0);
        u8* gb = out + (o0 - opad);
        const u32 full_b = total >> 4;
        for (u32 u = (opad ? 1u : 0u) + threadIdx.x; u < full_b; u += kUThreads) reinterpret_cast<uint4*>(gb)[u] = reinterpret_cast<const uint4*>(sm_out)[u];
        const u32 head_end = opad ? (total < 16u ? total : 16u) : 0u;
        for (u32 b = opad + threadIdx.x; b < head_end; b += kUThreads) gb[b] = sm_out[b];
        const u32 tail_a = full_b * 16u > head_end ? full_b * 16u : head_end;
        for (u32 b = tail_a + threadIdx.x; b < total; b += kUThreads) gb[b] = sm_out[b];
    }
}

// Details of the first failing record (one thread; runs only on error).
__global__ void k_diagnose(const u8* __restrict__ gaf, const u32* __restrict__ rec_start, LenTableView T,
                           const u64* __restrict__ out_off, PipelineMeta* __restrict__ meta) {
    const u32 r = meta->first_err;
    if (r == 0xFFFFFFFFu) return;
    const u32 s = rec_start[r], e = rec_start[r + 1];
    CountSink cs;
    u32 ea = 0, eb = 0;
    u32 st = convert_record_global(gaf + s, e - s - 1, T, cs, ea, eb);
    meta->err_status = st;
    meta->err_a = ea;
    meta->err_b = eb;
    meta->err_rec_start = s;
    meta->err_out_end = out_off[r + 1];
}

}  // namespace g2p

#include "g2p_fuse.cuh"
#include "g2p_par.cuh"
#include "g2p_filter.cuh"
