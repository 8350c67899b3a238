// g2p_par.cuh — token-parallel conversion of the records k_rec does not take (longer than 240 bytes, up to
// chromosome-scale assembly records): SURVEY.md §8a rows a4-a9 in the closed form of Appendix B.
//
// k_long gives one warp to one record and walks it from end to end: a 250 kB assembly record is a 12 ms chain of
// dependent loads, and a block that holds a few hundred of them leaves most of the GPU idle.  Here nothing is serial
// per record.  All such records of a block are processed together and every kernel is a flat data-parallel pass:
//
//   k_par_plan     thread / record      2 KiB tiles per record (aligned to 16 bytes of the text)         + u64 scan
//   k_par_tilemap  thread / record      tile -> (record, text position)
//   k_par_tabs     thread / 16 bytes    tab positions -> the record's (small) tab list
//   k_par_head     thread / record      columns 1-12, tags, path and cg spans, RecDesc                   (parse_gaf_record)
//   k_par_count    thread / 16 bytes    '>' '<' markers inside the path column, op letters inside the cg value
//   k_par_ranges   thread / record      its steps / ops in the flat arrays, its run of line descriptors  + u64 scans
//   k_par_fill     thread / 16 bytes    ... the positions of the markers and letters, in text order
//   k_par_steps    thread / path step   name, ":start-end", ONE table probe                              (gaf2paf_main.cpp:157-170)
//   k_par_ops      thread / CIGAR op    length and class -> (target, query, match, block) contributions   (for_each_cg)
//   uint4 scans                         exclusive prefix sums of both (mod 2^32: differences inside a record are exact,
//                                       records whose sums do not fit 31 bits are recognised by 64-bit totals and left alone)
//   k_par_totals   thread / record      summed step lengths, mirrored path interval of '-' records       (flip_gaf)
//   k_par_lines    thread / path step   boundary B_i -> lower_bound in the cumulative target length of the ops
//                                       (cigar_next_by_target as a binary search), sums by prefix differences, the
//                                       numbers of the step's PAF line, its length -> LineDesc         (gaf2paf :172-263)
//   u64 scan, k_par_finish, k_par_place line offsets inside the record, record sizes, status
//
// '-' records are index arithmetic on the same arrays (normalised step i = step ns-1-i, op j = op no-1-j).  The
// output is what k_long's size pass produces: out_off / status / RecDesc per record and one LineDesc per PAF line in
// the dense descriptor array printed by k_emit_lines<DENSE>; a record owns a run of slots, one per path step in
// output order, padded to a multiple of 32.  Only canonical records are converted (same definition as k_rec /
// k_long); the rest goes on to k_long and from there to the general kernel.
#pragma once
#include "g2p_long.cuh"

namespace g2p {

constexpr u32 kPTile = 2048, kPThreads = 128;   // 16 bytes per thread
constexpr u32 kPMaxTabs = 40;

struct __align__(16) ParRec {
    u32 r, s, len, ntabs;     // record ordinal, text start, length without '\n'
    u32 status;               // 0 ok, 1 not canonical (-> k_long)
    u32 pa, pb, ca, cb;       // path column and cg value, absolute text positions
    u32 s0, ns, g0, no;       // its steps / ops in the flat arrays
    i32 qs, ps, pe;           // ps / pe mirrored for '-' records by k_par_totals
    u32 total, rconst, minus;
    u32 slot0, nslots;        // descriptor run
    u32 pad0;
    unsigned long long sum_steps, sum_ops;   // 64-bit totals (overflow guards of the 32-bit prefix sums)
    u32 tabs[kPMaxTabs];
};

struct ParArgs {
    const u8* gaf;
    u64 n;
    const u32* rec_start;
    LenTableView T;
    const u32* list;        // records to convert (delegated by k_rec)
    u32 nlist;
    ParRec* recs;           // [nlist]
    u64* tile_base;         // [nlist + 1] exclusive tile offsets
    u64* tile_off;          // [ntiles + 1] exclusive (steps | ops << 32) offsets
    uint2* tile_map;        // [ntiles] record (list index), first text position
    u64* slot_scan;         // [nlist + 1] descriptor run starts (32-padded runs | small runs << 32)
    u32 ntiles;
    u32* spos; u32* srec;   // per step: marker position, record (list index)
    u32* opos; u32* orec;   // per op: letter position, record
    uint4* sval;            // per step: tlen, sa, se, nl | '<' << 16                        (k_par_steps)
    uint4* sx;              // per step (+1): {slen, 0, 0, 0}, scanned
    uint4* ox;              // per op (+1): {target, query, match, block}, scanned
    u64* lx;                // per step (+1), normalised order: line_len | emit << 32, scanned
    u32 nsteps, nops;
    // outputs shared with k_long
    u64* out_off; u32* status; RecDesc* rdesc; LineDesc* desc;
    u32* reject_list; u32* n_reject;
    u32* n_desc;            // k_long's blocks follow the runs of these records
    u32* n_desc2;
    u32 half, small_max;    // first slot of the upper half; records of at most small_max steps take their slots there
};

// ---- uint4 exclusive scan, four independent 32-bit lanes (mod 2^32); x[n] = totals -------------------
__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint4 sub4(uint4 a, uint4 b) { return make_uint4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ uint4 shfl_up4(uint4 v, int d) {
    return make_uint4(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d), __shfl_up_sync(0xffffffffu, v.z, d), __shfl_up_sync(0xffffffffu, v.w, d));
}
__device__ __forceinline__ uint4 block_excl_scan4(uint4 v, uint4& total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 incl = v;
    for (int o = 1; o < 32; o <<= 1) { const uint4 up = shfl_up4(incl, o); if (lane >= (u32)o) incl = add4(incl, up); }
    __shared__ uint4 wt[kScanThreads / 32];
    __syncthreads();
    if (lane == 31) wt[warp] = incl;
    __syncthreads();
    uint4 pre = make_uint4(0, 0, 0, 0), tot = pre;
    for (int i = 0; i < kScanThreads / 32; ++i) { if (i < (int)warp) pre = add4(pre, wt[i]); tot = add4(tot, wt[i]); }
    total = tot;
    return sub4(add4(pre, incl), v);
}
__global__ void __launch_bounds__(kScanThreads) k_scan4_reduce(const uint4* __restrict__ x, u32 n, uint4* __restrict__ block_sum) {
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint4 s = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < n) s = add4(s, x[base + i]);
    uint4 total;
    block_excl_scan4(s, total);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_scan4_blocks(uint4* __restrict__ block_sum, u32 nblocks, uint4* __restrict__ total_out) {
    __shared__ uint4 part[1024];
    const u32 per = (nblocks + 1023) / 1024;
    const u32 a = threadIdx.x * per, b = min(a + per, nblocks);
    uint4 s = make_uint4(0, 0, 0, 0);
    for (u32 i = a; i < b; ++i) s = add4(s, block_sum[i]);
    part[threadIdx.x] = s;
    __syncthreads();
    for (u32 o = 1; o < 1024; o <<= 1) {
        const uint4 v = threadIdx.x >= o ? part[threadIdx.x - o] : make_uint4(0, 0, 0, 0);
        __syncthreads();
        part[threadIdx.x] = add4(part[threadIdx.x], v);
        __syncthreads();
    }
    uint4 run = threadIdx.x ? part[threadIdx.x - 1] : make_uint4(0, 0, 0, 0);
    for (u32 i = a; i < b; ++i) { const uint4 c = block_sum[i]; block_sum[i] = run; run = add4(run, c); }
    if (threadIdx.x == 1023) *total_out = part[1023];
}
__global__ void __launch_bounds__(kScanThreads) k_scan4_apply(uint4* __restrict__ x, u32 n, const uint4* __restrict__ block_off) {
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint4 v[kScanItems];
    uint4 s = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { v[i] = base + i < n ? x[base + i] : make_uint4(0, 0, 0, 0); s = add4(s, v[i]); }
    uint4 total;
    uint4 run = add4(block_off[blockIdx.x], block_excl_scan4(s, total));
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) x[base + i] = run;
        run = add4(run, v[i]);
    }
}

// ---- plan: tiles per record ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_par_plan(const ParArgs a) {
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < a.nlist; k += gridDim.x * blockDim.x) {
        const u32 r = a.list[k];
        const u32 s = a.rec_start[r], len = a.rec_start[r + 1] - s - 1;
        ParRec& R = a.recs[k];
        R.r = r; R.s = s; R.len = len; R.ntabs = 0; R.status = 0; R.sum_steps = 0; R.sum_ops = 0; R.nslots = 0; R.slot0 = 0;
        a.tile_base[k] = ((u64)len + (s & 15u) + kPTile - 1) / kPTile;
    }
}
// tile -> record (list index) and absolute text position of its first byte (16-byte aligned)
__global__ void __launch_bounds__(256) k_par_tilemap(const ParArgs a) {
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < a.nlist; k += gridDim.x * blockDim.x) {
        const u32 t0 = (u32)a.tile_base[k], t1 = (u32)a.tile_base[k + 1], p0 = a.recs[k].s & ~15u;
        for (u32 t = t0; t < t1; ++t) a.tile_map[t] = make_uint2(k, p0 + (t - t0) * kPTile);
    }
}
__device__ __forceinline__ u32 par_tile_record(const ParArgs& a, u32 tile, u32& pos0) {
    const uint2 m = a.tile_map[tile];
    pos0 = m.y;
    return m.x;
}
// the calling thread's 16 aligned bytes and the bits of [lo, hi) (absolute positions) among them
struct ParChunk { u32 w0, w1, w2, w3, pos; };
__device__ __forceinline__ ParChunk par_chunk(const ParArgs& a, u32 tile_pos) {
    ParChunk c;
    c.pos = tile_pos + threadIdx.x * 16u;
    const uint4 v = (u64)c.pos < a.n ? ldg_vec_guarded(a.gaf, (u64)c.pos, a.n) : make_uint4(0, 0, 0, 0);
    c.w0 = v.x; c.w1 = v.y; c.w2 = v.z; c.w3 = v.w;
    return c;
}
__device__ __forceinline__ u32 par_range(const ParChunk& c, u32 lo, u32 hi) {
    const long long l = (long long)lo - (long long)c.pos, h = (long long)hi - (long long)c.pos;
    return range16((int)(l < -1 ? -1 : (l > 17 ? 17 : l)), (int)(h < -1 ? -1 : (h > 17 ? 17 : h)));
}
__device__ __forceinline__ u32 par_eq_mask(const ParChunk& c, u32 splat) {
    return movemask4(zero_bytes(c.w0 ^ splat)) | (movemask4(zero_bytes(c.w1 ^ splat)) << 4) | (movemask4(zero_bytes(c.w2 ^ splat)) << 8) |
           (movemask4(zero_bytes(c.w3 ^ splat)) << 12);
}
__device__ __forceinline__ u32 par_nondigit_mask(const ParChunk& c) {
    return movemask4(nondigit_bytes(c.w0)) | (movemask4(nondigit_bytes(c.w1)) << 4) | (movemask4(nondigit_bytes(c.w2)) << 8) |
           (movemask4(nondigit_bytes(c.w3)) << 12);
}

__global__ void __launch_bounds__(kPThreads) k_par_tabs(const ParArgs a) {
    __shared__ u32 s_k, s_pos;
    if (threadIdx.x == 0) { u32 p; s_k = par_tile_record(a, blockIdx.x, p); s_pos = p; }
    __syncthreads();
    ParRec& R = a.recs[s_k];
    const ParChunk c = par_chunk(a, s_pos);
    u32 m = par_eq_mask(c, 0x09090909u) & par_range(c, R.s, R.s + R.len);
    while (m) {
        const u32 b = (u32)__ffs((int)m) - 1u;
        m &= m - 1u;
        const u32 idx = atomicAdd(&R.ntabs, 1u);
        if (idx < kPMaxTabs) R.tabs[idx] = c.pos + b - R.s;
    }
}

// plain decimal (<= 9 digits) or '*' (-> -1); false if not canonical
__device__ __forceinline__ bool par_num(const u8* t, u32 a0, u32 b0, i32& v) {
    const u32 n = b0 - a0;
    if (n == 1 && t[a0] == '*') { v = -1; return true; }
    if (n == 0 || n > 9) return false;
    u32 x = 0;
    for (u32 k = a0; k < b0; ++k) { const u32 d = (u32)t[k] - '0'; if (d > 9) return false; x = x * 10u + d; }
    v = (i32)x;
    return true;
}

// columns 1-12 and tags of one record from its tab list (parse_gaf_record, gafkluge.hpp:84-204); RecDesc for k_emit_lines
__global__ void __launch_bounds__(128) k_par_head(const ParArgs a) {
    __shared__ u32 p10[10];
    if (threadIdx.x < 10) { u32 v = 1; for (u32 i = 0; i < threadIdx.x; ++i) v *= 10u; p10[threadIdx.x] = v; }
    __syncthreads();
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < a.nlist; k += gridDim.x * blockDim.x) {
        ParRec& R = a.recs[k];
        const u8* t = a.gaf + R.s;
        bool ok = false;
        do {
            const u32 nt = R.ntabs;
            if (nt < 12 || nt > kPMaxTabs || R.len == 0) break;   // 12 columns and at least one tag (cg)
            for (u32 i = 1; i < nt; ++i) {   // the atomics arrive in any order
                const u32 v = R.tabs[i];
                u32 j = i;
                while (j > 0 && R.tabs[j - 1] > v) { R.tabs[j] = R.tabs[j - 1]; --j; }
                R.tabs[j] = v;
            }
            auto fa = [&](u32 f) { return f ? R.tabs[f - 1] + 1u : 0u; };
            auto fb = [&](u32 f) { return f < nt ? R.tabs[f] : R.len; };
            if (t[0] == '*') break;   // (skipped lines, gaf2paf_main.cpp:360, are short; a long one is left to k_long)
            LineRec L;
            L.qn_b = fb(0);
            if (L.qn_b == 0) break;
            i32 qs, qe, plen, ps, pe, mapq;
            if (!par_num(t, fa(1), fb(1), L.qlen) || !par_num(t, fa(2), fb(2), qs) || !par_num(t, fa(3), fb(3), qe)) break;
            if (fb(4) - fa(4) != 1 || (t[fa(4)] != '+' && t[fa(4)] != '-')) break;
            const bool minus = t[fa(4)] == '-';
            const u32 pa = fa(5), pb = fb(5);
            if (pb <= pa || (t[pa] != '>' && t[pa] != '<')) break;   // bare stable names / empty paths: left to k_long
            if (!par_num(t, fa(6), fb(6), plen) || !par_num(t, fa(7), fb(7), ps) || !par_num(t, fa(8), fb(8), pe) ||
                !par_num(t, fa(9), fb(9), L.m) || !par_num(t, fa(10), fb(10), L.b) || !par_num(t, fa(11), fb(11), mapq)) break;
            (void)qe; (void)plen;
            L.mapq = mapq >= 255 ? -1 : mapq;   // gafkluge.hpp:176-183
            u32 ca = 0, cb = 0;
            L.tp_a = L.tp_b = L.rc_a = L.rc_b = 0;
            bool bad = false;
            for (u32 f = 12; f <= nt && !bad; ++f) {
                const u32 x = fa(f), y = fb(f);
                if (y == x) continue;   // empty field: ignored by the tag loop
                if (y - x < 5 || t[x + 2] != ':' || t[x + 4] != ':') { bad = true; break; }
                for (u32 g = 12; g < f; ++g) { const u32 x2 = fa(g); if (fb(g) != x2 && t[x2] == t[x] && t[x2 + 1] == t[x + 1]) bad = true; }   // duplicate tag
                if (t[x] == 'c' && t[x + 1] == 'g') { ca = x + 5; cb = y; }
                else if (t[x] == 't' && t[x + 1] == 'p') { L.tp_a = x + 3; L.tp_b = y; }
                else if (t[x] == 'r' && t[x + 1] == 'c') { L.rc_a = x + 3; L.rc_b = y; }
            }
            if (bad || cb == 0 || ca >= cb || qs < 0 || ps < 0 || pe < 0) break;
            L.gi_n = gi_fast(L.m, L.b, L.gi);
            if (L.gi_n == 0 || !rec_desc_fits(L)) break;
            R.pa = R.s + pa; R.pb = R.s + pb; R.ca = R.s + ca; R.cb = R.s + cb;
            R.qs = qs; R.ps = ps; R.pe = pe; R.minus = minus ? 1u : 0u;
            R.rconst = line_const_len(L, p10);
            store_rec_desc(a.rdesc + R.r, L);
            ok = true;
        } while (0);
        if (!ok) { R.status = 1; R.pa = R.pb = R.ca = R.cb = 0; }
    }
}

// marker bits inside the path column, op-letter bits inside the cg value
__device__ __forceinline__ void par_masks(const ParRec& R, const ParChunk& c, u32& mm, u32& om) {
    mm = om = 0;
    if (R.status) return;
    const u32 pr = par_range(c, R.pa, R.pb), cr = par_range(c, R.ca, R.cb);
    if (pr) mm = (par_eq_mask(c, 0x3E3E3E3Eu) | par_eq_mask(c, 0x3C3C3C3Cu)) & pr;
    if (cr) om = par_nondigit_mask(c) & cr;
}
__global__ void __launch_bounds__(kPThreads) k_par_count(const ParArgs a) {
    __shared__ u32 s_k, s_pos;
    __shared__ u64 ws[kPThreads / 32];
    if (threadIdx.x == 0) { u32 p; s_k = par_tile_record(a, blockIdx.x, p); s_pos = p; }
    __syncthreads();
    u32 mm, om;
    par_masks(a.recs[s_k], par_chunk(a, s_pos), mm, om);
    u64 c = (u64)__popc(mm) | ((u64)__popc(om) << 32);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { u64 s = 0; for (u32 i = 0; i < kPThreads / 32; ++i) s += ws[i]; a.tile_off[blockIdx.x] = s; }
}
// the record's ranges in the flat arrays; its descriptor run (one slot per step, padded to 32)
__global__ void __launch_bounds__(256) k_par_ranges(const ParArgs a) {
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < a.nlist; k += gridDim.x * blockDim.x) {
        ParRec& R = a.recs[k];
        const u64 b0 = a.tile_off[a.tile_base[k]], b1 = a.tile_off[a.tile_base[k + 1]], d = b1 - b0;
        R.s0 = (u32)b0; R.ns = (u32)d;
        R.g0 = (u32)(b0 >> 32); R.no = (u32)(d >> 32);
        if (!R.status && (R.ns == 0 || R.no == 0)) R.status = 1;
        // output bytes of a record must fit 32 bits: every line is rconst + at most 120 bytes + a piece of the CIGAR text
        if (!R.status && (u64)R.ns * (R.rconst + 120u) + R.len > 0xf0000000ULL) R.status = 1;
        // records of a few steps share 32-slot blocks (upper half of the array, like k_long's small batches)
        const bool small = R.ns <= a.small_max;
        R.nslots = R.status ? 0u : (small ? (R.ns + 3u) & ~3u : (R.ns + 31u) & ~31u);
        a.slot_scan[k] = small ? (u64)R.nslots << 32 : (u64)R.nslots;
    }
}
__global__ void __launch_bounds__(kPThreads) k_par_fill(const ParArgs a) {
    __shared__ u32 s_k, s_pos;
    __shared__ u64 ws[kPThreads / 32];
    if (threadIdx.x == 0) { u32 p; s_k = par_tile_record(a, blockIdx.x, p); s_pos = p; }
    __syncthreads();
    const u32 k = s_k;
    const ParChunk ch = par_chunk(a, s_pos);
    u32 mm, om;
    par_masks(a.recs[k], ch, mm, om);
    const u64 c = (u64)__popc(mm) | ((u64)__popc(om) << 32);
    u64 incl = c;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const u64 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += up; }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    u64 pre = 0;
    for (u32 i = 0; i < warp; ++i) pre += ws[i];
    const u64 base = a.tile_off[blockIdx.x] + pre + incl - c;
    u32 si = (u32)base, oi = (u32)(base >> 32);
    while (mm) { const u32 b = (u32)__ffs((int)mm) - 1u; mm &= mm - 1u; a.spos[si] = ch.pos + b; a.srec[si] = k; ++si; }
    while (om) { const u32 b = (u32)__ffs((int)om) - 1u; om &= om - 1u; a.opos[oi] = ch.pos + b; a.orec[oi] = k; ++oi; }
}
// (slot_scan scanned) -> the records' descriptor runs
__global__ void __launch_bounds__(256) k_par_slots(const ParArgs a) {
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < a.nlist; k += gridDim.x * blockDim.x) {
        ParRec& R = a.recs[k];
        R.slot0 = R.ns <= a.small_max ? a.half + (u32)(a.slot_scan[k] >> 32) : (u32)a.slot_scan[k];
    }
}

// adds v to *dst once per warp when all its lanes name the same destination (the common case), else per lane
__device__ __forceinline__ void par_sum64(unsigned long long* dst, u32 key, u32 v, bool active) {
    const u32 FULL = 0xffffffffu;
    const u32 k0 = __shfl_sync(FULL, key, 0);
    const bool lane0 = __shfl_sync(FULL, (u32)active, 0) != 0;
    const bool same = __all_sync(FULL, !active || key == k0) && lane0;
    if (same) {
        unsigned long long s = active ? v : 0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(FULL, s, o);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(dst, s);
    } else if (active && v) atomicAdd(dst, (unsigned long long)v);
}

// one path step: "[><]name[:start-end]", one table probe
__global__ void __launch_bounds__(128) k_par_steps(const ParArgs a) {
    const u32 nround = (a.nsteps + 31u) & ~31u;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < nround; s += gridDim.x * blockDim.x) {
        const bool in = s < a.nsteps;
        const u32 k = in ? a.srec[s] : 0u;
        ParRec& R = a.recs[k];
        uint4 v = make_uint4(0, 0, 0, 0);
        u32 slen = 0;
        bool ok = false;
        const bool live = in && !R.status;
        if (live) do {
            const u8* g = a.gaf;
            const u32 mp = a.spos[s];
            const u32 end = s + 1 < R.s0 + R.ns ? a.spos[s + 1] : R.pb;
            u32 e = mp + 1;
            while (e < end && g[e] != ':') ++e;
            const u32 nl = e - (mp + 1);
            if (nl == 0 || nl > 255) break;
            i64 tl64;
            if (!table_lookup(a.T, g + mp + 1, nl, tl64) || tl64 < 0 || tl64 > 0x7fffffffLL) break;
            u32 sa = 0, se = (u32)tl64;
            if (e < end) {   // ":start-end" (gafkluge.hpp:131-146), plain digits only
                u32 q = e + 1, x = 0;
                const u32 q1 = q;
                while (q < end && (u32)g[q] - '0' <= 9u && q - q1 < 10) { x = x * 10u + ((u32)g[q] - '0'); ++q; }
                if (q == q1 || q - q1 > 9 || q >= end || g[q] != '-') break;
                sa = x;
                const u32 q2 = ++q;
                x = 0;
                while (q < end && (u32)g[q] - '0' <= 9u && q - q2 < 10) { x = x * 10u + ((u32)g[q] - '0'); ++q; }
                if (q == q2 || q - q2 > 9 || q != end || x < sa) break;
                se = x;
            }
            slen = se - sa;
            v = make_uint4((u32)tl64, sa, se, nl | ((u32)(g[mp] == '<') << 16));
            ok = true;
        } while (0);
        if (live && !ok) atomicExch(&R.status, 1u);
        par_sum64(&R.sum_steps, k, slen, live && ok);
        if (in) { a.sval[s] = v; a.sx[s] = make_uint4(ok ? slen : 0u, 0, 0, 0); }
    }
}
// one CIGAR op: digits between the previous letter and this one (for_each_cg, gafkluge.hpp:226-239)
__global__ void __launch_bounds__(128) k_par_ops(const ParArgs a) {
    const u32 nround = (a.nops + 31u) & ~31u;
    for (u32 o = blockIdx.x * blockDim.x + threadIdx.x; o < nround; o += gridDim.x * blockDim.x) {
        const bool in = o < a.nops;
        const u32 k = in ? a.orec[o] : 0u;
        ParRec& R = a.recs[k];
        uint4 v = make_uint4(0, 0, 0, 0);
        const bool live = in && !R.status;
        bool ok = false;
        if (live) {
            const u8* g = a.gaf;
            const u32 lp = a.opos[o];
            const u32 ds = o > R.g0 ? a.opos[o - 1] + 1u : R.ca;
            const u32 nd = lp - ds;
            const u32 kc = (u32)g[lp] - '=';
            ok = kc < 28u && ((kOpMask >> kc) & 1u) && nd != 0 && nd <= 7 && !(nd > 1 && g[ds] == '0');
            u32 x = 0;
            if (ok) { for (u32 q = ds; q < lp; ++q) x = x * 10u + ((u32)g[q] - '0'); ok = x != 0; }
            if (ok && o + 1 == R.g0 + R.no && lp != R.cb - 1) ok = false;   // digits after the last op letter
            if (!ok) atomicExch(&R.status, 1u);
            else v = make_uint4(((kTargetMask >> kc) & 1u) ? x : 0u, ((kQueryMask >> kc) & 1u) ? x : 0u, ((kMatchMask >> kc) & 1u) ? x : 0u, x);
        }
        par_sum64(&R.sum_ops, k, v.w, live && ok);
        if (in) a.ox[o] = v;
    }
}
// summed step lengths; flip_gaf's mirrored path interval (gaf2paf_main.cpp:111-131)
__global__ void __launch_bounds__(256) k_par_totals(const ParArgs a) {
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < a.nlist; k += gridDim.x * blockDim.x) {
        ParRec& R = a.recs[k];
        if (R.status) continue;
        // sums that do not fit 31 bits: the 32-bit prefix differences would be wrong; k_long sums in 64 bits
        if (R.sum_steps > 0x7fffffffULL || R.sum_ops > 0x7fffffffULL) { R.status = 1; continue; }
        const u32 total = (u32)R.sum_steps;
        R.total = total;
        if (R.minus) { const i32 nps = (i32)total - R.pe, npe = (i32)total - R.ps; R.ps = nps; R.pe = npe; }
        if (R.ps < 0 || R.pe < R.ps) R.status = 1;
    }
}

// normalised views of the scanned arrays of one record
struct ParView {
    const uint4* sx; const uint4* ox; const u32* opos;
    u32 s0, ns, g0, no, total;
    uint4 obase, ototal;
    bool minus;
    // cumulative step length before normalised step i (0 <= i <= ns)
    __device__ __forceinline__ u32 cum_steps(u32 i) const { return minus ? total - (sx[s0 + ns - i].x - sx[s0].x) : sx[s0 + i].x - sx[s0].x; }
    // inclusive prefix sums over the ops in normalised order, op j (0 <= j < no)
    __device__ __forceinline__ uint4 incl(u32 j) const {
        if (!minus) return sub4(ox[g0 + j + 1], obase);
        return sub4(ototal, sub4(ox[g0 + no - 1 - j], obase));   // total - exclusive prefix of the original op no-1-j
    }
    __device__ __forceinline__ u32 orig(u32 j) const { return minus ? no - 1 - j : j; }
};
struct ParBoundary { u32 j, t, cq, cm, cb; bool cut, exh; };
// boundary B of the cumulative target length -> position in the op stream (as k_short's phase 5)
__device__ __forceinline__ ParBoundary par_boundary(const ParView& V, const u8* gaf, u32 B) {
    ParBoundary b;
    b.j = b.t = b.cq = b.cm = b.cb = 0; b.cut = b.exh = false;
    if (B == 0) return b;
    u32 lo = 0, hi = V.no;
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (V.incl(mid).x >= B) hi = mid; else lo = mid + 1; }
    b.j = lo;
    if (lo == V.no) { b.exh = true; return b; }
    const uint4 E = V.incl(lo);
    const uint4 P = lo ? V.incl(lo - 1) : make_uint4(0, 0, 0, 0);
    b.t = P.x;
    const u32 off = B - P.x;
    const u32 kc = (u32)gaf[V.opos[V.g0 + V.orig(lo)]] - '=';
    b.cq = P.y + (((kQueryMask >> kc) & 1u) ? off : 0u);
    b.cm = P.z + (((kMatchMask >> kc) & 1u) ? off : 0u);
    b.cb = P.w + off;
    b.cut = E.x > B;
    return b;
}

// one PAF line per path step (gaf2paf_main.cpp:157-263 in the closed form of SURVEY.md Appendix B)
__global__ void __launch_bounds__(128) k_par_lines(const ParArgs a) {
    __shared__ u32 p10[10];
    if (threadIdx.x < 10) { u32 v = 1; for (u32 i = 0; i < threadIdx.x; ++i) v *= 10u; p10[threadIdx.x] = v; }
    __syncthreads();
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < a.nsteps; s += gridDim.x * blockDim.x) {
        ParRec& R = a.recs[a.srec[s]];
        if (R.nslots == 0) { a.lx[s] = 0; continue; }   // rejected before it got a descriptor run (every step of it writes one zero)
        const bool rminus = R.minus != 0;
        // this thread owns ORIGINAL step s - s0; its normalised index:
        const u32 jo = s - R.s0, i = rminus ? R.ns - 1 - jo : jo;
        u32 line = 0, emit = 0;
        LineDesc* slot = a.desc + R.slot0 + i;
        if (!R.status) {
            ParView V;
            V.sx = a.sx; V.ox = a.ox; V.opos = a.opos;
            V.s0 = R.s0; V.ns = R.ns; V.g0 = R.g0; V.no = R.no; V.total = R.total; V.minus = rminus;
            V.obase = a.ox[R.g0];
            V.ototal = sub4(a.ox[R.g0 + R.no], V.obase);
            const uint4 sv = a.sval[s];
            const i32 tlen = (i32)sv.x, sa = (i32)sv.y, se = (i32)sv.z;
            const u32 nl = sv.w & 0xffffu;
            const bool rev = ((sv.w >> 16) & 1u) != (u32)rminus;
            const i32 slen = se - sa;
            const i32 W = R.pe - R.ps;
            const i32 so = i == 0 ? R.ps : 0;
            const bool last = i + 1 == R.ns;
            // B_i = cumulative quota of the steps before i: B_0 = 0, B_i = cum_steps(i) - ps, B_ns = W   (Appendix B.3)
            const u32 cs = i == 0 ? 0u : V.cum_steps(i);
            bool bad = i > 0 && cs < (u32)R.ps;
            const u32 B = i == 0 || bad ? 0u : cs - (u32)R.ps;
            i32 quota = slen - so, eo = 0;
            if (last) { quota = W - (i32)B; eo = slen - so - quota; }
            if (so < 0 || quota < 0 || eo < 0) bad = true;   // :178 assert / negative quota
            if (!bad && quota > 0) {
                const u32 eB = B + (u32)quota;
                const ParBoundary b0 = par_boundary(V, a.gaf, B), b1 = par_boundary(V, a.gaf, eB);
                if (b0.exh || b1.exh) bad = true;   // :80 assert(cur_len > target_len): CIGAR shorter than the path
                else {
                    const u32 q = b1.cq - b0.cq, nm = b1.cm - b0.cm, nb = b1.cb - b0.cb;
                    if (nm > 0) {   // gaf2paf_main.cpp:225
                        LineStep L;
                        L.rev = rev;
                        L.q0 = (u32)R.qs + b0.cq; L.q1 = L.q0 + q;
                        L.name_a = a.spos[s] + 1u - R.s; L.nl = nl; L.tlen = (u32)tlen;
                        L.ts = (u32)(sa + (rev ? eo : so)); L.te = (u32)(se - (rev ? so : eo));
                        L.nm = nm; L.nb = nb;
                        L.lenS = 0; L.codeS = 0; L.codeE = 0; L.mid_a = L.mid_b = 0;
                        L.mid_fwd = rev == rminus;   // text order == output order
                        L.lenE = eB - (b1.t > B ? b1.t : B);
                        const u32 jS = B == 0 ? 0u : (b0.cut ? b0.j : b0.j + 1u);
                        const bool cutS = B != 0 && b0.cut;
                        const u32 jE = b1.j;
                        u32 mS = jS;   // verbatim middle tokens: [mS, jE)
                        if (jS < jE) {
                            if (cutS) { L.lenS = V.incl(jS).x - B; L.codeS = a.gaf[a.opos[R.g0 + V.orig(jS)]]; mS = jS + 1; }
                            if (mS < jE) {
                                const u32 o1 = rminus ? R.no - jE : mS, o2 = rminus ? R.no - 1 - mS : jE - 1;   // original index range [o1, o2]
                                L.mid_a = (o1 ? a.opos[R.g0 + o1 - 1] + 1u : R.ca) - R.s;
                                L.mid_b = a.opos[R.g0 + o2] + 1u - R.s;
                            }
                        }
                        L.codeE = a.gaf[a.opos[R.g0 + V.orig(jE)]];
                        line = R.rconst + line_step_len(L, p10);
                        emit = 1;
                        store_line_desc(slot, R.r, 0u, line, L);   // loff: k_par_place
                    }
                }
            }
            if (bad) atomicExch(&R.status, 1u);
        }
        if (!emit) slot->rec = kDescInvalid;
        a.lx[R.s0 + i] = (u64)line | ((u64)emit << 32);
    }
}

// per record: output bytes, status, padding slots; rejected records go on to k_long
__global__ void __launch_bounds__(256) k_par_finish(const ParArgs a) {
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < a.nlist; k += gridDim.x * blockDim.x) {
        ParRec& R = a.recs[k];
        if (k == 0) { *a.n_desc = (u32)a.slot_scan[a.nlist]; *a.n_desc2 = (u32)(a.slot_scan[a.nlist] >> 32); }
        for (u32 i = R.ns; i < R.nslots; ++i) a.desc[R.slot0 + i].rec = kDescInvalid;
        if (R.status) {
            a.status[R.r] = ST_OK;   // overwritten by k_long / the general kernel
            a.out_off[R.r] = 0;
            a.reject_list[atomicAdd(a.n_reject, 1u)] = R.r;
        } else {
            const u64 d = a.lx[R.s0 + R.ns] - a.lx[R.s0];
            const u32 bytes = (u32)d;
            a.status[R.r] = ST_OK | ST_F_LONG | (bytes ? (u32)ST_F_DESC : 0u);
            a.out_off[R.r] = bytes;
        }
    }
}
// a step's line: its offset inside the record's output (or nothing left of it when the record was rejected late)
__global__ void __launch_bounds__(128) k_par_place(const ParArgs a) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < a.nsteps; s += gridDim.x * blockDim.x) {
        const ParRec& R = a.recs[a.srec[s]];
        if (R.nslots == 0) continue;
        const u32 i = s - R.s0;   // read as a normalised index: position s of lx, slot i of the run
        LineDesc* slot = a.desc + R.slot0 + i;
        if (R.status) { slot->rec = kDescInvalid; continue; }
        const u64 x0 = a.lx[s], x1 = a.lx[s + 1];
        if ((u32)((x1 - x0) >> 32) == 0) continue;   // no line
        slot->loff = (u32)(x0 - a.lx[R.s0]);
    }
}

}  // namespace g2p
