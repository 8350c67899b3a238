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
//   k_par_count    thread / 16 bytes    '>' '<' markers inside the path column, op letters inside the cg value            + u64, uint4 scans
//   k_par_ranges   thread / record      its steps / ops in the flat arrays, its run of line descriptors  + u64 scans
//   k_par_fill     thread / 16 bytes    ... the positions of the markers and letters, in text order
//   k_par_steps    thread / path step   name, ":start-end", ONE table probe                              (gaf2paf_main.cpp:157-170)
//   (k_par_count and k_par_fill also parse the CIGAR ops, for_each_cg: the first sums their (target, query, match, block)
//    contributions per tile, the second -- after a scan over the tiles -- writes every op's exclusive prefix sums.  All
//    sums are mod 2^32: differences inside a record are exact, and records whose sums do not fit 31 bits are
//    recognised by 64-bit totals and left alone)
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
    u32 pa, pb, ca, cb;       // path column and cg value, absolute text positions (one 16-byte load in the tile kernels)
    u32 r, s, len, ntabs;     // record ordinal, text start, length without '\n'
    u32 s0, ns, g0, no;       // its steps / ops in the flat arrays
    u32 status;               // 0 ok, 1 not canonical (-> k_long)
    i32 qs, ps, pe;           // ps / pe mirrored for '-' records by k_par_totals
    u32 total, rconst, minus;
    u32 slot0, nslots;        // descriptor run
    u32 pad0, pad1, pad2;
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
    uint4* tile_osum;       // [ntiles + 1] exclusive sums of the tiles' op contributions
    u32* spos; u32* srec;   // per step: marker position, record (list index)
    u32* opos;              // per op: letter position
    uint4* sval;            // per step: tlen, sa, se, nl | '<' << 16                        (k_par_steps)
    u64* sx;                // per step (+1): step length, scanned
    uint4* ox;              // per op (+1): exclusive prefix sums of {target, query, match, block}  (k_par_fill, from the tile sums)
    u32* ot;                // per op (+1): ox[].x alone, what the boundary search reads
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
    const uint2 m = __ldg(a.tile_map + tile);
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
    u32 t_pos;
    ParRec& R = a.recs[par_tile_record(a, blockIdx.x, t_pos)];
    const ParChunk c = par_chunk(a, t_pos);
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
            if ((u32)t[cb - 1] - '0' <= 9u) break;   // digits after the last op letter
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
__device__ __forceinline__ void par_masks(const uint4 sp, const ParChunk& c, u32& mm, u32& om) {
    mm = om = 0;   // (records k_par_head rejected have empty spans; a status set later must not change what is counted)
    const u32 pr = par_range(c, sp.x, sp.y), cr = par_range(c, sp.z, sp.w);
    if (pr) mm = (par_eq_mask(c, 0x3E3E3E3Eu) | par_eq_mask(c, 0x3C3C3C3Cu)) & pr;
    if (cr) om = par_nondigit_mask(c) & cr;
}
// The CTA's tile in shared memory behind a 16-byte halo (the text just before the tile), so that the digits of an op
// are read from there whichever thread or tile they begin in.  Returns the calling thread's chunk.
__device__ __forceinline__ ParChunk par_stage(const ParArgs& a, u32 tile_pos, u8* sm) {
    const ParChunk c = par_chunk(a, tile_pos);
    reinterpret_cast<uint4*>(sm)[1 + threadIdx.x] = make_uint4(c.w0, c.w1, c.w2, c.w3);
    if (threadIdx.x == 0) reinterpret_cast<uint4*>(sm)[0] = tile_pos >= 16u ? ldg_vec_guarded(a.gaf, (u64)tile_pos - 16u, a.n) : make_uint4(0, 0, 0, 0);
    __syncthreads();
    return c;
}
// One CIGAR op (for_each_cg, gafkluge.hpp:226-239), its letter at offset `off` of the staged tile.  The digits before it
// come from one unaligned 4-byte window when there are at most three (the byte before the first digit is a letter or
// the ':' of "cg:Z:"), else from a byte loop.  Returns length | class << 24 (kPOpT / Q / M: consumes target / query /
// counts as match; kPOpValid), or 0 if it is not a canonical op.
constexpr u32 kPOpT = 1u << 24, kPOpQ = 1u << 25, kPOpM = 1u << 26, kPOpValid = 1u << 27;
__device__ __forceinline__ u32 par_op(const u8* sm, u32 off) {
    const u32 kc = (u32)sm[off] - '=';
    const u32 o4 = off - 4u;
    const u32* q = reinterpret_cast<const u32*>(sm + (o4 & ~3u));
    const u32 W = __funnelshift_r(q[0], q[1], (o4 & 3u) * 8u);   // byte 3 = the last digit
    const u32 F = nondigit_bytes(W);
    u32 x, nd, first;
    if (F) {   // at most three digits
        nd = (u32)__clz((int)F) >> 3;
        const u32 D = W & 0x0F0F0F0Fu;
        const u32 d3 = D >> 24, d2 = (D >> 16) & 0xffu, d1 = (D >> 8) & 0xffu;
        x = d3 + (nd > 1 ? d2 * 10u : 0u) + (nd > 2 ? d1 * 100u : 0u);
        first = nd > 2 ? d1 : d2;
    } else {
        x = 0; nd = 0; first = 0;
        u32 p = 1;
        while (nd < 8u) {
            const u32 d = (u32)sm[off - 1u - nd] - '0';
            if (d > 9u) break;
            x += d * p; p *= 10u; first = d; ++nd;
        }
    }
    if (!(kc < 28u && ((kOpMask >> kc) & 1u)) || nd == 0 || nd > 7 || (nd > 1 && first == 0) || x == 0) return 0u;
    return x | kPOpValid | (((kTargetMask >> kc) & 1u) ? kPOpT : 0u) | (((kQueryMask >> kc) & 1u) ? kPOpQ : 0u) | (((kMatchMask >> kc) & 1u) ? kPOpM : 0u);
}
// ... its (target, query, match, block) contributions
__device__ __forceinline__ uint4 par_op_sums(u32 op) {
    const u32 x = op & 0xffffffu;
    return make_uint4(op & kPOpT ? x : 0u, op & kPOpQ ? x : 0u, op & kPOpM ? x : 0u, x);
}
// sum over the CTA, valid in thread 0
__device__ __forceinline__ uint4 par_block_sum4(uint4 v) {
    __shared__ uint4 ws4[kPThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, o); v.y += __shfl_down_sync(0xffffffffu, v.y, o);
        v.z += __shfl_down_sync(0xffffffffu, v.z, o); v.w += __shfl_down_sync(0xffffffffu, v.w, o);
    }
    if ((threadIdx.x & 31) == 0) ws4[threadIdx.x >> 5] = v;
    __syncthreads();
    uint4 s = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) for (u32 i = 0; i < kPThreads / 32; ++i) s = add4(s, ws4[i]);
    return s;
}
__global__ void __launch_bounds__(kPThreads) k_par_count(const ParArgs a) {
    __shared__ u64 ws[kPThreads / 32];
    __shared__ __align__(16) u8 s_text[16 + kPTile];
    u32 t_pos;
    ParRec& R = a.recs[par_tile_record(a, blockIdx.x, t_pos)];   // (every thread reads the same map entry: no barrier before the loads below)
    const uint4 spans = __ldg(reinterpret_cast<const uint4*>(&R));
    const ParChunk ch = par_stage(a, t_pos, s_text);
    u32 mm, om;
    par_masks(spans, ch, mm, om);
    u64 c = (u64)__popc(mm) | ((u64)__popc(om) << 32);
    uint4 sum = make_uint4(0, 0, 0, 0);
    bool bad = false;
    while (om) {
        const u32 b = (u32)__ffs((int)om) - 1u;
        om &= om - 1u;
        const u32 op = par_op(s_text, 16u + threadIdx.x * 16u + b);
        if (!op) bad = true;
        sum = add4(sum, par_op_sums(op));
    }
    if (bad) atomicExch(&R.status, 1u);   // (k_par_fill sees the same ops and contributes the same zeros)
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    const uint4 tot = par_block_sum4(sum);   // (contains the barrier that orders ws)
    if (threadIdx.x == 0) {
        u64 s = 0;
        for (u32 i = 0; i < kPThreads / 32; ++i) s += ws[i];
        a.tile_off[blockIdx.x] = s;
        a.tile_osum[blockIdx.x] = tot;
        if (tot.w) atomicAdd(&R.sum_ops, (unsigned long long)tot.w);   // (a tile's block length fits 32 bits: 1024 ops of < 10^7)
    }
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
    __shared__ u64 ws[kPThreads / 32];
    __shared__ __align__(16) u8 s_text[16 + kPTile];
    u32 t_pos;
    const u32 k = par_tile_record(a, blockIdx.x, t_pos);
    const uint4 spans = __ldg(reinterpret_cast<const uint4*>(a.recs + k));
    const ParChunk ch = par_stage(a, t_pos, s_text);
    u32 mm, om;
    par_masks(spans, ch, mm, om);
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
    // the ops: values as in k_par_count, prefix = tile prefix (scanned) + the threads before this one + the ops before this one
    // (parsed once; a canonical chunk holds at most 8 ops -- with more, one of them has no digits and the record is rejected)
    __shared__ u32 s_ops[kPThreads * 8];
    const u32 soff = 16u + threadIdx.x * 16u;
    uint4 vals = make_uint4(0, 0, 0, 0);
    {
        u32 j = 0;
        for (u32 m = om; m; ++j) {
            const u32 b = (u32)__ffs((int)m) - 1u;
            m &= m - 1u;
            const u32 op = par_op(s_text, soff + b);
            if (j < 8u) s_ops[threadIdx.x * 8u + j] = op;
            vals = add4(vals, par_op_sums(op));
        }
    }
    uint4 inc4 = vals;
    for (int o = 1; o < 32; o <<= 1) { const uint4 up = shfl_up4(inc4, o); if (lane >= (u32)o) inc4 = add4(inc4, up); }
    __shared__ uint4 ws4[kPThreads / 32];
    if (lane == 31) ws4[warp] = inc4;
    __syncthreads();
    uint4 run = add4(a.tile_osum[blockIdx.x], sub4(inc4, vals));
    for (u32 i = 0; i < warp; ++i) run = add4(run, ws4[i]);
    for (u32 j = 0; om; ++j) {
        const u32 b = (u32)__ffs((int)om) - 1u;
        om &= om - 1u;
        a.opos[oi] = ch.pos + b; a.ox[oi] = run; a.ot[oi] = run.x;
        run = add4(run, par_op_sums(j < 8u ? s_ops[threadIdx.x * 8u + j] : 0u));
        ++oi;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { const uint4 t = a.tile_osum[a.ntiles]; a.ox[a.nops] = t; a.ot[a.nops] = t.x; }
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

// one path step: "[><]name[:start-end]", one table probe.  (Flags instead of early exits: the probe, a dependent random
// access, must be issued once for the whole warp, not once per divergent path that leads to it.)
__global__ void __launch_bounds__(128) k_par_steps(const ParArgs a) {
    const u32 nround = (a.nsteps + 31u) & ~31u;
    const u8* g = a.gaf;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < nround; s += gridDim.x * blockDim.x) {
        const bool in = s < a.nsteps;
        const u32 k = in ? a.srec[s] : 0u;
        ParRec& R = a.recs[k];
        const bool live = in && !R.status;
        // ---- the token: name length, key
        u32 mp = 0, end = 0, nl = 0, mode = 0;   // mode 1: name of <= 16 bytes (its own key), 2: longer
        u64 k0 = 0, k1 = 0;
        if (live) {
            mp = a.spos[s];
            end = s + 1 < R.s0 + R.ns ? a.spos[s + 1] : R.pb;
            const u32 na = mp + 1, tl = end - na;
            bool scan = true;
            u32 e = na;
            if ((u64)mp + 28 <= a.n) {   // the first 16 bytes in registers: ':' by SWAR
                const u32* q = reinterpret_cast<const u32*>(g + (na & ~3u));
                const u32 sh = (na & 3u) * 8u;
                const u32 x0 = __ldg(q), x1 = __ldg(q + 1), x2 = __ldg(q + 2), x3 = __ldg(q + 3), x4 = __ldg(q + 4);
                u32 w0 = __funnelshift_r(x0, x1, sh), w1 = __funnelshift_r(x1, x2, sh), w2 = __funnelshift_r(x2, x3, sh), w3 = __funnelshift_r(x3, x4, sh);
                u32 cm = movemask4(zero_bytes(w0 ^ 0x3A3A3A3Au)) | (movemask4(zero_bytes(w1 ^ 0x3A3A3A3Au)) << 4) |
                         (movemask4(zero_bytes(w2 ^ 0x3A3A3A3Au)) << 8) | (movemask4(zero_bytes(w3 ^ 0x3A3A3A3Au)) << 12);
                cm &= tl >= 16 ? 0xffffu : ((1u << tl) - 1u);
                if (cm || tl <= 16) {
                    nl = cm ? (u32)__ffs((int)cm) - 1u : tl;
                    w0 = keep_bytes(w0, (int)nl); w1 = keep_bytes(w1, (int)nl - 4); w2 = keep_bytes(w2, (int)nl - 8); w3 = keep_bytes(w3, (int)nl - 12);
                    k0 = (u64)w0 | ((u64)w1 << 32); k1 = (u64)w2 | ((u64)w3 << 32);
                    mode = nl ? 1u : 0u;
                    scan = false;
                } else e = na + 16;
            }
            if (scan) {
                while (e < end && g[e] != ':') ++e;
                nl = e - na;
                mode = nl != 0 && nl <= 255 ? 2u : 0u;
            }
        }
        // ---- the probe
        i64 tl64 = 0;
        bool found = false;
        if (mode == 1) found = table_lookup_key16(a.T, k0, k1, nl, tl64);
        else if (mode == 2) found = table_lookup(a.T, g + mp + 1, nl, tl64);
        bool ok = found && tl64 >= 0 && tl64 <= 0x7fffffffLL;
        // ---- ":start-end" (gafkluge.hpp:131-146), plain digits only
        u32 sa = 0, se = (u32)tl64;
        const u32 e = mp + 1 + nl;
        if (ok && e < end) {
            u32 q = e + 1, x = 0;
            const u32 q1 = q;
            while (q < end && (u32)g[q] - '0' <= 9u && q - q1 < 10) { x = x * 10u + ((u32)g[q] - '0'); ++q; }
            const bool ok1 = q != q1 && q - q1 <= 9 && q < end && g[q] == '-';
            sa = x;
            const u32 q2 = ++q;
            x = 0;
            while (ok1 && q < end && (u32)g[q] - '0' <= 9u && q - q2 < 10) { x = x * 10u + ((u32)g[q] - '0'); ++q; }
            ok = ok1 && q != q2 && q - q2 <= 9 && q == end && x >= sa;
            se = x;
        }
        const u32 slen = ok ? se - sa : 0u;
        if (live && !ok) atomicExch(&R.status, 1u);
        par_sum64(&R.sum_steps, k, slen, live && ok);
        if (in) {
            a.sval[s] = ok ? make_uint4((u32)tl64, sa, se, nl | ((u32)(g[mp] == '<') << 16)) : make_uint4(0, 0, 0, 0);
            a.sx[s] = slen;
        }
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
    const u64* sx; const uint4* ox; const u32* ot; const u32* opos;
    u32 s0, ns, g0, no, total;
    uint4 obase, ototal;
    bool minus;
    // cumulative step length before normalised step i (0 <= i <= ns)
    __device__ __forceinline__ u32 cum_steps(u32 i) const { return minus ? total - (u32)(sx[s0 + ns - i] - sx[s0]) : (u32)(sx[s0 + i] - sx[s0]); }
    // inclusive prefix sums over the ops in normalised order, op j (0 <= j < no)
    __device__ __forceinline__ uint4 incl(u32 j) const {
        if (!minus) return sub4(ox[g0 + j + 1], obase);
        return sub4(ototal, sub4(ox[g0 + no - 1 - j], obase));   // total - exclusive prefix of the original op no-1-j
    }
    // ... its target component alone, from the compact array
    __device__ __forceinline__ u32 incl_t(u32 j) const {
        if (!minus) return __ldg(ot + g0 + j + 1) - obase.x;
        return ototal.x - (__ldg(ot + g0 + no - 1 - j) - obase.x);
    }
    __device__ __forceinline__ u32 orig(u32 j) const { return minus ? no - 1 - j : j; }
};
struct ParBoundary { u32 j, t, cq, cm, cb; bool cut, exh; };
// first op (normalised order) in [lo, hi) whose cumulative target length reaches B, else hi
__device__ __forceinline__ u32 par_lower_bound(const ParView& V, u32 B, u32 lo, u32 hi) {
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (V.incl_t(mid) >= B) hi = mid; else lo = mid + 1; }
    return lo;
}
// two of them side by side (independent loads in every round)
__device__ __forceinline__ void par_lower_bound2(const ParView& V, u32 B0, u32 B1, u32 lo, u32 hi, u32& r0, u32& r1) {
    u32 l0 = lo, h0 = hi, l1 = lo, h1 = hi;
    while (l0 < h0 || l1 < h1) {
        const u32 m0 = (l0 + h0) >> 1, m1 = (l1 + h1) >> 1;
        const u32 t0 = l0 < h0 ? V.incl_t(m0) : 0u, t1 = l1 < h1 ? V.incl_t(m1) : 0u;
        if (l0 < h0) { if (t0 >= B0) h0 = m0; else l0 = m0 + 1; }
        if (l1 < h1) { if (t1 >= B1) h1 = m1; else l1 = m1 + 1; }
    }
    r0 = l0; r1 = l1;
}
// the same for a whole warp that shares V, B0 and B1: 32 probes per round and bound, ~log32(no) rounds
__device__ __forceinline__ void par_lower_bound2_warp(const ParView& V, u32 B0, u32 B1, u32& r0, u32& r1) {
    const u32 FULL = 0xffffffffu, lane = threadIdx.x & 31u;
    u32 l0 = 0, h0 = V.no, l1 = 0, h1 = V.no;
    while (l0 < h0 || l1 < h1) {   // (uniform)
        const u32 s0 = (h0 - l0 + 31u) >> 5, s1 = (h1 - l1 + 31u) >> 5;
        u32 p0 = l0 + (lane + 1u) * s0 - 1u, p1 = l1 + (lane + 1u) * s1 - 1u;
        p0 = p0 < h0 ? p0 : h0 - 1u; p1 = p1 < h1 ? p1 : h1 - 1u;
        const bool a0 = l0 < h0, a1 = l1 < h1;
        const u32 t0 = a0 ? V.incl_t(p0) : 0u, t1 = a1 ? V.incl_t(p1) : 0u;
        const u32 m0 = __ballot_sync(FULL, a0 && t0 >= B0), m1 = __ballot_sync(FULL, a1 && t1 >= B1);
        if (a0) {
            if (m0 == 0) l0 = h0;   // no op of [l0, h0) reaches B0
            else {
                const int f = __ffs((int)m0) - 1;
                const u32 pf = __shfl_sync(FULL, p0, f), pp = __shfl_sync(FULL, p0, f ? f - 1 : 0);
                if (f) l0 = pp + 1u;
                h0 = pf;            // the answer lies in [l0, pf]
            }
        }
        if (a1) {
            if (m1 == 0) l1 = h1;
            else {
                const int f = __ffs((int)m1) - 1;
                const u32 pf = __shfl_sync(FULL, p1, f), pp = __shfl_sync(FULL, p1, f ? f - 1 : 0);
                if (f) l1 = pp + 1u;
                h1 = pf;
            }
        }
    }
    r0 = l0; r1 = l1;
}
// boundary B of the cumulative target length, j = its lower bound in the op stream -> the sums up to it (as k_short's phase 5)
__device__ __forceinline__ ParBoundary par_boundary(const ParView& V, const u8* gaf, u32 B, u32 j) {
    ParBoundary b;
    b.j = b.t = b.cq = b.cm = b.cb = 0; b.cut = b.exh = false;
    if (B == 0) return b;
    b.j = j;
    if (j == V.no) { b.exh = true; return b; }
    const uint4 E = V.incl(j);
    const uint4 P = j ? V.incl(j - 1) : make_uint4(0, 0, 0, 0);
    b.t = P.x;
    const u32 off = B - P.x;
    const u32 kc = (u32)gaf[V.opos[V.g0 + V.orig(j)]] - '=';
    b.cq = P.y + (((kQueryMask >> kc) & 1u) ? off : 0u);
    b.cm = P.z + (((kMatchMask >> kc) & 1u) ? off : 0u);
    b.cb = P.w + off;
    b.cut = E.x > B;
    return b;
}

// one PAF line per path step (gaf2paf_main.cpp:157-263 in the closed form of SURVEY.md Appendix B)
__global__ void __launch_bounds__(128) k_par_lines(const ParArgs a) {
    const u32 FULL = 0xffffffffu;
    __shared__ u32 p10[10];
    if (threadIdx.x < 10) { u32 v = 1; for (u32 i = 0; i < threadIdx.x; ++i) v *= 10u; p10[threadIdx.x] = v; }
    __syncthreads();
    const u32 nround = (a.nsteps + 31u) & ~31u;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < nround; s += gridDim.x * blockDim.x) {
        const bool in = s < a.nsteps;
        const u32 k = in ? a.srec[s] : 0xffffffffu;
        ParRec& R = a.recs[in ? k : 0u];
        const bool has_run = in && R.nslots != 0;   // else: rejected before it got a descriptor run
        const bool rminus = R.minus != 0;
        // this thread owns ORIGINAL step s - s0; its normalised index:
        const u32 jo = s - R.s0, i = rminus ? R.ns - 1 - jo : jo;
        u32 line = 0, emit = 0;
        LineDesc* slot = a.desc + R.slot0 + i;
        ParView V;
        V.sx = a.sx; V.ox = a.ox; V.ot = a.ot; V.opos = a.opos;
        V.s0 = R.s0; V.ns = R.ns; V.g0 = R.g0; V.no = R.no; V.total = R.total; V.minus = rminus;
        V.obase = V.ototal = make_uint4(0, 0, 0, 0);
        i32 sa = 0, se = 0, tlen = 0, so = 0, eo = 0, quota = 0;
        u32 nl = 0, B = 0, eB = 0;
        bool rev = false, bad = false;
        const bool live = has_run && !R.status;
        if (has_run) {   // (also for lanes whose record was rejected meanwhile: the warp-wide search below reads through every lane's view)
            V.obase = a.ox[R.g0];
            V.ototal = sub4(a.ox[R.g0 + R.no], V.obase);
        }
        if (live) {
            const uint4 sv = a.sval[s];
            tlen = (i32)sv.x; sa = (i32)sv.y; se = (i32)sv.z;
            nl = sv.w & 0xffffu;
            rev = ((sv.w >> 16) & 1u) != (u32)rminus;
            const i32 slen = se - sa;
            const i32 W = R.pe - R.ps;
            so = i == 0 ? R.ps : 0;
            const bool last = i + 1 == R.ns;
            // B_i = cumulative quota of the steps before i: B_0 = 0, B_i = cum_steps(i) - ps, B_ns = W   (Appendix B.3)
            const u32 cs = i == 0 ? 0u : V.cum_steps(i);
            bad = i > 0 && cs < (u32)R.ps;
            B = i == 0 || bad ? 0u : cs - (u32)R.ps;
            quota = slen - so;
            if (last) { quota = W - (i32)B; eo = slen - so - quota; }
            if (so < 0 || quota < 0 || eo < 0) bad = true;   // :178 assert / negative quota
            eB = B + (u32)(quota > 0 ? quota : 0);
        }
        const bool need = live && !bad && quota > 0;
        // 32 consecutive steps of one record: their boundaries lie between those of the first and the last of them.  The
        // warp brackets them together (32 probes per round), then every lane searches its two boundaries in the bracket
        // (a few hundred bytes of `ot`), side by side.
        u32 lo = 0, hi = V.no;
        if (__all_sync(FULL, k == __shfl_sync(FULL, k, 0)) && __any_sync(FULL, need)) {
            u32 bmin = need ? B : 0xffffffffu, bmax = need ? eB : 0u;
            for (int o = 16; o > 0; o >>= 1) {
                const u32 x0 = __shfl_xor_sync(FULL, bmin, o), x1 = __shfl_xor_sync(FULL, bmax, o);
                bmin = x0 < bmin ? x0 : bmin;
                bmax = x1 > bmax ? x1 : bmax;
            }
            par_lower_bound2_warp(V, bmin, bmax, lo, hi);
        }
        if (need) {
            u32 j0, j1;
            par_lower_bound2(V, B, eB, lo, hi, j0, j1);
            const ParBoundary b0 = par_boundary(V, a.gaf, B, j0), b1 = par_boundary(V, a.gaf, eB, j1);
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
                        if (cutS) { L.lenS = V.incl_t(jS) - B; L.codeS = a.gaf[a.opos[R.g0 + V.orig(jS)]]; mS = jS + 1; }
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
        if (live && bad) atomicExch(&R.status, 1u);
        if (has_run) {
            if (!emit) slot->rec = kDescInvalid;
            a.lx[R.s0 + i] = (u64)line | ((u64)emit << 32);
        } else if (in) a.lx[s] = 0;   // (every step of the record writes one zero)
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
