// g2p_rec.cuh — size pass for short records, one thread per record.
//
// k_rec replaces k_short's size pass on the throughput path.  k_short spends ~690 warp
// instructions per record on coordination (SWAR classification of every byte into three masks,
// shuffle scans, scatter of token positions, five group votes) with 18 of 32 lanes active.  Here a
// record is parsed by ONE thread with a plain scalar walk over its text (~210 warp instructions per
// record, 20 lanes active), and the kernel is built so that the 32 lanes of a warp stay in the same
// loop at the same time:
//
//   * order: the CTA's 256 records are counting-sorted by length, so that the lanes of a warp hold
//     records of similar shape (a warp pays the maximum trip count over its lanes);
//   * staging: the records are copied from global memory with 128-bit coalesced loads (8 lanes per
//     record) into per-record shared-memory slots of an ODD number of 32-bit words, so that lane
//     i's record starts in bank (i * stride) mod 32: byte loads of the 32 lanes at equal progress
//     hit 32 different banks;
//   * the walk is a fixed sequence of short loops (one per column, one per tag, one per path step,
//     one per CIGAR op), never a state machine: lanes re-converge after every loop;
//   * pass S: the path column is tokenised forwards once, every step name is probed once in the
//     lengths table, and two words per step are left in a dead part of the record's own slot (the
//     columns and tags between the path and the CIGAR are not needed again);
//   * '+' records then walk steps and CIGAR forwards, '-' records backwards (flip_gaf,
//     gaf2paf_main.cpp:92-131, is index arithmetic) -- one instruction stream for both strands;
//   * the CIGAR is cut at step boundaries by streaming (cigar_next_by_target, gaf2paf_main.cpp:71-90):
//     ops are taken until the step's target quota is reached, the op that crosses the boundary is
//     split and its remainder starts the next step.
//
// Outputs are exactly k_short's: per record the PAF byte count, line count, status and a RecDesc,
// per PAF line a LineDesc in the record's own kSMaxLines slots; the line map and k_emit_lines are
// unchanged.  Only canonical records are converted (same definition as k_short, see g2p_short.cuh);
// anything else is appended to the delegate list for k_par / k_long / the general kernel.
#pragma once
#include "g2p_short.cuh"

namespace g2p {

constexpr int kRThreads = 256;          // records per CTA
constexpr u32 kRStage = 8;              // lanes per record in the staging copy
constexpr u32 kRMaxTags = 8;
constexpr u32 kRMinChunks = 6, kRMaxChunks = 16;   // 16-byte chunks per record slot (host picks from the mean record length)
#ifndef G2P_REC_CTAS
#define G2P_REC_CTAS 4
#endif
// Variants measured on B200 (tools/gpu_quick.sh): loop-free decoding of <= 3-digit numbers and
// four-bytes-per-trip scans.  0 = the plain byte loops.
#ifndef G2P_REC_FAST_OP
#define G2P_REC_FAST_OP 0
#endif
#ifndef G2P_REC_FAST_NUM
#define G2P_REC_FAST_NUM 0
#endif
#ifndef G2P_REC_STAGE_PASSES
#define G2P_REC_STAGE_PASSES 2   /* staging passes whose loads are in flight together (1, 2, 4 or 8) */
#endif
static_assert(kRMaxChunks <= 2 * kRStage && (kRThreads / (kRThreads / kRStage)) % G2P_REC_STAGE_PASSES == 0, "staging copies a record in two rounds of kRStage lanes");
#ifndef G2P_REC_SORT
#define G2P_REC_SORT 1   /* order the CTA's records by length before assigning them to threads */
#endif
#ifndef G2P_REC_SCAN4
#define G2P_REC_SCAN4 1
#endif

__host__ __device__ __forceinline__ u32 rec_slot_words(u32 chunks) { return 4u * chunks + 1u; }   // odd: conflict-free lane stride
static inline size_t rec_smem(u32 chunks) { return (size_t)kRThreads * rec_slot_words(chunks) * 4u + 32u; }

// Slot capacity for an input of n bytes in nrec records: ~1.4x the mean record, so that the bulk of a
// short-read file fits while four CTAs stay resident per SM; longer records go to k_long.
static inline u32 rec_chunks_for(u64 n, u32 nrec) {
    const u64 mean = nrec ? n / nrec : 0;
    u64 c = (mean * 7 / 5 + 15) / 16 + 1;
    if (c < kRMinChunks) c = kRMinChunks;
    if (c > kRMaxChunks) c = kRMaxChunks;
    return (u32)c;
}

struct RecArgs {
    ShortArgs s;
    u32 chunks;        // slot capacity in 16-byte chunks: a record is taken if its staged span (phase + bytes + '\n') fits
    const u32* perm;   // records ordered by length class (k_len_hist / k_len_scatter), or null: input order + in-CTA sort
};

// ---- global ordering of the records by length class -----------------------------------------
// A CTA of k_rec converts 256 consecutive entries of `perm`; with the records bucketed by length
// (64 classes of 4 bytes) these have the same length class, so (a) the lanes of a warp run the same
// loop trip counts and (b) the warps of a CTA finish together -- an in-CTA sort alone gives (a) but
// makes the warp with the longest records hold the CTA's shared memory ~25 % longer than the mean.
// Counting sort: per-CTA histograms (bin-major matrix), the pipeline's u64 scan kernels, scatter.
constexpr u32 kLenBins = 64, kLenSortRecs = 2048;
__device__ __forceinline__ u32 len_bin(u32 bytes_with_nl) { return (bytes_with_nl < 255u ? bytes_with_nl : 255u) >> 2; }
__global__ void __launch_bounds__(256) k_len_hist(const u32* __restrict__ rec_start, u32 nrec, u32 ncta, u64* __restrict__ M) {
    __shared__ u32 h[kLenBins];
    if (threadIdx.x < kLenBins) h[threadIdx.x] = 0;
    __syncthreads();
    for (u32 k = 0; k < kLenSortRecs / 256u; ++k) {
        const u32 r = blockIdx.x * kLenSortRecs + k * 256u + threadIdx.x;
        if (r < nrec) atomicAdd(&h[len_bin(rec_start[r + 1] - rec_start[r])], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kLenBins) M[(size_t)threadIdx.x * ncta + blockIdx.x] = h[threadIdx.x];
}
__global__ void __launch_bounds__(256) k_len_scatter(const u32* __restrict__ rec_start, u32 nrec, u32 ncta, const u64* __restrict__ M,
                                                     u32* __restrict__ perm) {
    __shared__ u32 base[kLenBins];
    if (threadIdx.x < kLenBins) base[threadIdx.x] = (u32)M[(size_t)threadIdx.x * ncta + blockIdx.x];
    __syncthreads();
    for (u32 k = 0; k < kLenSortRecs / 256u; ++k) {
        const u32 r = blockIdx.x * kLenSortRecs + k * 256u + threadIdx.x;
        if (r < nrec) perm[atomicAdd(&base[len_bin(rec_start[r + 1] - rec_start[r])], 1u)] = r;
    }
}

__device__ __forceinline__ u32 haszero16(u32 x) { return (x - 0x00010001u) & ~x & 0x80008000u; }

// First position >= p holding a tab or the record's '\n'; four bytes per trip, loaded together.
__device__ __forceinline__ u32 rec_scan_field(const u8* rt, u32 p) {
#if !G2P_REC_SCAN4
    u32 c;
    while ((c = rt[p]) != '\t' && c != '\n') ++p;
    return p;
#else
    for (;;) {
        const u32 c0 = rt[p], c1 = rt[p + 1], c2 = rt[p + 2], c3 = rt[p + 3];
        const bool t0 = c0 - 9u <= 1u, t1 = c1 - 9u <= 1u, t2 = c2 - 9u <= 1u, t3 = c3 - 9u <= 1u;
        if (t0 || t1 || t2 || t3) return p + (t0 ? 0u : (t1 ? 1u : (t2 ? 2u : 3u)));
        p += 4;
    }
#endif
}
// First position >= p holding ':', '>' or '<' (pass S; the caller planted a '>' after the path column).
__device__ __forceinline__ u32 rec_scan_step(const u8* rt, u32 p) {
#if !G2P_REC_SCAN4
    u32 c;
    while ((c = rt[p]) != ':' && c != '>' && c != '<') ++p;
    return p;
#else
    for (;;) {
        const u32 c0 = rt[p], c1 = rt[p + 1], c2 = rt[p + 2], c3 = rt[p + 3];
        const bool t0 = c0 == ':' || c0 == '>' || c0 == '<', t1 = c1 == ':' || c1 == '>' || c1 == '<';
        const bool t2 = c2 == ':' || c2 == '>' || c2 == '<', t3 = c3 == ':' || c3 == '>' || c3 == '<';
        if (t0 || t1 || t2 || t3) return p + (t0 ? 0u : (t1 ? 1u : (t2 ? 2u : 3u)));
        p += 4;
    }
#endif
}
// Pass S: the path column, forwards, once per record.  Every step token "[><]name[:start-end]" is
// scanned with one loop (name end and token end together), its name is probed in the lengths table,
// and two words per step are left in `stab` (a dead part of the record's own slot):
//   stab[2i]   = target length | interval flag << 31
//   stab[2i+1] = marker position | name length << 8 | token end << 16
// The caller has replaced the tab after the path column by '>' so that the scan needs no bound.
// Returns false if a token is not canonical or a name is unknown (-> delegate); `total` = sum of
// the step lengths (flip_gaf's path_target_len, gaf2paf_main.cpp:111-127).
__device__ __forceinline__ bool rec_steps(const LenTableView& T, const u8* rt, const u32 pa, const u32 pb, const bool prefixed, u32* stab,
                                          const u32 cap, u32& ns_out, u64& total_out) {
    u32 ns = 0, mp = prefixed ? pa : pa - 1;
    u64 total = 0;
    for (;;) {
        const u32 name_a = mp + 1;
        u32 j = pb;
        u8 c = 0;
        if (prefixed) {
            j = rec_scan_step(rt, name_a);   // rt[pb] == '>'
            c = rt[j];
        }
        const u32 nl = j - name_a;
        if (nl == 0 || nl > 16) return false;
        u32 w0, w1, w2, w3;
        lds16_unaligned(rt + name_a, w0, w1, w2, w3);
        w0 = keep_bytes(w0, (int)nl); w1 = keep_bytes(w1, (int)nl - 4);
        w2 = keep_bytes(w2, (int)nl - 8); w3 = keep_bytes(w3, (int)nl - 12);
        u32 slen = 0;
        const bool interval = prefixed && c == ':';
        if (interval) {   // ":start-end" (gafkluge.hpp:131-146), plain digits only
            u32 k = j + 1, x = 0, d;
            const u32 k1 = k;
            while ((d = (u32)rt[k] - '0') <= 9u) { x = x * 10u + d; ++k; }
            if (k == k1 || k - k1 > 9 || rt[k] != '-') return false;
            const u32 sa = x;
            ++k; x = 0;
            const u32 k2 = k;
            while ((d = (u32)rt[k] - '0') <= 9u) { x = x * 10u + d; ++k; }
            if (k == k2 || k - k2 > 9 || (rt[k] != '>' && rt[k] != '<') || x < sa) return false;
            slen = x - sa;
            j = k;
        }
        i64 tl64;
        if (!table_lookup_key16(T, (u64)w0 | ((u64)w1 << 32), (u64)w2 | ((u64)w3 << 32), nl, tl64) || tl64 < 0 || tl64 > 0x7fffffffLL) return false;
        if (!interval) slen = (u32)tl64;
        total += slen;
        if (ns >= cap) return false;
        stab[2 * ns] = (u32)tl64 | (interval ? 0x80000000u : 0u);
        stab[2 * ns + 1] = (mp & 0xffu) | (nl << 8) | (j << 16);
        ++ns;
        if (j >= pb) break;
        mp = j;
    }
    ns_out = ns; total_out = total;
    return true;
}

// One CIGAR token in walk direction.  Forward: "digits letter" starts at cp, cp advances past the
// letter.  Backward: the token ends at cp (exclusive), cp retreats to its first digit.  [ts, te) is
// the token's text span.  Tokens of up to three digits (all of a short read's) are decoded without
// a loop from the five bytes next to the cursor, loaded together: one shared-memory round trip per
// op instead of one per digit, and the same instruction stream for '+' and '-' records.
#if G2P_REC_FAST_OP == 1
#define G2P_REC_SLOW_INLINE G2P_NOINLINE
#else
#define G2P_REC_SLOW_INLINE __forceinline__
#endif
__device__ G2P_REC_SLOW_INLINE bool rec_fetch_op_slow(const u8* rt, const bool minus, u32& cp, u32& x, u32& kc, u32& ts, u32& te) {
    u32 k = cp, letter = 0;
    if (minus) { letter = rt[cp - 1]; k = cp - 2; }   // rt[ca - 1] == ':' stops the backward digit walk
    const u32 k0 = k;
    const u32 step = minus ? 0xffffffffu : 1u;
    u32 v = 0, mul = 1, d, edge = 1, firstd = 1;
    bool any = false;
    while ((d = (u32)rt[k] - '0') <= 9u) {
        v = minus ? v + d * mul : v * 10u + d;
        mul *= 10u;
        if (!any) firstd = d;
        any = true;
        edge = d;
        k += step;
    }
    const u32 nd = minus ? k0 - k : k - k0;
    const u32 lead = minus ? edge : firstd;   // most significant digit
    if (minus) { ts = k + 1; te = cp; cp = k + 1; }
    else { letter = rt[k]; ts = cp; te = k + 1; cp = k + 1; }
    kc = letter - '=';
    x = v;
    return nd != 0 && nd <= 7 && !(nd > 1 && lead == 0) && v != 0 && kc < 28u && ((kOpMask >> kc) & 1u);
}
__device__ __forceinline__ bool rec_fetch_op(const u8* rt, const bool minus, u32& cp, u32& x, u32& kc, u32& ts, u32& te) {
#if !G2P_REC_FAST_OP
    return rec_fetch_op_slow(rt, minus, cp, x, kc, ts, te);
#else
    const u32 step = minus ? 0xffffffffu : 1u;
    const u32 base = minus ? cp - 1 : cp;
    const u32 b0 = rt[base], b1 = rt[base + step], b2 = rt[base + 2u * step], b3 = rt[base + 3u * step], b4 = rt[base + 4u * step];
    // digits in walk order: forward b0 b1 b2 (b3), backward b1 b2 b3 (b4) after the letter b0
    const u32 d0 = (minus ? b1 : b0) - '0', d1 = (minus ? b2 : b1) - '0', d2 = (minus ? b3 : b2) - '0', d3 = (minus ? b4 : b3) - '0';
    if (d0 <= 9u && d1 <= 9u && d2 <= 9u && d3 <= 9u) return rec_fetch_op_slow(rt, minus, cp, x, kc, ts, te);   // >= 4 digits
    const u32 nd = d0 > 9u ? 0u : (d1 > 9u ? 1u : (d2 > 9u ? 2u : 3u));
    const u32 v2 = minus ? d1 * 10u + d0 : d0 * 10u + d1;
    const u32 v3 = minus ? d2 * 100u + d1 * 10u + d0 : d0 * 100u + d1 * 10u + d2;
    const u32 v = nd == 1 ? d0 : (nd == 2 ? v2 : v3);
    const u32 lead = minus ? (nd == 1 ? d0 : (nd == 2 ? d1 : d2)) : d0;   // most significant digit
    const u32 letter = minus ? b0 : (nd == 1 ? b1 : (nd == 2 ? b2 : b3));
    if (minus) { te = cp; ts = cp - nd - 1u; cp = ts; }
    else { ts = cp; te = cp + nd + 1u; cp = te; }
    kc = letter - '=';
    x = v;
    return nd != 0 && !(nd > 1 && lead == 0) && v != 0 && kc < 28u && ((kOpMask >> kc) & 1u);
#endif
}

// Numeric column at p: plain digits (<= 9 of them) or '*' (-> -1, string_to_int gafkluge.hpp:30), then a
// tab.  Up to three digits are decoded without a loop.  Returns false if not canonical.
__device__ __forceinline__ bool rec_num(const u8* rt, u32& p, i32& v) {
#if !G2P_REC_FAST_NUM
    if (rt[p] == '*') { v = -1; ++p; }
    else {
        u32 x = 0, d;
        const u32 p0 = p;
        while ((d = (u32)rt[p] - '0') <= 9u) { x = x * 10u + d; ++p; }
        if (p == p0 || p - p0 > 9) return false;
        v = (i32)x;
    }
    return rt[p++] == '\t';
#else
    const u32 c0 = rt[p], c1 = rt[p + 1], c2 = rt[p + 2], c3 = rt[p + 3];
    const u32 d0 = c0 - '0', d1 = c1 - '0', d2 = c2 - '0', d3 = c3 - '0';
    if (d0 <= 9u && d1 <= 9u && d2 <= 9u && d3 <= 9u) {   // >= 4 digits
        u32 x = 0, d;
        const u32 p0 = p;
        while ((d = (u32)rt[p] - '0') <= 9u) { x = x * 10u + d; ++p; }
        v = (i32)x;
        const bool ok = p - p0 <= 9 && rt[p] == '\t';
        ++p;
        return ok;
    }
    const u32 nd = d0 > 9u ? 0u : (d1 > 9u ? 1u : (d2 > 9u ? 2u : 3u));
    const u32 x = nd == 1 ? d0 : (nd == 2 ? d0 * 10u + d1 : d0 * 100u + d1 * 10u + d2);
    const u32 term = nd == 0 ? c1 : (nd == 1 ? c1 : (nd == 2 ? c2 : c3));
    const bool star = c0 == '*';
    v = star ? -1 : (i32)x;
    p += (star ? 1u : nd) + 1u;
    return (star || nd != 0) && term == '\t';
#endif
}

// The record walk after pass S: steps (from `stab`) and ops in normalised order ('-' records walk
// both backwards).  Returns false to delegate.
struct RecOut {
    u32 size, nlines;
};
__device__ __forceinline__ bool rec_walk(const ShortArgs& a, const u8* rt, const u32 r, const bool minus, const u32 rconst, const u32* p10,
                                         const bool prefixed, const u32* stab, const u32 ns, const u64 total, const u32 ca, const u32 cb,
                                         const i32 qs, i32 ps, i32 pe, RecOut& out) {
    if (minus) {   // flip_gaf: mirror the path interval about the summed step lengths (gaf2paf_main.cpp:128-131)
        if (total > 0x7fffffffULL) return false;
        const i32 nps = (i32)total - pe, npe = (i32)total - ps;
        ps = nps; pe = npe;
    }
    const i32 W = pe - ps;
    const u32 cend = minus ? ca : cb;   // CIGAR cursor and where it ends
    u32 cp = minus ? cb : ca;
    u32 rem = 0, remk = 0;              // unconsumed part of the op cut by the previous boundary
    u32 qcur = 0, tbc = 0;              // query / target bases consumed by the steps so far
    u32 size = 0, nlines = 0;
    for (u32 i = 0; i < ns; ++i) {
        const u32 idx = minus ? ns - 1 - i : i;
        const u32 sA = stab[2 * idx], sB = stab[2 * idx + 1];
        const u32 mp = sB & 0xffu, nl = (sB >> 8) & 0xffu;
        const i32 tlen = (i32)(sA & 0x7fffffffu);
        i32 sa = 0, se = tlen;
        if (sA & 0x80000000u) {   // interval: the digits were validated by pass S
            u32 k = mp + nl + 2, x = 0, d;
            while ((d = (u32)rt[k] - '0') <= 9u) { x = x * 10u + d; ++k; }
            sa = (i32)x;
            ++k; x = 0;
            while ((d = (u32)rt[k] - '0') <= 9u) { x = x * 10u + d; ++k; }
            se = (i32)x;
        }
        const bool rev = (prefixed && rt[mp] == '<') != minus;
        const bool last = i + 1 == ns;
        const i32 slen = se - sa;
        // ---- quota (gaf2paf_main.cpp:176-182)
        const i32 so = i == 0 ? ps : 0;
        i32 quota = slen - so, eo = 0;
        if (last) { quota = W - (i32)tbc; eo = slen - so - quota; }
        if (so < 0 || quota < 0 || eo < 0) return false;
        if (quota > 0) {
            // ---- take `quota` target bases of CIGAR (cigar_next_by_target, gaf2paf_main.cpp:71-90)
            u32 need = (u32)quota, q = 0, nm = 0, nb = 0;
            LineStep L;
            L.lenS = 0; L.codeS = 0; L.mid_a = 0; L.mid_b = 0; L.lenE = 0; L.codeE = 0;
            bool done = false;
            if (rem) {   // the remainder of a cut op is target-consuming by construction
                const u32 take = rem < need ? rem : need;
                if ((kQueryMask >> remk) & 1u) q += take;
                if ((kMatchMask >> remk) & 1u) nm += take;
                nb += take;
                if (rem >= need) { L.lenE = need; L.codeE = (u8)(remk + '='); rem -= need; done = true; }
                else { L.lenS = rem; L.codeS = (u8)(remk + '='); need -= rem; rem = 0; }
            }
            while (!done) {
                if (cp == cend) return false;   // :80 assert: CIGAR shorter than the path
                u32 x, kc, ts, tte;
                if (!rec_fetch_op(rt, minus, cp, x, kc, ts, tte)) return false;
                const bool tgt = (kTargetMask >> kc) & 1u;
                if (tgt && x >= need) {
                    L.lenE = need; L.codeE = (u8)(kc + '=');
                    rem = x - need; remk = kc;
                    x = need;
                    done = true;
                } else {
                    if (tgt) need -= x;
                    if (L.mid_b == 0) { L.mid_a = ts; L.mid_b = tte; }
                    else if (minus) L.mid_a = ts;
                    else L.mid_b = tte;
                }
                if ((kQueryMask >> kc) & 1u) q += x;
                if ((kMatchMask >> kc) & 1u) nm += x;
                nb += x;
            }
            if (nm > 0) {   // gaf2paf_main.cpp:225
                if (nlines >= kSMaxLines) return false;
                L.rev = rev;
                L.mid_fwd = rev == minus;
                L.q0 = (u32)qs + qcur; L.q1 = L.q0 + q;
                L.name_a = mp + 1; L.nl = nl; L.tlen = (u32)tlen;
                L.ts = (u32)(sa + (rev ? eo : so)); L.te = (u32)(se - (rev ? so : eo));
                L.nm = nm; L.nb = nb;
                const u32 line = rconst + line_step_len(L, p10);
                store_line_desc(a.sdesc + (size_t)r * kSMaxLines + nlines, r, size, line, L);
                size += line;
                ++nlines;
            }
            qcur += q;
            tbc += (u32)quota;
        }
    }
    // the reference parses the whole CIGAR before anything else: what the path left over must be valid too
    while (cp != cend) {
        u32 x, kc, ts, tte;
        if (!rec_fetch_op(rt, minus, cp, x, kc, ts, tte)) return false;
    }
    out.size = size; out.nlines = nlines;
    return true;
}

__global__ void __launch_bounds__(kRThreads, G2P_REC_CTAS) k_rec(const RecArgs ra) {
    G2P_DYN_SMEM(smem);
    __shared__ u32 p10[10];
    const ShortArgs& a = ra.s;
    if (threadIdx.x < 10) {
        u32 v = 1;
        for (u32 i = 0; i < threadIdx.x; ++i) v *= 10u;
        p10[threadIdx.x] = v;
    }
    const u32 C = ra.chunks, SW = rec_slot_words(C);
    u32* slots = reinterpret_cast<u32*>(smem);
    const u32 r0 = blockIdx.x * (u32)kRThreads;
    // ---- order the CTA's records by length (counting sort on the byte count): record rl goes to slot
    // s_slot[rl], thread t converts the record of slot t.  The lanes of a warp then hold records of
    // similar length, hence with similar numbers of steps and ops and similar field widths, so the
    // trip counts of the walk's loops (a warp pays the maximum over its lanes) stay close to the mean.
    __shared__ u32 s_hist[kRThreads];
    __shared__ u16 s_slot[kRThreads], s_rec[kRThreads];
    __shared__ u32 s_wsum[kRThreads / 32];
#if G2P_REC_SORT
    if (ra.perm == nullptr) {
        const u32 t = threadIdx.x, r = r0 + t;
        s_hist[t] = 0;
        __syncthreads();
        u32 key = 255, pos = 0;
        if (r < a.nrec) {
            const u32 l = a.rec_start[r + 1] - a.rec_start[r];
            key = l < 255u ? l : 255u;
        }
        pos = atomicAdd(&s_hist[key], 1u);
        __syncthreads();
        // exclusive scan of the 256 bins
        const u32 h = s_hist[t];
        u32 inc = h;
        for (int d = 1; d < 32; d <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, inc, d); if ((t & 31u) >= (u32)d) inc += v; }
        if ((t & 31u) == 31u) s_wsum[t >> 5] = inc;
        __syncthreads();
        u32 base = 0;
        for (u32 w = 0; w < (t >> 5); ++w) base += s_wsum[w];
        __syncthreads();
        s_hist[t] = base + inc - h;
        __syncthreads();
        const u32 slot = s_hist[key] + pos;
        s_slot[t] = (u16)slot;
        s_rec[slot] = (u16)t;
    } else { s_slot[threadIdx.x] = (u16)threadIdx.x; s_rec[threadIdx.x] = (u16)threadIdx.x; }
    __syncthreads();
#else
    s_slot[threadIdx.x] = (u16)threadIdx.x; s_rec[threadIdx.x] = (u16)threadIdx.x;
    __syncthreads();
#endif
    // ---- stage: 8 lanes per record, 32 records per pass, 128-bit coalesced loads; the loads of two
    // passes (up to four vectors per thread) are issued before the first store, so that the
    // rec_start -> text -> shared-memory chains of the passes overlap
    {
        const u32 gl = threadIdx.x & (kRStage - 1), grp = threadIdx.x / kRStage;
        constexpr u32 kPer = kRThreads / kRStage;   // records per pass
#pragma unroll 1
        for (u32 pass = 0; pass < (u32)kRThreads / kPer; pass += G2P_REC_STAGE_PASSES) {
            uint4 v0[G2P_REC_STAGE_PASSES], v1[G2P_REC_STAGE_PASSES];
            u32 nchv[G2P_REC_STAGE_PASSES];
#pragma unroll
            for (int u = 0; u < G2P_REC_STAGE_PASSES; ++u) {
                const u32 rl = (pass + (u32)u) * kPer + grp;
                nchv[u] = 0;
                if (r0 + rl < a.nrec) {
                    const u32 r = ra.perm ? ra.perm[r0 + rl] : r0 + rl;
                    const u32 s = a.rec_start[r], e = a.rec_start[r + 1];
                    const u32 A = s & ~15u;
                    const u32 nch = (e - A + 15u) >> 4;   // the record and its '\n'
                    if (e - s > 1u && e - s - 1u <= kSLimit && nch <= C) {
                        nchv[u] = nch;
                        if (gl < nch) v0[u] = ldg_vec_guarded(a.gaf, (u64)A + 16u * gl, a.n);
                        if (gl + kRStage < nch) v1[u] = ldg_vec_guarded(a.gaf, (u64)A + 16u * (gl + kRStage), a.n);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < G2P_REC_STAGE_PASSES; ++u) {
                const u32 rl = (pass + (u32)u) * kPer + grp;
                u32* dst = slots + (size_t)s_slot[rl] * SW;
                if (gl < nchv[u]) { dst[4 * gl] = v0[u].x; dst[4 * gl + 1] = v0[u].y; dst[4 * gl + 2] = v0[u].z; dst[4 * gl + 3] = v0[u].w; }
                if (gl + kRStage < nchv[u]) {
                    const u32 c = gl + kRStage;
                    dst[4 * c] = v1[u].x; dst[4 * c + 1] = v1[u].y; dst[4 * c + 2] = v1[u].z; dst[4 * c + 3] = v1[u].w;
                }
            }
        }
    }
    __syncthreads();
    if (r0 + s_rec[threadIdx.x] >= a.nrec) return;
    const u32 r = ra.perm ? ra.perm[r0 + s_rec[threadIdx.x]] : r0 + s_rec[threadIdx.x];
    const u32 s = a.rec_start[r], e = a.rec_start[r + 1];
    const u32 len = e - s - 1, sh = s & 15u;
    u8* rt = reinterpret_cast<u8*>(slots + (size_t)threadIdx.x * SW) + sh;

    bool ok = false, skip = false;
    RecOut out;
    out.size = 0; out.nlines = 0;
    LineRec R;
    R.gi = 0; R.gi_n = 0; R.qn_b = 0; R.qlen = R.mapq = R.m = R.b = 0; R.tp_a = R.tp_b = R.rc_a = R.rc_b = 0;
    do {
        if (len == 0 || len > kSLimit || ((e - (s & ~15u) + 15u) >> 4) > C) break;   // (k_emit_lines stages kSLimit bytes per record)
        rt[len] = '\n';   // sentinel also for an unterminated last line
        if (rt[0] == '*') { skip = true; ok = true; break; }   // gaf2paf_main.cpp:360
        // ---- columns 1..12 (parse_gaf_record, gafkluge.hpp:84-183)
        u32 p = 0;
        u8 c;
        p = rec_scan_field(rt, 0);
        c = rt[p];
        if (c != '\t' || p == 0) break;
        R.qn_b = p;
        ++p;
        bool bad = false;
        i32 qs, qe, plen, ps, pe, mapq;
        if (!rec_num(rt, p, R.qlen)) break;
        if (!rec_num(rt, p, qs)) break;
        if (!rec_num(rt, p, qe)) break;
        c = rt[p];   // strand
        if ((c != '+' && c != '-') || rt[p + 1] != '\t') break;
        const bool minus = c == '-';
        p += 2;
        const u32 pa = p;   // path
        p = rec_scan_field(rt, p);
        if (rt[p] != '\t' || p == pa) break;
        const u32 pb = p;
        ++p;
        if (!rec_num(rt, p, plen)) break;
        if (!rec_num(rt, p, ps)) break;
        if (!rec_num(rt, p, pe)) break;
        if (!rec_num(rt, p, R.m)) break;
        if (!rec_num(rt, p, R.b)) break;
        if (!rec_num(rt, p, mapq)) break;
        (void)qe; (void)plen;
        R.mapq = mapq >= 255 ? -1 : mapq;   // gafkluge.hpp:176-183
        // ---- optional tags (gafkluge.hpp:185-202): XX:T:value, no duplicates
        u32 ca = 0, cb = 0;
        u32 ka = 0, kb = 0, kc_ = 0, kd = 0, ntags = 0;
        for (;;) {
            const u32 fa = p;
            const u8 c0 = rt[p], c1 = rt[p + 1];
            if (c0 == '\t' || c0 == '\n' || c0 == ':' || c1 == '\t' || c1 == '\n' || c1 == ':') { bad = true; break; }
            const u8 c3 = rt[p + 3];
            if (rt[p + 2] != ':' || c3 == '\t' || c3 == '\n' || c3 == ':' || rt[p + 4] != ':') { bad = true; break; }
            const u32 key = (u32)c0 | ((u32)c1 << 8);
            const u32 kk = key * 0x00010001u;
            if (haszero16(ka ^ kk) | haszero16(kb ^ kk) | haszero16(kc_ ^ kk) | haszero16(kd ^ kk)) { bad = true; break; }
            if (++ntags > kRMaxTags) { bad = true; break; }
            kd = (kd << 16) | (kc_ >> 16); kc_ = (kc_ << 16) | (kb >> 16); kb = (kb << 16) | (ka >> 16); ka = (ka << 16) | key;
            p = rec_scan_field(rt, p + 5);
            c = rt[p];
            if (key == ((u32)'c' | ((u32)'g' << 8))) { ca = fa + 5; cb = p; }
            else if (key == ((u32)'t' | ((u32)'p' << 8))) { R.tp_a = fa + 3; R.tp_b = p; }
            else if (key == ((u32)'r' | ((u32)'c' << 8))) { R.rc_a = fa + 3; R.rc_b = p; }
            if (c == '\n') break;
            ++p;
        }
        if (bad || p != len) break;
        if (cb == 0 || ca >= cb || qs < 0 || ps < 0 || pe < 0) break;
        const u8 pc0 = rt[pa];
        const bool prefixed = pc0 == '>' || pc0 == '<';
        if (!prefixed && pb - pa == 1 && pc0 == '*') break;   // empty path: left to the general kernel
        R.gi_n = gi_fast(R.m, R.b, R.gi);
        if (R.gi_n == 0 || !rec_desc_fits(R)) break;
        const u32 rconst = line_const_len(R, p10);
        // step table: the larger of the two dead regions of the slot, after the path column up to
        // "cg:Z" or after the CIGAR (rt[ca - 1] == ':' must survive: it stops the backward CIGAR walk)
        u32 ta = (sh + pb + 1 + 3) & ~3u, tb = sh + ca - 1;
        {
            const u32 ta2 = (sh + cb + 1 + 3) & ~3u, tb2 = sh + len;
            if (tb2 > ta2 && (tb <= ta || tb2 - ta2 > tb - ta)) { ta = ta2; tb = tb2; }
        }
        const u32 cap = tb > ta ? (tb - ta) >> 3 : 0u;
        u32* stab = reinterpret_cast<u32*>(rt - sh + ta);
        rt[pb] = '>';   // bounds the token scan of pass S
        u32 ns;
        u64 total;
        if (!rec_steps(a.T, rt, pa, pb, prefixed, stab, cap, ns, total)) break;
        ok = rec_walk(a, rt, r, minus, rconst, p10, prefixed, stab, ns, total, ca, cb, qs, ps, pe, out);
    } while (0);

    if (!ok) {
        a.status[r] = ST_OK;   // overwritten by k_long / the general kernel
        a.out_off[r] = 0;
        a.line_off[r] = 0;
        a.deleg_list[atomicAdd(a.n_deleg, 1u)] = r;
    } else {
        const bool fast = !skip && out.size != 0;
        a.status[r] = (skip ? (u32)ST_SKIP : (u32)ST_OK) | ST_F_FAST | (fast ? (u32)ST_F_DESC : 0u);
        a.out_off[r] = out.size;
        a.line_off[r] = fast ? out.nlines : 0u;
        if (fast) store_rec_desc(a.rdesc + r, R);
    }
}

}  // namespace g2p
