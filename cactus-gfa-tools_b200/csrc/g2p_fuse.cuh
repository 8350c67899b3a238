// g2p_fuse.cuh — the one-pass conversion kernel for short records.
//
// k_fuse reads the GAF text ONCE and writes the PAF text ONCE.  A CTA takes the next byte tile of the input
// (ticket counter), brings it into shared memory with one TMA bulk copy, and does everything for the records
// that START in the tile without touching global memory again except for the lengths-table probes:
//
//   A  tile load        128-bit coalesced loads of [tile - 16, tile + kFTile + kFTail) into shared memory (a TMA bulk copy with
//                       G2P_FUSE_TMA_LOAD=1)
//   B  line index       SWAR '\n' compare over 128-bit shared loads, block scan of the counts -> record starts
//                       (the getline loop, gaf2paf_main.cpp:357-363)
//   C  columns + tags   one thread per record, k_rec's scalar walk (parse_gaf_record, gafkluge.hpp:84-204);
//                       counts the path steps; block scan -> the record's line slots
//   D  steps + CIGAR    same thread: every step token is probed once in the lengths table, then steps and ops
//                       are walked in normalised order ('-' records backwards: flip_gaf, gaf2paf_main.cpp:92-131)
//                       with the streaming cut of cigar_next_by_target (gaf2paf_main.cpp:71-90); one 48-byte
//                       slot per step holds the numbers of its PAF line (gaf2paf_main.cpp:157-263) and its length
//   H  offsets          block scan of the line lengths; ONE decoupled look-back over the tiles' output sizes gives
//                       the tile's offset in the output
//   I  format + store   one thread per PAF line formats its line left to right with a 32-bit word accumulator
//                       (numbers of < 1000 and their separator are one append) into a staging buffer that is
//                       co-aligned with the output, which then leaves with one TMA bulk store per round
//
// Nothing else is written: no record index, no line descriptors, no per-record sizes, no scans in global
// memory.  The kernel converts only canonical records (same definition as k_rec; additionally the columns it
// copies verbatim -- query length, matches, block length -- must be plain decimals without leading zeros) of at
// most kFLimit bytes.  Anything else sets FuseMeta::fallback: the host then runs the general pipeline
// (k_rec / k_long / k_convert_list + scans + k_emit_lines), which owns every error path of the reference.
#pragma once
#include "g2p_rec.cuh"

namespace g2p {

#ifndef G2P_FUSE_TMA_LOAD
#define G2P_FUSE_TMA_LOAD 0
#endif
constexpr int kFThreads = 256;
constexpr u32 kFLimit = 1000;              // longest record (bytes, without '\n') converted here
constexpr u32 kFTail = 1024;               // bytes after the tile its last record may extend into (>= kFLimit + 1)
// Shared-memory layout of one configuration: TILE input bytes per CTA, room for MAXREC records and MAXSLOTS path
// steps, CTAS resident CTAs per SM (which fixes the shared-memory budget; what the tables leave is the staging buffer).
template <u32 TILE, u32 MAXREC, u32 MAXSLOTS, int CTAS, bool DIRECT = false>
struct FuseCfg {
    static constexpr u32 kTile = TILE, kMaxRec = MAXREC, kMaxSlots = MAXSLOTS;
    static constexpr int kCtas = CTAS;
    static constexpr bool kDirect = DIRECT;   // lines are written straight to global memory (32-bit words), no staging buffer / TMA store
    static constexpr u32 kUnits = TILE / 16 / kFThreads;   // 16-byte vectors per thread in the newline scan
    static_assert(TILE % (16 * kFThreads) == 0, "the newline scan gives every thread whole vectors");
    static_assert(16 + TILE + kFTail + 64 < 65536, "text positions are 16-bit");
    static constexpr u32 kBudget = (227u * 1024u) / CTAS - 1024u - 512u;              // per CTA: 1 KB reserved by the driver + static shared
    static constexpr u32 kTextBytes = 16 + TILE + kFTail + 48;
    static constexpr u32 kOffStart = kTextBytes;                                      // u16 start[MAXREC + 2]
    static constexpr u32 kOffLidx = (kOffStart + 2 * (MAXREC + 2) + 3) & ~3u;         // u16 lidx[MAXSLOTS]: slots that print a line, in order
    static constexpr u32 kOffInfo = (kOffLidx + 2 * MAXSLOTS + 15) & ~15u;            // uint4 info[2 * MAXREC]
    static constexpr u32 kOffSlots = kOffInfo + 32 * MAXREC;                          // uint4 slots[3 * MAXSLOTS]
    static constexpr u32 kOffOff = kOffSlots + 48 * MAXSLOTS;                         // u32 off[MAXSLOTS + 1]
    static constexpr u32 kOffStage = (kOffOff + 4 * (MAXSLOTS + 1) + 15) & ~15u;
    static_assert(DIRECT ? kBudget >= kOffStage + 64 : kBudget > kOffStage + 8192, "no room for the staging buffer");
    static constexpr u32 kStage = DIRECT ? 32u : ((kBudget - kOffStage - 32) & ~15u);   // staged PAF bytes per round
    static constexpr size_t kSmem = kOffStage + kStage + 32;
};
// Configurations.  0-2 and 5 are tuning alternatives for ordinary short-read input (G2P_FUSE_CFG); 3 and 4 keep the largest
// tables on smaller tiles: the host moves to them when a tile reports kFuseTooManyRecords / kFuseTooManySteps
// (denser input: shorter records or more steps per byte).
typedef FuseCfg<32768, 352, 832, 2> FuseCfg0;
typedef FuseCfg<24576, 288, 704, 2> FuseCfg1;
typedef FuseCfg<16384, 192, 480, 3> FuseCfg2;
typedef FuseCfg<16384, 352, 832, 2> FuseCfg3;
typedef FuseCfg<8192, 352, 832, 2> FuseCfg4;
typedef FuseCfg<28672, 288, 736, 2> FuseCfg5;
typedef FuseCfg<24576, 288, 704, 3, true> FuseCfg6;
typedef FuseCfg<32768, 352, 832, 2, true> FuseCfg7;
typedef FuseCfg<16384, 192, 480, 4, true> FuseCfg8;
constexpr int kFuseCfgs = 9, kFuseCfgDense = 3;
static inline u32 fuse_cfg_tile(int cfg) {
    static const u32 t[kFuseCfgs] = {FuseCfg0::kTile, FuseCfg1::kTile, FuseCfg2::kTile, FuseCfg3::kTile, FuseCfg4::kTile, FuseCfg5::kTile, FuseCfg6::kTile, FuseCfg7::kTile, FuseCfg8::kTile};
    return t[cfg];
}
static inline size_t fuse_cfg_smem(int cfg) {
    static const size_t t[kFuseCfgs] = {FuseCfg0::kSmem, FuseCfg1::kSmem, FuseCfg2::kSmem, FuseCfg3::kSmem, FuseCfg4::kSmem, FuseCfg5::kSmem, FuseCfg6::kSmem, FuseCfg7::kSmem, FuseCfg8::kSmem};
    return t[cfg];
}
// the configuration to try after `cfg` reported a capacity overflow (-1: none left)
static inline int fuse_cfg_denser(int cfg) { return cfg == 3 ? 4 : (cfg == 4 ? -1 : kFuseCfgDense); }

enum : u32 { kFuseNotConvertible = 1u, kFuseTooManyRecords = 2u, kFuseTooManySteps = 4u,
             kFuseTimeoutLoad = 16u, kFuseTimeoutLookback = 32u };   // a spin wait gave up (never expected; reported, the general pipeline takes over)

struct FuseMeta {
    u32 n_records;    // records seen (sum over the tiles)
    u32 fallback;     // kFuse* reasons (or-ed): the result is void; capacity reasons alone -> run again with a smaller tile, else the general pipeline
    u32 overflow;     // the output did not fit out_cap (out_total is still exact): grow and run again
    u32 n_lines;      // PAF lines written
    u64 out_total;    // bytes of PAF
    u64 pad;
};

struct FuseArgs {
    const u8* gaf;
    u64 n;
    u32 ntiles;
    LenTableView T;
    u8* out;
    u64 out_cap;
    u64* tile_status;   // look-back words, zeroed before the launch (flag in the top two bits, kIdxFlagAgg / kIdxFlagPre)
    u32* ticket;        // zeroed before the launch
    FuseMeta* meta;     // zeroed before the launch
};

// ---- per-record constants of the lines, 32 bytes in shared memory ---------------------------
//   w0: start (text position of the record) | pfx_len << 16   ("qname\tqlen\t" is copied verbatim)
//   w1: m_a | m_len << 16 | b_len << 24                        (columns 10 / 11, copied verbatim)
//   w2: b_a | (mapq + 1) << 16                                 (mapq -1 .. 254)
//   w3: tp_a | tp_len << 16          w4: rc_a | rc_len << 16   ("type:value" spans, len 0 = absent)
//   w5: gi (0 .. 1000: floor(m / b * 1000 + 0.5))     w6: bytes of a line of this record that do not depend on the step
struct FRec {
    u32 start, pfx_len, m_a, m_len, b_a, b_len, tp_a, tp_len, rc_a, rc_len, gi, rconst;
    i32 mapq;
};
__device__ __forceinline__ void frec_store(uint4* dst, const FRec& r) {
    dst[0] = make_uint4(r.start | (r.pfx_len << 16), r.m_a | (r.m_len << 16) | (r.b_len << 24), r.b_a | ((u32)(r.mapq + 1) << 16), r.tp_a | (r.tp_len << 16));
    dst[1] = make_uint4(r.rc_a | (r.rc_len << 16), r.gi, r.rconst, 0u);
}
__device__ __forceinline__ void frec_load(const uint4* src, FRec& r) {
    const uint4 a = src[0], b = src[1];
    r.start = a.x & 0xffffu; r.pfx_len = a.x >> 16;
    r.m_a = a.y & 0xffffu; r.m_len = (a.y >> 16) & 0xffu; r.b_len = a.y >> 24;
    r.b_a = a.z & 0xffffu; r.mapq = (i32)(a.z >> 16) - 1;
    r.tp_a = a.w & 0xffffu; r.tp_len = a.w >> 16;
    r.rc_a = b.x & 0xffffu; r.rc_len = b.x >> 16;
    r.gi = b.y; r.rconst = b.z;
}

// ---- one line slot (48 bytes) ------------------------------------------------------------------
//   v0: tlen, ts, te, q0        v1: q1, nm, nb, lenS
//   v2: lenE, mid_a | mid_len << 16, name_pos | nl << 16 | flags << 24, rec | codeS << 16 | codeE << 24
// Before the walk v0.x holds tlen | interval << 31 and v2.z the step token (marker position, name length).
constexpr u32 kFSlotRev = 1u, kFSlotMidFwd = 2u;

// ---- D1: the path column -> one slot per step token, in normalised order (step j of the text becomes slot
// (minus ? ns - 1 - j : j): flip_gaf, gaf2paf_main.cpp:92-110, is index arithmetic).  Only positions are recorded
// (v2.z = marker position | name length << 16, v0.x = interval flag << 31); the table probe and the interval digits
// are left to one thread per slot (fuse_probe).  The caller has planted a '>' after the path column.
// CONVERGENCE: these per-record functions are called by ALL lanes of a warp (`live` = the lane has a record) and
// never leave a loop by `return`: an early return inside divergent code moves the reconvergence point of the
// compiler's BSSY / BSYNC pairs to the end of the function, after which the lanes of a warp run one by one (ncu:
// 3 of 32 lanes active in the op walk).  Failures set a flag; every data-dependent loop runs `while any lane is in
// it` (__any_sync is also the reconvergence point of the trip).
__device__ __forceinline__ bool fuse_tokens(const bool live, const u8* rt, const u32 rtpos, const u32 pa, const u32 pb, const bool prefixed,
                                            const bool minus, uint4* slots, const u32 ns) {
    const u32 FULL = 0xffffffffu;
    u32 j = 0, mp = prefixed ? pa : pa - 1;
    bool going = live, ok = true;
    while (__any_sync(FULL, going)) {
        if (going) {
            const u32 name_a = mp + 1;
            u32 e = pb;
            bool interval = false;
            if (prefixed) {
                e = rec_scan_step(rt, name_a);   // rt[pb] == '>'
                interval = rt[e] == ':';
            }
            const u32 nl = e - name_a;
            if (interval) e = rec_scan_step(rt, e + 1);   // the token ends at the next marker
            // (a second ':' inside the token is left to the general kernel)
            if (nl == 0 || nl > 255 || j >= ns || rt[e] == ':') { ok = false; going = false; }
            else {
                uint4* sl = slots + 3u * (minus ? ns - 1u - j : j);
                sl[0].x = interval ? 0x80000000u : 0u;
                sl[2].z = (rtpos + mp) | (nl << 16);
                ++j;
                if (e >= pb) going = false;
                mp = e;
            }
        }
    }
    return ok && (!live || j == ns);
}

// ---- E: one thread per slot: the step's name is probed once in the lengths table (gaf2paf_main.cpp:162-167) and an
// interval's ":start-end" (gafkluge.hpp:131-146; plain digits only) is decoded.  v0 = {tlen | interval << 31, start, end, 0}.
__device__ __forceinline__ bool fuse_probe(const LenTableView& T, const u8* text, uint4* sl) {
    const u32 z = sl[2].z;
    const u8* name = text + (z & 0xffffu) + 1u;
    const u32 nl = (z >> 16) & 0xffu;
    i64 tl64;
    if (nl <= 16u) {
        u32 w0, w1, w2, w3;
        lds16_unaligned(name, w0, w1, w2, w3);
        w0 = keep_bytes(w0, (int)nl); w1 = keep_bytes(w1, (int)nl - 4);
        w2 = keep_bytes(w2, (int)nl - 8); w3 = keep_bytes(w3, (int)nl - 12);
        if (!table_lookup_key16(T, (u64)w0 | ((u64)w1 << 32), (u64)w2 | ((u64)w3 << 32), nl, tl64)) return false;
    } else if (!table_lookup(T, name, nl, tl64)) return false;   // long names: 128-bit hash + arena compare
    if (tl64 < 0 || tl64 > 0x7fffffffLL) return false;
    u32 sa = 0, se = (u32)tl64;
    const u32 ivl = sl[0].x & 0x80000000u;
    if (ivl) {
        const u8* q = name + nl + 1u;
        u32 k = 0, x = 0, d;
        while ((d = (u32)q[k] - '0') <= 9u) { x = x * 10u + d; ++k; }
        if (k == 0 || k > 9 || q[k] != '-') return false;
        sa = x;
        const u32 k2 = ++k;
        x = 0;
        while ((d = (u32)q[k] - '0') <= 9u) { x = x * 10u + d; ++k; }
        if (k == k2 || k - k2 > 9 || (q[k] != '>' && q[k] != '<') || x < sa) return false;
        se = x;
    }
    sl[0] = make_uint4((u32)tl64 | ivl, sa, se, 0u);
    return true;
}

// First position >= p holding a tab or '\n' (the record's own newline, or the virtual one, bounds the scan): aligned
// 32-bit words, SWAR compare, four bytes per trip.
__device__ __forceinline__ u32 fuse_scan_field(const u8* rt, const u32 p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(rt + p);
    const u32* q = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
    const u32 off = (u32)(a & 3u);
    u32 w = *q | ((1u << (8u * off)) - 1u);   // the bytes before p read as 0xFF
    u32 base = p - off;
    for (;;) {
        const u32 m = zero_bytes(w ^ 0x09090909u) | zero_bytes(w ^ 0x0A0A0A0Au);
        if (m) return base + (((u32)__ffs((int)m) - 1u) >> 3);
        base += 4u;
        w = *++q;
    }
}

// One CIGAR token in walk direction, loop-free for up to four digits (all of a short read's).  Forward: "digits
// letter" starts at cp; backward: the token ends at cp (exclusive).  One unaligned 32-bit window of the text holds the
// digits either way -- forward the four bytes at cp (digits first), backward the four bytes before the letter (digits
// last); SWAR finds how many of them are digits, the window is shifted so that the last digit sits in the top byte,
// and the value is four multiply-adds.  Tokens of five or more digits take the byte loop of k_rec.
__device__ __forceinline__ u32 lds32_unaligned(const u8* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const u32* q = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
    return __funnelshift_r(q[0], q[1], (u32)(a & 3u) * 8u);
}
__device__ __forceinline__ bool fuse_fetch_op(const u8* rt, const bool minus, u32& cp, u32& x, u32& kc, u32& ts, u32& te) {
    const u32 w = lds32_unaligned(rt + (minus ? cp - 5u : cp));   // (the bytes before "cg:Z:" belong to the record: always readable)
    const u32 nd_mask = nondigit_bytes(w);                        // 0x80 in every byte that is not '0'..'9'
    // forward: digits are the leading bytes; backward: the trailing ones
    const u32 nd = minus ? ((u32)__clz((int)nd_mask) >> 3) : (nd_mask ? ((u32)__ffs((int)nd_mask) - 1u) >> 3 : 4u);
    const u32 letter = rt[minus ? cp - 1u : cp + nd];
    if (nd == 4u && (minus ? (u32)rt[cp - 6u] - '0' <= 9u : letter - '0' <= 9u)) {   // >= 5 digits (rare)
        u32 cp2 = cp, x2, kc2, ts2, te2;
        const bool r = rec_fetch_op_slow(rt, minus, cp2, x2, kc2, ts2, te2);
        cp = cp2; x = x2; kc = kc2; ts = ts2; te = te2;
        return r;
    }
    // last digit in the top byte, zeros below the first digit
    const u32 al = nd ? (minus ? w >> (8u * (4u - nd)) << (8u * (4u - nd)) : w << (8u * (4u - nd))) : 0u;
    const u32 v = ((al >> 24) & 15u) + 10u * ((al >> 16) & 15u) + 100u * ((al >> 8) & 15u) + 1000u * (al & 15u);
    const u32 lead = nd ? (al >> (8u * (4u - nd))) & 0xffu : 0u;   // most significant digit (as a character)
    if (minus) { te = cp; ts = cp - nd - 1u; cp = ts; }
    else { ts = cp; te = cp + nd + 1u; cp = te; }
    kc = letter - '=';
    x = v;
    return nd != 0 && !(nd > 1 && lead == '0') && v != 0 && kc < 28u && ((kOpMask >> kc) & 1u);
}

// ---- D2: the record walk, one thread per record.  (a) per step: strand, quota and clips (gaf2paf_main.cpp:157-182);
// (b) ONE loop over the CIGAR ops, in normalised order, that opens / closes the steps as their target quota fills up
// (cigar_next_by_target, gaf2paf_main.cpp:71-90): every trip fetches at most one op, so the lanes of a warp -- records
// with similar op counts -- stay in the same loop instead of diverging over nested step / op / digit loops.  Slot i
// (normalised step i) receives the numbers of its PAF line; kFSlotEmit marks the steps that print one.
constexpr u32 kFSlotEmit = 4u;
__device__ __forceinline__ bool fuse_walk2(const bool live, const u8* rt, const u32 rtpos, const u8* text, const u32 rec, const bool minus,
                                           const bool prefixed, uint4* slots, const u32 ns_in, const u32 ca, const u32 cb, const i32 qs, i32 ps, i32 pe) {
    const u32 FULL = 0xffffffffu;
    const u32 ns = live ? ns_in : 0u;
    bool ok = true;
    // (a) step lengths, the mirrored path interval of '-' records (flip_gaf, gaf2paf_main.cpp:111-131), quotas
    u64 total = 0;
    for (u32 i = 0; __any_sync(FULL, i < ns); ++i)
        if (i < ns) { const uint4 v = slots[3u * i]; total += v.z - v.y; }
    if (minus) {
        if (total > 0x7fffffffULL) ok = false;
        const i32 nps = (i32)total - pe, npe = (i32)total - ps;
        ps = nps; pe = npe;
    }
    const i32 W = pe - ps;
    u32 tbc = 0;
    for (u32 i = 0; __any_sync(FULL, i < ns); ++i) {
        if (i < ns) {
            uint4* sl = slots + 3u * i;
            const uint4 v = sl[0];
            const u32 z = sl[2].z;
            const i32 tlen = (i32)(v.x & 0x7fffffffu), sa = (i32)v.y, se = (i32)v.z;
            const bool rev = (prefixed && text[z & 0xffffu] == '<') != minus;
            const i32 slen = se - sa;
            const i32 so = i == 0 ? ps : 0;
            i32 quota = slen - so, eo = 0;
            if (i + 1 == ns) { quota = W - (i32)tbc; eo = slen - so - quota; }
            if (so < 0 || quota < 0 || eo < 0) { ok = false; quota = 0; }
            tbc += (u32)quota;
            sl[0] = make_uint4((u32)tlen, (u32)(sa + (rev ? eo : so)), (u32)(se - (rev ? so : eo)), 0u);
            sl[1].x = (u32)quota;
            sl[2].z = ((z & 0xffffu) + 1u) | (z & 0x00ff0000u) | ((rev ? kFSlotRev : 0u) << 24) | ((rev == minus ? kFSlotMidFwd : 0u) << 24);
        }
    }
    // (b) ops
    const u32 cend = minus ? ca : cb;
    u32 cp = minus ? cb : ca;
    u32 rem = 0, remk = 0;       // unconsumed part of the op cut by the previous boundary
    u32 qcur = (u32)qs;          // query position at the start of the open step
    u32 i = 0, need = 0;         // open step and what is left of its quota (0: no step open)
    u32 q = 0, nm = 0, nb = 0, lenS = 0, codeS = 0, lenE = 0, codeE = 0, mid_a = 0, mid_b = 0;
    bool going = live && ok;
    while (__any_sync(FULL, going)) {
        if (going) {
            bool fetch = true, close = false;
            if (need == 0) {   // open the next step
                fetch = false;
                if (i == ns) going = false;
                else {
                    need = slots[3u * i + 1u].x;
                    q = nm = nb = lenS = codeS = lenE = codeE = mid_a = mid_b = 0;
                    if (need == 0) ++i;   // no target bases left for it: no line, nothing consumed
                    else {
                        fetch = true;
                        if (rem) {   // the remainder of a cut op is target-consuming by construction
                            const u32 take = rem < need ? rem : need;
                            if ((kQueryMask >> remk) & 1u) q += take;
                            if ((kMatchMask >> remk) & 1u) nm += take;
                            nb += take;
                            if (rem >= need) { lenE = need; codeE = remk + '='; rem -= need; close = true; fetch = false; }
                            else { lenS = rem; codeS = remk + '='; need -= rem; rem = 0; }
                        }
                    }
                }
            }
            if (fetch) {   // one op
                u32 x = 0, kc = 0, ts = 0, tte = 0;
                // (:80 assert: CIGAR shorter than the path)
                if (cp == cend || !fuse_fetch_op(rt, minus, cp, x, kc, ts, tte)) { ok = false; going = false; }
                else {
                    const bool tgt = (kTargetMask >> kc) & 1u;
                    if (tgt && x >= need) {
                        lenE = need; codeE = kc + '=';
                        rem = x - need; remk = kc;
                        x = need;
                        close = true;
                    } else {
                        if (tgt) need -= x;
                        if (mid_b == 0) { mid_a = ts; mid_b = tte; }
                        else if (minus) mid_a = ts;
                        else mid_b = tte;
                    }
                    if ((kQueryMask >> kc) & 1u) q += x;
                    if ((kMatchMask >> kc) & 1u) nm += x;
                    nb += x;
                }
            }
            if (close) {
                uint4* sl = slots + 3u * i;
                if (nm > 0) {   // gaf2paf_main.cpp:225
                    const u32 mid_len = mid_b > mid_a ? mid_b - mid_a : 0u;
                    sl[0].w = qcur;
                    sl[1] = make_uint4(qcur + q, nm, nb, lenS);
                    sl[2].x = lenE;
                    sl[2].y = (rtpos + mid_a) | (mid_len << 16);
                    sl[2].z |= kFSlotEmit << 24;
                    sl[2].w = rec | (codeS << 16) | (codeE << 24);
                }
                qcur += q;
                need = 0;
                ++i;
            }
        }
    }
    // the reference parses the whole CIGAR before anything else: what the path left over must be valid too
    going = live && ok && cp != cend;
    while (__any_sync(FULL, going)) {
        if (going) {
            u32 x, kc, ts, tte;
            if (!fuse_fetch_op(rt, minus, cp, x, kc, ts, tte)) { ok = false; going = false; }
            else if (cp == cend) going = false;
        }
    }
    return ok;
}

// ---- G: one thread per slot: the byte length of its PAF line (0: the step prints none)
__device__ __forceinline__ u32 fuse_line_len(const uint4* sl, const uint4* info, const u32* p10) {
    const uint4 v2 = sl[2];
    if (!((v2.z >> 24) & kFSlotEmit)) return 0u;
    const uint4 v0 = sl[0], v1 = sl[1];
    const u32 rconst = info[2u * (v2.w & 0xffffu) + 1u].z;
    const u32 codeS = (v2.w >> 16) & 0xffu;
    u32 n = ((v2.z >> 16) & 0xffu) + dlen_u32(v0.w, p10) + dlen_u32(v1.x, p10) + dlen_u32(v0.x, p10) + dlen_u32(v0.y, p10) + dlen_u32(v0.z, p10) +
            dlen_u32(v1.y, p10) + dlen_u32(v1.z, p10) + dlen_u32(v2.x, p10) + 1u;
    if (codeS) n += dlen_u32(v1.w, p10) + 1u;
    return rconst + n + (v2.y >> 16);
}

// One PAF line (paf.hpp:83-95 + gaf2paf_main.cpp:228-256), left to right, into `dst` (any alignment) .. dst + len.
__device__ __forceinline__ void fuse_write_line(u8* dst, const u32 len, const u8* text, const FRec& R, const uint4 v0, const uint4 v1, const uint4 v2) {
    const u8* rt = text + R.start;
    // head: the bytes up to the first word boundary go out one by one (the word is shared with the previous line)
    const u32 head = (4u - ((u32)reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
    for (u32 i = 0; i < head; ++i) dst[i] = rt[i];
    WEmit E;
    E.wp = reinterpret_cast<u32*>(dst + head);
    E.lo = 0; E.sh = 0;
    E.copy(rt + head, R.pfx_len - head);                       // qname \t qlen \t   (>= 4 bytes)
    E.num(v0.w, '\t');                                         // query start
    E.num(v1.x, '\t');                                         // query end
    const u32 flags = v2.z >> 24;
    const bool rev = (flags & kFSlotRev) != 0;
    E.put((rev ? (u32)'-' : (u32)'+') | ((u32)'\t' << 8), 2u);
    {                                                          // target name
        const u32 nl = (v2.z >> 16) & 0xffu;
        if (nl <= 8u) E.copy_small(text + (v2.z & 0xffffu), nl); else E.copy(text + (v2.z & 0xffffu), nl);
    }
    E.put('\t', 1u);
    E.num(v0.x, '\t');                                         // target length
    E.num(v0.y, '\t');                                         // target start
    E.num(v0.z, '\t');                                         // target end
    E.num(v1.y, '\t');                                         // matches
    E.num(v1.z, '\t');                                         // block length
    if (R.mapq < 0) E.put((u32)'-' | ((u32)'1' << 8) | ((u32)'\t' << 16), 3u);
    else E.num((u32)R.mapq, '\t');
    // (every item below ends with the tab that separates it from the next one)
    if (R.tp_len) {
        E.put((u32)'t' | ((u32)'p' << 8) | ((u32)':' << 16), 3u);
        if (R.tp_len <= 8u) E.copy_small(rt + R.tp_a, R.tp_len); else E.copy(rt + R.tp_a, R.tp_len);
        E.put('\t', 1u);
    }
    if (R.rc_len) {
        E.put((u32)'r' | ((u32)'c' << 8) | ((u32)':' << 16), 3u);
        E.copy(rt + R.rc_a, R.rc_len);
        E.put('\t', 1u);
    }
    E.put4((u32)'g' | ((u32)'m' << 8) | ((u32)':' << 16) | ((u32)'i' << 24));
    E.put(':', 1u);
    if (R.m_len <= 8u) E.copy_small(rt + R.m_a, R.m_len); else E.copy(rt + R.m_a, R.m_len);
    E.put4((u32)'\t' | ((u32)'g' << 8) | ((u32)'l' << 16) | ((u32)':' << 24));
    E.put((u32)'i' | ((u32)':' << 8), 2u);
    if (R.b_len <= 8u) E.copy_small(rt + R.b_a, R.b_len); else E.copy(rt + R.b_a, R.b_len);
    E.put4((u32)'\t' | ((u32)'g' << 8) | ((u32)'i' << 16) | ((u32)':' << 24));
    E.put((u32)'f' | ((u32)':' << 8), 2u);
    if (R.gi == 0u || R.gi == 1000u) E.put(R.gi ? (u32)'1' : (u32)'0', 1u);
    else {   // "0." + three decimals, trailing zeros dropped (printf("%g"), gaf2paf_main.cpp:248-253)
        const u32 d0 = R.gi / 100u, r = R.gi - d0 * 100u, d1 = r / 10u, d2 = r - d1 * 10u;
        const bool more = (d1 | d2) != 0;
        E.put((u32)'0' | ((u32)'.' << 8) | ((u32)('0' + d0) << 16) | (more ? (u32)('0' + d1) << 24 : 0u), more ? 4u : 3u);
        if (d2) E.put('0' + d2, 1u);
    }
    E.put4((u32)'\t' | ((u32)'c' << 8) | ((u32)'g' << 16) | ((u32)':' << 24));
    E.put((u32)'Z' | ((u32)':' << 8), 2u);
    // CIGAR pieces, reversed for '<' steps (gaf2paf_main.cpp:184-211)
    const u32 codeS = (v2.w >> 16) & 0xffu, codeE = v2.w >> 24;
    if (rev) E.num(v2.x, codeE);
    else if (codeS) E.num(v1.w, codeS);
    const u32 mid_len = v2.y >> 16;
    if (mid_len) {
        const u8* mid = text + (v2.y & 0xffffu);
        if (flags & kFSlotMidFwd) E.copy(mid, mid_len);
        else {   // tokens in reverse order: walk the span backwards, one "digits letter" token at a time
            u32 e = mid_len;
            while (e) {
                u32 s = e - 1;
                while (s > 0 && mid[s - 1] <= '9') --s;   // the previous token's letter (> '9') ends the digits
                E.copy(mid + s, e - s);
                e = s;
            }
        }
    }
    if (rev) { if (codeS) E.num(v1.w, codeS); }
    else E.num(v2.x, codeE);
    E.put('\n', 1u);
    // tail: pending bytes (< 4) one by one
    u8* tp = reinterpret_cast<u8*>(E.wp);
    for (u32 i = 0; 8u * i < E.sh; ++i) tp[i] = (u8)(E.lo >> (8u * i));
    (void)len;
}

template <class C>
__global__ void __launch_bounds__(kFThreads, C::kCtas) k_fuse(const FuseArgs a) {
    constexpr u32 kFTile = C::kTile, kFUnits = C::kUnits, kFStage = C::kStage, kFMaxRec = C::kMaxRec, kFMaxSlots = C::kMaxSlots;
    G2P_DYN_SMEM(smem);
    __shared__ u32 p10[10];
    __shared__ u32 s_tile, s_flag, s_end;
    __shared__ u32 s_w[kFThreads / 32];
    __shared__ u64 s_obase;
#if G2P_FUSE_TMA_LOAD && !defined(G2P_HOSTSIM)
    __shared__ __align__(8) u64 s_bar;
#endif
    const u32 FULL = 0xffffffffu;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid < 10) {
        u32 v = 1;
        for (u32 i = 0; i < tid; ++i) v *= 10u;
        p10[tid] = v;
    }
    if (tid == 0) {
        s_tile = atomicAdd(a.ticket, 1u);
        s_flag = 0; s_end = 0xffffffffu;
#if G2P_FUSE_TMA_LOAD && !defined(G2P_HOSTSIM)
        mbar_init(&s_bar, 1);
        fence_mbar_init();
#endif
    }
    __syncthreads();
    const u32 tile = s_tile;
    volatile u64* st = a.tile_status;
    volatile u32* g_fallback = &a.meta->fallback;
    if (*g_fallback) {   // the result is void already: only keep the look-back chain alive
        if (tid == 0) st[tile] = kIdxFlagPre;
        return;
    }
    u8* text = smem;   // text[16 + p] = gaf[base + p]
    u16* s_start = reinterpret_cast<u16*>(smem + C::kOffStart);
    uint4* s_info = reinterpret_cast<uint4*>(smem + C::kOffInfo);
    uint4* s_slots = reinterpret_cast<uint4*>(smem + C::kOffSlots);
    u32* s_off = reinterpret_cast<u32*>(smem + C::kOffOff);
    u16* s_lidx = reinterpret_cast<u16*>(smem + C::kOffLidx);
    u8* s_stage = smem + C::kOffStage;

    // ---------------- A: tile load
    const u64 base = (u64)tile * kFTile;
    {
        const u64 src0 = tile ? base - 16 : 0;
        const u32 dst0 = tile ? 0u : 16u;
        const u64 room = (u64)(tile ? 16u : 0u) + kFTile + kFTail;
        const u32 avail = (u32)(a.n - src0 < room ? a.n - src0 : room);
        const u32 full = avail & ~15u;
#if G2P_FUSE_TMA_LOAD && !defined(G2P_HOSTSIM)
        if (tid == 0 && full) { mbar_expect_tx(&s_bar, full); bulk_g2s(text + dst0, a.gaf + src0, full, &s_bar); }
#else
        // 128-bit coalesced loads (source and destination are 16-byte aligned).  The TMA bulk copy this replaces
        // (G2P_FUSE_TMA_LOAD=1) is no faster here and, on B200 / driver 580, its mbarrier was seen not to complete
        // (rarely, in launches whose CTAs mostly leave at once: inputs with long records, concurrent streams) -- see DESIGN.md.
        for (u32 i = tid; i < (full >> 4); i += kFThreads)
            reinterpret_cast<uint4*>(text + dst0)[i] = __ldg(reinterpret_cast<const uint4*>(a.gaf + src0) + i);
#endif
        if (tid < avail - full) text[dst0 + full + tid] = a.gaf[src0 + full + tid];   // last partial vector
        if (tid < 32) text[dst0 + avail + tid] = '\n';   // virtual newline after an unterminated last line (kFTextBytes has the room)
        if (!tile && tid < 16) text[tid] = '\n';         // the first record starts at position 0
#if G2P_FUSE_TMA_LOAD && !defined(G2P_HOSTSIM)
        if (full) {
            u32 spins = 0;
            while (!mbar_try_wait(&s_bar, 0)) {
                if (++spins > (1u << 22)) { atomicOr(&a.meta->fallback, (u32)(kFuseNotConvertible | kFuseTimeoutLoad)); break; }
            }
        }
#endif
    }
    __syncthreads();

    // ---------------- B: line index (record starts inside the tile, in text positions relative to text + 16)
    const u32 limit = (u32)(a.n - base < (u64)kFTile ? a.n - base : (u64)kFTile);   // record starts are < limit
    u32 R;
    {
        u32 m[kFUnits][4];
        u32 cnt = 0;
#pragma unroll
        for (u32 k = 0; k < kFUnits; ++k) {
            const u32 q0 = 16u * (tid * kFUnits + k);
            const uint4 v = *reinterpret_cast<const uint4*>(text + 16 + q0);
            m[k][0] = nl_bits(v.x); m[k][1] = nl_bits(v.y); m[k][2] = nl_bits(v.z); m[k][3] = nl_bits(v.w);
            if (q0 + 17u > limit) {   // a newline at q starts a record at q + 1 only if q + 1 < limit
#pragma unroll
                for (u32 j = 0; j < 4; ++j)
                    for (u32 b = 0; b < 4; ++b)
                        if (q0 + 4u * j + b + 1u >= limit) m[k][j] &= ~(0x80u << (8u * b));
            }
            cnt += __popc(m[k][0]) + __popc(m[k][1]) + __popc(m[k][2]) + __popc(m[k][3]);
        }
        u32 incl = cnt;
        for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(FULL, incl, o); if (lane >= (u32)o) incl += up; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        u32 wpre = 0, total = 0;
        for (u32 i = 0; i < kFThreads / 32; ++i) { if (i < warp) wpre += s_w[i]; total += s_w[i]; }
        const u32 head = limit > 0 && text[15] == '\n' ? 1u : 0u;
        R = head + total;
        u32 rank = head + wpre + incl - cnt;
        if (R <= kFMaxRec) {
            if (tid == 0 && head) s_start[0] = 0;
#pragma unroll
            for (u32 k = 0; k < kFUnits; ++k) {
                const u32 q0 = 16u * (tid * kFUnits + k);
#pragma unroll
                for (u32 j = 0; j < 4; ++j) {
                    u32 mm = m[k][j];
                    while (mm) {
                        const u32 bit = (u32)__ffs((int)mm) - 1u;
                        mm &= mm - 1u;
                        s_start[rank++] = (u16)(q0 + 4u * j + (bit >> 3) + 1u);
                    }
                }
            }
        }
        // the end of the last record: the first newline at or after position limit - 1 (warp 0, one vector per lane and trip)
        if (warp == 0 && R > 0 && R <= kFMaxRec) {
            const u32 from = limit - 1u;
            u32 found = 0xffffffffu;
            for (u32 u0 = from >> 4; found == 0xffffffffu && 16u * u0 < from + kFLimit + 2u; u0 += 32) {
                const u32 q0 = 16u * (u0 + lane);
                u32 mine = 0xffffffffu;
                if (q0 < kFTile + kFTail + 16u) {
                    const uint4 v = *reinterpret_cast<const uint4*>(text + 16 + q0);
                    const u32 w[4] = {v.x, v.y, v.z, v.w};
                    for (int j = 3; j >= 0; --j) {
                        u32 mm = nl_bits(w[j]);
                        while (mm) {
                            const u32 bit = 31u - (u32)__clz((int)mm);
                            mm &= ~(1u << bit);
                            const u32 q = q0 + 4u * (u32)j + (bit >> 3);
                            if (q >= from) mine = q;   // descending scan: the last assignment is the smallest q >= from
                        }
                    }
                }
                for (int o = 16; o > 0; o >>= 1) { const u32 t2 = __shfl_xor_sync(FULL, mine, o); mine = t2 < mine ? t2 : mine; }
                found = mine;
            }
            if (lane == 0) s_end = found;
        }
    }
    __syncthreads();
    if (R > kFMaxRec) {   // too many records for the tile's tables
        if (tid == 0) { atomicOr(&a.meta->fallback, (u32)kFuseTooManyRecords); st[tile] = kIdxFlagPre; }
        return;
    }
    if (tid == 0 && R) s_start[R] = s_end == 0xffffffffu ? (u16)(s_start[R - 1] + kFLimit + 2u) : (u16)(s_end + 1u);
    __syncthreads();

    // ---------------- C + D: one thread per record (passes of kFThreads records; the line slots of a pass follow
    // those of the pass before)
    u32 nslots = 0;
    bool bad = false, too_many = false;
    for (u32 r0 = 0; r0 < R; r0 += kFThreads) {
        const u32 rec = r0 + tid;
        const bool have = rec < R;
        bool ok = !have, skip = false;
        u32 ns = 0;
        // state of the record between the two halves of the pass
        const u8* rt = text + 16;
        u32 rtpos = 16, len = 0, pa = 0, pb = 0, ca = 0, cb = 0;
        i32 qs = 0, ps = 0, pe = 0;
        bool minus = false, prefixed = false;
        if (have) do {
            const u32 s = s_start[rec], e = s_start[rec + 1];
            rtpos = 16u + s;
            rt = text + rtpos;
            len = e - s - 1u;
            if (len == 0 || len > kFLimit) break;
            if (rt[0] == '*') { skip = true; ok = true; break; }   // gaf2paf_main.cpp:360
            // ---- columns 1..12 (parse_gaf_record, gafkluge.hpp:84-183); the record's own '\n' (or the virtual one) ends every scan
            FRec F;
            u32 p = fuse_scan_field(rt, 0);
            u8 c = rt[p];
            if (c != '\t' || p == 0) break;
            ++p;
            i32 qlen, qe, plen, m, b, mapq;
            const u32 qlen_a = p;
            if (!rec_num(rt, p, qlen)) break;
            if (qlen < 0 || (rt[qlen_a] == '0' && p - qlen_a > 2)) break;   // copied verbatim: plain decimal, no leading zero
            F.pfx_len = p;
            if (!rec_num(rt, p, qs)) break;
            if (!rec_num(rt, p, qe)) break;
            c = rt[p];   // strand
            if ((c != '+' && c != '-') || rt[p + 1] != '\t') break;
            minus = c == '-';
            p += 2;
            pa = p;   // path
            p = fuse_scan_field(rt, p);
            if (rt[p] != '\t' || p == pa) break;
            pb = p;
            ++p;
            if (!rec_num(rt, p, plen)) break;
            if (!rec_num(rt, p, ps)) break;
            if (!rec_num(rt, p, pe)) break;
            F.m_a = p;
            if (!rec_num(rt, p, m)) break;
            F.m_len = p - 1u - F.m_a;
            if (m < 0 || (rt[F.m_a] == '0' && F.m_len > 1)) break;
            F.b_a = p;
            if (!rec_num(rt, p, b)) break;
            F.b_len = p - 1u - F.b_a;
            if (b < 0 || (rt[F.b_a] == '0' && F.b_len > 1)) break;
            if (!rec_num(rt, p, mapq)) break;
            (void)qe; (void)plen;
            F.mapq = mapq >= 255 ? -1 : mapq;   // gafkluge.hpp:176-183
            // ---- optional tags (gafkluge.hpp:185-202): XX:T:value, no duplicates
            u32 tp_a = 0, tp_b = 0, rc_a = 0, rc_b = 0;
            u32 ka = 0, kb = 0, kc_ = 0, kd = 0, ntags = 0;
            bool tbad = false;
            for (;;) {
                const u32 fa = p;
                const u8 c0 = rt[p], c1 = rt[p + 1];
                if (c0 == '\t' || c0 == '\n' || c0 == ':' || c1 == '\t' || c1 == '\n' || c1 == ':') { tbad = true; break; }
                const u8 c3 = rt[p + 3];
                if (rt[p + 2] != ':' || c3 == '\t' || c3 == '\n' || c3 == ':' || rt[p + 4] != ':') { tbad = true; break; }
                const u32 key = (u32)c0 | ((u32)c1 << 8);
                const u32 kk = key * 0x00010001u;
                if (haszero16(ka ^ kk) | haszero16(kb ^ kk) | haszero16(kc_ ^ kk) | haszero16(kd ^ kk)) { tbad = true; break; }
                if (++ntags > kRMaxTags) { tbad = true; break; }
                kd = (kd << 16) | (kc_ >> 16); kc_ = (kc_ << 16) | (kb >> 16); kb = (kb << 16) | (ka >> 16); ka = (ka << 16) | key;
                p = fuse_scan_field(rt, p + 5);
                c = rt[p];
                if (key == ((u32)'c' | ((u32)'g' << 8))) { ca = fa + 5; cb = p; }
                else if (key == ((u32)'t' | ((u32)'p' << 8))) { tp_a = fa + 3; tp_b = p; }
                else if (key == ((u32)'r' | ((u32)'c' << 8))) { rc_a = fa + 3; rc_b = p; }
                if (c == '\n') break;
                ++p;
            }
            if (tbad || p != len) break;
            if (cb == 0 || ca >= cb || qs < 0 || ps < 0 || pe < 0) break;
            const u8 pc0 = rt[pa];
            prefixed = pc0 == '>' || pc0 == '<';
            if (!prefixed && pb - pa == 1 && pc0 == '*') break;   // empty path: left to the general kernel
            // gi = floor(m / b * 1000 + 0.5) / 1000 printed with %g (gaf2paf_main.cpp:248-253).  In integers:
            // K = (2000 m + b) / (2 b).  The reference's three rounded double operations can only differ from it when
            // m / b * 1000 + 0.5 is an integer exactly (any other value is >= 1 / (2 b) > 2e-10 away from one, the rounding
            // error is < 4e-13): those ties, and large m, take the double path of k_rec.
            u32 gk, gi_n;
            if (b <= 0) { gk = 0; gi_n = 1; }
            else {
                bool exact = false;
                gk = 0;
                if (m < (1 << 20)) {
                    const u32 num = 2000u * (u32)m + (u32)b, den = 2u * (u32)b;
                    gk = num / den;
                    exact = num - gk * den != 0u;
                }
                if (!exact) {
                    u64 pack;
                    const u32 n1 = gi_fast(m, b, pack);
                    if (n1 == 0) break;
                    gk = n1 == 1 ? ((u32)(pack & 0xff) == '1' ? 1000u : 0u) : 100u * ((u32)(pack >> 16) & 0xfu) + 10u * ((u32)(pack >> 24) & 0xfu) + ((u32)(pack >> 32) & 0xfu);
                }
                if (gk > 1000u) break;
                const u32 r100 = gk % 100u;
                gi_n = gk == 0u || gk == 1000u ? 1u : (gk % 10u ? 5u : (r100 ? 4u : 3u));
            }
            F.gi = gk;
            // bytes of a line that do not depend on the step (line_const_len of k_rec; the verbatim columns by their text length)
            F.rconst = F.pfx_len - 2u + 12u + dlen_i32(F.mapq, p10) + (tp_b ? 4u + (tp_b - tp_a) : 0u) + (rc_b ? 4u + (rc_b - rc_a) : 0u) +
                       6u + F.m_len + 6u + F.b_len + 6u + gi_n + 6u + 1u;
            F.start = rtpos;
            F.tp_a = tp_a; F.tp_len = tp_b - tp_a; F.rc_a = rc_a; F.rc_len = rc_b - rc_a;
            frec_store(s_info + 2u * rec, F);
            // ---- count the path steps
            if (prefixed) {   // '<' = 0x3C, '>' = 0x3E: (byte & 0xFD) == 0x3C, four bytes per trip on aligned words
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(rt + pa), a1 = reinterpret_cast<uintptr_t>(rt + pb);
                const u32* q = reinterpret_cast<const u32*>(a0 & ~(uintptr_t)3);
                const u32* qe2 = reinterpret_cast<const u32*>((a1 + 3u) & ~(uintptr_t)3);
                const u32 head_skip = (u32)(a0 & 3u), tail_skip = (u32)((0u - (u32)a1) & 3u);
                for (const u32* w = q; w < qe2; ++w) {
                    u32 mk = zero_bytes((*w & 0xFDFDFDFDu) ^ 0x3C3C3C3Cu);
                    if (w == q) mk &= 0xffffffffu << (8u * head_skip);
                    if (w + 1 == qe2 && tail_skip) mk &= 0xffffffffu >> (8u * tail_skip);
                    ns += (u32)__popc(mk);
                }
            } else ns = 1;
            ok = true;
        } while (0);
        if (!ok) { ns = 0; bad = true; }
        // ---- block scan of the step counts -> first slot of the record
        u32 incl = ns;
        for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(FULL, incl, o); if (lane >= (u32)o) incl += up; }
        __syncthreads();   // (s_w is free again)
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        u32 wpre = 0, total = 0;
        for (u32 i = 0; i < kFThreads / 32; ++i) { if (i < warp) wpre += s_w[i]; total += s_w[i]; }
        const u32 slot0 = nslots + wpre + incl - ns;
        const u32 pass_slot0 = nslots;
        nslots += total;
        if (nslots > kFMaxSlots) { too_many = true; break; }   // uniform: every thread sees the same total
        const bool live = have && ok && !skip;
        // D1: step tokens -> slots (all lanes: see CONVERGENCE)
        if (live) const_cast<u8*>(rt)[pb] = '>';   // bounds the token scans
        if (!fuse_tokens(live, rt, rtpos, pa, pb, prefixed, minus, s_slots + 3u * slot0, ns)) bad = true;
        if (bad) s_flag = 1;
        __syncthreads();
        if (s_flag) break;   // uniform
        // E: one thread per step slot: table probe + interval digits
        for (u32 sidx = pass_slot0 + tid; sidx < nslots; sidx += kFThreads)
            if (!fuse_probe(a.T, text, s_slots + 3u * sidx)) s_flag = 1;
        __syncthreads();
        if (s_flag) break;   // uniform
        // D2: quotas + the op walk
        if (!fuse_walk2(live, rt, rtpos, text, rec, minus, prefixed, s_slots + 3u * slot0, ns, ca, cb, qs, ps, pe)) bad = true;
    }
    if (bad) s_flag = 1;
    __syncthreads();
    if (s_flag || too_many) {
        if (tid == 0) { atomicOr(&a.meta->fallback, s_flag ? (u32)kFuseNotConvertible : (u32)kFuseTooManySteps); st[tile] = kIdxFlagPre; }
        return;
    }

    // ---------------- G: one thread per slot: line lengths
    for (u32 sidx = tid; sidx < nslots; sidx += kFThreads) s_off[sidx] = fuse_line_len(s_slots + 3u * sidx, s_info, p10);
    __syncthreads();

    // ---------------- H: line lengths -> offsets inside the tile and the list of the slots that print a line (one
    // block scan over (lines << 32 | bytes)); then the tile's offset in the output (look-back)
    u32 tile_bytes, nlines;
    {
        const u32 per = (nslots + kFThreads - 1) / kFThreads;
        const u32 a0 = tid * per < nslots ? tid * per : nslots, a1 = a0 + per < nslots ? a0 + per : nslots;
        u64 sum = 0;
        for (u32 i = a0; i < a1; ++i) { const u32 l = s_off[i]; sum += (u64)l | ((u64)(l != 0) << 32); }
        u64 incl = sum;
        for (int o = 1; o < 32; o <<= 1) { const u64 up = __shfl_up_sync(FULL, incl, o); if (lane >= (u32)o) incl += up; }
        __shared__ u64 s_w64[kFThreads / 32];
        if (lane == 31) s_w64[warp] = incl;
        __syncthreads();
        u64 wpre = 0, total = 0;
        for (u32 i = 0; i < kFThreads / 32; ++i) { if (i < warp) wpre += s_w64[i]; total += s_w64[i]; }
        u64 run = wpre + incl - sum;
        u32 rb = (u32)run, rl = (u32)(run >> 32);
        // s_off[i] becomes the exclusive offset; the length stays recoverable as off[i + 1] - off[i] (off[nslots] = total)
        for (u32 i = a0; i < a1; ++i) {
            const u32 l = s_off[i];
            s_off[i] = rb;
            rb += l;
            if (l) s_lidx[rl++] = (u16)i;
        }
        tile_bytes = (u32)total;
        nlines = (u32)(total >> 32);
        if (tid == 0) s_off[nslots] = tile_bytes;   // (no thread's slot range reaches index nslots)
    }
    if (warp == 0) {
        u64 prefix = 0;
        if (tile > 0) {
            if (lane == 0) st[tile] = kIdxFlagAgg | (u64)tile_bytes;
            __syncwarp();
            int j = (int)tile - 1;
            for (;;) {
                const int idx = j - (int)lane;
                u64 w = kIdxFlagPre;   // tiles before the first: prefix 0
                if (idx >= 0) {
                    u32 spins = 0;
                    while (((w = st[idx]) >> 62) == 0) {
                        __nanosleep(64);
                        if (++spins > (1u << 22) || *g_fallback) {   // the result is void anyway once the fallback flag is up: do not wait for anyone
                            if (spins > (1u << 22)) atomicOr(&a.meta->fallback, (u32)(kFuseNotConvertible | kFuseTimeoutLookback));
                            w = kIdxFlagPre;
                            break;
                        }
                    }
                }
                const u32 pre = __ballot_sync(FULL, (w >> 62) == 2);
                const u32 upto = pre ? (u32)__ffs((int)pre) - 1u : 31u;   // lanes 0..upto contribute
                u64 val = lane <= upto ? (w & kIdxValMask) : 0;
                for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(FULL, val, o);
                prefix += __shfl_sync(FULL, val, 0);
                if (pre) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            st[tile] = kIdxFlagPre | (prefix + tile_bytes);
            s_obase = prefix;
            atomicAdd(&a.meta->n_records, R);
            if (nlines) atomicAdd(&a.meta->n_lines, nlines);
            if (tile == a.ntiles - 1) a.meta->out_total = prefix + tile_bytes;
            if (prefix + tile_bytes > a.out_cap) atomicExch(&a.meta->overflow, 1u);
        }
    }
    __syncthreads();
    const u64 obase = s_obase;
    if (obase + tile_bytes > a.out_cap || *g_fallback) return;   // nothing may be written (the host grows the buffer / runs the general pipeline)

    if (C::kDirect) {   // ---------------- I (direct): every line straight to its place in the output
        for (u32 k = tid; k < nlines; k += kFThreads) {
            const u32 slot = s_lidx[k];
            const u32 o = s_off[slot], l = s_off[slot + 1] - o;
            const uint4* sl = s_slots + 3u * slot;
            const uint4 v0 = sl[0], v1 = sl[1], v2 = sl[2];
            FRec F;
            frec_load(s_info + 2u * (v2.w & 0xffffu), F);
            fuse_write_line(a.out + obase + o, l, text, F, v0, v1, v2);
        }
        return;
    }
    // ---------------- I: format and store, in rounds of at most kFThreads lines and kFStage bytes
    u32 done = 0;
    while (done < nlines) {
        const u32 off0 = s_off[s_lidx[done]];
        const u32 pad = (u32)((obase + off0) & 15u);
        const u32 k = done + tid;
        u32 slot = 0, o = 0, l = 0;
        bool fits = false;
        if (k < nlines) {
            slot = s_lidx[k];
            o = s_off[slot];
            l = s_off[slot + 1] - o;
            fits = o + l - off0 + pad <= kFStage;
        }
        // lines that fit form a prefix of the round (offsets are monotone): count them
        const u32 bal = __ballot_sync(FULL, fits);
        __syncthreads();   // the previous round's staging buffer has been read (its issuer waited before this barrier)
        if (lane == 0) s_w[warp] = (u32)__popc(bal);
        __syncthreads();
        u32 cnt = 0;
        for (u32 i = 0; i < kFThreads / 32; ++i) cnt += s_w[i];
        if (cnt == 0) {   // a single line longer than the staging buffer: not a short record after all
            if (tid == 0) atomicOr(&a.meta->fallback, (u32)kFuseNotConvertible);
            return;
        }
        if (fits) {
            const uint4* sl = s_slots + 3u * slot;
            const uint4 v0 = sl[0], v1 = sl[1], v2 = sl[2];
            FRec F;
            frec_load(s_info + 2u * (v2.w & 0xffffu), F);
            fuse_write_line(s_stage + pad + (o - off0), l, text, F, v0, v1, v2);
        }
        const u32 last = s_lidx[done + cnt - 1u];
        const u32 bytes = s_off[last + 1u] - off0;
        u8* gb = a.out + (obase + off0 - pad);   // 16-byte aligned
        const u32 total = pad + bytes;
        const u32 full_b = total >> 4, first_b = pad ? 1u : 0u;
#if !defined(G2P_HOSTSIM)
        fence_async_smem();
        __syncthreads();
        if (tid == 0 && full_b > first_b) { bulk_s2g(gb + 16u * first_b, s_stage + 16u * first_b, 16u * (full_b - first_b)); bulk_commit(); }
#else
        __syncthreads();
        for (u32 u = first_b + tid; u < full_b; u += kFThreads) reinterpret_cast<uint4*>(gb)[u] = reinterpret_cast<const uint4*>(s_stage)[u];
#endif
        const u32 head_end = pad ? (total < 16u ? total : 16u) : 0u;
        for (u32 b2 = pad + tid; b2 < head_end; b2 += kFThreads) gb[b2] = s_stage[b2];
        const u32 tail_a = full_b * 16u > head_end ? full_b * 16u : head_end;
        for (u32 b2 = tail_a + tid; b2 < total; b2 += kFThreads) gb[b2] = s_stage[b2];
#if !defined(G2P_HOSTSIM)
        if (tid == 0) bulk_wait_read0();   // the staging buffer must outlive the copy's reads
#endif
        done += cnt;
    }
}

}  // namespace g2p
