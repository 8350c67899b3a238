// g2p_core.cuh — per-record GAF -> PAF conversion, written once for device code.
//
// Everything in this header is `__host__ __device__` so that the *same* code the
// sm_100a kernels run can also be instantiated by the host-side simulator under
// tests/hostsim (CPU-only CI; it is test infrastructure, the product library only
// ever launches the CUDA kernels).
//
// What it computes (reference: gaf2paf_main.cpp:92-264, gafkluge.hpp:84-239,
// paf.hpp:83-95; closed form in SURVEY.md Appendix B): one minigraph GAF record is
// cut at every path-step boundary, coordinates are lifted through the name->length
// table, and one PAF line per step with >0 matches is produced.
//
// Design: a record is never materialised into lists.  Two cursors stream over the
// record text -- one over the path column's '>'/'<' step tokens and one over the
// cg:Z: CIGAR -- forwards for '+' records and backwards for '-' records (which is
// what flip_gaf, gaf2paf_main.cpp:92-131, amounts to).  State is O(1) per record,
// so the same function serves 150-byte short-read records and 300 kB assembly
// records.  The walk is templated on an output sink: CountSink yields the exact
// byte length (pass 1 of the emitter), StoreSink writes the bytes (pass 2).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define G2P_HD __host__ __device__ __forceinline__
#define G2P_HD_NOINLINE __host__ __device__ __noinline__
#else
#define G2P_HD inline
#define G2P_HD_NOINLINE
#endif

namespace g2p {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;
typedef int32_t i32;

// ---------------------------------------------------------------------------------
// Per-record status.  Low byte = code, bits 8..15 = auxiliary (GAF column number).
// Codes < 16 are the reference's `exit(1)` paths, codes >= 16 are its aborts
// (live asserts and uncaught exceptions -> SIGABRT, rc 134).
// ---------------------------------------------------------------------------------
enum : u32 {
    ST_OK = 0,
    ST_ERR_NAME = 1,          // gaf2paf_main.cpp:118,163  "unable to find X in lengths map"
    ST_ERR_NOCG = 2,          // gaf2paf_main.cpp:365-368
    ST_ABORT_COLUMN = 16,     // gafkluge.hpp:94   runtime_error "Error parsing GAF column N"
    ST_ABORT_STRAND = 17,     // gafkluge.hpp:115
    ST_ABORT_RANGE = 18,      // gafkluge.hpp:141  "Error parsing GAF range of TOKEN"
    ST_ABORT_STOL = 19,       // std::invalid_argument from stol/stoi
    ST_ABORT_STOL_RANGE = 20, // std::out_of_range from stol/stoi
    ST_ABORT_TAG = 21,        // gafkluge.hpp:191  "Unable to parse optional tag X"
    ST_ABORT_DUPTAG = 22,     // gafkluge.hpp:197  "Duplicate optional field found: X"
    ST_ABORT_CIGAR = 23,      // gafkluge.hpp:232-234 (assert / stol inside for_each_cg)
    ST_ABORT_ASSERT = 24,     // gaf2paf_main.cpp:80,101,136,178 asserts
    ST_SKIP = 255             // '*' line (gaf2paf_main.cpp:360): no output, not an error
};
G2P_HD bool st_is_abort(u32 st) { return (st & 0xff) >= 16 && (st & 0xff) != ST_SKIP; }
G2P_HD bool st_is_error(u32 st) {   // code 3 is gaf2unstable's multi-contig warning, not an error
    const u32 c = st & 0xff;
    return c == ST_ERR_NAME || c == ST_ERR_NOCG || (c >= 16 && c != ST_SKIP);
}

// ---------------------------------------------------------------------------------
// name -> length table (reference: unordered_map<string,int64_t>, gaf2paf_main.cpp:22-45).
// Open addressing, linear probing, 32-byte slots = one DRAM/L2 sector per probe.
// Names of <= 16 bytes are compared exactly through (k0,k1,name_len) alone; longer
// names carry a 128-bit hash in (k0,k1) and are verified against the name arena.
// ---------------------------------------------------------------------------------
struct __attribute__((aligned(32))) LenSlot {
    u64 k0, k1;
    i64 length;
    u32 name_off;
    u32 name_len;   // 0xFFFFFFFF = empty
};
static const u32 kEmptySlot = 0xFFFFFFFFu;

struct LenTableView {
    const LenSlot* slots;
    const u8* arena;
    u32 nslots;
};

G2P_HD u64 mix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// Key of a name: the bytes themselves (little-endian, zero padded) when len <= 16,
// otherwise a 128-bit fold of its 16-byte blocks.
G2P_HD void name_key(const u8* s, u32 len, u64& k0, u64& k1) {
    u64 a = 0, b = 0;
    if (len <= 16) {
        u32 n0 = len < 8 ? len : 8;
        for (u32 i = 0; i < n0; ++i) a |= (u64)s[i] << (8 * i);
        for (u32 i = 8; i < len; ++i) b |= (u64)s[i] << (8 * (i - 8));
    } else {
        u64 h0 = 0x9e3779b97f4a7c15ULL, h1 = 0xc2b2ae3d27d4eb4fULL;
        for (u32 base = 0; base < len; base += 16) {
            u64 x = 0, y = 0;
            for (u32 i = 0; i < 8 && base + i < len; ++i) x |= (u64)s[base + i] << (8 * i);
            for (u32 i = 8; i < 16 && base + i < len; ++i) y |= (u64)s[base + i] << (8 * (i - 8));
            h0 = mix64(h0 ^ x ^ ((h1 << 1) | (h1 >> 63)));
            h1 = mix64(h1 ^ y ^ h0);
        }
        a = h0; b = h1;
    }
    k0 = a; k1 = b;
}

// Slot of a key: 32-bit multiply/xor-shift mixing (cheap on the device: the probe, not the
// hash, should dominate a lookup).
G2P_HD u32 slot_index(u64 k0, u64 k1, u32 len, u32 nslots) {
    u32 x = (u32)k0 * 0x9E3779B1u ^ (u32)(k0 >> 32) * 0x85EBCA77u ^ (u32)k1 * 0xC2B2AE3Du ^ (u32)(k1 >> 32) * 0x27D4EB2Fu ^
            (len * 0x165667B1u);
    x ^= x >> 15; x *= 0x2C1B3C6Du;
    x ^= x >> 12; x *= 0x297A2D39u;
    x ^= x >> 15;
    return (u32)(((u64)x * (u64)nslots) >> 32);
}

// Probe with a ready key for names of <= 16 bytes (the key is the name itself, so no arena
// compare is needed).  Returns true and the length when present.
G2P_HD bool table_lookup_key16(const LenTableView& T, u64 k0, u64 k1, u32 len, i64& length) {
    if (T.nslots == 0) return false;
    u32 idx = slot_index(k0, k1, len, T.nslots);
    for (;;) {
        const LenSlot* sl = T.slots + idx;
#if defined(__CUDA_ARCH__)
        const ulonglong2 lo = __ldg(reinterpret_cast<const ulonglong2*>(sl));
        const ulonglong2 hi = __ldg(reinterpret_cast<const ulonglong2*>(sl) + 1);
        const u64 s_k0 = lo.x, s_k1 = lo.y;
        const i64 s_len = (i64)hi.x;
        const u32 s_nlen = (u32)(hi.y >> 32);
#else
        const u64 s_k0 = sl->k0, s_k1 = sl->k1;
        const i64 s_len = sl->length;
        const u32 s_nlen = sl->name_len;
#endif
        if (s_nlen == kEmptySlot) return false;
        if (s_nlen == len && s_k0 == k0 && s_k1 == k1) { length = s_len; return true; }
        idx = idx + 1 == T.nslots ? 0 : idx + 1;
    }
}

// Returns true and the length when `name` is in the table.
G2P_HD bool table_lookup(const LenTableView& T, const u8* name, u32 len, i64& length) {
    if (T.nslots == 0) return false;
    u64 k0, k1;
    name_key(name, len, k0, k1);
    u32 idx = slot_index(k0, k1, len, T.nslots);
    for (;;) {
        const LenSlot* sl = T.slots + idx;
#if defined(__CUDA_ARCH__)
        const ulonglong2 lo = __ldg(reinterpret_cast<const ulonglong2*>(sl));
        const ulonglong2 hi = __ldg(reinterpret_cast<const ulonglong2*>(sl) + 1);
        const u64 s_k0 = lo.x, s_k1 = lo.y;
        const i64 s_len = (i64)hi.x;
        const u32 s_off = (u32)hi.y, s_nlen = (u32)(hi.y >> 32);
#else
        const u64 s_k0 = sl->k0, s_k1 = sl->k1;
        const i64 s_len = sl->length;
        const u32 s_off = sl->name_off, s_nlen = sl->name_len;
#endif
        if (s_nlen == kEmptySlot) return false;
        if (s_nlen == len && s_k0 == k0 && s_k1 == k1) {
            bool same = true;
            if (len > 16) {
                const u8* a = T.arena + s_off;
                for (u32 i = 0; i < len; ++i) {
                    if (a[i] != name[i]) { same = false; break; }
                }
            }
            if (same) { length = s_len; return true; }
        }
        idx = idx + 1 == T.nslots ? 0 : idx + 1;
    }
}

// ---------------------------------------------------------------------------------
// std::stol semantics (gafkluge.hpp:33, :143-144, :234; gaf2paf_main.cpp:35):
// optional isspace prefix, optional sign, >=1 digit, trailing bytes ignored.
// ---------------------------------------------------------------------------------
G2P_HD bool is_space(u8 c) { return c == ' ' || (c >= 9 && c <= 13); }

G2P_HD_NOINLINE u32 stol_general(const u8* s, u32 a, u32 b, i64& out) {
    u32 i = a;
    while (i < b && is_space(s[i])) ++i;
    bool neg = false;
    if (i < b && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
    u32 nd = 0;
    u64 v = 0;
    bool over = false;
    const u64 lim = neg ? 9223372036854775808ULL : 9223372036854775807ULL;
    for (; i < b; ++i, ++nd) {
        u32 d = (u32)s[i] - '0';
        if (d > 9) break;
        if (v > (lim - d) / 10) over = true;
        if (!over) v = v * 10 + d;
    }
    if (nd == 0) return ST_ABORT_STOL;
    if (over) return ST_ABORT_STOL_RANGE;
    out = neg ? (i64)(0 - v) : (i64)v;
    return ST_OK;
}

G2P_HD u32 stol_span(const u8* s, u32 a, u32 b, i64& out) {
    u32 n = b - a;
    if (n >= 1 && n <= 18) {
        u64 v = 0;
        u32 i = a;
        for (; i < b; ++i) {
            u32 d = (u32)s[i] - '0';
            if (d > 9) break;
            v = v * 10 + d;
        }
        if (i == b) { out = (i64)v; return ST_OK; }
    }
    return stol_general(s, a, b, out);
}

// gafkluge.hpp:32-34 string_to_int: "*" -> -1
G2P_HD u32 gaf_int(const u8* s, u32 a, u32 b, i64& out) {
    if (b - a == 1 && s[a] == '*') { out = -1; return ST_OK; }
    return stol_span(s, a, b, out);
}

// ---------------------------------------------------------------------------------
// decimal formatting
// ---------------------------------------------------------------------------------
G2P_HD u32 dec_len_u64(u64 v) {
    if (v < 10000000000ULL) {
        u32 w = (u32)(v < 4294967296ULL ? v : 4294967295ULL);
        if (v < 4294967296ULL) {
            if (w < 10u) return 1;
            if (w < 100u) return 2;
            if (w < 1000u) return 3;
            if (w < 10000u) return 4;
            if (w < 100000u) return 5;
            if (w < 1000000u) return 6;
            if (w < 10000000u) return 7;
            if (w < 100000000u) return 8;
            if (w < 1000000000u) return 9;
        }
        return 10;
    }
    u32 n = 10;
    u64 p = 10000000000ULL;
    while (n < 20 && v >= p) { ++n; if (n < 20) p *= 10; }
    return n;
}
G2P_HD u32 dec_len_i64(i64 v) {
    return v < 0 ? 1 + dec_len_u64(0 - (u64)v) : dec_len_u64((u64)v);
}

// Writes v right-aligned ending at dst[n-1] where n = dec_len; returns n.
G2P_HD u32 dec_write_u64(u8* dst, u64 v) {
    u32 n = dec_len_u64(v);
    u32 i = n;
    while (v >= 4294967296ULL) { u64 q = v / 10; dst[--i] = (u8)('0' + (u32)(v - q * 10)); v = q; }
    u32 w = (u32)v;
    do { u32 q = w / 10; dst[--i] = (u8)('0' + (w - q * 10)); w = q; } while (i > 0);
    return n;
}

// ---------------------------------------------------------------------------------
// "%g" with 6 significant digits == what `ostream << double` prints at default
// precision (gaf2paf_main.cpp:253).  Exact: works on the binary value with 128-bit
// integers, round-half-even like glibc.  Valid for 0 and 1e-5 <= |v| < 1e22, which
// covers every value floor(x*1000+0.5)/1000 can take for int64 inputs.
// ---------------------------------------------------------------------------------
G2P_HD u32 fmt_g6(double v, u8* out) {
    union { double d; u64 u; } cv;
    cv.d = v;
    u32 n = 0;
    if (cv.u >> 63) { out[n++] = '-'; cv.u &= ~(1ULL << 63); }
    double av = cv.d;
    if (av == 0.0) { out[n++] = '0'; return n; }
    // binary decomposition av = M * 2^E
    int be = (int)((cv.u >> 52) & 0x7ff);
    u64 M = cv.u & ((1ULL << 52) - 1);
    int E;
    if (be == 0) { E = -1074; } else { M |= 1ULL << 52; E = be - 1075; }
    // decimal exponent X = floor(log10(av))
    int X;
    {
        const double p10[] = {1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10,
                              1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        X = -6;
        for (int k = 0; k < 28; ++k) { if (av >= p10[k]) X = k - 5; }
    }
    typedef unsigned __int128 u128;
    u64 D = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int p = 5 - X;   // D = round(av * 10^p)
        u128 num = (u128)M, den = 1;
        if (E >= 0) num <<= E; else den <<= (-E);
        if (p >= 0) { for (int k = 0; k < p; ++k) num *= 10; }
        else { for (int k = 0; k < -p; ++k) den *= 10; }
        u128 q = num / den, rem = num - q * den;
        u128 twice = rem * 2;
        if (twice > den || (twice == den && (q & 1))) q += 1;
        D = (u64)q;
        if (D >= 1000000ULL) { X += 1; continue; }   // rounded up to 10^6: renormalise
        break;
    }
    // D has exactly 6 digits (100000..999999)
    u8 dg[6];
    { u32 w = (u32)D; for (int k = 5; k >= 0; --k) { dg[k] = (u8)('0' + w % 10); w /= 10; } }
    int nsig = 6;
    while (nsig > 1 && dg[nsig - 1] == '0') --nsig;
    if (X < -4 || X >= 6) {
        out[n++] = dg[0];
        if (nsig > 1) { out[n++] = '.'; for (int k = 1; k < nsig; ++k) out[n++] = dg[k]; }
        out[n++] = 'e';
        int ax = X;
        if (ax < 0) { out[n++] = '-'; ax = -ax; } else out[n++] = '+';
        if (ax >= 100) { out[n++] = (u8)('0' + ax / 100); ax %= 100; }
        out[n++] = (u8)('0' + ax / 10);
        out[n++] = (u8)('0' + ax % 10);
    } else if (X >= 0) {
        for (int k = 0; k <= X; ++k) out[n++] = dg[k];
        if (nsig > X + 1) { out[n++] = '.'; for (int k = X + 1; k < nsig; ++k) out[n++] = dg[k]; }
    } else {
        out[n++] = '0'; out[n++] = '.';
        for (int k = 0; k < -X - 1; ++k) out[n++] = '0';
        for (int k = 0; k < nsig; ++k) out[n++] = dg[k];
    }
    return n;
}

// gi:f: value of a record (gaf2paf_main.cpp:248-253).  Separately rounded IEEE
// operations (no FMA contraction) so that the double equals the x86-64 reference's.
G2P_HD u32 fmt_gi(i64 m, i64 b, u8* out) {
    if (b <= 0) { out[0] = '0'; return 1; }
#if defined(__CUDA_ARCH__)
    double x = __ddiv_rn((double)m, (double)b);
    double k = floor(__dadd_rn(__dmul_rn(x, 1000.0), 0.5));
#else
    volatile double x = (double)m / (double)b;
    volatile double t = x * 1000.0;
    volatile double t2 = t + 0.5;
    double k = __builtin_floor(t2);
#endif
    if (k >= 0.0 && k <= 1000.0) {
        // k/1000 printed with %g: strip trailing zeros of the 3 decimals
        u32 ki = (u32)k;
        if (ki == 0) { out[0] = '0'; return 1; }
        if (ki == 1000) { out[0] = '1'; return 1; }
        u32 d0 = ki / 100, d1 = (ki / 10) % 10, d2 = ki % 10;
        out[0] = '0'; out[1] = '.';
        out[2] = (u8)('0' + d0); out[3] = (u8)('0' + d1); out[4] = (u8)('0' + d2);
        return d2 ? 5 : (d1 ? 4 : 3);
    }
#if defined(__CUDA_ARCH__)
    double g = __ddiv_rn(k, 1000.0);
#else
    volatile double g = k / 1000.0;
#endif
    return fmt_g6(g, out);
}

// ---------------------------------------------------------------------------------
// CIGAR op classes (gaf2paf_main.cpp:50-56).  Op letters all lie in ['=', 'X'].
// ---------------------------------------------------------------------------------
#define G2P_OPBIT(c) (1u << ((c) - '='))
static const u32 kOpMask = G2P_OPBIT('M') | G2P_OPBIT('I') | G2P_OPBIT('D') | G2P_OPBIT('N') | G2P_OPBIT('S') |
                           G2P_OPBIT('H') | G2P_OPBIT('P') | G2P_OPBIT('X') | G2P_OPBIT('=');
static const u32 kQueryMask = G2P_OPBIT('M') | G2P_OPBIT('I') | G2P_OPBIT('S') | G2P_OPBIT('=') | G2P_OPBIT('X');
static const u32 kTargetMask = G2P_OPBIT('M') | G2P_OPBIT('D') | G2P_OPBIT('N') | G2P_OPBIT('=') | G2P_OPBIT('X');
static const u32 kMatchMask = G2P_OPBIT('M') | G2P_OPBIT('=');
G2P_HD bool op_in(u32 mask, u8 c) {
    u32 k = (u32)c - '=';
    return k < 28 && ((mask >> k) & 1u);
}

// ---------------------------------------------------------------------------------
// Parsed record header (reference: GafRecord, gafkluge.hpp:56-79, minus the heap).
// All spans are byte offsets into the record text.
// ---------------------------------------------------------------------------------
struct RecHdr {
    i64 qlen, qs, ps, pe, m, b;   // cols 2,3,8,9,10,11 ('*' -> -1); ps/pe already mirrored for '-' records
    i32 mapq;                     // col 12 ('*' or >=255 -> -1)
    u32 qn_b;                     // query name = [0, qn_b)
    u32 path_a, path_b;           // col 6
    u32 tp_a, tp_b;               // "type:value" of tp tag (tp_b == 0: absent)
    u32 rc_a, rc_b;               // same for rc
    u32 cg_a, cg_b;               // cg value (has_cg says whether the tag exists)
    u8 minus;                     // strand == '-'
    u8 prefixed;                  // path starts with '>' or '<'
    u8 empty_path;                // path == "*"
    u8 has_cg;
};

struct StepTok {
    u32 name_a, name_b;
    i64 start, end;
    u8 rev, is_interval;
};

// One step token [p,q) of a prefixed path (gafkluge.hpp:123-147).  `checked` adds the
// reference's failure modes; the walk re-parses already validated tokens without.
template <bool CHECKED>
G2P_HD u32 parse_step_token(const u8* r, u32 p, u32 q, StepTok& t) {
    t.rev = r[p] == '<';
    u32 colon = p + 1;
    while (colon < q && r[colon] != ':') ++colon;
    t.name_a = p + 1;
    t.name_b = colon;
    if (colon == q) { t.is_interval = 0; t.start = 0; t.end = 0; return ST_OK; }
    t.is_interval = 1;
    u32 dash = colon + 1;
    while (dash < q && r[dash] != '-') ++dash;
    if (CHECKED && dash == q) return ST_ABORT_RANGE;
    // start = stol(substr(colon+1, dash-colon)): the dash itself is inside the substring
    u32 st = stol_span(r, colon + 1, dash, t.start);
    if (CHECKED && st != ST_OK) {
        // only digits-free prefixes fail; "…-" with no digits is invalid_argument
        return st;
    }
    st = stol_span(r, dash + 1, q, t.end);
    if (CHECKED && st != ST_OK) return st;
    return ST_OK;
}

// first '>' or '<' in [from, end), else end: four bytes at a time once the address is word-aligned (this scan is a
// quarter of the instructions of the gaf2unstable kernels when done byte by byte)
G2P_HD u32 next_marker(const u8* r, u32 from, u32 end) {
    while (from < end && (reinterpret_cast<uintptr_t>(r + from) & 3u) != 0) {
        if (r[from] == '>' || r[from] == '<') return from;
        ++from;
    }
    while (from + 4 <= end) {
        const u32 w = *reinterpret_cast<const u32*>(r + from);
        const u32 x = w ^ 0x3E3E3E3Eu, y = w ^ 0x3C3C3C3Cu;   // a zero byte where the text has '>' / '<'
        const u32 m = (~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) | ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y)) & 0x80808080u;
        if (m) {
            u32 k = 0;
            while (!((m >> (8 * k + 7)) & 1u)) ++k;
            return from + k;
        }
        from += 4;
    }
    while (from < end && r[from] != '>' && r[from] != '<') ++from;
    return from;
}

// ---------------------------------------------------------------------------------
// Header parse == parse_gaf_record (gafkluge.hpp:84-204) + the checks main() and
// flip_gaf make before any output of the record (gaf2paf_main.cpp:360-370, 92-131).
// On ST_ERR_NAME, [ea,eb) is the missing name.
// ---------------------------------------------------------------------------------
G2P_HD u32 parse_header(const u8* r, u32 len, const LenTableView& T, RecHdr& h, u32& ea, u32& eb) {
    if (len > 0 && r[0] == '*') return ST_SKIP;
    u32 pos = 0;
    bool eof = false;
    u32 a = 0, b = 0;
    i64 tmp;
    u32 st;
    bool have_miss = false;
    u32 miss_a = 0, miss_b = 0;
    i64 path_total = 0;

#define G2P_NEXT_COL(col)                                             \
    do {                                                              \
        if (eof) return ST_ABORT_COLUMN | ((col) << 8);               \
        a = pos; b = a;                                               \
        while (b < len && r[b] != '\t') ++b;                          \
        if (b < len) pos = b + 1; else { pos = len; eof = true; }     \
        if (b == a) return ST_ABORT_COLUMN | ((col) << 8);            \
    } while (0)

    G2P_NEXT_COL(1);
    h.qn_b = b;
    G2P_NEXT_COL(2);
    st = gaf_int(r, a, b, h.qlen); if (st) return st;
    G2P_NEXT_COL(3);
    st = gaf_int(r, a, b, h.qs); if (st) return st;
    G2P_NEXT_COL(4);
    st = gaf_int(r, a, b, tmp); if (st) return st;
    G2P_NEXT_COL(5);
    if (b - a != 1 || (r[a] != '+' && r[a] != '-' && r[a] != '*')) return ST_ABORT_STRAND;
    const u8 strand = r[a];
    h.minus = strand == '-';
    G2P_NEXT_COL(6);
    h.path_a = a; h.path_b = b;
    h.prefixed = (r[a] == '<' || r[a] == '>');
    h.empty_path = (!h.prefixed && b - a == 1 && r[a] == '*');
    if (h.prefixed) {
        // validate every token now: the reference parses the whole path before any output
        u32 p = a;
        while (p < b) {
            u32 q = next_marker(r, p + 1, b);
            StepTok t;
            st = parse_step_token<true>(r, p, q, t);
            if (st) return st;
            if (h.minus) {
                // flip_gaf: path length from the steps (gaf2paf_main.cpp:111-127)
                if (t.is_interval) {
                    path_total += t.end - t.start;
                } else {
                    i64 l;
                    if (table_lookup(T, r + t.name_a, t.name_b - t.name_a, l)) path_total += l;
                    else { have_miss = true; miss_a = t.name_a; miss_b = t.name_b; }  // reversed order: last one wins
                }
            }
            p = q;
        }
    } else if (!h.empty_path && h.minus) {
        i64 l;
        if (table_lookup(T, r + a, b - a, l)) path_total += l;
        else { have_miss = true; miss_a = a; miss_b = b; }
    }
    G2P_NEXT_COL(7);
    st = gaf_int(r, a, b, tmp); if (st) return st;
    G2P_NEXT_COL(8);
    st = gaf_int(r, a, b, h.ps); if (st) return st;
    G2P_NEXT_COL(9);
    st = gaf_int(r, a, b, h.pe); if (st) return st;
    G2P_NEXT_COL(10);
    st = gaf_int(r, a, b, h.m); if (st) return st;
    G2P_NEXT_COL(11);
    st = gaf_int(r, a, b, h.b); if (st) return st;
    G2P_NEXT_COL(12);
    if (b - a == 1 && r[a] == '*') {
        h.mapq = -1;
    } else {
        st = stol_span(r, a, b, tmp); if (st) return st;
        if (tmp > 2147483647LL || tmp < -2147483648LL) return ST_ABORT_STOL_RANGE;   // stoi
        h.mapq = tmp >= 255 ? -1 : (i32)tmp;
    }
#undef G2P_NEXT_COL

    // optional fields (gafkluge.hpp:185-202)
    h.tp_a = h.tp_b = h.rc_a = h.rc_b = h.cg_a = h.cg_b = 0;
    h.has_cg = 0;
    u64 seen = 0;          // 64-bit sketch of tag names; exact check only on a sketch hit
    bool maybe_dup = false;
    const u32 tags_from = pos;
    while (!eof) {
        a = pos; b = a;
        while (b < len && r[b] != '\t') ++b;
        if (b < len) pos = b + 1; else { pos = len; eof = true; }
        if (b == a) continue;
        u32 c1 = a;
        while (c1 < b && r[c1] != ':') ++c1;
        u32 c2 = c1 + 1;
        while (c2 < b && r[c2] != ':') ++c2;
        if (b - a < 5 || c1 >= b || c2 >= b) return ST_ABORT_TAG;
        u32 kl = c1 - a;
        u32 hk = 0;
        for (u32 i = a; i < c1; ++i) hk = hk * 31 + r[i];
        u64 bit = 1ULL << ((hk ^ (hk >> 6) ^ kl) & 63);
        if (seen & bit) maybe_dup = true;
        seen |= bit;
        if (kl == 2) {
            u8 x = r[a], y = r[a + 1];
            if (x == 'c' && y == 'g') { h.has_cg = 1; h.cg_a = c2 + 1; h.cg_b = b; }
            else if (x == 't' && y == 'p') { h.tp_a = c1 + 1; h.tp_b = b; }
            else if (x == 'r' && y == 'c') { h.rc_a = c1 + 1; h.rc_b = b; }
        }
    }
    if (maybe_dup) {
        // exact pairwise comparison of tag names, in field order like the reference
        u32 p1 = tags_from;
        while (p1 < len) {
            u32 e1 = p1;
            while (e1 < len && r[e1] != '\t') ++e1;
            if (e1 > p1) {
                u32 k1 = p1;
                while (r[k1] != ':') ++k1;
                u32 p0 = tags_from;
                while (p0 < p1) {
                    u32 e0 = p0;
                    while (r[e0] != '\t') ++e0;
                    if (e0 > p0) {
                        u32 k0 = p0;
                        while (r[k0] != ':') ++k0;
                        if (k0 - p0 == k1 - p1) {
                            bool same = true;
                            for (u32 i = 0; i < k1 - p1; ++i) if (r[p0 + i] != r[p1 + i]) { same = false; break; }
                            if (same) return ST_ABORT_DUPTAG;
                        }
                    }
                    p0 = e0 + 1;
                }
            }
            p1 = e1 + 1;
        }
    }

    if (!h.has_cg) return ST_ERR_NOCG;

    // cg tokens: the reference materialises the whole CIGAR before any output (for_each_cg,
    // gafkluge.hpp:226-239): each token runs up to the next of "MIDNSHPX=" and its length is
    // std::stol of what precedes the letter -- leading blanks, a sign and trailing junk are
    // accepted, no digits or > int64 is an uncaught exception, a missing letter an assert.
    {
        u32 i = h.cg_a;
        while (i < h.cg_b) {
            u32 j = i;
            while (j < h.cg_b && !op_in(kOpMask, r[j])) ++j;
            if (j >= h.cg_b) return ST_ABORT_CIGAR;
            i64 v;
            const u32 cst = stol_general(r, i, j, v);
            if (cst) return cst;
            i = j + 1;
        }
    }

    if (strand == '*') return ST_ABORT_ASSERT;   // gaf2paf_main.cpp:136 assert(strand == '+')
    if (h.minus) {
        if (h.cg_a == h.cg_b) return ST_ABORT_ASSERT;                  // :101 assert(!cigar.empty())
        if (have_miss) { ea = miss_a; eb = miss_b; return ST_ERR_NAME; }   // :117-120
        i64 ns = path_total - h.pe, ne = path_total - h.ps;           // :128-131
        h.ps = ns; h.pe = ne;
    }
    return ST_OK;
}

// ---------------------------------------------------------------------------------
// CIGAR cursor.  `pos` is the next unread byte when reading forwards and the
// exclusive end of the unread region when reading backwards.  (rem, code) is the
// tail of an op cut by the previous step (cigar_cut, gaf2paf_main.cpp:59-68).
// ---------------------------------------------------------------------------------
struct OpCur {
    i64 rem;
    u32 pos;
    u8 code;
};

template <bool BWD>
G2P_HD void op_read(const u8* r, u32 lo, u32 hi, u32& pos, i64& l, u8& c) {
    // token = text up to (forwards) / back to (backwards) the neighbouring op letter; the CIGAR
    // has been validated by parse_header, so every token ends in a letter and std::stol succeeds
    if (!BWD) {
        u32 j = pos;
        while (j < hi && !op_in(kOpMask, r[j])) ++j;
        l = 0;
        stol_span(r, pos, j, l);
        c = r[j];
        pos = j + 1;
    } else {
        c = r[--pos];
        u32 j = pos;
        while (j > lo && !op_in(kOpMask, r[j - 1])) --j;
        l = 0;
        stol_span(r, j, pos, l);
        pos = j;
    }
}

struct StepOps {
    i64 q, t, nm, nb;       // query / target / matching / all bases of the step's pieces
    i64 first_len, last_len;
    u32 cglen;              // bytes of the pieces printed as <len><op>
    u32 n_new;              // ops newly read from the text (excludes the carried-over tail)
    u8 had_pending, cut, first_code;
};

// cigar_next_by_target (gaf2paf_main.cpp:71-90) without the list: take pieces until
// `quota` target bases are covered, cutting the last op if it overshoots.
template <bool BWD>
G2P_HD u32 consume_step(const u8* r, u32 cg_a, u32 cg_b, OpCur& c, i64 quota, StepOps& s) {
    s.q = s.t = s.nm = s.nb = 0;
    s.first_len = s.last_len = 0;
    s.cglen = 0; s.n_new = 0;
    s.had_pending = 0; s.cut = 0; s.first_code = 0;
    i64 cur = 0;
    while (cur < quota) {
        i64 l;
        u8 code;
        const bool pend = c.rem > 0;
        if (pend) {
            l = c.rem; code = c.code; c.rem = 0;
            s.had_pending = 1; s.first_code = code;
        } else {
            if (BWD ? (c.pos <= cg_a) : (c.pos >= cg_b)) return ST_ABORT_ASSERT;   // :80 assert(cur_len > target_len)
            op_read<BWD>(r, cg_a, cg_b, c.pos, l, code);
            ++s.n_new;
        }
        const u32 k = (u32)code - '=';
        if ((kTargetMask >> k) & 1u) {
            if (l > quota - cur) {
                const i64 use = quota - cur;
                c.rem = l - use; c.code = code;
                l = use;
                s.cut = 1;
            }
            cur += l;
            s.t += l;
        }
        if (pend) s.first_len = l;
        s.last_len = l;
        if ((kQueryMask >> k) & 1u) s.q += l;
        if ((kMatchMask >> k) & 1u) s.nm += l;
        s.nb += l;
        s.cglen += dec_len_i64(l) + 1;
    }
    return ST_OK;
}

// ---------------------------------------------------------------------------------
// Output sinks
// ---------------------------------------------------------------------------------
struct CountSink {
    static const bool kCount = true;
    u64 n;
    G2P_HD CountSink() : n(0) {}
    G2P_HD void bytes(const u8*, u32 k) { n += k; }
    G2P_HD void ch(u8) { n += 1; }
    G2P_HD void dec(i64 v) { n += dec_len_i64(v); }
    G2P_HD void udec(u64 v) { n += dec_len_u64(v); }
    G2P_HD void skip(u32 k) { n += k; }
};

struct StoreSink {
    static const bool kCount = false;
    u8* p;
    G2P_HD explicit StoreSink(u8* dst) : p(dst) {}
    G2P_HD void bytes(const u8* s, u32 k) { for (u32 i = 0; i < k; ++i) p[i] = s[i]; p += k; }
    G2P_HD void ch(u8 c) { *p++ = c; }
    G2P_HD void udec(u64 v) { p += dec_write_u64(p, v); }
    G2P_HD void dec(i64 v) {
        if (v < 0) { *p++ = '-'; udec(0 - (u64)v); } else udec((u64)v);
    }
};

// Pieces of one step printed as <len><op>…, in consumption order (replay from the
// cursor state at step start) or reversed (walk the text the other way from the
// cursor state at step end): gaf2paf_main.cpp:184-211.
template <bool BWD, class Sink>
G2P_HD void emit_pieces(const u8* r, u32 cg_a, u32 cg_b, const OpCur& c0, const OpCur& c1, i64 quota,
                        const StepOps& s, bool reversed, Sink& S) {
    if (!reversed) {
        OpCur c = c0;
        i64 cur = 0;
        while (cur < quota) {
            i64 l;
            u8 code;
            if (c.rem > 0) { l = c.rem; code = c.code; c.rem = 0; }
            else op_read<BWD>(r, cg_a, cg_b, c.pos, l, code);
            if (op_in(kTargetMask, code)) {
                if (l > quota - cur) l = quota - cur;
                cur += l;
            }
            S.dec(l);
            S.ch(code);
        }
    } else {
        u32 pos = c1.pos;
        for (u32 i = 0; i < s.n_new; ++i) {
            i64 l;
            u8 code;
            op_read<!BWD>(r, cg_a, cg_b, pos, l, code);
            if (i == 0 && s.cut) l = s.last_len;
            S.dec(l);
            S.ch(code);
        }
        if (s.had_pending) { S.dec(s.first_len); S.ch(s.first_code); }
    }
}

// ---------------------------------------------------------------------------------
// The step walk == gaf2paf() (gaf2paf_main.cpp:134-264) over a parsed header.
// Returns the record status; on ST_ERR_NAME the lines of earlier steps have been
// produced (the reference has written them before it exits) and [ea,eb) names the
// missing sequence.  Aborts produce no output for the record: the caller discards
// whatever the sink received.
// ---------------------------------------------------------------------------------
template <bool BWD, class Sink>
G2P_HD u32 walk_steps(const u8* r, const RecHdr& h, const LenTableView& T, Sink& S, u32& ea, u32& eb) {
    if (h.empty_path) return ST_OK;
    const i64 W = h.pe - h.ps;
    OpCur cur;
    cur.rem = 0; cur.code = 0;
    cur.pos = BWD ? h.cg_b : h.cg_a;
    i64 tbc = 0, qbc = 0;
    bool first = true;

    // per-record constants of every line
    u8 gi[16];
    const u32 gi_n = fmt_gi(h.m, h.b, gi);
    u32 const_len = 0;
    if constexpr (Sink::kCount) {
        // qname \t qlen \t . \t . \t strand \t name \t tlen \t ts \t te \t nm \t nb \t mapq  = 11 tabs + 1 strand char
        const_len = h.qn_b + dec_len_i64(h.qlen) + 12 + dec_len_i64((i64)h.mapq);
        if (h.tp_b) const_len += 4 + (h.tp_b - h.tp_a);
        if (h.rc_b) const_len += 4 + (h.rc_b - h.rc_a);
        const_len += 6 + dec_len_i64(h.m) + 6 + dec_len_i64(h.b) + 6 + gi_n + 6 + 1;
    }

    u32 p = BWD ? h.path_b : h.path_a;   // BWD: exclusive end of unread tokens; FWD: start of next token
    for (;;) {
        StepTok t;
        bool is_last;
        if (!h.prefixed) {
            t.name_a = h.path_a; t.name_b = h.path_b;
            t.rev = 0; t.is_interval = 0; t.start = t.end = 0;
            is_last = true;
        } else if (!BWD) {
            u32 q = next_marker(r, p + 1, h.path_b);
            parse_step_token<false>(r, p, q, t);
            p = q;
            is_last = q >= h.path_b;
        } else {
            u32 q = p;
            --p;
            while (r[p] != '>' && r[p] != '<') --p;
            parse_step_token<false>(r, p, q, t);
            is_last = p <= h.path_a;
        }
        bool rev = (t.rev != 0) != (h.minus != 0);

        i64 tlen;
        if (!table_lookup(T, r + t.name_a, t.name_b - t.name_a, tlen)) {   // :162-165
            ea = t.name_a; eb = t.name_b;
            return ST_ERR_NAME;
        }
        i64 sa = t.start, se = t.end;
        if (!t.is_interval) { sa = 0; se = tlen; }                         // :170-174
        i64 so = first ? h.ps : 0;                                         // :176
        i64 eo = is_last ? tbc + (se - sa) - W - so : 0;                   // :177
        if (so < 0 || eo < 0) return ST_ABORT_ASSERT;                      // :178
        const i64 quota = (se - eo) - (sa + so);                           // :182
        if (quota < 0) return ST_ABORT_ASSERT;   // reference: undefined behaviour (walks off the list head)

        const OpCur c0 = cur;
        StepOps s;
        u32 st = consume_step<BWD>(r, h.cg_a, h.cg_b, cur, quota, s);
        if (st) return st;

        if (rev) { i64 x = so; so = eo; eo = x; }                          // :184-186
        if (s.nm > 0) {                                                    // :225
            const i64 q0 = h.qs + qbc;
            const i64 ts = sa + so, te = se - eo;
            if constexpr (Sink::kCount) {
                S.skip(const_len + (t.name_b - t.name_a) + s.cglen);
                S.dec(q0); S.dec(q0 + s.q); S.dec(tlen); S.dec(ts); S.dec(te); S.dec(s.nm); S.dec(s.nb);
            } else {
                S.bytes(r, h.qn_b); S.ch('\t');
                S.dec(h.qlen); S.ch('\t');
                S.dec(q0); S.ch('\t');
                S.dec(q0 + s.q); S.ch('\t');
                S.ch(rev ? '-' : '+'); S.ch('\t');
                S.bytes(r + t.name_a, t.name_b - t.name_a); S.ch('\t');
                S.dec(tlen); S.ch('\t');
                S.dec(ts); S.ch('\t');
                S.dec(te); S.ch('\t');
                S.dec(s.nm); S.ch('\t');
                S.dec(s.nb); S.ch('\t');
                S.dec((i64)h.mapq);
                if (h.tp_b) { S.ch('\t'); S.ch('t'); S.ch('p'); S.ch(':'); S.bytes(r + h.tp_a, h.tp_b - h.tp_a); }
                if (h.rc_b) { S.ch('\t'); S.ch('r'); S.ch('c'); S.ch(':'); S.bytes(r + h.rc_a, h.rc_b - h.rc_a); }
                S.ch('\t'); S.ch('g'); S.ch('m'); S.ch(':'); S.ch('i'); S.ch(':'); S.dec(h.m);
                S.ch('\t'); S.ch('g'); S.ch('l'); S.ch(':'); S.ch('i'); S.ch(':'); S.dec(h.b);
                S.ch('\t'); S.ch('g'); S.ch('i'); S.ch(':'); S.ch('f'); S.ch(':'); S.bytes(gi, gi_n);
                S.ch('\t'); S.ch('c'); S.ch('g'); S.ch(':'); S.ch('Z'); S.ch(':');
                emit_pieces<BWD>(r, h.cg_a, h.cg_b, c0, cur, quota, s, rev, S);
                S.ch('\n');
            }
        }
        qbc += s.q;
        tbc += s.t;
        first = false;
        if (is_last) break;
    }
    return ST_OK;
}

// Whole record: header + walk.  `len` excludes the newline.
template <class Sink>
G2P_HD u32 convert_record(const u8* r, u32 len, const LenTableView& T, Sink& S, u32& ea, u32& eb) {
    RecHdr h;
    ea = eb = 0;
    u32 st = parse_header(r, len, T, h, ea, eb);
    if (st == ST_SKIP) return ST_SKIP;
    if (st != ST_OK) return st;
    return h.minus ? walk_steps<true>(r, h, T, S, ea, eb) : walk_steps<false>(r, h, T, S, ea, eb);
}

}  // namespace g2p
