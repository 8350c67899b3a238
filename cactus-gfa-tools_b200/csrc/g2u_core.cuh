// g2u_core.cuh — per-record stable -> unstable GAF rewrite (gaf2unstable), host+device.
//
// Reference: gaf2unstable_main.cpp:70-175 (get_unstable_interval, gaf2unstable),
// gafkluge.hpp:84-204 (parse) and :274-323 (re-serialisation).  For each path step
// `>contig:s-e` the nodes of `contig` overlapping [s,e) replace the step (reversed for
// '<'), a bare stable contig in the path column is resolved through path_start/path_end
// and rewrites those columns, an rc:Z:<reference contig> tag is set when all nodes lie
// on one reference contig, and the record is printed back with its optional fields in
// tag-name order (the reference keeps them in a std::map).
//
// Same streaming/sink structure as g2p_core.cuh: nothing is materialised per record; a
// step is a binary search in the contig's node array (sorted by SO offset).
#pragma once
#include "g2p_core.cuh"

namespace g2p {

enum : u32 {
    ST_WARN_MULTIREF = 3   // not an error: "Target path spans multiple reference contigs" (stderr warning)
};

struct __attribute__((aligned(32))) UNode {
    i64 offset;     // SO
    i64 cum;        // sum of lengths of the contig's nodes before this one
    u32 length;     // sequence length
    u32 name_off;
    u32 name_len;
    i32 ref;        // reference contig id, -1 = node unknown to rgfa2contig
};

struct UnstableView {
    LenTableView contigs;      // stable contig name -> contig index
    const u32* contig_begin;   // [ncontig + 1] into nodes
    const UNode* nodes;
    const u8* node_names;
    const u32* ref_off;        // [nref + 1] into ref_names
    const u8* ref_names;
};

constexpr u32 kUMaxTags = 16;
struct URecHdr {
    i64 qlen, qs, qe, plen, ps, pe, m, b;
    i32 mapq;
    u32 qn_b, path_a, path_b, tags_from;
    u8 strand, prefixed, empty_path;
    // the optional fields, collected while they are validated (so that neither the duplicate check nor the output in
    // name order has to scan the text again for every field); ntags > kUMaxTags: more than fit, the text is rescanned
    u32 ntags;
    u32 tag_a[kUMaxTags];
    u16 tag_n[kUMaxTags];   // bytes of the whole field
    u8 tag_k[kUMaxTags];    // bytes of its name
};

// Columns 1-12 and validation of the optional fields: parse_gaf_record (gafkluge.hpp:84-204).
//
// The parse -- and, below, the output of a record -- is cut into PHASES that carry their state in a struct.  On the
// host (and for one record at a time) they simply run one after the other: u_parse_header, unstable_record.  The
// kernels (k_unstable*, g2p_kernels.cuh) call the same phases with a warp barrier after each and run the two
// data-dependent loops (path steps, optional fields) as warp-uniform loops: per-record code like this diverges a
// little more with every data-dependent inner loop, and the lanes of a warp only meet again where ALL of them arrive
// -- without the barriers 4 of 32 lanes were active per instruction (profiles/r02_k_unstable.txt).
struct UParse { u32 pos; bool eof; };

#define G2U_NEXT_COL(col)                                             \
    do {                                                              \
        if (c.eof) return ST_ABORT_COLUMN | ((col) << 8);             \
        a = c.pos; b = a;                                             \
        while (b < len && r[b] != '\t') ++b;                          \
        if (b < len) c.pos = b + 1; else { c.pos = len; c.eof = true; } \
        if (b == a) return ST_ABORT_COLUMN | ((col) << 8);            \
    } while (0)

// '*' lines, columns 1-5
G2P_HD u32 u_hdr_a(const u8* r, u32 len, URecHdr& h, UParse& c) {
    if (len > 0 && r[0] == '*') return ST_SKIP;
    c.pos = 0; c.eof = false;
    u32 a = 0, b = 0, st;
    G2U_NEXT_COL(1); h.qn_b = b;
    G2U_NEXT_COL(2); st = gaf_int(r, a, b, h.qlen); if (st) return st;
    G2U_NEXT_COL(3); st = gaf_int(r, a, b, h.qs); if (st) return st;
    G2U_NEXT_COL(4); st = gaf_int(r, a, b, h.qe); if (st) return st;
    G2U_NEXT_COL(5);
    if (b - a != 1 || (r[a] != '+' && r[a] != '-' && r[a] != '*')) return ST_ABORT_STRAND;
    h.strand = r[a];
    return ST_OK;
}
// column 6: the path and the syntax of its steps
G2P_HD u32 u_hdr_b(const u8* r, u32 len, URecHdr& h, UParse& c) {
    u32 a = 0, b = 0, st = ST_OK;
    G2U_NEXT_COL(6);
    h.path_a = a; h.path_b = b;
    h.prefixed = (r[a] == '<' || r[a] == '>');
    h.empty_path = (!h.prefixed && b - a == 1 && r[a] == '*');
    if (h.prefixed) {   // (no return from inside a loop, here and below: on the device a return there moves the point where the
        u32 p = a;      // warp's lanes meet again to the end of the function)
        while (p < b && st == ST_OK) {
            u32 q = next_marker(r, p + 1, b);
            StepTok t;
            st = parse_step_token<true>(r, p, q, t);
            p = q;
        }
    }
    return st;
}
// columns 7-12
G2P_HD u32 u_hdr_c(const u8* r, u32 len, URecHdr& h, UParse& c) {
    u32 a = 0, b = 0, st;
    i64 tmp;
    G2U_NEXT_COL(7); st = gaf_int(r, a, b, h.plen); if (st) return st;
    G2U_NEXT_COL(8); st = gaf_int(r, a, b, h.ps); if (st) return st;
    G2U_NEXT_COL(9); st = gaf_int(r, a, b, h.pe); if (st) return st;
    G2U_NEXT_COL(10); st = gaf_int(r, a, b, h.m); if (st) return st;
    G2U_NEXT_COL(11); st = gaf_int(r, a, b, h.b); if (st) return st;
    G2U_NEXT_COL(12);
    if (b - a == 1 && r[a] == '*') {
        h.mapq = -1;
    } else {
        st = stol_span(r, a, b, tmp); if (st) return st;
        if (tmp > 2147483647LL || tmp < -2147483648LL) return ST_ABORT_STOL_RANGE;
        h.mapq = tmp >= 255 ? -1 : (i32)tmp;
    }
    h.tags_from = c.pos;
    return ST_OK;
}
#undef G2U_NEXT_COL
// optional fields: syntax + duplicate names (exact, pairwise: records carry a handful)
G2P_HD u32 u_hdr_d(const u8* r, u32 len, URecHdr& h, UParse& c) {
    u32 a = 0, b = 0;
    bool eof = c.eof;
    u32 p1 = c.pos;
    u32 tag_st = ST_OK;
    h.ntags = 0;
    while (!eof && tag_st == ST_OK) {
        a = p1; b = a;
        while (b < len && r[b] != '\t') ++b;
        if (b < len) p1 = b + 1; else { p1 = len; eof = true; }
        if (b == a) continue;
        u32 c1 = a;
        while (c1 < b && r[c1] != ':') ++c1;
        u32 c2 = c1 + 1;
        while (c2 < b && r[c2] != ':') ++c2;
        if (b - a < 5 || c1 >= b || c2 >= b) { tag_st = ST_ABORT_TAG; break; }
        // compare with every earlier field
        const u32 kn = c1 - a;
        if (h.ntags <= kUMaxTags) {
            for (u32 j = 0; j < h.ntags && j < kUMaxTags; ++j) {
                if (h.tag_k[j] == kn) {
                    bool same = true;
                    for (u32 i = 0; i < kn; ++i) if (r[h.tag_a[j] + i] != r[a + i]) { same = false; break; }
                    if (same) tag_st = ST_ABORT_DUPTAG;
                }
            }
            if (h.ntags < kUMaxTags && kn <= 255u && b - a <= 65535u) {
                h.tag_a[h.ntags] = a; h.tag_n[h.ntags] = (u16)(b - a); h.tag_k[h.ntags] = (u8)kn;
                ++h.ntags;
            } else h.ntags = kUMaxTags + 1;
            continue;
        }
        u32 p0 = h.tags_from;
        while (p0 < a && tag_st == ST_OK) {
            u32 e0 = p0;
            while (r[e0] != '\t') ++e0;
            if (e0 > p0) {
                u32 k0 = p0;
                while (r[k0] != ':') ++k0;
                if (k0 - p0 == c1 - a) {
                    bool same = true;
                    for (u32 i = 0; i < c1 - a; ++i) if (r[p0 + i] != r[a + i]) { same = false; break; }
                    if (same) tag_st = ST_ABORT_DUPTAG;
                }
            }
            p0 = e0 + 1;
        }
    }
    return tag_st;
}
G2P_HD u32 u_parse_header(const u8* r, u32 len, URecHdr& h) {
    UParse c;
    u32 st = u_hdr_a(r, len, h, c);
    if (st == ST_OK) st = u_hdr_b(r, len, h, c);
    if (st == ST_OK) st = u_hdr_c(r, len, h, c);
    if (st == ST_OK) st = u_hdr_d(r, len, h, c);
    return st;
}

// get_unstable_interval (gaf2unstable_main.cpp:70-107): node index range [i0, i1) of the
// contig covering [start, end), with the reference's assertions.
G2P_HD u32 u_interval(const UnstableView& V, const u8* name, u32 name_len, i64 start, i64 end, u32& i0, u32& i1) {
    i64 cidx;
    if (!table_lookup(V.contigs, name, name_len, cidx)) return ST_ABORT_ASSERT;   // :72 assert(lookup.count(contig))
    const u32 lo = V.contig_begin[cidx], hi = V.contig_begin[cidx + 1];
    // upper_bound(start): first node with offset > start
    u32 a = lo, b = hi;
    while (a < b) { u32 m = (a + b) >> 1; if (V.nodes[m].offset > start) b = m; else a = m + 1; }
    if (a == lo) return ST_ABORT_ASSERT;                                           // :78
    i0 = a - 1;
    // lower_bound(end): first node with offset >= end
    a = lo; b = hi;
    while (a < b) { u32 m = (a + b) >> 1; if (V.nodes[m].offset >= end) b = m; else a = m + 1; }
    if (a == lo) return ST_ABORT_ASSERT;                                           // :83
    i1 = a;
    if (i0 >= i1) return ST_ABORT_ASSERT;   // reference: indexes an empty vector (undefined behaviour)
    const UNode& first = V.nodes[i0];
    const UNode& last = V.nodes[i1 - 1];
    i64 ui_len = (last.cum + (i64)last.length) - first.cum;
    ui_len -= start - first.offset;                                                // :95-99
    if (ui_len > end - start) {                                                    // :100-104
        if ((i64)last.length - (ui_len - (end - start)) <= 0) return ST_ABORT_ASSERT;
        ui_len = end - start;
    }
    if (ui_len != end - start) return ST_ABORT_ASSERT;                             // :105
    return ST_OK;
}

// lexicographic (unsigned byte) comparison of two tag names, like std::string::operator<
G2P_HD int u_key_cmp(const u8* x, u32 xn, const u8* y, u32 yn) {
    u32 n = xn < yn ? xn : yn;
    for (u32 i = 0; i < n; ++i) {
        if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    }
    return xn == yn ? 0 : (xn < yn ? -1 : 1);
}

template <class Sink>
G2P_HD void u_put_int(Sink& S, i64 v) {   // gafkluge.hpp:27-29 int_to_string
    if (v == -1) S.ch('*'); else S.dec(v);
}

// gaf2unstable (gaf2unstable_main.cpp:109-175) + operator<<(GafRecord) (gafkluge.hpp:288-323), in phases (see above)
struct URun {
    u32 p;                 // next step token of a prefixed path
    u32 fail;              // first failure (the record aborts the reference)
    i32 ref_first;
    bool multi, any_step;
    bool steps_left, tags_left;
    bool set_rc, rc_done, have_last;
    u32 last_a, last_n;
};

// columns 1-5; an empty path prints its six '*' columns here
template <class Sink>
G2P_HD void u_out_head(const u8* r, URecHdr& h, URun& u, Sink& S) {
    S.bytes(r, h.qn_b); S.ch('\t');
    u_put_int(S, h.qlen); S.ch('\t');
    u_put_int(S, h.qs); S.ch('\t');
    u_put_int(S, h.qe); S.ch('\t');
    S.ch(h.strand); S.ch('\t');
    u.p = h.path_a; u.fail = ST_OK; u.ref_first = -1; u.multi = false; u.any_step = false;
    u.steps_left = !h.empty_path; u.tags_left = false;
    if (h.empty_path) for (int k = 0; k < 6; ++k) { S.ch('*'); S.ch('\t'); }
}
// one path step -> its nodes
template <class Sink>
G2P_HD void u_out_step(const u8* r, const UnstableView& V, URecHdr& h, URun& u, Sink& S) {
    StepTok t;
    bool is_last;
    u32 st = ST_OK;
    if (!h.prefixed) {
        t.name_a = h.path_a; t.name_b = h.path_b; t.rev = 0; t.is_interval = 0; t.start = t.end = 0;
        is_last = true;
    } else {
        u32 q = next_marker(r, u.p + 1, h.path_b);
        parse_step_token<false>(r, u.p, q, t);
        is_last = q >= h.path_b;
        if (!t.is_interval && !(u.p == h.path_a && is_last)) st = ST_ABORT_ASSERT;   // :116 assert(path.size() == 1)
        u.p = q;
    }
    u32 i0 = 0, i1 = 0;
    if (st == ST_OK) {
        if (!t.is_interval) {
            st = u_interval(V, r + t.name_a, t.name_b - t.name_a, h.ps, h.pe, i0, i1);
            if (st == ST_OK) {
                const i64 path_len = h.pe - h.ps;                     // :119-127
                h.ps -= V.nodes[i0].offset;
                h.pe = h.ps + path_len;
                const UNode& last = V.nodes[i1 - 1];
                h.plen = (last.cum + (i64)last.length) - V.nodes[i0].cum;
            }
        } else {
            st = u_interval(V, r + t.name_a, t.name_b - t.name_a, t.start, t.end, i0, i1);
        }
    }
    if (st == ST_OK) {
        const u32 cnt = i1 - i0;
        for (u32 k = 0; k < cnt; ++k) {
            const UNode& nd = V.nodes[t.rev ? i1 - 1 - k : i0 + k];   // :135-137
            S.ch(t.rev ? '<' : '>');
            S.bytes(V.node_names + nd.name_off, nd.name_len);
            if (nd.ref < 0) { st = ST_ABORT_ASSERT; break; }          // :160-161 node_id / partition lookup
            if (!u.any_step) { u.ref_first = nd.ref; u.any_step = true; }
            else if (nd.ref != u.ref_first) u.multi = true;
        }
    }
    if (st != ST_OK) u.fail = st;
    if (st != ST_OK || is_last) u.steps_left = false;
}
// columns 7-12
template <class Sink>
G2P_HD void u_out_mid(URecHdr& h, URun& u, Sink& S) {
    if (u.fail) return;
    if (!h.empty_path) {
        S.ch('\t');
        u_put_int(S, h.plen); S.ch('\t');
        u_put_int(S, h.ps); S.ch('\t');
        u_put_int(S, h.pe); S.ch('\t');
        u_put_int(S, h.m); S.ch('\t');
        u_put_int(S, h.b); S.ch('\t');
    }
    S.dec(h.mapq == -1 ? 255 : (i64)h.mapq);
    // optional fields in tag-name order, rc overridden when exactly one reference contig (:172-174)
    u.set_rc = u.any_step && !u.multi;
    u.rc_done = !u.set_rc;
    u.last_a = 0; u.last_n = 0; u.have_last = false;
    u.tags_left = true;
}
// the next optional field in name order (or the synthetic rc tag)
template <class Sink>
G2P_HD void u_out_tag(const u8* r, u32 len, const UnstableView& V, URecHdr& h, URun& u, Sink& S) {
    const u8 rc_key[2] = {'r', 'c'};
    // smallest key strictly greater than the last one printed
    bool found = false;
    u32 best_a = 0, best_b = 0, best_k = 0;
    u32 p0 = h.tags_from;
    if (h.ntags <= kUMaxTags) {   // from the collected fields
        for (u32 j = 0; j < h.ntags; ++j) {
            const u32 ta = h.tag_a[j], kn = h.tag_k[j];
            const bool is_rc = u.set_rc && kn == 2 && r[ta] == 'r' && r[ta + 1] == 'c';
            if (!is_rc && (!u.have_last || u_key_cmp(r + ta, kn, r + u.last_a, u.last_n) > 0) &&
                (!found || u_key_cmp(r + ta, kn, r + best_a, best_k) < 0)) {
                found = true; best_a = ta; best_b = ta + h.tag_n[j]; best_k = kn;
            }
        }
        p0 = len;
    }
    while (p0 < len) {
        u32 e0 = p0;
        while (e0 < len && r[e0] != '\t') ++e0;
        if (e0 > p0) {
            u32 k0 = p0;
            while (r[k0] != ':') ++k0;
            const u32 kn = k0 - p0;
            const bool is_rc = u.set_rc && kn == 2 && r[p0] == 'r' && r[p0 + 1] == 'c';
            if (!is_rc && (!u.have_last || u_key_cmp(r + p0, kn, r + u.last_a, u.last_n) > 0) &&
                (!found || u_key_cmp(r + p0, kn, r + best_a, best_k) < 0)) {
                found = true; best_a = p0; best_b = e0; best_k = kn;
            }
        }
        p0 = e0 + 1;
    }
    // does the synthetic rc tag come before the candidate?
    if (!u.rc_done && (!found || u_key_cmp(rc_key, 2, r + best_a, best_k) < 0)) {
        S.ch('\t'); S.ch('r'); S.ch('c'); S.ch(':'); S.ch('Z'); S.ch(':');
        const u32 o0 = V.ref_off[u.ref_first], o1 = V.ref_off[u.ref_first + 1];
        S.bytes(V.ref_names + o0, o1 - o0);
        u.rc_done = true;
        return;   // `last` unchanged: rc never equals an input key here (input rc is skipped)
    }
    if (!found) { u.tags_left = false; return; }
    S.ch('\t');
    S.bytes(r + best_a, best_b - best_a);
    u.last_a = best_a; u.last_n = best_k; u.have_last = true;
}
template <class Sink>
G2P_HD u32 u_out_end(URun& u, Sink& S) {
    S.ch('\n');
    return u.fail ? u.fail : (u.multi ? (u32)ST_WARN_MULTIREF : (u32)ST_OK);
}

template <class Sink>
G2P_HD u32 unstable_record(const u8* r, u32 len, const UnstableView& V, Sink& S, u32& ea, u32& eb) {
    ea = eb = 0;
    URecHdr h;
    u32 st = u_parse_header(r, len, h);
    if (st != ST_OK) return st;
    URun u;
    u_out_head(r, h, u, S);
    while (u.steps_left) u_out_step(r, V, h, u, S);
    u_out_mid(h, u, S);
    while (u.tags_left) u_out_tag(r, len, V, h, u, S);
    return u_out_end(u, S);
}

}  // namespace g2p
