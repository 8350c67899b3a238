// g2p_table.hpp — host-side build of the name -> length open-addressing table.
//
// Mirrors get_len_map (reference gaf2paf_main.cpp:22-45): rows are getline()-split,
// tokens are split on TAB only with empty tokens dropped (split_delims, paf.hpp:31-47),
// rows with fewer than two tokens are ignored, length = std::stol(token[1]) and a
// later duplicate name overwrites an earlier one.  The table is built once per run on
// the host and uploaded; the kernels only probe it (g2p_core.cuh: table_lookup).
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "g2p_core.cuh"

namespace g2p {

struct HostLenTable {
    std::vector<LenSlot> slots;
    std::vector<u8> arena;
    u64 n_entries = 0;

    LenTableView view() const {
        LenTableView v;
        v.slots = slots.data();
        v.arena = arena.data();
        v.nslots = (u32)slots.size();
        return v;
    }

    void reserve_for(u64 n_names) {
        // load factor <= 1/3 while the table stays L2-sized (short probe chains keep the lanes of a warp
        // together: every extra probe of one lane is an extra trip for the whole warp), <= 0.6 beyond
        u64 want = n_names <= 1000000ULL ? n_names * 3 + 16 : n_names * 5 / 3 + 16;
        if (want > 0xFFFFFFF0ULL) want = 0xFFFFFFF0ULL;
        slots.assign((size_t)want, LenSlot{0, 0, 0, 0, kEmptySlot});
        n_entries = 0;
    }

    // insert or overwrite
    void put(const u8* name, u32 len, i64 length) {
        u64 k0, k1;
        name_key(name, len, k0, k1);
        u32 n = (u32)slots.size();
        u32 idx = slot_index(k0, k1, len, n);
        for (;;) {
            LenSlot& s = slots[idx];
            if (s.name_len == kEmptySlot) {
                s.k0 = k0; s.k1 = k1; s.length = length; s.name_len = len;
                s.name_off = (u32)arena.size();
                if (len > 16) arena.insert(arena.end(), name, name + len);
                ++n_entries;
                return;
            }
            if (s.name_len == len && s.k0 == k0 && s.k1 == k1 &&
                (len <= 16 || std::memcmp(arena.data() + s.name_off, name, len) == 0)) {
                s.length = length;
                return;
            }
            idx = idx + 1 == n ? 0 : idx + 1;
        }
    }
};

// Parses a lengths TSV held in memory.  Returns ST_OK, or the abort status the
// reference would die with (std::stol throwing inside get_len_map -> SIGABRT).
inline u32 build_len_table(const char* tsv, size_t n, HostLenTable& out) {
    const u8* s = reinterpret_cast<const u8*>(tsv);
    // pass 1: count rows to size the table
    u64 rows = 0;
    for (size_t i = 0; i < n; ++i) rows += s[i] == '\n';
    if (n > 0 && s[n - 1] != '\n') ++rows;
    out.reserve_for(rows);
    out.arena.clear();
    // the arena must never be empty (device pointer arithmetic on a valid base)
    out.arena.reserve(64);
    size_t pos = 0;
    while (pos < n) {
        size_t eol = pos;
        while (eol < n && s[eol] != '\n') ++eol;
        // first two non-empty TAB-separated tokens
        size_t t0a = 0, t0b = 0, t1a = 0, t1b = 0;
        int ntok = 0;
        size_t i = pos;
        while (i < eol && ntok < 2) {
            while (i < eol && s[i] == '\t') ++i;
            if (i >= eol) break;
            size_t a = i;
            while (i < eol && s[i] != '\t') ++i;
            if (ntok == 0) { t0a = a; t0b = i; } else { t1a = a; t1b = i; }
            ++ntok;
        }
        if (ntok == 2) {
            i64 v = 0;
            // spans are relative to a row base so that 32-bit offsets suffice
            u32 st = stol_span(s + t1a, 0, (u32)(t1b - t1a), v);
            if (st != ST_OK) return st;
            if (t0b - t0a > 0x7fffffffULL) return ST_ABORT_STOL_RANGE;
            out.put(s + t0a, (u32)(t0b - t0a), v);
        }
        pos = eol + 1;
    }
    if (out.arena.empty()) out.arena.push_back(0);
    return ST_OK;
}

}  // namespace g2p
