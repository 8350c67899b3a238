// g2p_errfmt.hpp — the stderr line the reference prints for a failed record (g2p_format_error of the C-ABI).
// Host-only, shared by libg2p.so and by the CPU stub of the C-ABI that CPU-only CI links the executables
// against (tests/hostsim/g2p_stub_capi.cpp).
#pragma once
#include <cstdio>
#include <string>

#include "../../include/g2p.h"

namespace g2p_errfmt {

inline int format_error(const g2p_result* res, const char* gaf, size_t n, char* buf, size_t cap) {
    if (!res || !buf || cap == 0) return G2P_E_ARG;
    buf[0] = 0;
    if (res->rec_status == G2P_REC_ERR_NAME) {   // gaf2paf_main.cpp:118,163
        std::string name;
        if (gaf && res->err_name_off + res->err_name_len <= n) name.assign(gaf + res->err_name_off, res->err_name_len);
        std::snprintf(buf, cap, "[gaf2paf] error: unable to find %s in lengths map\n", name.c_str());
    } else if (res->rec_status == G2P_REC_ERR_NOCG) {   // gaf2paf_main.cpp:365-368
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

}  // namespace g2p_errfmt
