// gaf2unstable — drop-in command line for the reference tool of the same name
// (reference gaf2unstable_main.cpp:177-301): same options (and option quirks), same stdout
// bytes, same stderr messages and exit codes; the per-record rewrite runs on a B200 through
// the C-ABI of libg2p.so (include/g2p.h), the rGFA tables are built once on the host.
//
//   gaf2unstable [options] <gaf>
//     -g, --rgfa FILE   (uncompressed) minigraph rGFA
//     -o FILE           write "node<TAB>length" for every rGFA node (input of gaf2paf -l)
//
// Environment: G2P_DEVICE=K (device ordinal), G2P_CHUNK_MB=M (bytes of GAF per GPU call).
#include <fcntl.h>
#include <getopt.h>
#include <signal.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/g2p.h"

namespace {

void help(char** argv) {
    fprintf(stderr,
            "usage: %s [options] <gaf> \n"
            "Replace stable sequences in path steps, ex >chr1:500-1000, with the unstable graph node names, ex >s1:1-100>s2:100-600\n"
            "\n"
            "options: \n"
            "    -g, --rGFA FILE           (uncompressed) minigraph rGFA, required to look up unstable mappings\n"
            "    -o, --out-lengths FILE    Output lengths of all minigraph sequences in given file (can be passed to gaf2paf)\n",
            argv[0]);
}

bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    fclose(f);
    return true;
}

void write_all(const char* p, size_t n) {
    while (n) {
        ssize_t k = ::write(1, p, n);
        if (k < 0) {
            if (errno == EINTR) continue;
            _exit(1);
        }
        p += k;
        n -= (size_t)k;
    }
}

long env_long(const char* k, long dflt) {
    const char* v = getenv(k);
    return v && *v ? strtol(v, nullptr, 10) : dflt;
}

}  // namespace

int main(int argc, char** argv) {
    std::string rgfa_path, node_lengths_path;
    int c;
    optind = 1;
    while (true) {
        // the long form of -o maps to a value the switch does not handle (reference :196-226)
        static const struct option long_options[] = {
            {"help", no_argument, 0, 'h'}, {"rgfa", required_argument, 0, 'g'}, {"out-lengths", required_argument, 0, '0'}, {0, 0, 0, 0}};
        int option_index = 0;
        c = getopt_long(argc, argv, "hg:o:", long_options, &option_index);
        if (c == -1) break;
        switch (c) {
            case 'h':
                // falls through to `rgfa_path = optarg` with optarg == NULL in the reference: SIGSEGV
                raise(SIGSEGV);
                return 139;
            case 'g': rgfa_path = optarg; break;
            case 'o': node_lengths_path = optarg; break;
            case '?':
                help(argv);
                exit(1);
            default: abort();
        }
    }
    if (argc <= 1) { help(argv); return 1; }
    if (optind >= argc) {
        fprintf(stderr, "[gaf2unstable] error: too few arguments\n");
        help(argv);
        return 1;
    }
    const std::string in_gaf_path = argv[optind++];
    if (optind < argc - 1) {
        fprintf(stderr, "[gaf2unstable] error: too many arguments\n");
        help(argv);
        return 1;
    }
    if (rgfa_path.empty()) {
        fprintf(stderr, "[gaf2unstable] error: -g option required\n");
        return 1;
    }
    FILE* f = in_gaf_path == "-" ? stdin : fopen(in_gaf_path.c_str(), "rb");
    if (!f) {
        fprintf(stderr, "[gaf2unstable] error: unable to open input: %s\n", in_gaf_path.c_str());
        return 1;
    }
    std::string rgfa;
    if (!read_file(rgfa_path, rgfa)) {
        fprintf(stderr, "[gaf2unstable] error: Could not open %s\n", rgfa_path.c_str());
        return 1;
    }
    {   // the reference maps the file O_RDWR and asserts on failure (gfakluge.hpp:570-573)
        int fd = open(rgfa_path.c_str(), O_RDWR);
        if (fd == -1) abort();
        close(fd);
    }

    g2p_ctx* ctx = nullptr;
    if (g2p_create((int)env_long("G2P_DEVICE", 0), &ctx) != G2P_OK) {
        fprintf(stderr, "[gaf2unstable] error: no usable CUDA device (this build has no CPU path)\n");
        return 1;
    }
    int ref_rc = 0;
    std::vector<char> msg(1 << 16);
    int rc = g2p_load_rgfa(ctx, rgfa.data(), rgfa.size(), &ref_rc, msg.data(), msg.size());
    if (rc == G2P_E_TABLE) {
        fputs(msg.data(), stderr);
        if (ref_rc == 1) return 1;
        abort();
    }
    if (rc != G2P_OK) { fprintf(stderr, "[gaf2unstable] error: %s\n", g2p_last_error(ctx)); return 1; }
    std::string().swap(rgfa);

    if (!node_lengths_path.empty()) {
        FILE* o = fopen(node_lengths_path.c_str(), "wb");
        if (!o) {
            fprintf(stderr, "[gaf2unstable] error: unable to open output: %s\n", node_lengths_path.c_str());
            return 1;
        }
        const char* tsv = nullptr;
        size_t n = 0;
        g2p_rgfa_node_lengths(ctx, &tsv, &n);
        fwrite(tsv, 1, n, o);
        fclose(o);
    }

    const size_t chunk = (size_t)std::max(1L, env_long("G2P_CHUNK_MB", 256)) << 20;
    size_t cap = chunk + (1 << 20);
    char* buf = static_cast<char*>(g2p_host_alloc(cap));
    if (!buf) { fprintf(stderr, "[gaf2unstable] error: cannot allocate pinned host memory\n"); return 1; }
    std::string carry;
    bool eof = false;
    std::vector<char> wmsg(1 << 20);
    while (!eof) {
        if (carry.size() + chunk > cap) {
            size_t want = carry.size() + chunk + (1 << 20);
            char* nb = static_cast<char*>(g2p_host_alloc(want));
            if (!nb) { fprintf(stderr, "[gaf2unstable] error: cannot allocate pinned host memory\n"); return 1; }
            g2p_host_free(buf);
            buf = nb; cap = want;
        }
        memcpy(buf, carry.data(), carry.size());
        size_t have = carry.size();
        carry.clear();
        while (have < cap - 1) {
            size_t want = std::min(chunk, cap - 1 - have);
            size_t k = fread(buf + have, 1, want, f);
            have += k;
            if (k < want) { eof = true; break; }
            if (have >= chunk) break;
        }
        if (!eof) {
            size_t cut = have;
            while (cut > 0 && buf[cut - 1] != '\n') --cut;
            if (cut == 0) {   // a single line longer than the chunk: keep reading it
                carry.assign(buf, have);
                if (carry.size() >= 0xF0000000ULL) { fprintf(stderr, "[gaf2unstable] error: line longer than 4 GiB\n"); return 1; }
                continue;
            }
            carry.assign(buf + cut, have - cut);
            have = cut;
        }
        const char* out = nullptr;
        g2p_result res;
        rc = g2p_unstable_host(ctx, buf, have, &out, &res);
        if (rc != G2P_OK) { fprintf(stderr, "[gaf2unstable] error: GPU conversion failed: %s\n", g2p_last_error(ctx)); return 1; }
        const g2p_warn* warns = nullptr;
        size_t nw = 0;
        g2p_unstable_warnings(ctx, &warns, &nw);
        for (size_t i = 0; i < nw; ++i) {
            if (wmsg.size() < warns[i].out_len + 4096) wmsg.resize(warns[i].out_len + 4096);
            g2p_format_unstable_warning(ctx, out + warns[i].out_off, warns[i].out_len, wmsg.data(), wmsg.size());
            fputs(wmsg.data(), stderr);
        }
        write_all(out, res.out_bytes);
        if (res.rec_status != G2P_REC_OK) {
            fprintf(stderr, "terminate: gaf2unstable cannot convert record %llu of this block (status %u)\n", (unsigned long long)res.err_record,
                    res.rec_status);
            fflush(stderr);
            abort();   // reference: assert / uncaught exception -> SIGABRT
        }
    }
    if (f != stdin) fclose(f);
    g2p_host_free(buf);
    g2p_destroy(ctx);
    return 0;
}
