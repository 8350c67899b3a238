// gaf2unstable — drop-in command line for the reference tool of the same name
// (reference gaf2unstable_main.cpp:177-301): same options (and option quirks), same stdout
// bytes, same stderr messages and exit codes; the per-record rewrite runs on a B200 through
// the C-ABI of libg2p.so (include/g2p.h), the rGFA tables are built once on the host.
//
//   gaf2unstable [options] <gaf>
//     -g, --rgfa FILE   (uncompressed) minigraph rGFA
//     -o FILE           write "node<TAB>length" for every rGFA node (input of gaf2paf -l)
//
// Environment: G2P_DEVICE=K (device ordinal), G2P_CHUNK_MB=M (bytes of GAF per GPU call; default 16 for short records, up to 512 for long ones),
// G2P_IO_THREADS=T.  Host side: the reader -> converter -> writer pipeline of cli_pipeline.hpp.
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

#include "cli_pipeline.hpp"

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
    {   // the reference opens the GAF before it reads the rGFA (gaf2unstable_main.cpp:248-262)
        if (in_gaf_path != "-") {
            int fd = open(in_gaf_path.c_str(), O_RDONLY);
            if (fd == -1) {
                fprintf(stderr, "[gaf2unstable] error: unable to open input: %s\n", in_gaf_path.c_str());
                return 1;
            }
            close(fd);
        }
    }
    std::string rgfa;
    if (!cli::read_file(rgfa_path, rgfa)) {
        fprintf(stderr, "[gaf2unstable] error: Could not open %s\n", rgfa_path.c_str());
        return 1;
    }
    {   // the reference maps the file O_RDWR and asserts on failure (gfakluge.hpp:570-573)
        int fd = open(rgfa_path.c_str(), O_RDWR);
        if (fd == -1) abort();
        close(fd);
    }

    g2p_ctx* ctx = nullptr;
    if (g2p_create((int)cli::env_long("G2P_DEVICE", 0), &ctx) != G2P_OK) {
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

    size_t chunk = (size_t)std::max(1L, cli::env_long("G2P_CHUNK_MB", 16)) << 20;
    if (cli::env_long("G2P_CHUNK_BYTES", 0) > 0) chunk = (size_t)cli::env_long("G2P_CHUNK_BYTES", 0);   // tests: tiny chunks
    cli::Pipeline P;
    P.tool = "gaf2unstable";
    P.chunk_bytes = chunk;
    P.chunk_auto = !getenv("G2P_CHUNK_MB") && !getenv("G2P_CHUNK_BYTES");
    P.ctx.push_back(ctx);
    P.convert = [](g2p_ctx* cx, cli::Chunk& c) {
        int r = g2p_unstable_host(cx, c.buf, c.n, &c.out, &c.res);
        if (r != G2P_OK) return r;
        // the reference prints its multi-contig warning while it converts the record: keep the texts with the chunk
        const g2p_warn* warns = nullptr;
        size_t nw = 0;
        g2p_unstable_warnings(cx, &warns, &nw);
        std::vector<char> wmsg(1 << 16);
        for (size_t i = 0; i < nw; ++i) {
            if (wmsg.size() < warns[i].out_len + 4096) wmsg.resize(warns[i].out_len + 4096);
            g2p_format_unstable_warning(cx, c.out + warns[i].out_off, warns[i].out_len, wmsg.data(), wmsg.size());
            c.err_text += wmsg.data();
        }
        return r;
    };
    P.on_record_error = [](cli::Chunk& c) {
        fprintf(stderr, "terminate: gaf2unstable cannot convert record %llu of this block (status %u)\n", (unsigned long long)c.res.err_record,
                c.res.rec_status);
        fflush(stderr);
        abort();   // reference: assert / uncaught exception -> SIGABRT
    };
    const int prc = P.run({in_gaf_path});
    g2p_destroy(ctx);
    return prc;
}
