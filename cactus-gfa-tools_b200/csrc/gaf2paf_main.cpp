// gaf2paf — drop-in command line for the reference tool of the same name
// (reference gaf2paf_main.cpp:266-377): same options, same stdout bytes, same stderr
// messages and exit codes; the per-record work runs on B200 GPUs through the C-ABI of
// libg2p.so (include/g2p.h).
//
//   gaf2paf [options] <gaf> [gaf2] [gaf3] [...] > output.paf
//     -l, --lengths FILE   TSV with contig length as first two columns (.fai will do)
//
// Environment (argv stays identical to the reference):
//   G2P_GPUS=N        shard the input over the first N GPUs of the box by newline-aligned
//                     byte ranges; outputs are concatenated in input order (default 1)
//   G2P_DEVICE=K      first device ordinal (default 0)
//   G2P_CHUNK_MB=M    bytes of GAF per GPU call (default 256)
//   G2P_STATS=1       timing summary on stderr
#include <getopt.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/g2p.h"

namespace {

void help(char** argv) {
    fprintf(stderr,
            "usage: %s [options] <gaf> [gaf2] [gaf3] [...] > output.paf\n"
            "Convert minigraph GAF to PAF\n"
            "\n"
            "options: \n"
            "    -l, --lengths FILE      TSV with contig length as first two columns (.fai will do).\n",
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
            _exit(1);   // downstream closed: nothing sensible left to do
        }
        p += k;
        n -= (size_t)k;
    }
}

struct Gpu {
    g2p_ctx* ctx = nullptr;
    char* buf = nullptr;     // pinned input chunk
    size_t cap = 0, n = 0;
    const char* out = nullptr;
    g2p_result res;
    int rc = 0;
};

long env_long(const char* k, long dflt) {
    const char* v = getenv(k);
    return v && *v ? strtol(v, nullptr, 10) : dflt;
}

}  // namespace

int main(int argc, char** argv) {
    std::string lengths_path;
    int c;
    optind = 1;
    while (true) {
        static const struct option long_options[] = {{"help", no_argument, 0, 'h'}, {"lengths", required_argument, 0, 'l'}, {0, 0, 0, 0}};
        int option_index = 0;
        c = getopt_long(argc, argv, "h:l:", long_options, &option_index);
        if (c == -1) break;
        switch (c) {
            case 'l': lengths_path = optarg; break;
            case 'h':
            case '?':
                help(argv);
                exit(1);
            default: abort();
        }
    }
    if (argc <= 1) { help(argv); return 1; }
    if (optind >= argc) {
        fprintf(stderr, "[gaf2paf] error: too few arguments\n");
        help(argv);
        return 1;
    }
    if (lengths_path.empty()) {
        fprintf(stderr, "[gaf2paf] error: -l must be specified to produce valid PAF\n");
        return 1;
    }
    std::string tsv;
    if (!read_file(lengths_path, tsv)) {
        fprintf(stderr, "[gaf2paf] error: unable to open %s\n", lengths_path.c_str());
        return 1;
    }
    std::vector<std::string> in_paths;
    while (optind < argc) in_paths.push_back(argv[optind++]);

    const int ngpu = (int)std::max(1L, env_long("G2P_GPUS", 1));
    const int dev0 = (int)env_long("G2P_DEVICE", 0);
    const size_t chunk = (size_t)std::max(1L, env_long("G2P_CHUNK_MB", 256)) << 20;
    const bool stats = env_long("G2P_STATS", 0) != 0;

    std::vector<Gpu> gpus(ngpu);
    for (int g = 0; g < ngpu; ++g) {
        int rc = g2p_create(dev0 + g, &gpus[g].ctx);
        if (rc != G2P_OK) {
            fprintf(stderr, "[gaf2paf] error: no usable CUDA device %d (this build has no CPU path)\n", dev0 + g);
            return 1;
        }
        rc = g2p_load_lengths(gpus[g].ctx, tsv.data(), tsv.size());
        if (rc == G2P_E_TABLE) {
            // get_len_map: std::stol throws -> terminate (reference gaf2paf_main.cpp:35)
            fprintf(stderr, "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  stol\n");
            abort();
        }
        if (rc != G2P_OK) { fprintf(stderr, "[gaf2paf] error: %s\n", g2p_last_error(gpus[g].ctx)); return 1; }
        gpus[g].cap = chunk + (1 << 20);
        gpus[g].buf = static_cast<char*>(g2p_host_alloc(gpus[g].cap));
        if (!gpus[g].buf) { fprintf(stderr, "[gaf2paf] error: cannot allocate pinned host memory\n"); return 1; }
    }

    double t_gpu_ms = 0;
    uint64_t tot_rec = 0, tot_in = 0, tot_out = 0;
    auto t0 = std::chrono::steady_clock::now();

    for (const std::string& in_path : in_paths) {
        FILE* f = in_path == "-" ? stdin : fopen(in_path.c_str(), "rb");
        if (!f) {
            fprintf(stderr, "[gaf2paf] error: unable to open input: %s\n", in_path.c_str());
            return 1;
        }
        std::string carry;   // bytes after the last newline of the previous chunk
        bool eof = false;
        while (!eof) {
            // fill one newline-aligned chunk per GPU
            int used = 0;
            for (int g = 0; g < ngpu && !eof; ++g) {
                Gpu& G = gpus[g];
                G.n = 0;
                if (carry.size() + chunk > G.cap) {
                    size_t want = carry.size() + chunk + (1 << 20);
                    char* nb = static_cast<char*>(g2p_host_alloc(want));
                    if (!nb) { fprintf(stderr, "[gaf2paf] error: cannot allocate pinned host memory\n"); return 1; }
                    g2p_host_free(G.buf);
                    G.buf = nb; G.cap = want;
                }
                memcpy(G.buf, carry.data(), carry.size());
                size_t have = carry.size();
                carry.clear();
                while (have < G.cap - 1) {
                    size_t want = std::min(chunk, G.cap - 1 - have);
                    size_t k = fread(G.buf + have, 1, want, f);
                    have += k;
                    if (k < want) { eof = true; break; }
                    if (have >= chunk) break;
                }
                if (!eof) {
                    // cut after the last newline; the rest starts the next chunk
                    size_t cut = have;
                    while (cut > 0 && G.buf[cut - 1] != '\n') --cut;
                    if (cut == 0) {
                        // a single line longer than the chunk: keep reading it
                        carry.assign(G.buf, have);
                        --g;
                        if (carry.size() >= 0xF0000000ULL) { fprintf(stderr, "[gaf2paf] error: line longer than 4 GiB\n"); return 1; }
                        continue;
                    }
                    carry.assign(G.buf + cut, have - cut);
                    have = cut;
                }
                G.n = have;
                used = g + 1;
            }
            // convert the chunks concurrently, one host thread per GPU
            std::vector<std::thread> th;
            for (int g = 0; g < used; ++g) {
                th.emplace_back([&, g]() {
                    Gpu& G = gpus[g];
                    G.rc = g2p_convert_host(G.ctx, G.buf, G.n, &G.out, &G.res);
                });
            }
            for (auto& t : th) t.join();
            // emit in input order; stop at the first failing record like the reference
            for (int g = 0; g < used; ++g) {
                Gpu& G = gpus[g];
                if (G.rc != G2P_OK) {
                    fprintf(stderr, "[gaf2paf] error: GPU conversion failed: %s\n", g2p_last_error(G.ctx));
                    return 1;
                }
                write_all(G.out, G.res.out_bytes);
                t_gpu_ms += G.res.device_ms;
                tot_rec += G.res.n_records; tot_in += G.n; tot_out += G.res.out_bytes;
                if (G.res.rec_status != G2P_REC_OK) {
                    char msg[1024];
                    g2p_format_error(&G.res, G.buf, G.n, msg, sizeof msg);
                    fputs(msg, stderr);
                    fflush(stderr);
                    if (G.res.rec_status >= G2P_REC_ABORT) abort();   // reference: SIGABRT (assert / uncaught exception)
                    return 1;
                }
            }
        }
        if (f != stdin) fclose(f);
    }
    if (stats) {
        double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[gaf2paf] stats: records=%llu in=%llu B out=%llu B device=%.3f ms wall=%.3f s gpus=%d\n",
                (unsigned long long)tot_rec, (unsigned long long)tot_in, (unsigned long long)tot_out, t_gpu_ms, wall, ngpu);
    }
    for (auto& G : gpus) { g2p_host_free(G.buf); g2p_destroy(G.ctx); }
    return 0;
}
