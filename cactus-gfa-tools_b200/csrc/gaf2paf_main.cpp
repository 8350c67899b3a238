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
//   G2P_CHUNK_MB=M    bytes of GAF per GPU call (default: 16 for short records, up to 512 for long ones)
//   G2P_IO_THREADS=T  threads of the parallel pread / pwrite of regular files (default: half the cores, <= 16)
//   G2P_STATS=1       timing summary on stderr
//
// The host side is a reader -> per-GPU converter -> writer pipeline (cli_pipeline.hpp): the file reads,
// the GPU calls and the stdout writes of consecutive chunks overlap, and stdout stays in input order.
#include <getopt.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "cli_pipeline.hpp"

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
    if (!cli::read_file(lengths_path, tsv)) {
        fprintf(stderr, "[gaf2paf] error: unable to open %s\n", lengths_path.c_str());
        return 1;
    }
    std::vector<std::string> in_paths;
    while (optind < argc) in_paths.push_back(argv[optind++]);

    const int ngpu = (int)std::max(1L, cli::env_long("G2P_GPUS", 1));
    const int dev0 = (int)cli::env_long("G2P_DEVICE", 0);
    size_t chunk = (size_t)std::max(1L, cli::env_long("G2P_CHUNK_MB", 16)) << 20;
    if (cli::env_long("G2P_CHUNK_BYTES", 0) > 0) chunk = (size_t)cli::env_long("G2P_CHUNK_BYTES", 0);   // tests: tiny chunks
    const bool stats = cli::env_long("G2P_STATS", 0) != 0;

    cli::Pipeline P;
    P.tool = "gaf2paf";
    P.chunk_bytes = chunk;
    P.chunk_auto = !getenv("G2P_CHUNK_MB") && !getenv("G2P_CHUNK_BYTES");
    // one context per GPU, created concurrently (CUDA context creation is the bulk of the start-up time)
    P.ctx.assign(ngpu, nullptr);
    std::vector<int> crc(ngpu, G2P_OK), lrc(ngpu, G2P_OK);
    {
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g)
            th.emplace_back([&, g] {
                crc[g] = g2p_create(dev0 + g, &P.ctx[g]);
                if (crc[g] == G2P_OK) lrc[g] = g2p_load_lengths(P.ctx[g], tsv.data(), tsv.size());
            });
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < ngpu; ++g) {
        if (crc[g] != G2P_OK) {
            fprintf(stderr, "[gaf2paf] error: no usable CUDA device %d (this build has no CPU path)\n", dev0 + g);
            return 1;
        }
        if (lrc[g] == G2P_E_TABLE) {
            // get_len_map: std::stol throws -> terminate (reference gaf2paf_main.cpp:35)
            fprintf(stderr, "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  stol\n");
            abort();
        }
        if (lrc[g] != G2P_OK) { fprintf(stderr, "[gaf2paf] error: %s\n", g2p_last_error(P.ctx[g])); return 1; }
    }
    P.convert = [](g2p_ctx* ctx, cli::Chunk& c) { return g2p_convert_host(ctx, c.buf, c.n, &c.out, &c.res); };
    P.on_record_error = [](cli::Chunk& c) {
        // everything the reference had written before it stopped is on stdout; now its message and exit status
        char msg[1024];
        g2p_format_error(&c.res, c.buf, c.n, msg, sizeof msg);
        fputs(msg, stderr);
        fflush(stderr);
        if (c.res.rec_status >= G2P_REC_ABORT) abort();   // reference: SIGABRT (assert / uncaught exception)
        _exit(1);
    };
    auto t0 = std::chrono::steady_clock::now();
    if (stats) fprintf(stderr, "[gaf2paf] stats: contexts ready\n");
    const int rc = P.run(in_paths);
    if (stats) {
        double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[gaf2paf] stats: records=%llu in=%llu B out=%llu B device=%.3f ms wall=%.3f s gpus=%d\n",
                (unsigned long long)P.tot_rec, (unsigned long long)P.tot_in, (unsigned long long)P.tot_out, P.device_ms, wall, ngpu);
    }
    for (g2p_ctx* ctx : P.ctx) g2p_destroy(ctx);
    return rc;
}
