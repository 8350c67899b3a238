// gaffilter — drop-in command line for the reference tool of the same name (reference gaffilter_main.cpp:68-352):
// same options, same stdout bytes, same stderr lines and exit codes; loading, overlap search and the dominance test
// run on a B200 through g2p_filter_host (include/g2p.h, csrc/g2p_filter.cuh).
//
//   gaffilter [options] <gaf> > output.gaf
//     -r N  -m N  -o N  -q N  -b N  -i N  -p        (see help)
//
// Environment: G2P_DEVICE=K (device ordinal).  The whole input is one call (< 4 GiB), like the reference, which holds
// every record in memory before it prints the first one.
#include <getopt.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "cli_pipeline.hpp"

using namespace std;

static void help(char** argv) {
    cerr << "usage: " << argv[0] << " [options] <gaf> > output.gaf" << endl
         << "Filter GAF record if its query interval overlaps another query interval and\n"
         << "  1) the record is secondary and the overlapping record is primary or\n"
         << "  2) the record's MAPQ is lower than {ratio, see -r} times the overlapping record's MAPQ or\n"
         << "  3) the record's block length is less than {ratio, see -r} times larger than the overlapping record's block length (and its MAPQ isn't higher)" << endl
         << "  Also: the -o option can be used to mimic mzgaf2paf's query overlap filter" << endl
         << endl
         << "options: " << endl
         << "    -r, --ratio N                   If two query blocks overlap, and one is Nx bigger than the other, the bigger one is kept (otherwise both deleted) [0]" << endl
         << "    -m, --min-overlap N             Ignore overlaps that consitute <N% of the length [0]" << endl
         << "    -o, --min-overlap-length N      If >= 2 query regions with size >= N overlap, ignore the query region.  If 1 query region with size >= N overlaps any regions of size <= N, ignore the smaller ones only. Works separate to -r/-m but can be used in conjunction with them to combine the two filters (0 = disable) [0]" << endl
         << "    -q, --min-mapq N                Don't let an interval with MAPQ < N cause something to be filtered out" << endl
         << "    -b, --min-block-length N        Don't let an interval with block length < N cause something to be filtered out" << endl
         << "    -i, --min-identity N            Don't let an interval with identity < N cause something to be filtered out" << endl
         << "    -p, --paf                       Input is PAF, not GAF" << endl;
}

int main(int argc, char** argv) {
    g2p_filter_params P;
    memset(&P, 0, sizeof P);
    int c;
    optind = 1;
    while (true) {
        static const struct option long_options[] = {
            {"help", no_argument, 0, 'h'},
            {"ratio", required_argument, 0, 'r'},
            {"min-overlap", required_argument, 0, 'm'},
            {"min-overlap-length", required_argument, 0, 'o'},
            {"min-block-length", required_argument, 0, 'b'},
            {"min-mapq", required_argument, 0, 'q'},
            {"min-identity", required_argument, 0, 'i'},
            {"paf", no_argument, 0, 'p'},
            {0, 0, 0, 0}};
        int option_index = 0;
        c = getopt_long(argc, argv, "h:r:m:po:b:q:i:", long_options, &option_index);
        if (c == -1) break;
        switch (c) {
            case 'r': P.ratio = stof(optarg); break;             // (std::stof like the reference: a float widened to double)
            case 'm': P.min_overlap_pct = stof(optarg); break;
            case 'o': P.min_overlap_len = std::stol(optarg); break;
            case 'p': P.is_paf = 1; break;
            case 'b': P.min_block_len = std::stol(optarg); break;
            case 'i': P.min_identity = std::stof(optarg); break;
            case 'q': P.min_mapq = std::stol(optarg); break;
            case 'h':
            case '?':
                help(argv);
                exit(1);
            default: abort();
        }
    }
    if (argc <= 1) { help(argv); return 1; }
    if (P.ratio == 0 && P.min_overlap_len == 0) {
        cerr << "[gaffilter] error: at least one of -r or -o must be used to specify filter" << endl;
        return 1;
    }
    if (optind >= argc) {
        cerr << "[gaffilter] error: too few arguments" << endl;
        help(argv);
        return 1;
    }
    const string gaf_path = argv[optind++];
    string text;
    if (gaf_path == "-") {
        char buf[1 << 16];
        size_t k;
        while ((k = fread(buf, 1, sizeof buf, stdin)) > 0) text.append(buf, k);
    } else if (!cli::read_file(gaf_path, text)) {
        cerr << "[gaffilter] error: unable to open input: " << gaf_path << endl;
        return 1;
    }
    g2p_ctx* ctx = nullptr;
    if (g2p_create((int)cli::env_long("G2P_DEVICE", 0), &ctx) != G2P_OK) {
        cerr << "[gaffilter] error: no usable CUDA device (this build has no CPU path)" << endl;
        return 1;
    }
    const char* out = nullptr;
    g2p_filter_result res;
    const int rc = g2p_filter_host(ctx, text.data(), text.size(), &P, &out, &res);
    if (rc != G2P_OK) { cerr << "[gaffilter] error: " << g2p_last_error(ctx) << endl; return 1; }
    if (res.rec_status != G2P_REC_OK) {
        // the reference dies while it loads the records (assert / uncaught exception): nothing was printed yet
        cerr << "terminate: gaffilter cannot parse line " << res.err_record << " (status " << res.rec_status << ")" << endl;
        abort();
    }
    cerr << "[gaffilter]: Loaded " << res.n_loaded << (P.is_paf ? " PAF" : " GAF") << " records" << endl;
    cerr << "[gaffilter]: Constructed interval trees" << endl;
    cli::write_seq(1, out, res.out_bytes);
    cerr << "[gaffilter]: filtered " << res.n_filtered << " / " << res.n_loaded << ". total block lengths filtered: " << (long long)res.filtered_len << endl;   // (an int64_t in the reference: a sum beyond 2^63 prints negative)
    g2p_destroy(ctx);
    return 0;
}
