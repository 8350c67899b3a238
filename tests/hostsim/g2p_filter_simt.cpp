// g2p_filter_simt — the gaffilter device pipeline (csrc/g2p_filter.cuh: parse, radix sort, running maximum, sweep, emit)
// executed on the CPU by the SIMT emulator.  TEST INFRASTRUCTURE ONLY; the launch sequence mirrors run_filter (g2p_capi.cu).
//
// usage: g2p_filter_simt [-p] [-r R] [-m M] [-o N] [-b N] [-q N] [-i X] <gaf|->      (stdout / stderr / exit code as gaffilter)
#define HS_IMPLEMENTATION
#include "cuda_shim.hpp"

#include <string>

#include "../../cactus-gfa-tools_b200/csrc/g2p_kernels.cuh"

using namespace g2p;

int main(int argc, char** argv) {
    FilterParams P{0, 0, 0, 0, 0, 0, 0};
    const char* path = nullptr;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-p") P.is_paf = 1;
        else if (a == "-r" && i + 1 < argc) P.ratio = std::stof(argv[++i]);
        else if (a == "-m" && i + 1 < argc) P.min_overlap_pct = std::stof(argv[++i]);
        else if (a == "-i" && i + 1 < argc) P.min_identity = std::stof(argv[++i]);
        else if (a == "-o" && i + 1 < argc) P.min_overlap_len = std::stol(argv[++i]);
        else if (a == "-b" && i + 1 < argc) P.min_block_len = std::stol(argv[++i]);
        else if (a == "-q" && i + 1 < argc) P.min_mapq = std::stol(argv[++i]);
        else path = argv[i];
    }
    if (!path) return 1;
    std::string text_s;
    {
        FILE* f = std::strcmp(path, "-") == 0 ? stdin : std::fopen(path, "rb");
        if (!f) return 1;
        char buf[1 << 16];
        size_t k;
        while ((k = std::fread(buf, 1, sizeof buf, f)) > 0) text_s.append(buf, k);
    }
    const u64 n = text_s.size();
    std::vector<uint4> tb((n + 15) / 16 + 4);
    u8* text = reinterpret_cast<u8*>(tb.data());
    std::memcpy(text, text_s.data(), n);
    PipelineMeta meta;
    std::memset(&meta, 0, sizeof meta);
    const u32 ntiles = (u32)((n + kIdxTile - 1) / kIdxTile);
    std::vector<u32> tiles(ntiles + 1);
    if (ntiles) hs::launch(dim3(ntiles), dim3(kIdxThreads), 0, [&] { k_count_lines(text, n, tiles.data()); });
    hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan_tiles(tiles.data(), ntiles, text, n, &meta); });
    const u32 nrec = meta.n_records;
    std::vector<u32> rec(nrec + 2, 0);
    if (ntiles) hs::launch(dim3(ntiles), dim3(kIdxThreads), 0, [&] { k_fill_lines(text, n, tiles.data(), rec.data(), &meta); });
    FilterMeta fm;
    std::memset(&fm, 0, sizeof fm);
    fm.first_err = 0xFFFFFFFFu;
    fm.assert_rec = 0xFFFFFFFFu;
    if (nrec) {
        const u32 nblk = (nrec + kRsTile - 1) / kRsTile, nh = 16u * nblk;
        const u32 nscan_h = (nh + kScanTile - 1) / kScanTile, nscan_r = (nrec + kScanTile - 1) / kScanTile;
        std::vector<FRow> rows(nrec);
        std::vector<u64> k1(nrec), k2(nrec), hist(nh + 2), blocks(std::max(nscan_h, nscan_r) * 2), off(nrec + 1);
        std::vector<u32> v1(nrec), v2(nrec);
        std::vector<i64> pmax(nrec);
        std::vector<u8> keep(nrec, 0);
        u64 *keys = k1.data(), *keys2 = k2.data();
        u32 *vals = v1.data(), *vals2 = v2.data();
        FilterArgs fa{text, rec.data(), nrec, P, rows.data(), keys, vals, &fm};
        hs::launch(dim3(4), dim3(128), 0, [&] { k_filter_parse(fa); });
        for (u32 shift = 0; shift < 64; shift += 4) {
            hs::launch(dim3(nblk), dim3(kRsThreads), 0, [&] { k_rs_hist(keys, nrec, shift, hist.data(), nblk); });
            hs::launch(dim3(nscan_h), dim3(kScanThreads), 0, [&] { k_scan_reduce(hist.data(), nh, blocks.data()); });
            hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan_blocks(blocks.data(), nscan_h, hist.data() + nh + 1); });
            hs::launch(dim3(nscan_h), dim3(kScanThreads), 0, [&] { k_scan_apply(hist.data(), nh, blocks.data(), hist.data() + nh + 1); });
            hs::launch(dim3(nblk), dim3(kRsThreads), 0, [&] { k_rs_scatter(keys, vals, keys2, vals2, nrec, shift, hist.data(), nblk); });
            std::swap(keys, keys2);
            std::swap(vals, vals2);
        }
        for (u32 i = 1; i < nrec; ++i) if (keys[i - 1] > keys[i]) { std::fprintf(stderr, "g2p_filter_simt: sort order broken at %u\n", i); return 99; }
        {
            const u32 nseg = (nrec + kSegTile - 1) / kSegTile;
            std::vector<i64> tile_val(nseg), carry(nseg);
            std::vector<u32> tile_flag(nseg);
            hs::launch(dim3(4), dim3(256), 0, [&] { k_filter_ends(keys, vals, rows.data(), nrec, pmax.data()); });
            hs::launch(dim3(nseg), dim3(kSegThreads), 0, [&] { k_segmax<false>(keys, nrec, pmax.data(), tile_val.data(), tile_flag.data(), nullptr); });
            hs::launch(dim3(1), dim3(32), 0, [&] { k_segmax_tiles(tile_val.data(), tile_flag.data(), nseg, carry.data()); });
            hs::launch(dim3(nseg), dim3(kSegThreads), 0, [&] { k_segmax<true>(keys, nrec, pmax.data(), tile_val.data(), tile_flag.data(), carry.data()); });
        }
        hs::launch(dim3(4), dim3(128), 0, [&] { k_filter_sweep(keys, vals, rows.data(), pmax.data(), nrec, P, keep.data(), &fm); });
        hs::launch(dim3(4), dim3(128), 0, [&] { k_filter_emit<false>(text, rec.data(), nrec, P.is_paf, keep.data(), off.data(), nullptr); });
        hs::launch(dim3(nscan_r), dim3(kScanThreads), 0, [&] { k_scan_reduce(off.data(), nrec, blocks.data()); });
        hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan_blocks(blocks.data(), nscan_r, &fm.out_total); });
        hs::launch(dim3(nscan_r), dim3(kScanThreads), 0, [&] { k_scan_apply(off.data(), nrec, blocks.data(), &fm.out_total); });
        if (fm.first_err == 0xFFFFFFFFu && fm.unsupported) { std::fprintf(stderr, "unsupported query_start\n"); return 98; }
        if (fm.first_err != 0xFFFFFFFFu) {
            hs::launch(dim3(1), dim3(1), 0, [&] { k_filter_diagnose(fa); });
            std::fprintf(stderr, "abort: line %u status %u\n", fm.first_err, fm.err_status & 0xff);
            return 134;
        }
        if (fm.assert_rec != 0xFFFFFFFFu) { std::fprintf(stderr, "abort: assertion of the filter loop on record %u\n", fm.assert_rec); return 134; }
        std::vector<u8> out(fm.out_total + 64);
        hs::launch(dim3(4), dim3(128), 0, [&] { k_filter_emit<true>(text, rec.data(), nrec, P.is_paf, keep.data(), off.data(), out.data()); });
        std::fprintf(stderr, "[gaffilter]: Loaded %u %s records\n[gaffilter]: Constructed interval trees\n", fm.n_loaded, P.is_paf ? "PAF" : "GAF");
        std::fwrite(out.data(), 1, fm.out_total, stdout);
    } else std::fprintf(stderr, "[gaffilter]: Loaded 0 %s records\n[gaffilter]: Constructed interval trees\n", P.is_paf ? "PAF" : "GAF");
    std::fprintf(stderr, "[gaffilter]: filtered %u / %u. total block lengths filtered: %lld\n", fm.n_filtered, fm.n_loaded, (long long)fm.filtered_len);
    return 0;
}
