// g2p_simt — the device pipeline (k_fuse; else line index, k_rec / k_long / k_convert_list, scans, k_emit_lines)
// executed on the CPU by the SIMT emulator in cuda_shim.hpp, with the library's default configuration
// (G2P_FUSE=0, G2P_ONE_PASS_INDEX=1, G2P_LEN_SORT=1, G2P_SIZE_KERNEL=short select the variants, as in g2p_create).
//
// TEST INFRASTRUCTURE ONLY (never linked into libg2p.so or the executables).  The launch
// sequence below mirrors g2p_convert_device (g2p_capi.cu); the kernels are the product's own
// sources compiled for the host.  It lets CPU-only CI run the warp-cooperative kernels against
// the oracle, and lets a kernel bug be debugged with gdb/ASan instead of GPU round trips.
//
// usage: g2p_simt -l lengths.tsv <gaf|-> [gaf2 ...]     (stdout / exit code as gaf2paf)
//        G2P_SIMT_STATS=1 prints "records delegated" to stderr
#define HS_IMPLEMENTATION
#include "cuda_shim.hpp"

#include <string>

#include "../../cactus-gfa-tools_b200/csrc/g2p_kernels.cuh"
#include "../../cactus-gfa-tools_b200/csrc/g2p_table.hpp"

using namespace g2p;

static bool slurp(const char* path, std::string& out) {
    FILE* f = std::strcmp(path, "-") == 0 ? stdin : std::fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t k;
    while ((k = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    if (f != stdin) std::fclose(f);
    return true;
}

int main(int argc, char** argv) {
    const char* lengths = nullptr;
    std::vector<const char*> inputs;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "-l") && i + 1 < argc) lengths = argv[++i];
        else inputs.push_back(argv[i]);
    }
    if (!lengths || inputs.empty()) { std::fprintf(stderr, "usage: g2p_simt -l lengths.tsv <gaf> ...\n"); return 1; }
    std::string tsv;
    if (!slurp(lengths, tsv)) { std::fprintf(stderr, "[gaf2paf] error: unable to open %s\n", lengths); return 1; }
    HostLenTable table;
    if (build_len_table(tsv.data(), tsv.size(), table) != ST_OK) { std::fprintf(stderr, "abort: lengths table\n"); return 134; }
    const LenTableView T = table.view();

    std::string gaf_s;
    for (size_t i = 0; i < inputs.size(); ++i) {
        if (!slurp(inputs[i], gaf_s)) { std::fprintf(stderr, "[gaf2paf] error: unable to open input: %s\n", inputs[i]); return 1; }
        // a line never spans two files; the LAST file keeps an unterminated last line as it is (rec_start[nrec] = n + 1,
        // virtual newline), like the library sees it
        if (i + 1 < inputs.size() && !gaf_s.empty() && gaf_s.back() != '\n') gaf_s.push_back('\n');
    }
    const u64 n = gaf_s.size();
    // 16-byte aligned copy with slack, like a device allocation
    std::vector<uint4> gaf_buf((n + 15) / 16 + 4);
    u8* gaf = reinterpret_cast<u8*>(gaf_buf.data());
    std::memcpy(gaf, gaf_s.data(), n);

    // ---- the one-pass kernel (run_fused of g2p_capi.cu).  G2P_FUSE as in g2p_create: 0 never, 2 always, default: when the
    // mean record length exceeds 200 bytes
    const int fuse_mode = std::getenv("G2P_FUSE") ? std::atoi(std::getenv("G2P_FUSE")) : 1;
    u64 n_nl = 0;
    for (u64 i = 0; i < n; ++i) n_nl += gaf[i] == '\n';
    if (n && (fuse_mode == 2 || (fuse_mode == 1 && n / (n_nl ? n_nl : 1) > 200))) {
        int cfg = std::getenv("G2P_FUSE_CFG") ? std::atoi(std::getenv("G2P_FUSE_CFG")) : 6;
        u64 cap = std::getenv("G2P_FUSE_OUT_CAP") ? (u64)std::atoll(std::getenv("G2P_FUSE_OUT_CAP")) : n * 3 + (1u << 20);
        bool grown = false;
        for (;;) {
            const u32 tile = fuse_cfg_tile(cfg);
            const u32 ftiles = (u32)((n + tile - 1) / tile);
            std::vector<u64> fstat(ftiles + 1, 0);
            u32 fticket = 0;
            FuseMeta fm;
            std::memset(&fm, 0, sizeof fm);
            std::vector<u8> fout(cap + 256, 0xEE);
            FuseArgs fa{gaf, n, ftiles, T, fout.data(), cap, fstat.data(), &fticket, &fm};
            switch (cfg) {
                case 0: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg0::kSmem, [&] { k_fuse<FuseCfg0>(fa); }); break;
                case 1: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg1::kSmem, [&] { k_fuse<FuseCfg1>(fa); }); break;
                case 2: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg2::kSmem, [&] { k_fuse<FuseCfg2>(fa); }); break;
                case 3: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg3::kSmem, [&] { k_fuse<FuseCfg3>(fa); }); break;
                case 4: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg4::kSmem, [&] { k_fuse<FuseCfg4>(fa); }); break;
                case 5: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg5::kSmem, [&] { k_fuse<FuseCfg5>(fa); }); break;
                case 6: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg6::kSmem, [&] { k_fuse<FuseCfg6>(fa); }); break;
                case 7: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg7::kSmem, [&] { k_fuse<FuseCfg7>(fa); }); break;
                default: hs::launch(dim3(ftiles), dim3(kFThreads), FuseCfg8::kSmem, [&] { k_fuse<FuseCfg8>(fa); }); break;
            }
            if (fm.fallback) {
                if (!(fm.fallback & kFuseNotConvertible) && fuse_cfg_denser(cfg) >= 0) { cfg = fuse_cfg_denser(cfg); continue; }
                break;
            }
            if (!fm.overflow) {
                if (std::getenv("G2P_SIMT_STATS")) std::fprintf(stderr, "g2p_simt: k_fuse converted %u records, %u lines, %llu bytes out (%u tiles of %u bytes)\n", fm.n_records, fm.n_lines, (unsigned long long)fm.out_total, ftiles, tile);
                std::fwrite(fout.data(), 1, fm.out_total, stdout);
                std::fflush(stdout);
                return 0;
            }
            if (grown) break;
            cap = fm.out_total + 256;
            grown = true;
        }
        if (std::getenv("G2P_SIMT_STATS")) std::fprintf(stderr, "g2p_simt: k_fuse fell back to the general pipeline\n");
    }

    PipelineMeta meta;
    std::memset(&meta, 0, sizeof meta);
    const u32 ntiles = (u32)((n + kIdxTile - 1) / kIdxTile);
    std::vector<u32> rec;
    u32 nrec = 0;
    bool indexed = false;
    if (ntiles && std::getenv("G2P_ONE_PASS_INDEX") && std::atoi(std::getenv("G2P_ONE_PASS_INDEX")) != 0) {   // default: the counting kernels, like the library
        const u64 cap = std::getenv("G2P_SIMT_INDEX_CAP") ? (u64)std::atol(std::getenv("G2P_SIMT_INDEX_CAP")) : n / 32 + 1024;
        rec.assign(cap + 2, 0xDEADBEEFu);
        std::vector<u64> tstat(ntiles + 1, 0);
        u32* ticket = reinterpret_cast<u32*>(&tstat[ntiles]);
        hs::launch(dim3(ntiles), dim3(kIdxThreads), 0, [&] { k_index1(gaf, n, ntiles, tstat.data(), ticket, rec.data(), (u32)cap, &meta); });
        nrec = meta.n_records;
        indexed = (u64)nrec + 2 <= cap;
    }
    if (!indexed) {
        std::vector<u32> tiles(ntiles + 1);
        if (ntiles) hs::launch(dim3(ntiles), dim3(kIdxThreads), 0, [&] { k_count_lines(gaf, n, tiles.data()); });
        hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan_tiles(tiles.data(), ntiles, gaf, n, &meta); });
        nrec = meta.n_records;
        rec.assign(nrec + 2, 0);
        if (ntiles) hs::launch(dim3(ntiles), dim3(kIdxThreads), 0, [&] { k_fill_lines(gaf, n, tiles.data(), rec.data(), &meta); });
    }
    if (nrec == 0) return 0;

    std::vector<u32> status(nrec), list(nrec), list2(nrec);
    std::vector<u64> off(nrec + 1);
    const u32 nscan = (nrec + kScanTile - 1) / kScanTile;
    std::vector<u64> blocks(nscan);
    const u32 ncta = (nrec + kShortRecsPerCta - 1) / kShortRecsPerCta;
    const u32 nlist = std::min<u32>((nrec + kListThreads - 1) / kListThreads, 8u);
    // G2P_SIMT_DESC_CAP=<slots> shrinks k_long's descriptor array to exercise the overflow fallback
    u32 desc_cap = std::getenv("G2P_SIMT_DESC_CAP") ? (u32)std::atol(std::getenv("G2P_SIMT_DESC_CAP")) : (u32)(n / 8) + 2048;
    std::vector<LineDesc> desc(desc_cap + 1), sdesc((size_t)nrec * kSMaxLines);
    std::vector<RecDesc> rdesc(nrec);
    std::vector<u64> loff(nrec + 1);
    ShortArgs sa{gaf, n, rec.data(), nrec, T, off.data(), loff.data(), status.data(), list.data(), &meta.n_deleg, sdesc.data(), rdesc.data()};
    // G2P_SIZE_KERNEL=short selects the 8-lanes-per-record kernel, like the library; default: thread-per-record k_rec
    const char* sk = std::getenv("G2P_SIZE_KERNEL");
    if (sk && !std::strcmp(sk, "short")) hs::launch(dim3(ncta), dim3(kSThreads), kShortSmem, [&] { k_short<kSG>(sa); });
    else {
        const u32 chunks = std::getenv("G2P_REC_CHUNKS") ? (u32)std::atoi(std::getenv("G2P_REC_CHUNKS")) : rec_chunks_for(n, nrec);
        std::vector<u32> perm(nrec);
        const u32* permp = nullptr;
        if (std::getenv("G2P_LEN_SORT") && std::atoi(std::getenv("G2P_LEN_SORT")) != 0) {   // off by default, as run_pipeline (g2p_capi.cu)
            const u32 nsort = (nrec + kLenSortRecs - 1) / kLenSortRecs, nm = kLenBins * nsort;
            const u32 nscan_m = (nm + kScanTile - 1) / kScanTile;
            std::vector<u64> m((size_t)nm + 1 + nscan_m);
            u64* mb = m.data() + nm + 1;
            hs::launch(dim3(nsort), dim3(256), 0, [&] { k_len_hist(rec.data(), nrec, nsort, m.data()); });
            hs::launch(dim3(nscan_m), dim3(kScanThreads), 0, [&] { k_scan_reduce(m.data(), nm, mb); });
            hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan_blocks(mb, nscan_m, m.data() + nm); });
            hs::launch(dim3(nscan_m), dim3(kScanThreads), 0, [&] { k_scan_apply(m.data(), nm, mb, m.data() + nm); });
            hs::launch(dim3(nsort), dim3(256), 0, [&] { k_len_scatter(rec.data(), nrec, nsort, m.data(), perm.data()); });
            permp = perm.data();
        }
        RecArgs ra{sa, chunks, permp};
        hs::launch(dim3((nrec + kRThreads - 1) / kRThreads), dim3(kRThreads), rec_smem(chunks), [&] { k_rec(ra); });
    }
    LongArgs la{gaf, n, rec.data(), T, off.data(), status.data(), nullptr, list.data(), &meta.n_deleg, list2.data(), &meta.n_deleg2,
                desc.data(), rdesc.data(), &meta.n_desc, &meta.n_desc2, desc_cap, 8u, &meta.legacy_long, &meta.long_cursor};
    // ---- the token-parallel kernels (run_par of g2p_capi.cu) take k_rec's delegates first; G2P_PAR=0: k_long takes them all
    std::vector<u32> list3(nrec);
    u32 par_steps = 0, par_ops = 0;
    if (meta.n_deleg && !(std::getenv("G2P_PAR") && std::atoi(std::getenv("G2P_PAR")) == 0)) do {
        auto scan64 = [&](u64* x, u32 cnt) {
            const u32 nb = (cnt + kScanTile - 1) / kScanTile;
            std::vector<u64> bs(nb + 1);
            hs::launch(dim3(nb), dim3(kScanThreads), 0, [&] { k_scan_reduce(x, cnt, bs.data()); });
            hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan_blocks(bs.data(), nb, x + cnt); });
            hs::launch(dim3(nb), dim3(kScanThreads), 0, [&] { k_scan_apply(x, cnt, bs.data(), x + cnt); });
        };
        auto scan4 = [&](uint4* x, u32 cnt) {
            const u32 nb = (cnt + kScanTile - 1) / kScanTile;
            std::vector<uint4> bs(nb + 1);
            hs::launch(dim3(nb), dim3(kScanThreads), 0, [&] { k_scan4_reduce(x, cnt, bs.data()); });
            hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan4_blocks(bs.data(), nb, x + cnt); });
            hs::launch(dim3(nb), dim3(kScanThreads), 0, [&] { k_scan4_apply(x, cnt, bs.data()); });
        };
        const u32 npl = meta.n_deleg;
        std::vector<ParRec> precs(npl);
        std::vector<u64> tile_base(npl + 1), slot_scan(npl + 1);
        ParArgs pa;
        std::memset(&pa, 0, sizeof pa);
        pa.gaf = gaf; pa.n = n; pa.rec_start = rec.data(); pa.T = T; pa.list = list.data(); pa.nlist = npl;
        pa.recs = precs.data(); pa.tile_base = tile_base.data(); pa.slot_scan = slot_scan.data();
        pa.out_off = off.data(); pa.status = status.data(); pa.rdesc = rdesc.data(); pa.desc = desc.data();
        pa.reject_list = list3.data(); pa.n_reject = &meta.n_reject; pa.n_desc = &meta.n_desc; pa.n_desc2 = &meta.n_desc2; pa.small_max = 8;
        const u32 grec = (npl + 255) / 256;
        hs::launch(dim3(grec), dim3(256), 0, [&] { k_par_plan(pa); });
        scan64(tile_base.data(), npl);
        const u64 ptiles = tile_base[npl];
        if (ptiles == 0 || ptiles > 0xFFFFFF00ULL) break;
        pa.ntiles = (u32)ptiles;
        std::vector<u64> tile_off(ptiles + 1);
        std::vector<uint4> tile_osum(ptiles + 1);
        pa.tile_off = tile_off.data(); pa.tile_osum = tile_osum.data();
        std::vector<uint2> tile_map(ptiles);
        pa.tile_map = tile_map.data();
        hs::launch(dim3(grec), dim3(256), 0, [&] { k_par_tilemap(pa); });
        hs::launch(dim3(pa.ntiles), dim3(kPThreads), 0, [&] { k_par_tabs(pa); });
        hs::launch(dim3((npl + 127) / 128), dim3(128), 0, [&] { k_par_head(pa); });
        hs::launch(dim3(pa.ntiles), dim3(kPThreads), 0, [&] { k_par_count(pa); });
        scan64(tile_off.data(), pa.ntiles);
        scan4(tile_osum.data(), pa.ntiles);
        hs::launch(dim3(grec), dim3(256), 0, [&] { k_par_ranges(pa); });
        scan64(slot_scan.data(), npl);
        pa.nsteps = (u32)tile_off[ptiles]; pa.nops = (u32)(tile_off[ptiles] >> 32);
        if (pa.nsteps == 0 || pa.nops == 0) break;
        {   // room for the records' descriptor runs in both halves (k_long's blocks follow them); bounded by 12 bytes of descriptors per input byte
            const u64 big = (u32)slot_scan[npl], small = slot_scan[npl] >> 32;
            const u64 need = 2 * std::max<u64>(std::max<u64>(big, small) + 2048, desc_cap / 2u);
            if (std::getenv("G2P_SIMT_DESC_CAP") ? need > desc_cap : need > 3 * n / 16 + 8192) break;
            if (need > desc_cap) { desc_cap = (u32)need; desc.resize(desc_cap + 1); la.desc = desc.data(); la.desc_cap = desc_cap; pa.desc = desc.data(); }
        }
        pa.half = desc_cap / 2u; pa.small_max = 8;
        std::vector<u32> spos(pa.nsteps), srec(pa.nsteps), opos(pa.nops), ot(pa.nops + 1);
        std::vector<uint4> sval(pa.nsteps), ox(pa.nops + 1);
        std::vector<u64> lx(pa.nsteps + 1), sx(pa.nsteps + 1);
        pa.spos = spos.data(); pa.srec = srec.data(); pa.opos = opos.data(); pa.ot = ot.data();
        pa.sval = sval.data(); pa.sx = sx.data(); pa.ox = ox.data(); pa.lx = lx.data();
        hs::launch(dim3(grec), dim3(256), 0, [&] { k_par_slots(pa); });
        hs::launch(dim3(pa.ntiles), dim3(kPThreads), 0, [&] { k_par_fill(pa); });
        hs::launch(dim3(std::min<u32>((pa.nsteps + 127) / 128, 8u)), dim3(128), 0, [&] { k_par_steps(pa); });
        scan64(sx.data(), pa.nsteps);
        hs::launch(dim3(grec), dim3(256), 0, [&] { k_par_totals(pa); });
        hs::launch(dim3(std::min<u32>((pa.nsteps + 127) / 128, 8u)), dim3(128), 0, [&] { k_par_lines(pa); });
        scan64(lx.data(), pa.nsteps);
        hs::launch(dim3(grec), dim3(256), 0, [&] { k_par_finish(pa); });
        hs::launch(dim3(std::min<u32>((pa.nsteps + 127) / 128, 8u)), dim3(128), 0, [&] { k_par_place(pa); });
        par_steps = pa.nsteps; par_ops = pa.nops;
        la.list = list3.data(); la.n_list = &meta.n_reject;
    } while (0);
    const u32 nlong = 2;
    hs::launch(dim3(nlong), dim3(kLThreads), long_smem<false>(), [&] { k_long<false>(la); });
    hs::launch(dim3(nlist), dim3(kListThreads), 0, [&] { k_convert_list<false>(gaf, rec.data(), T, off.data(), status.data(), nullptr, &meta, list2.data(), &meta.n_deleg2); });
    std::vector<u64> blocks2(nscan);
    std::vector<LineMapEnt> map((size_t)nrec * kSMaxLines);
    hs::launch(dim3(nscan), dim3(kScanThreads), 0, [&] { k_scan_reduce2(off.data(), loff.data(), nrec, blocks.data(), blocks2.data()); });
    hs::launch(dim3(1), dim3(1024), 0, [&] { k_scan_blocks2(blocks.data(), blocks2.data(), nscan, &meta.out_total, &meta.lines_total); });
    hs::launch(dim3(nscan), dim3(kScanThreads), 0, [&] {
        k_scan_apply2(off.data(), loff.data(), nrec, blocks.data(), blocks2.data(), &meta.out_total, &meta.lines_total, rec.data(), map.data());
    });
    std::vector<u8> out(meta.out_total + 256, 0xEE);
    la.out = out.data();
    if (meta.lines_total) {
        const u32 nl = (u32)meta.lines_total;
        EmitArgs ea{gaf, n, rec.data(), off.data(), sdesc.data(), map.data(), rdesc.data(), status.data(), nl, out.data()};
        hs::launch(dim3((nl + kEThreads - 1) / kEThreads), dim3(kEThreads), kEmitSmem, [&] { k_emit_lines<false>(ea); });
    }
    const u32 half = desc_cap / 2u;
    const u32 n_slots = std::min<u32>(meta.n_desc, half);
    if (n_slots) {
        EmitArgs ea{gaf, n, rec.data(), off.data(), desc.data(), nullptr, rdesc.data(), status.data(), n_slots, out.data()};
        hs::launch(dim3((n_slots + kEThreads - 1) / kEThreads), dim3(kEThreads), kEmitSmem, [&] { k_emit_lines<true>(ea); });
    }
    const u32 n_slots2 = std::min<u32>(meta.n_desc2, desc_cap - half);
    if (n_slots2) {
        EmitArgs ea{gaf, n, rec.data(), off.data(), desc.data() + half, nullptr, rdesc.data(), status.data(), n_slots2, out.data()};
        hs::launch(dim3((n_slots2 + kEThreads - 1) / kEThreads), dim3(kEThreads), kEmitSmem, [&] { k_emit_lines<true>(ea); });
    }
    if (meta.legacy_long) hs::launch(dim3(nlong), dim3(kLThreads), long_smem<true>(), [&] { k_long<true>(la); });
    if (meta.n_deleg2)
        hs::launch(dim3(nlist), dim3(kListThreads), 0, [&] { k_convert_list<true>(gaf, rec.data(), T, off.data(), status.data(), out.data(), &meta, list2.data(), &meta.n_deleg2); });
    u64 out_bytes = meta.out_total;
    if (meta.first_err != 0xFFFFFFFFu) {
        hs::launch(dim3(1), dim3(1), 0, [&] { k_diagnose(gaf, rec.data(), T, off.data(), &meta); });
        out_bytes = meta.err_out_end;
    }
    if (std::getenv("G2P_SIMT_STATS")) std::fprintf(stderr, "g2p_simt: %u records, %u to k_par (%u steps, %u ops), %u to k_long, %u to the general kernel, %llu short lines, %u + %u long line slots (cap %u), %llu bytes out\n", nrec, meta.n_deleg, par_steps, par_ops, par_steps ? meta.n_reject : meta.n_deleg, meta.n_deleg2, (unsigned long long)meta.lines_total, meta.n_desc, meta.n_desc2, desc_cap, (unsigned long long)out_bytes);
    std::fwrite(out.data(), 1, out_bytes, stdout);
    std::fflush(stdout);
    if (meta.first_err != 0xFFFFFFFFu) {
        const u32 st = meta.err_status & 0xff;
        if (st == ST_ERR_NAME) {
            std::string nm((const char*)gaf + meta.err_rec_start + meta.err_a, meta.err_b - meta.err_a);
            std::fprintf(stderr, "[gaf2paf] error: unable to find %s in lengths map\n", nm.c_str());
            return 1;
        }
        if (st == ST_ERR_NOCG) {
            std::fprintf(stderr, "[gaf2paf] error: cg cigar not found. This tool only works on output of minigraph -c\n");
            return 1;
        }
        std::fprintf(stderr, "abort: record %u status %u aux %u\n", meta.first_err, st, (meta.err_status >> 8) & 0xff);
        return 134;
    }
    return 0;
}
