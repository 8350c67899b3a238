#!/bin/bash
# round-2 run L: the whole GPU suite, one bench line per workload, config 4 at its stated size through the executable,
# gaffilter timing, launch lists + full captures of the dominant kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2l_pytest.log
tail -3 gpurun_out/r2l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc $?"
for w in short tagged mixed stable medium asm unstable; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2l_bench_$w.json 2> gpurun_out/r2l_bench_$w.err
  echo "bench $w rc $?"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2l_reference_short.json 2> gpurun_out/r2l_reference_short.err; echo "reference arm rc $?"
# config 4 at its stated size: 100 000 assembly-scale records (~25 GB of GAF) through the drop-in executable
( time ./build/gafgen asm 100000 /dev/shm/c4.gaf /dev/shm/c4.tsv ) > gpurun_out/r2l_config4.txt 2>&1
ls -l /dev/shm/c4.gaf >> gpurun_out/r2l_config4.txt
for rep in 1 2; do ( time G2P_STATS=1 ./cactus-gfa-tools_b200/bin/gaf2paf -l /dev/shm/c4.tsv /dev/shm/c4.gaf > /dev/null ) >> gpurun_out/r2l_config4.txt 2>&1; done
head -c 300000000 /dev/shm/c4.gaf | head -n 1000 > /dev/shm/c4_sample.gaf
( time ./oracle/_ref/gaf2paf /dev/shm/c4_sample.gaf -l /dev/shm/c4.tsv > /dev/shm/c4_ref.paf ) >> gpurun_out/r2l_config4.txt 2>&1
./cactus-gfa-tools_b200/bin/gaf2paf -l /dev/shm/c4.tsv /dev/shm/c4_sample.gaf | cmp - /dev/shm/c4_ref.paf && echo "config4 sample (1000 records): executable output identical to the reference" >> gpurun_out/r2l_config4.txt
rm -f /dev/shm/c4*
# gaffilter on the PAF of 2 M short-read records (device time from the C-ABI, reference wall clock)
python - > gpurun_out/r2l_filter.txt 2>&1 <<'P'
import sys, time, os, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import cactus_gfa_tools_b200 as g2p, helpers as H
text, lengths = H.gen_filter_case(5, n_records=600000, n_queries=200000, paf=True)
cv = g2p.Converter(0)
par = g2p.Converter.filter_params(paf=True, ratio=2)
for _ in range(3):
    out, res = cv.filter_host(text, par)
t0 = time.perf_counter(); out, res = cv.filter_host(text, par); wall = time.perf_counter() - t0
open("/dev/shm/f.paf", "wb").write(text)
t0 = time.perf_counter(); r = subprocess.run([os.path.join(H.REF_BIN, "gaffilter"), "/dev/shm/f.paf", "-p", "-r", "2"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL); tref = time.perf_counter() - t0
os.unlink("/dev/shm/f.paf")
print("gaffilter -p -r 2: %d PAF lines (%d B) -> %d kept; device %.3f ms (%d launches), host call %.1f ms; reference %.2f s; identical: %s"
      % (res.n_loaded, len(text), res.n_loaded - res.n_filtered, res.device_ms, res.gpu_launches, wall * 1e3, tref, out == r.stdout))
P
cat gpurun_out/r2l_filter.txt
S="python bench.py --records 1000000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
$S > gpurun_out/r2l_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches_short1M.csv $S > gpurun_out/r2l_ncu_list.log 2>&1
echo "ncu list short rc $?"
$S > gpurun_out/r2l_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_rec|k_emit_lines" -s 6 -c 2 -f -o gpurun_out/r2l_short $S > gpurun_out/r2l_ncu_full.log 2>&1
echo "ncu full short rc $?"
T="python bench.py --workload tagged --records 500000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
$T > gpurun_out/r2l_plain_t.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches_tagged500k.csv $T > gpurun_out/r2l_ncu_list_t.log 2>&1
$T > gpurun_out/r2l_plain_t2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fuse" -s 3 -c 1 -f -o gpurun_out/r2l_kfuse_tagged $T > gpurun_out/r2l_ncu_full_t.log 2>&1
echo "ncu full tagged rc $?"
cp cactus-gfa-tools_b200/csrc/g2p_fuse.cuh gpurun_out/r2l_g2p_fuse.cuh
