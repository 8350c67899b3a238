#!/bin/bash
# round-2 run G: hybrid dispatch -- full GPU suite, one bench line per workload, launch lists + full captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc $?"
for w in short tagged mixed stable medium asm unstable; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err
  echo "bench $w rc $?"
done
S="python bench.py --records 1000000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
$S > gpurun_out/r2g_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_short1M.csv $S > gpurun_out/r2g_ncu_list.log 2>&1
echo "ncu list short rc $?"
T="python bench.py --workload tagged --records 500000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
$T > gpurun_out/r2g_plain_t.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_tagged500k.csv $T > gpurun_out/r2g_ncu_list_t.log 2>&1
echo "ncu list tagged rc $?"
$T > gpurun_out/r2g_plain_t2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fuse" -s 3 -c 1 -f -o gpurun_out/r2g_kfuse_tagged $T > gpurun_out/r2g_ncu_full_t.log 2>&1
echo "ncu full tagged rc $?"
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
