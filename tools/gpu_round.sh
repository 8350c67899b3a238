#!/usr/bin/env bash
# One gpurun call: GPU tests, smoke, bench (short + asm), then ncu launch list and a full
# capture of the conversion kernels on a reduced record count.  Outputs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench short"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "rc=$?"; cat gpurun_out/bench_short.json; tail -3 gpurun_out/bench_short.err
echo "== bench asm"; timeout 1200 python bench.py --workload asm --steps 3 --warmup 3 > gpurun_out/bench_asm.json 2> gpurun_out/bench_asm.err; echo "rc=$?"; cat gpurun_out/bench_asm.json; tail -3 gpurun_out/bench_asm.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
  CMD="python bench.py --records 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  echo "== ncu launch list"
  $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "rc=$?"
  echo "== ncu full"
  $CMD > gpurun_out/plain2.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_short -s 6 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
  echo "rc=$?"
fi
ls -la gpurun_out
