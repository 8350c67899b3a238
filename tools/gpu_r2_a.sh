#!/bin/bash
# round-2 run A: the full GPU test suite, then one bench line per workload shape (profiles/r02_bench_*.json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt; df -h /dev/shm /tmp >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
for w in short tagged mixed stable medium asm; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err
  echo "bench $w rc $?"; head -c 600 gpurun_out/r2a_bench_$w.json; echo
done
