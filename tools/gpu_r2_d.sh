#!/bin/bash
# round-2 run D: k_fuse with warp-converged loops; N2 + hpp20 GPU tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_hpp20.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
Q="--no-cli --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
for c in 0 1 2; do G2P_FUSE_CFG=$c timeout 300 python bench.py --workload short $Q > gpurun_out/r2d_short_cfg$c.json 2>&1; echo "cfg $c rc $?"; done
G2P_FUSE_CFG=0 timeout 300 python bench.py --workload tagged $Q > gpurun_out/r2d_tagged_cfg0.json 2>&1
S="python bench.py --records 1000000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
$S > gpurun_out/r2d_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fuse" -s 3 -c 1 -f -o gpurun_out/r2d_kfuse $S > gpurun_out/r2d_ncu_full.log 2>&1
echo "ncu full rc $?"
