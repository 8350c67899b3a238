#!/bin/bash
# round-2 run F: k_fuse writing lines straight to global memory (no staging), 3 and 4 CTAs per SM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
Q="--no-cli --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
for c in 0 6 7 8; do G2P_FUSE_CFG=$c timeout 300 python bench.py --workload short $Q > gpurun_out/r2f_short_cfg$c.json 2>&1; echo "cfg $c rc $?"; done
for c in 0 6 8; do G2P_FUSE_CFG=$c timeout 300 python bench.py --workload tagged $Q > gpurun_out/r2f_tagged_cfg$c.json 2>&1; done
G2P_FUSE_CFG=6 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or golden or synthetic" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2f_pytest.log; tail -2 gpurun_out/r2f_pytest.log
timeout 300 python bench.py --workload unstable --steps 5 --warmup 3 > gpurun_out/r2f_bench_unstable.json 2> gpurun_out/r2f_bench_unstable.err; echo "unstable rc $?"
S="python bench.py --records 1000000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
export G2P_FUSE_CFG=6
$S > gpurun_out/r2f_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fuse" -s 3 -c 1 -f -o gpurun_out/r2f_kfuse $S > gpurun_out/r2f_ncu_full.log 2>&1
echo "ncu full rc $?"
cp cactus-gfa-tools_b200/csrc/g2p_fuse.cuh gpurun_out/r2f_g2p_fuse.cuh
