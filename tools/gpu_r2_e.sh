#!/bin/bash
# round-2 run E: k_fuse (SWAR op fetch, branch-free appends): configurations; ncu of the best-looking one
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or golden or fuzz or synthetic" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
Q="--no-cli --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
for c in 0 5 1 2; do G2P_FUSE_CFG=$c timeout 300 python bench.py --workload short $Q > gpurun_out/r2e_short_cfg$c.json 2>&1; echo "cfg $c rc $?"; done
for c in 0 5; do G2P_FUSE_CFG=$c timeout 300 python bench.py --workload tagged $Q > gpurun_out/r2e_tagged_cfg$c.json 2>&1; done
timeout 300 python bench.py --workload unstable --steps 5 --warmup 3 > gpurun_out/r2e_bench_unstable.json 2> gpurun_out/r2e_bench_unstable.err; echo "unstable rc $?"
S="python bench.py --records 1000000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
export G2P_FUSE_CFG=5
$S > gpurun_out/r2e_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fuse" -s 3 -c 1 -f -o gpurun_out/r2e_kfuse $S > gpurun_out/r2e_ncu_full.log 2>&1
echo "ncu full rc $?"
