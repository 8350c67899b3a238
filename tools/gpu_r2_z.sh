#!/bin/bash
# round-2 run Z: gaffilter emit kernels in warp-synchronised phases, word-at-a-time marker scan: tests, timings, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2z_pytest.log; tail -3 gpurun_out/r2z_pytest.log
python tools/filter_bench.py > gpurun_out/r2z_filter_paf.txt 2>&1; cat gpurun_out/r2z_filter_paf.txt
python tools/filter_bench.py --gaf > gpurun_out/r2z_filter_gaf.txt 2>&1; cat gpurun_out/r2z_filter_gaf.txt
timeout 400 python bench.py --workload unstable --steps 10 --warmup 3 > gpurun_out/r2z_bench_unstable.json 2> gpurun_out/r2z_bench_unstable.err
echo "unstable rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2z_bench_unstable.json | head -1) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2z_bench_unstable.json | head -1)"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file gpurun_out/r2z_launches_filter_paf.csv python tools/filter_bench.py --no-ref --reps 1 > gpurun_out/r2z_ncu_paf.log 2>&1; echo "list paf rc $?"
