#!/bin/bash
# round-2 run B: k_fuse on the GPU -- tests, bench lines, tile variants, ncu launch list + full capture, CLI chunk sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --workload short --steps 10 --warmup 3 > gpurun_out/r2b_bench_short.json 2> gpurun_out/r2b_bench_short.err; echo "bench short rc $?"
head -c 1500 gpurun_out/r2b_bench_short.json; echo
Q="--no-cli --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
for t in 16384 8192; do G2P_FUSE_TILE=$t timeout 300 python bench.py --workload short $Q > gpurun_out/r2b_short_tile$t.json 2>&1; echo "tile $t rc $?"; done
G2P_FUSE=0 timeout 300 python bench.py --workload short $Q > gpurun_out/r2b_short_nofuse.json 2>&1
for w in tagged mixed; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cli > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err; echo "bench $w rc $?"; done
S="python bench.py --records 1000000 --steps 2 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
$S > gpurun_out/r2b_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_short1M.csv $S > gpurun_out/r2b_ncu_list.log 2>&1
echo "ncu list rc $?"
$S > gpurun_out/r2b_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fuse" -s 3 -c 1 -f -o gpurun_out/r2b_kfuse $S > gpurun_out/r2b_ncu_full.log 2>&1
echo "ncu full rc $?"
# CLI chunk sweep on the 10 M-record file
./build/gafgen short 10000000 /dev/shm/r2b.gaf /dev/shm/r2b.tsv > /dev/null 2>&1
for mb in 16 32 64 128; do for rep in 1 2; do G2P_CHUNK_MB=$mb G2P_STATS=1 ./cactus-gfa-tools_b200/bin/gaf2paf -l /dev/shm/r2b.tsv /dev/shm/r2b.gaf 2>&1 >/dev/null | grep "wall=" | sed "s/^/chunk $mb: /"; done; done | tee gpurun_out/r2b_cli_chunks.txt
rm -f /dev/shm/r2b.gaf /dev/shm/r2b.tsv
