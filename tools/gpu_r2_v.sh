#!/bin/bash
# round-2 run V (final numbers): the whole GPU suite, smoke, one bench line per workload, the reference arm, config 4 at its
# stated size through the executable, gaffilter timing, launch lists, full captures of the k_par kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2v_pytest.log
tail -3 gpurun_out/r2v_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc $?"
for w in short tagged mixed stable medium asm unstable; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2v_bench_$w.json 2> gpurun_out/r2v_bench_$w.err
  echo "bench $w rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_bench_$w.json | head -1)"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2v_reference_short.json 2> gpurun_out/r2v_reference_short.err; echo "reference arm rc $?"
# config 4 at its stated size: 100 000 assembly-scale records (~25 GB of GAF) through the drop-in executable
( time ./build/gafgen asm 100000 /dev/shm/c4.gaf /dev/shm/c4.tsv ) > gpurun_out/r2v_config4.txt 2>&1
ls -l /dev/shm/c4.gaf >> gpurun_out/r2v_config4.txt
for rep in 1 2; do ( time G2P_STATS=1 ./cactus-gfa-tools_b200/bin/gaf2paf -l /dev/shm/c4.tsv /dev/shm/c4.gaf > /dev/null ) >> gpurun_out/r2v_config4.txt 2>&1; done
( time G2P_PAR=0 G2P_STATS=1 ./cactus-gfa-tools_b200/bin/gaf2paf -l /dev/shm/c4.tsv /dev/shm/c4.gaf > /dev/null ) >> gpurun_out/r2v_config4.txt 2>&1
head -c 300000000 /dev/shm/c4.gaf | head -n 1000 > /dev/shm/c4_sample.gaf
( time ./oracle/_ref/gaf2paf /dev/shm/c4_sample.gaf -l /dev/shm/c4.tsv > /dev/shm/c4_ref.paf ) >> gpurun_out/r2v_config4.txt 2>&1
./cactus-gfa-tools_b200/bin/gaf2paf -l /dev/shm/c4.tsv /dev/shm/c4_sample.gaf | cmp - /dev/shm/c4_ref.paf && echo "config4 sample (1000 records): executable output identical to the reference" >> gpurun_out/r2v_config4.txt
rm -f /dev/shm/c4*
for w in short asm stable; do
  S="python bench.py --workload $w --steps 1 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
  [ $w = short ] && S="$S --records 1000000"
  $S > gpurun_out/r2v_plain_$w.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_launches_$w.csv $S > gpurun_out/r2v_ncu_list_$w.log 2>&1
  echo "ncu list $w rc $?"
done
S="python bench.py --workload asm --records 1000 --steps 1 --warmup 1 --no-cli --no-e2e --no-cpu-baseline"
for k in k_par_count k_par_fill k_par_lines k_par_steps; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -f -o gpurun_out/r2v_$k $S > gpurun_out/r2v_ncu_$k.log 2>&1
  echo "ncu $k rc $?"
done
cp cactus-gfa-tools_b200/csrc/g2p_par.cuh gpurun_out/r2v_g2p_par.cuh
