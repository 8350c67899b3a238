#!/bin/bash
# multi-GPU run: N = $1.  One logical input sharded by newline-aligned byte ranges (bench.py), the executable with G2P_GPUS=N,
# the N-GPU PCIe floor, and the 2-GPU CLI test.
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m${N}_topo.txt 2>&1
nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv >> gpurun_out/r2m${N}_topo.txt 2>&1
nproc >> gpurun_out/r2m${N}_topo.txt; free -g >> gpurun_out/r2m${N}_topo.txt
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "cli_sharded" > gpurun_out/r2m2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2m2_pytest.log; tail -2 gpurun_out/r2m2_pytest.log; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/pcie_floor.py > gpurun_out/r2m${N}_pcie_floor.json 2> gpurun_out/r2m${N}_pcie_floor.err; echo "floor rc $?"; cat gpurun_out/r2m${N}_pcie_floor.json
for w in short mixed; do
  timeout 900 $TR bench.py --gpus $N --workload $w --steps 10 --warmup 3 > gpurun_out/r2m${N}_bench_$w.json 2> gpurun_out/r2m${N}_bench_$w.err; echo "bench $w N=$N rc $?"
  head -c 700 gpurun_out/r2m${N}_bench_$w.json; echo
done
