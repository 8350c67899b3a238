#!/usr/bin/env bash
# Round-end evidence: GPU tests, smoke, bench (short, asm, reference arm), ncu launch list + full captures.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench short"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "rc=$?"; cat gpurun_out/bench_short.json
echo "== bench asm"; timeout 1200 python bench.py --workload asm --steps 3 --warmup 3 > gpurun_out/bench_asm.json 2> gpurun_out/bench_asm.err; echo "rc=$?"; cat gpurun_out/bench_asm.json
echo "== reference arm"; timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err; echo "rc=$?"; cat gpurun_out/ref_arm.json
if [ "${SKIP_NCU:-0}" != "1" ]; then SKIP_ASM=${SKIP_ASM:-0} bash tools/gpu_ncu_rec.sh; fi
