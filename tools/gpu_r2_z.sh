#!/bin/bash
# round-2 final run: the whole GPU suite, smoke, the default bench line + the reference arm, gaffilter timings (also one
# query sequence with 600 k alignments: the assembly-contig shape the segmented scan is for)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2f_smoke.log
python tools/filter_bench.py > gpurun_out/r2f_filter_paf.txt 2>&1; cat gpurun_out/r2f_filter_paf.txt
python tools/filter_bench.py --gaf > gpurun_out/r2f_filter_gaf.txt 2>&1; cat gpurun_out/r2f_filter_gaf.txt
python tools/filter_bench.py --gaf --queries 3 --records 300000 --no-ref > gpurun_out/r2f_filter_contigs.txt 2>&1; cat gpurun_out/r2f_filter_contigs.txt
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_reference.json 2> gpurun_out/r2f_reference.err; echo "reference rc $?"
