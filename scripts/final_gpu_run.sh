#!/bin/bash
# Round-end measurement pass on the GPU box (one GPU): tests, smoke, bench (both arms), launch list, ncu summaries of every
# config kernel, latency, parity report.  Everything lands in gpurun_out/r02_final_*.
#   ASIF_GIT_HEAD=<commit> bash scripts/final_gpu_run.sh [skip-parity]
mkdir -p gpurun_out
O=gpurun_out/r02_final
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > ${O}_box.txt; echo "host cores: $(nproc)" >> ${O}_box.txt
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > ${O}_gpu_tests.log; tail -3 ${O}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; tail -1 ${O}_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "bench-ref rc=$?"
timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > ${O}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > ${O}_ncu_launches.log 2>&1; echo "ncu-launches rc=$?"
bash scripts/ncu_configs.sh c1 c2 c3a c3b c4 c5
NSTATES=10000000 bash scripts/ncu_configs.sh c2 > /dev/null 2>&1 && mv gpurun_out/r02_c2_kernel_summary.json ${O}_c2_1e7_summary.json && mv gpurun_out/r02_c2_kernel_summary.md ${O}_c2_1e7_summary.md; rm -f gpurun_out/r02_c2_source.csv
NSTATES=1000000 bash scripts/ncu_configs.sh c3a c5 > /dev/null 2>&1; for c in c3a c5; do mv gpurun_out/r02_${c}_kernel_summary.json ${O}_${c}_1e6_summary.json; mv gpurun_out/r02_${c}_kernel_summary.md ${O}_${c}_1e6_summary.md; mv gpurun_out/r02_${c}_source.csv ${O}_${c}_1e6_source.csv; done
bash scripts/ncu_configs.sh c2 c3a c5 > /dev/null 2>&1
python scripts/filter_latency.py > ${O}_filter_latency.json 2> ${O}_filter_latency.err; tail -2 ${O}_filter_latency.json | cut -c1-400
python scripts/c1_sweep.py 1e6 1e7 1e8 > ${O}_c1_sweep.jsonl 2>&1
python scripts/bench_c3a.py c3a rb c5 c5roll > ${O}_c3a_c5.jsonl 2>&1
if [ "$1" != "skip-parity" ]; then timeout 2400 python scripts/parity_report.py > ${O}_parity_report.jsonl 2> ${O}_parity_report.err; echo "parity rc=$?"; fi
du -sh gpurun_out
