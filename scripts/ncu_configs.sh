#!/bin/bash
# usage (on the GPU box): bash scripts/ncu_configs.sh c3a c1 ...   -> gpurun_out/prof_<cfg>.ncu-rep (launch 4 of the config's kernel,
# i.e. after three warm-up launches), each only after the same command has exited 0 without ncu.  Read here with scripts/ncu_summary.py.
mkdir -p gpurun_out
for c in "$@"; do
  python scripts/profile_cfg.py $c > gpurun_out/plain_$c.log 2>&1 || { echo "$c: plain run failed"; tail -3 gpurun_out/plain_$c.log; continue; }
  cat gpurun_out/plain_$c.log
  ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|ckpt_kernel' -s 3 -c 1 -f -o gpurun_out/prof_$c \
      python scripts/profile_cfg.py $c > gpurun_out/ncu_$c.log 2>&1; echo "$c ncu rc=$?"
done
