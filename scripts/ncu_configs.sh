#!/bin/bash
# usage (on the GPU box): bash scripts/ncu_configs.sh c3a c1 ...
# For every config: one plain run (must exit 0), then one `ncu --set full` capture of launch 4 of its filter kernel (after three
# warm-up launches).  The .ncu-rep is condensed ON THE BOX (gpurun copies back at most 64 MiB): gpurun_out/r02_<cfg>_kernel_summary.{json,md}
# (scripts/ncu_summary.py) and gpurun_out/r02_<cfg>_source.csv (per-instruction page); KEEP_REP=1 keeps the report as well.
mkdir -p gpurun_out
for c in "$@"; do
  python scripts/profile_cfg.py $c $NSTATES > gpurun_out/plain_$c.log 2>&1 || { echo "$c: plain run failed"; tail -3 gpurun_out/plain_$c.log; continue; }
  cat gpurun_out/plain_$c.log
  ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|ckpt_kernel|deferred_kernel' -s 3 -c 1 -f -o gpurun_out/prof_$c \
      python scripts/profile_cfg.py $c $NSTATES > gpurun_out/ncu_$c.log 2>&1; echo "$c ncu rc=$?"
  python scripts/ncu_summary.py gpurun_out/prof_$c.ncu-rep gpurun_out/r02_${c}_kernel > /dev/null 2>&1
  ncu -i gpurun_out/prof_$c.ncu-rep --page source --csv > gpurun_out/r02_${c}_source.csv 2>/dev/null
  [ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_$c.ncu-rep
done
