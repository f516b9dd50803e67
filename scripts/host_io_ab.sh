#!/bin/bash
# A/B of the host-memory transfer modes (ASIF_B200_HOST_IO) on the bench workload: e2e states/s per mode
mkdir -p gpurun_out
for rep in 1 2; do
for m in staged out inout; do
  ASIF_B200_HOST_IO=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_io_${m}_$rep.json 2> gpurun_out/bench_io_${m}_$rep.err
  python - $m $rep <<'PY'
import json, sys
m, rep = sys.argv[1:3]
try:
    d = json.loads([l for l in open('gpurun_out/bench_io_%s_%s.json' % (m, rep)) if l.startswith('{')][-1])
    print('%-7s rep %s value %.4g e2e %.4g (%.3f ms/step) same=%s' % (m, rep, d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['matches_device_resident']))
except Exception as ex:
    print(m, rep, 'failed', ex)
PY
done
done
