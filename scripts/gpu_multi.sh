#!/bin/bash
# N-GPU bench exactly as the driver launches it
N=${1:-2}
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/multi_$N.err | tee gpurun_out/multi_$N.json | cut -c1-200
grep -o '"e2e": {[^}]*}' gpurun_out/multi_$N.json
tail -3 gpurun_out/multi_$N.err
if [ "$2" != "noref" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-400
fi
