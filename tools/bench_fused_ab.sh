#!/bin/bash
# A/B of the one-launch overlapping step against the multi-launch schedule on N GPUs of one box (run under gpurun --gpus N)
N=${1:-2}
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "launches", d["gpu_launches"], "sustained", round(d["sustained"]["ms_per_step"],4), d["config"]["kernel"], d["clocks"]["sm_mhz"], d["sustained"]["clocks"]["sm_mhz"])'
python bench.py --no-cpu --no-configs --no-e2e 2>/dev/null | python -c "$show" "N=1"
for rep in ${REPS:-1 2}; do
for f in 1 0; do
PDB200_P2P_FUSED=$f python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu --no-configs --no-e2e 2>/dev/null | python -c "$show" "N=$N fused=$f"
done; done
