#!/bin/bash
# where does the weak-scaling loss come from?  (run under gpurun --gpus 2)
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "sustained", round(d["sustained"]["ms_per_step"],4), d["clocks"]["sm_mhz"], flush=True)'
B="python bench.py --no-cpu --no-configs --no-e2e"
for rep in 1 2; do
CUDA_VISIBLE_DEVICES=0 $B 2>/dev/null | python -c "$show" "gpu0 alone"
CUDA_VISIBLE_DEVICES=1 $B 2>/dev/null | python -c "$show" "gpu1 alone"
(CUDA_VISIBLE_DEVICES=0 $B 2>/dev/null | python -c "$show" "gpu0 (gpu1 busy)") &
CUDA_VISIBLE_DEVICES=1 $B 2>/dev/null | python -c "$show" "gpu1 (gpu0 busy)"
wait
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu --no-configs --no-e2e 2>/dev/null | python -c "$show" "N=2 coupled"
done
