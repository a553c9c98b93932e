#!/bin/bash
# 4-GPU box: NCCL exactness of the peer tail at world 2 and 4, training bench at 4 ranks
O=gpurun_out/final4gpu; mkdir -p $O
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_peer_tail.py -q --timeout 500 -rf > $O/pytest_4gpu.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|skipped" $O/pytest_4gpu.log | cut -c1-400 | head -10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
timeout 400 $TR bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_4gpu.json 2> $O/bench_4gpu.err
echo "--- bench 4gpu rc=$?"; python -c "
import json;d=json.load(open('$O/bench_4gpu.json'));print(d['value']/1e6,d['ms_per_step'],d['tail_ms'])"; grep -i "symmetric\|error\|Traceback\|unavailable" $O/bench_4gpu.err | head -5
