#!/bin/bash
# 2-GPU box, end of round 2: NCCL world-2 exactness test of the peer tail, training bench (peer tail), sharded frame render, block IDW
O=gpurun_out/final2gpu; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_peer_tail.py -q --timeout 800 -rf > $O/pytest_2gpu.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest_2gpu.log | cut -c1-600 | head -10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "--- bench 2gpu rc=$?"; python -c "
import json;d=json.load(open('$O/bench_2gpu.json'));print(d['value']/1e6,d['ms_per_step'],d['tail_ms'])"; grep -i "symmetric\|error\|Traceback" $O/bench_2gpu.err | head -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_1gpu.json 2> $O/bench_1gpu.err; python -c "
import json;d=json.load(open('$O/bench_1gpu.json'));print('1gpu same box',d['value']/1e6,d['ms_per_step'],d['tail_ms'])"
timeout 600 $TR bench.py --gpus 2 --workload garden --steps 3 --warmup 3 > $O/bench_2gpu_garden.json 2> $O/bench_2gpu_garden.err
echo "--- garden 2gpu rc=$?"; cut -c1-300 $O/bench_2gpu_garden.json; tail -2 $O/bench_2gpu_garden.err
timeout 600 $TR bench.py --gpus 2 --workload missionbay --steps 3 --warmup 3 > $O/bench_2gpu_missionbay.json 2> $O/bench_2gpu_missionbay.err
echo "--- missionbay 2gpu rc=$?"; cut -c1-300 $O/bench_2gpu_missionbay.json; tail -2 $O/bench_2gpu_missionbay.err
