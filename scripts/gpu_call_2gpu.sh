#!/bin/bash
# 2-GPU box: NCCL world-2 exactness test of the peer tail, training bench (peer vs pipelined tail), sharded frame render, block IDW
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_peer_tail.py -q --timeout 800 -rf > gpurun_out/pytest_r02_2gpu.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|symmetric memory" gpurun_out/pytest_r02_2gpu.log | cut -c1-600 | head -30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for mode in peer pipelined; do
UBN_BENCH_TAIL=$mode timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_2gpu_$mode.json 2> gpurun_out/bench_r02_2gpu_$mode.err
echo "--- bench 2gpu tail=$mode rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_2gpu_$mode.json'));print(d['value']/1e6,d['ms_per_step'],d['tail_ms'])"; grep -i "symmetric\|error\|Traceback" gpurun_out/bench_r02_2gpu_$mode.err | head -5
done
timeout 600 $TR bench.py --gpus 2 --workload garden --steps 3 --warmup 3 > gpurun_out/bench_r02_2gpu_garden.json 2> gpurun_out/bench_r02_2gpu_garden.err
echo "--- garden 2gpu rc=$?"; cut -c1-1200 gpurun_out/bench_r02_2gpu_garden.json; tail -2 gpurun_out/bench_r02_2gpu_garden.err
timeout 600 $TR bench.py --gpus 2 --workload missionbay --steps 3 --warmup 3 > gpurun_out/bench_r02_2gpu_missionbay.json 2> gpurun_out/bench_r02_2gpu_missionbay.err
echo "--- missionbay 2gpu rc=$?"; cut -c1-1200 gpurun_out/bench_r02_2gpu_missionbay.json; tail -2 gpurun_out/bench_r02_2gpu_missionbay.err
