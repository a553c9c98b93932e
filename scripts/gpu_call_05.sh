#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_peer_tail.py -q --timeout 600 -rf -k "rgbnet or peer or training or golden" > gpurun_out/pytest_r02_e.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" gpurun_out/pytest_r02_e.log | cut -c1-400 | head -40
for mode in fused tc3; do
UBN_RGBNET_BWD_MODE=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_e_$mode.json 2> gpurun_out/bench_r02_e_$mode.err
echo "--- bench bwd=$mode rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_e_$mode.json'));print(d['ms_per_step'],d['tail_ms']['value'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_r02_e_$mode.err
done
timeout 600 python bench.py --workload garden --steps 3 --warmup 3 > gpurun_out/bench_r02_e_garden1.json 2> gpurun_out/bench_r02_e_garden1.err
echo "--- garden rc=$?"; cat gpurun_out/bench_r02_e_garden1.json | cut -c1-1500; tail -3 gpurun_out/bench_r02_e_garden1.err
