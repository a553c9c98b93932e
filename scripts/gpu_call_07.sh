#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rf > gpurun_out/pytest_r02_g.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" gpurun_out/pytest_r02_g.log | cut -c1-500 | head -40
for mode in fused tc3; do
UBN_RGBNET_BWD_MODE=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r02_g_$mode.json 2> gpurun_out/bench_r02_g_$mode.err
echo "--- bench bwd=$mode rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_g_$mode.json'));print(d['ms_per_step'],d['tail_ms']['value'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_r02_g_$mode.err
done
timeout 600 python bench.py --workload garden --steps 3 --warmup 3 > gpurun_out/bench_r02_g_garden1.json 2> gpurun_out/bench_r02_g_garden1.err
echo "--- garden rc=$?"; cat gpurun_out/bench_r02_g_garden1.json | cut -c1-1500; tail -3 gpurun_out/bench_r02_g_garden1.err
for fk in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel $fk > gpurun_out/bench_r02_g_fk$fk.json 2> gpurun_out/bench_r02_g_fk$fk.err
echo "--- bench feature-kernel=$fk rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_g_fk$fk.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"; tail -2 gpurun_out/bench_r02_g_fk$fk.err
done
timeout 600 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r02_g_bicycle.json 2> gpurun_out/bench_r02_g_bicycle.err
echo "--- bicycle rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_g_bicycle.json'));print(d['ms_per_step'],d['rho'],d['roofline']['all_kernels_ms'],d['roofline']['all_kernels_frac'])"
timeout 600 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel 2 > gpurun_out/bench_r02_g_bicycle_fk2.json 2> gpurun_out/bench_r02_g_bicycle_fk2.err
echo "--- bicycle fk2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_g_bicycle_fk2.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'])"
