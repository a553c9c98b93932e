#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rf > gpurun_out/pytest_r02_h.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|\[tma\]\|\[psnr tf32x1\]" gpurun_out/pytest_r02_h.log | cut -c1-500 | head -40
for fk in 0 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel $fk > gpurun_out/bench_r02_h_fk$fk.json 2> gpurun_out/bench_r02_h_fk$fk.err
echo "--- bench feature-kernel=$fk rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_h_fk$fk.json'));print(d['ms_per_step'],d['tail_ms']['value'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"; tail -2 gpurun_out/bench_r02_h_fk$fk.err
done
for fk in 0 2; do
timeout 600 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel $fk > gpurun_out/bench_r02_h_bicycle_fk$fk.json 2> gpurun_out/bench_r02_h_bicycle_fk$fk.err
echo "--- bicycle fk=$fk rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_h_bicycle_fk$fk.json'));print(d['ms_per_step'],d['rho'],d['roofline']['all_kernels_ms'],d['roofline']['all_kernels_frac'])"
done
for tma in "" "--no-tma"; do
timeout 600 python bench.py --workload garden --steps 3 --warmup 3 $tma > gpurun_out/bench_r02_h_garden1$tma.json 2> gpurun_out/bench_r02_h_garden1$tma.err
echo "--- garden $tma rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_h_garden1$tma.json'));print(d['ms_per_step'],d['value'],d['tma_feature_read'],d['check'])"
done
UBN_RGBNET_MODE=tc1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r02_h_tf32x1.json 2> gpurun_out/bench_r02_h_tf32x1.err
echo "--- bench tf32x1 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_h_tf32x1.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'])"
