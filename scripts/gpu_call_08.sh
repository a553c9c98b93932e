#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rf --deselect "tests/test_gpu_models.py::test_fused_rgbnet_vs_torch[tc3+fused]" --deselect "tests/test_gpu_models.py::test_fused_rgbnet_vs_torch[tc1+fused]" > gpurun_out/pytest_r02_h.log 2>&1
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
# the warp-specialised backward (new this call): its own short-fused invocation so that a hang cannot take the rest down
timeout 240 python -m pytest "tests/test_gpu_models.py::test_fused_rgbnet_vs_torch" -q -k "fused" -rf > gpurun_out/pytest_r02_h_ws.log 2>&1
echo "--- ws pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" gpurun_out/pytest_r02_h_ws.log | cut -c1-400 | head
UBN_RGBNET_BWD_MODE=fused timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r02_h_ws.json 2> gpurun_out/bench_r02_h_ws.err
echo "--- bench ws rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_h_ws.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'])"
