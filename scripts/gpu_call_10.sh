#!/bin/bash
O=gpurun_out/call10; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_render.py -q --timeout 300 -rf -k "rgbnet or tma or render or golden or training" > $O/pytest.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|\[tma\]" $O/pytest.log | cut -c1-400 | head -20
for mode in tc3 tc3w4; do
UBN_RGBNET_MODE=$mode UBN_RGBNET_BWD_MODE=fused timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_$mode.json 2> $O/bench_$mode.err
echo "--- bench fwd=$mode rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$mode.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"; tail -2 $O/bench_$mode.err
done
for tma in "" "--no-tma"; do
timeout 300 python bench.py --workload garden --steps 3 --warmup 3 $tma > $O/garden$tma.json 2> $O/garden$tma.err
echo "--- garden $tma rc=$?"; python -c "
import json;d=json.load(open('$O/garden$tma.json'));print(d['ms_per_step'],d['value'],d['tma_feature_read'],d['check'])"
done
UBN_RGBNET_BWD_MODE=fused timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_shade_bwd_fused_ws|k_shade_fwd_tc|k_shade_dw2_tc|k_march_feature_v3" -s 12 -c 4 -o $O/rgbnet -f python bench.py --only-timed --steps 1 --warmup 3 --no-reference-gpu > $O/ncu_full.log 2>&1; ls -la $O/*.ncu-rep; tail -3 $O/ncu_full.log
