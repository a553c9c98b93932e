#!/bin/bash
O=gpurun_out/call11; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_models.py -q --timeout 300 -rf -k "rgbnet or golden or training" > $O/pytest.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-400 | head -20
for bm in fused fused4; do
UBN_RGBNET_BWD_MODE=$bm timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_$bm.json 2> $O/bench_$bm.err
echo "--- bench bwd=$bm rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$bm.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"; tail -2 $O/bench_$bm.err
done
UBN_RGBNET_BWD_MODE=fused timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_shade_bwd_fused_ws" -s 3 -c 1 -o $O/ws -f python bench.py --only-timed --steps 1 --warmup 3 --no-reference-gpu > $O/ncu_full.log 2>&1; ls -la $O/*.ncu-rep; tail -3 $O/ncu_full.log
