#!/bin/bash
O=gpurun_out/call15; mkdir -p $O
timeout 60 scripts/_bin/trace_ws 0 > $O/trace_ws_rowmajor.txt 2>&1; timeout 60 scripts/_bin/trace_ws 4 > $O/trace_ws.txt 2>&1; head -1 $O/trace_ws_rowmajor.txt; head -7 $O/trace_ws.txt | cut -c1-600
timeout 600 python -m pytest tests/test_gpu_models.py -q --timeout 300 -rf -k "rgbnet or golden or training or psnr" > $O/pytest.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-400 | head -20
for bm in fused; do
UBN_RGBNET_BWD_MODE=$bm timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_$bm.json 2> $O/bench_$bm.err
echo "--- bench bwd=$bm rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$bm.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"; tail -2 $O/bench_$bm.err
done
timeout 300 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_bicycle.json 2> $O/bench_bicycle.err
python -c "
import json;d=json.load(open('$O/bench_bicycle.json'));print('bicycle',d['ms_per_step'],d['roofline']['all_kernels_ms'])"
