# dW2 rebuilt from ReLU masks (ws column warps ballot them), equal-cell merge in the slab-major k0 scatter
O=gpurun_out/call20; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_at_size.py -q --timeout 600 -x -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-300 | head -20
for dw in masks h2; do
UBN_RGBNET_DW2=$dw timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_$dw.json 2> $O/bench_$dw.err
echo "--- bench dw2=$dw rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$dw.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench_$dw.err
done
timeout 300 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_bicycle.json 2> $O/bench_bicycle.err; python -c "
import json;d=json.load(open('$O/bench_bicycle.json'));print('bicycle',d['ms_per_step'],d['roofline']['all_kernels_ms'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 2 --no-reference-gpu > $O/ncu_launch.log 2>&1; grep -E "dw2|bwd_slab|density|fused_ws|fwd_tc" $O/launches.csv | tail -7 | cut -d, -f5,13-15 | cut -c1-60,200-260
