# db2 summed by the dW2 launch; neighbour-cell face merge in the k0 scatter
O=gpurun_out/call24; mkdir -p $O
timeout 60 scripts/_bin/trace_ws 4 > $O/trace_ws.txt 2>&1; sed -n 1,7p $O/trace_ws.txt | cut -c1-700
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_at_size.py -q --timeout 300 -x -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-300 | head -20
for h in 1 0; do
UBN_RGBNET_MASKS=$h timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_$h.json 2> $O/bench_$h.err
echo "--- bench h1=$h rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$h.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench_$h.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 2 --no-reference-gpu > $O/ncu_launch.log 2>&1; grep -E "tc::" $O/launches.csv | tail -3 | awk -F'","' '{print substr($5,1,45), $NF}'
