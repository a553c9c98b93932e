# H1 saved transposed for the dW2 launch
O=gpurun_out/call33; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_at_size.py -q --timeout 600 -x -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-400 | head
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench.json 2> $O/bench.err
echo "--- bench rc=$?"; python -c "
import json;d=json.load(open('$O/bench.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 2 --no-reference-gpu > $O/ncu_launch.log 2>&1; grep -E "tc::" $O/launches.csv | tail -3 | awk -F'","' '{print substr($5,1,45), $NF}'
