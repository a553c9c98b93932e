# two-phase run-merging density scatter A/B
O=gpurun_out/call27; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_at_size.py tests/test_gpu_callers_unchanged.py -q --timeout 600 -x -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-300 | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench.json 2> $O/bench.err
echo "--- bench rc=$?"; python -c "
import json;d=json.load(open('$O/bench.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench.err
timeout 300 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_bicycle.json 2> $O/bench_bicycle.err; python -c "
import json;d=json.load(open('$O/bench_bicycle.json'));print('bicycle',d['ms_per_step'],d['roofline']['all_kernels_ms'])"
