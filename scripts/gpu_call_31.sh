# k0 scatter: pairs of neighbours taken 16 apart, reductions issued alternately; default gather = v4 for single-slab grids
O=gpurun_out/call31; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_models.py -q --timeout 600 -x -rf -k "feature_kernel or golden or training" > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-400 | head
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench.json 2> $O/bench.err
echo "--- bench rc=$?"; python -c "
import json;d=json.load(open('$O/bench.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench.err
timeout 300 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_bicycle.json 2> $O/bench_bicycle.err; python -c "
import json;d=json.load(open('$O/bench_bicycle.json'));print('bicycle',d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"
