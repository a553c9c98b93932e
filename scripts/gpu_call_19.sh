# fast density kernels now taken for odd-sized slabs; slab-major scatter default; backward scatter overlap A/B
O=gpurun_out/call19; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_at_size.py tests/test_gpu_callers_unchanged.py -q --timeout 600 -x -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-300 | head -20
for ov in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --bwd-overlap $ov > $O/bench_ov$ov.json 2> $O/bench_ov$ov.err
echo "--- bench overlap=$ov rc=$?"; python -c "
import json;d=json.load(open('$O/bench_ov$ov.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench_ov$ov.err
done
timeout 300 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --bwd-overlap 1 > $O/bench_bicycle_ov1.json 2> $O/bench_bicycle.err; python -c "
import json;d=json.load(open('$O/bench_bicycle_ov1.json'));print('bicycle ov1',d['ms_per_step'],d['roofline']['all_kernels_ms'])"
