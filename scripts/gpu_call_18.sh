# fast density kernels (compile-time slabs, clamped cells, sincosf, branch-free pair reductions); slab-major scatter with x-range split
O=gpurun_out/call18; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_at_size.py tests/test_gpu_ops.py tests/test_gpu_callers_unchanged.py -q --timeout 600 -x -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-300 | head -20
for fk in 3 4 5; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel $fk > $O/bench_fk$fk.json 2> $O/bench_fk$fk.err
echo "--- bench fk=$fk rc=$?"; python -c "
import json;d=json.load(open('$O/bench_fk$fk.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench_fk$fk.err
done
timeout 300 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/bench_bicycle.json 2> $O/bench_bicycle.err; python -c "
import json;d=json.load(open('$O/bench_bicycle.json'));print('bicycle',d['ms_per_step'],d['roofline']['all_kernels_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_march_feature_bwd_slab|k_march_density" -s 9 -c 3 -o $O/march_kernels -f python bench.py --only-timed --steps 1 --warmup 2 --no-reference-gpu --feature-kernel 4 > $O/ncu_full.log 2>&1; ls -la $O/*.ncu-rep
