# gather variant 6 (8 samples x 3 quad lanes per instruction) vs variant 3
O=gpurun_out/call30; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_models.py -q --timeout 600 -x -rf -k "feature_kernel or golden" > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-400 | head
for fk in 3 6; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel $fk > $O/bench_fk$fk.json 2> $O/bench_fk$fk.err
echo "--- bench fk=$fk rc=$?"; python -c "
import json;d=json.load(open('$O/bench_fk$fk.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench_fk$fk.err
done
timeout 300 python bench.py --workload bicycle --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel 6 > $O/bench_bicycle_fk6.json 2> $O/bench_bicycle.err; python -c "
import json;d=json.load(open('$O/bench_bicycle_fk6.json'));print('bicycle fk6',d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"
timeout 300 python bench.py --workload garden --steps 3 --warmup 3 --feature-kernel 6 > $O/bench_garden_fk6.json 2> $O/bench_garden.err; cut -c1-200 $O/bench_garden_fk6.json
