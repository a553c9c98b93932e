# k0 scatter tuning A/B: 3 (baseline) / 7 (8 resident blocks + pre-scaled weights) / 8 (groups of 8) / 9 (pre-scaled weights)
O=gpurun_out/call36; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_models.py -q --timeout 300 -x -rf -k "feature_kernel" > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-300 | head -5
for fk in 3 7 8 9 3; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-reduced-precision --feature-kernel $fk > $O/bench_fk$fk.json 2> $O/bench_fk$fk.err
python -c "
import json;d=json.load(open('$O/bench_fk$fk.json'));print('fk=$fk',d['ms_per_step'],d['roofline']['all_kernels_ms']['march_feature_bwd'])"
done
