# dW2 short accumulation chains + running sum in TMEM; density-bwd scan through smem; slab-major k0 scatter A/B
O=gpurun_out/call17; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_at_size.py -q --timeout 600 -x -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-300 | head -20
timeout 60 scripts/_bin/probe_sincos | tee $O/probe_sincos.json
for fk in 1 3; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --feature-kernel $fk > $O/bench_fk$fk.json 2> $O/bench_fk$fk.err
echo "--- bench fk=$fk rc=$?"; python -c "
import json;d=json.load(open('$O/bench_fk$fk.json'));print(d['ms_per_step'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['tail_ms']['value'])"; tail -2 $O/bench_fk$fk.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 2 --no-reference-gpu --feature-kernel 3 > $O/ncu_launch.log 2>&1; grep -E "dw2|bwd_slab|density_bwd|loss_finish|fused_ws" $O/launches.csv | tail -6 | cut -d, -f5,13-15 | cut -c1-200
