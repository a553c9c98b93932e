# Round-2 evidence run (session 2) on ONE B200: tests, smoke, bench lines, parity report, ncu launch list + --set full of the hot kernels.
O=gpurun_out/final2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $O/bench_truck.json 2> $O/bench_truck.err; cut -c1-300 $O/bench_truck.json
timeout 400 python bench.py --workload bicycle --no-cpu-baseline > $O/bench_bicycle.json 2> $O/bench_bicycle.err; cut -c1-200 $O/bench_bicycle.json
timeout 300 python bench.py --workload garden --steps 3 --warmup 3 > $O/bench_garden.json 2> $O/bench_garden.err; cut -c1-200 $O/bench_garden.json
timeout 300 python bench.py --workload garden --steps 3 --warmup 3 --tma > $O/bench_garden_tma.json 2> $O/bench_garden_tma.err; cut -c1-200 $O/bench_garden_tma.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 3 --no-reference-gpu > $O/ncu_launch.log 2>&1; wc -l $O/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_march_feature_v3|k_shade_bwd_fused|k_shade_fwd_tc|k_tv_adam|k_shade_dw2_tc|k_march_feature_v2|k_march_density" -s 24 -c 10 -o $O/new_kernels -f python bench.py --only-timed --steps 1 --warmup 3 --no-reference-gpu > $O/ncu_full.log 2>&1; ls -la $O/*.ncu-rep
timeout 600 python scripts/parity_at_size_report.py > $O/parity_at_size.jsonl 2> $O/parity_at_size.err; wc -l $O/parity_at_size.jsonl
