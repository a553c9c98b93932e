# Round-2 FINAL evidence run on ONE B200 (gpurun): tests, smoke, bench lines, render workloads, parity report, ncu launch list + --set full
# of the hot kernels.  Outputs under gpurun_out/final3/ (summaries are copied into profiles/ by scripts/collect_final_r02.sh).
O=gpurun_out/final3; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
timeout 700 python bench.py > $O/bench_truck.json 2> $O/bench_truck.err; cut -c1-300 $O/bench_truck.json
timeout 400 python bench.py --workload bicycle --no-cpu-baseline > $O/bench_bicycle.json 2> $O/bench_bicycle.err; cut -c1-200 $O/bench_bicycle.json
timeout 300 python bench.py --workload garden --steps 3 --warmup 3 > $O/bench_garden.json 2> $O/bench_garden.err; cut -c1-200 $O/bench_garden.json
timeout 300 python bench.py --workload garden --steps 3 --warmup 3 --tma > $O/bench_garden_tma.json 2> $O/bench_garden_tma.err; cut -c1-200 $O/bench_garden_tma.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 3 --no-reference-gpu > $O/ncu_launch.log 2>&1; wc -l $O/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_march_feature_v3|k_march_feature_bwd_slab|k_shade_bwd_fused|k_shade_fwd_tc|k_tv_adam|k_shade_dw2|k_march_density" -s 24 -c 8 -o $O/hot_kernels -f python bench.py --only-timed --steps 1 --warmup 3 --no-reference-gpu > $O/ncu_full.log 2>&1; ls -la $O/*.ncu-rep
timeout 600 python scripts/parity_at_size_report.py > $O/parity_at_size.jsonl 2> $O/parity_at_size.err; wc -l $O/parity_at_size.jsonl
