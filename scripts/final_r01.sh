# Round-end evidence run on ONE B200 (gpurun): tests, smoke, bench lines, reference arm, ncu launch list + --set full of the
# kernels that changed last.  Outputs under gpurun_out/final/ (copied into profiles/ by hand).
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > $O/bench_truck.json 2> $O/bench_truck.err; cut -c1-260 $O/bench_truck.json
timeout 400 python bench.py --workload bicycle --no-cpu-baseline > $O/bench_bicycle.json 2> $O/bench_bicycle.err; cut -c1-200 $O/bench_bicycle.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-200 $O/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 3 > $O/ncu_launch.log 2>&1; wc -l $O/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_total_variation_stream|k_march_density_bwd" -s 6 -c 2 -o $O/tail_kernels -f python bench.py --only-timed --steps 1 --warmup 3 > $O/ncu_full.log 2>&1; ls -la $O/*.ncu-rep
