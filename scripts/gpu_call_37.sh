# last refresh: whole gpu suite, smoke, default bench line, launch list (k0 scatter with 8 resident blocks + pre-scaled weights)
O=gpurun_out/final4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
timeout 700 python bench.py > $O/bench_truck.json 2> $O/bench_truck.err; cut -c1-260 $O/bench_truck.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --only-timed --steps 1 --warmup 3 --no-reference-gpu > $O/ncu_launch.log 2>&1; wc -l $O/launches.csv
