# copy the summaries of the final evidence run (gpurun_out/final3) into profiles/ (tracked)
O=gpurun_out/final3
cp $O/bench_truck.json profiles/r02_bench_truck_final.json
cp $O/bench_bicycle.json profiles/r02_bench_bicycle_final.json
cp $O/bench_garden.json profiles/r02_bench_garden_1gpu.json
cp $O/bench_garden_tma.json profiles/r02_bench_garden_1gpu_tma.json
cp $O/launches.csv profiles/r02_launches_truck_final.csv
grep -v "^\[ref-cuda\]" $O/parity_at_size.jsonl > profiles/r02_parity_at_size.jsonl
tail -5 $O/pytest.log > profiles/r02_pytest_gpu_tail_final.txt; cat $O/smoke.txt >> profiles/r02_pytest_gpu_tail_final.txt
python scripts/ncu_summary.py $O/hot_kernels.ncu-rep > profiles/r02_ncu_full_summary_final.txt
python scripts/make_traffic_json.py $O/hot_kernels.ncu-rep
