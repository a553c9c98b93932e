# bench JSON line check (forward roofline key, TF32x1 key)
O=gpurun_out/call35; mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline --no-reference-gpu > $O/bench.json 2> $O/bench.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench.json'));print(d['ms_per_step'],d['fwd_only'],d['reduced_precision_tf32x1']['ms_per_step'])"; tail -2 $O/bench.err
timeout 400 python bench.py --workload bicycle --no-cpu-baseline --no-reference-gpu > $O/bench_bicycle.json 2> $O/bench_bicycle.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_bicycle.json'));print(d['ms_per_step'],d['fwd_only'])"
