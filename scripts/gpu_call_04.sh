#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rf > gpurun_out/pytest_r02_d.log 2>&1
echo "--- pytest rc=$?"; grep -n "parity-at-size" gpurun_out/pytest_r02_d.log | cut -c1-1200; grep -n "^E  \|passed\|failed\|FAILED" gpurun_out/pytest_r02_d.log | cut -c1-400 | head -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_d_peer.json 2> gpurun_out/bench_r02_d_peer.err
echo "--- bench peer rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_d_peer.json'));print(d['ms_per_step'],d['tail_ms'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'],d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_r02_d_peer.err
UBN_BENCH_TAIL=pipelined timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_d_pipe.json 2> gpurun_out/bench_r02_d_pipe.err
echo "--- bench pipelined rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_d_pipe.json'));print(d['ms_per_step'],d['tail_ms'])"
