#!/bin/bash
mkdir -p gpurun_out
( for t in 0 1; do for a in 0 5 6 7; do for b in 0 5 6 7; do
    if [ $t = 1 ] && [ $a != 0 ]; then continue; fi
    timeout 30 scripts/bin/probe_umma_layouts $t $a $b || echo "{\"probe\": \"umma_layout\", \"args\": \"$t $a $b\", \"rc\": $?}"
  done; done; done ) > gpurun_out/probe_umma_b.jsonl 2>&1
echo "--- probe"; cat gpurun_out/probe_umma_b.jsonl | cut -c1-220
timeout 1200 python scripts/parity_at_size_report.py > gpurun_out/parity_r02_b.jsonl 2> gpurun_out/parity_r02_b.err
echo "--- parity rc=$?"; cat gpurun_out/parity_r02_b.jsonl; tail -5 gpurun_out/parity_r02_b.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/pytest_r02_b.log 2>&1
echo "--- pytest rc=$?"; tail -30 gpurun_out/pytest_r02_b.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err
echo "--- bench rc=$?"; cat gpurun_out/bench_r02_b.json | cut -c1-3000; tail -3 gpurun_out/bench_r02_b.err
