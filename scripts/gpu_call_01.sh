#!/bin/bash
# first GPU call of round 2: UMMA layout probe, parity report at size, full gpu test suite, short bench
mkdir -p gpurun_out
nvidia-smi -L
( for t in 0 1; do for a in 0 1 2 3 4; do for b in 0 1 2 3 4; do
    if [ $t = 1 ] && [ $a != 0 ]; then continue; fi
    timeout 30 scripts/bin/probe_umma_layouts $t $a $b || echo "{\"probe\": \"umma_layout\", \"args\": \"$t $a $b\", \"rc\": $?}"
  done; done; done ) > gpurun_out/probe_umma.jsonl 2>&1
echo "--- probe"; cat gpurun_out/probe_umma.jsonl | cut -c1-200
timeout 1200 python scripts/parity_at_size_report.py > gpurun_out/parity_r02_a.jsonl 2> gpurun_out/parity_r02_a.err
echo "--- parity rc=$?"; cat gpurun_out/parity_r02_a.jsonl; tail -5 gpurun_out/parity_r02_a.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/pytest_r02_a.log 2>&1
echo "--- pytest rc=$?"; tail -40 gpurun_out/pytest_r02_a.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err
echo "--- bench rc=$?"; cat gpurun_out/bench_r02_a.json | cut -c1-3000
