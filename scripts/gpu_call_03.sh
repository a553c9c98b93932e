#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/probe_mean_order.py > gpurun_out/probe_mean_order.jsonl 2>&1
echo "--- mean order"; cat gpurun_out/probe_mean_order.jsonl | cut -c1-600
timeout 1500 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_callers_unchanged.py -q --timeout 900 -rf > gpurun_out/pytest_r02_c.log 2>&1
echo "--- pytest rc=$?"; grep -n "parity-at-size\|^E  \|passed\|failed" gpurun_out/pytest_r02_c.log | cut -c1-6000
