#!/bin/bash
# 4-GPU box: peer-tail NCCL exactness at world 2 and 4 with the k0 check first and diagnostics
O=gpurun_out/final4gpu; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_tail.py -q --timeout 500 -rf -k "nccl" > $O/pytest_4gpu_b.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  .*rank\|passed\|failed\|FAILED\|skipped" $O/pytest_4gpu_b.log | cut -c1-500 | head -10
