#!/bin/bash
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
from unboundednerfpytorch_b200 import _cabi; print('abi', _cabi.load().ubn_abi_version(), 'feature kernel', _cabi.load().ubn_get_feature_kernel())"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rf --deselect "tests/test_gpu_models.py::test_fused_rgbnet_vs_torch[tc3+fused]" --deselect "tests/test_gpu_models.py::test_fused_rgbnet_vs_torch[tc1+fused]" > gpurun_out/pytest_r02_i.log 2>&1
echo "--- pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|\[tma\]\|\[psnr tf32x1\]" gpurun_out/pytest_r02_i.log | cut -c1-500 | head -40
timeout 240 python -m pytest "tests/test_gpu_models.py::test_fused_rgbnet_vs_torch" -q -k "fused" -rf > gpurun_out/pytest_r02_i_ws.log 2>&1
echo "--- ws pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" gpurun_out/pytest_r02_i_ws.log | cut -c1-400 | head
for mode in fused4 fused; do
UBN_RGBNET_BWD_MODE=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r02_i_$mode.json 2> gpurun_out/bench_r02_i_$mode.err
echo "--- bench bwd=$mode rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_r02_i_$mode.json'));print(d['ms_per_step'],d['tail_ms']['value'],d['roofline']['all_kernels_ms'],d['fwd_only']['ms_per_step'])"; tail -2 gpurun_out/bench_r02_i_$mode.err
done
timeout 900 python scripts/parity_at_size_report.py > gpurun_out/parity_r02_i.jsonl 2> gpurun_out/parity_r02_i.err
echo "--- parity rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/parity_r02_i.jsonl'):
    d=json.loads(l)
    if 'config' in d:
        print(d['config'], 'flips', d.get('flips'), {k:(round(v['rel_scale'],9) if isinstance(v,dict) and 'rel_scale' in v else v) for k,v in d.items() if k in ('rgb_marched','depth','weights','raw_alpha','raw_density','raw_rgb','grad density.grid','grad k0.grid','refself density.grid','n_relu_ambiguous')}, {k:v for k,v in d.items() if k.startswith('truth')})
PY
