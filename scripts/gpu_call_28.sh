# density-grid gradient against the fp64 scatter of the reference's per-sample gradients: per-sample scatter vs run-merging scatter
O=gpurun_out/call28; mkdir -p $O
timeout 900 python - > $O/density_truth.txt 2>&1 <<'PY'
import torch, json
from tests import parity_at_size as P
from tests.util import ref_ext
from unboundednerfpytorch_b200 import ops
for variant in (0, 1):
    ops.set_density_scatter(variant)
    out, _, _ = P.compare('truck_dense', torch.device('cuda:0'), ext=ref_ext())
    print(json.dumps({'variant': variant, 'grad density.grid': out['grad density.grid'], 'refself': out['refself density.grid'], 'truth density.grid': out['truth density.grid']}))
    del out; torch.cuda.empty_cache()
PY
cat $O/density_truth.txt | grep variant | cut -c1-900
timeout 1200 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_models.py -q --timeout 600 -rf > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $O/pytest.log | cut -c1-400 | head -20
