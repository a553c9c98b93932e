"""profiles/traffic.json from ncu --set full captures: per bench.py kernel key, dram__bytes_read.sum + dram__bytes_write.sum per launch.

    python scripts/make_traffic_json.py gpurun_out/final3/hot_kernels.ncu-rep [more.ncu-rep ...]

bench.py reads the file into roofline.traffic (labelled static: the capture is not repeated inside the timed run).  A key that is
served by two launches (rgbnet_bwd = first backward launch + dW2 launch) gets the sum; when a kernel appears several times in the
captures the last one wins."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# substring of the kernel name -> bench.py key
KEYS = [('k_march_density_fwd', 'march_density_fwd'), ('k_march_density_bwd', 'march_density_bwd'),
        ('k_march_feature_v3', 'march_feature_fwd'), ('k_march_feature_bwd_slab', 'march_feature_bwd'),
        ('k_shade_fwd_tc', 'rgbnet_fwd'), ('k_shade_bwd_fused_ws', 'rgbnet_bwd_launch1'), ('k_shade_dw2', 'rgbnet_bwd_dw2'),
        ('k_tv_adam_peer', 'tv_adam_peer'), ('k_tv_adam_stream', 'tv_adam_stream')]


def unit_scale(u):
    return {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u]


def main(paths):
    out = {}
    for path in paths:
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        r, w = idx['dram__bytes_read.sum'], idx['dram__bytes_write.sum']
        for d in data:
            name = d[idx['Kernel Name']]
            for sub, key in KEYS:
                if sub in name:
                    out[key] = int(float(d[r]) * unit_scale(units[r]) + float(d[w]) * unit_scale(units[w]))
    if 'rgbnet_bwd_launch1' in out and 'rgbnet_bwd_dw2' in out:
        out['rgbnet_bwd'] = out['rgbnet_bwd_launch1'] + out['rgbnet_bwd_dw2']
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1:])
