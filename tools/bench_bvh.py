"""LBVH rebuild alone on cuda:0 (CUDA events, 256 MiB L2 flush before every timed build, median of 20):
    python tools/bench_bvh.py [C2 C5 ...]
Algorithmic bytes: 424 B per triangle."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth  # noqa: E402


def main(names):
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    out = {}
    for name in names:
        v, f = synth.make_mesh(synth.CONFIGS[name])
        vt, ft = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
        w = R.restirbvhWorker(vt, ft)
        for _ in range(3):
            w.update_mesh(vt, ft)
        ts = []
        for _ in range(20):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            w.update_mesh(vt, ft)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        F = f.shape[0]
        out[name] = {"triangles": F, "ms": ms, "alg_GBps": 424 * F / ms / 1e6, "frac_of_hbm_peak": 424 * F / ms / 1e6 / peaks["hbm_gbs"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1:] or ["C2", "C5"])
