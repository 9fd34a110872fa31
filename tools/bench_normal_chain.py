"""Times the G-buffer normal chain at BASELINE config 2 sizes on one GPU (CUDA events, L2 flushed between launches):
auto_normals forward / backward (V = 250 k, F = 500 k), prepare_shading_normal forward / backward and the barycentric
gradient scatter (n = 800 x 800 pixels on the real primary-ray G-buffer of the C2 view), each with its algorithmic bytes
and the fraction of the measured HBM peak.  Prints one JSON line; also times the reference's own kernels
(oracle/_ref) for prepare_shading_normal when that library is present.

    python tools/bench_normal_chain.py [reps]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R, slangpy_shim  # noqa: E402


def main(reps):
    dev = torch.device("cuda", 0)
    cfg = synth.CONFIGS["C2"]
    W, H = cfg["W"], cfg["H"]
    n = W * H
    v, f = synth.make_mesh(cfg)
    vert, tri = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    V, F = vert.shape[0], tri.shape[0]
    k = slangpy_shim.get_kernels()
    worker = R.restirbvhWorker(vert, tri)
    worker.update_mesh(vert, tri)
    ro, rd = synth.camera_rays(W, H)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    vsum, vnrm = torch.empty_like(vert), torch.empty_like(vert)
    k.vertex_normals_fwd(vert, tri, vsum, vnrm)
    occ, depth = torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
    pos, smooth, geom = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
    prim, bary = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, 2, device=dev)
    ws = slangpy_shim.workspace(dev, n)
    k.gbuffer_primary(worker.packed, ro, rd, occ, pos, geom, depth, prim, bary, ws=ws)                        # face normals
    k.gbuffer_primary(worker.packed, ro, rd, occ, pos, smooth, depth, prim, bary, vnormal=vnrm, tri=tri, ws=ws)  # interpolated
    eye = ro[:1].contiguous()
    pert = torch.tensor([[0.0, 0.0, 1.0]], device=dev)
    tng = torch.zeros(n, 3, device=dev)
    ins = [pos, eye, pert, smooth, tng, geom]
    out, go = torch.empty(n, 3, device=dev), torch.randn(n, 3, device=dev)
    grads = [torch.empty(n, 3, device=dev) if i in (0, 3, 5) else None for i in range(6)]  # what renderer.py:1013 needs
    gvn, gv, gvn_in = torch.zeros(V, 3, device=dev), torch.zeros(V, 3, device=dev), torch.randn(V, 3, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = 6551.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    cases = {
        # name: (callable, algorithmic bytes)
        "vertex_normals_fwd": (lambda: k.vertex_normals_fwd(vert, tri, vsum, vnrm), F * 12 + V * 12 + V * 12 * 3),
        "vertex_normals_bwd": (lambda: k.vertex_normals_bwd(vert, tri, vsum, gvn_in, gv), F * 12 + V * 12 * 4),
        "shading_normal_fwd": (lambda: k.shading_normal_fwd(n, ins, 1, 1, out), n * 12 * 5),
        "shading_normal_bwd": (lambda: k.shading_normal_bwd(n, ins, 1, 1, go, grads), n * 12 * 8),
        "interpolate_bwd_normals": (lambda: k.interpolate_bwd(go, prim, bary, tri, gvn), n * 24 + int((occ > 0).sum()) * 12),
    }
    try:
        from oracle import ref as REF
        if REF.available():
            cases["reference_normal_cu_fwd"] = (lambda: REF.prepare_shading_normal_fwd(W, H, ins, 1, 1), n * 12 * 5)
            cases["reference_normal_cu_bwd"] = (lambda: REF.prepare_shading_normal_bwd(W, H, ins, 1, 1, go), n * 12 * 11)
    except Exception:
        pass
    res = {}
    for name, (fn, nbytes) in cases.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts))
        res[name] = {"us": round(us, 2), "alg_MB": round(nbytes / 1e6, 2), "GBps": round(nbytes / us / 1e3, 1),
                     "frac_hbm": round(nbytes / us / 1e3 / peak, 3)}
    print(json.dumps({"config": "C2 sizes: V=%d F=%d n=%d, foreground %d" % (V, F, n, int((occ > 0).sum())),
                      "hbm_peak_GBps": peak, "l2": "256 MiB flush before every timed launch", "kernels": res}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
