"""Manual robustness check at BASELINE config 5 sizes (2 M triangles, 2048 x 2048, 2k x 1k envmap): LBVH rebuild,
G-buffer, a short spp loop with backward; prints timings and sanity statistics.  Not part of the test-suite (seconds of
GPU time, 10+ GB of buffers).

    python tools/check_large.py [spp]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R, slangpy_shim  # noqa: E402


def main(spp):
    cfg = synth.CONFIGS["C5"]
    dev = torch.device("cuda", 0)
    W, H, mb = cfg["W"], cfg["H"], cfg["max_bounce"]
    n = W * H
    v, f = synth.make_mesh(cfg)
    env = torch.from_numpy(synth.envmap(*cfg["env"])).to(dev).requires_grad_(True)
    ro, rd = synth.camera_rays(W, H)
    vert, tri = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    worker = R.restirbvhWorker(vert, tri)
    k = slangpy_shim.get_kernels()
    mods = R.load_m_for_restir(W, H, device=dev, max_bounce=mb)
    mat = synth.ProceduralMaterial(0.0)
    for it in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        worker.update_mesh(vert, tri)
        torch.cuda.synchronize()
        t1 = time.time()
        occ, depth = torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
        pos, nrm = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
        prim, bary = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, 2, device=dev)
        k.gbuffer_primary(worker.packed, torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev), occ, pos, nrm, depth, prim,
                          bary, ws=slangpy_shim.workspace(dev, n))
        kdks = mat.sample_no_di_dense(pos) * occ
        normal = nrm.requires_grad_(True)
        kd = kdks[:, 0:3].contiguous().requires_grad_(True)
        rs = kdks[:, 4:6].contiguous().requires_grad_(True)
        outs = R.run_restir_di_with_pt(False, 1, 1, 1, mat, None, worker, *mods, env, occ, normal, depth, kd, rs,
                                       torch.from_numpy(rd).to(dev), pos, None, None, None, None, W, H, spp, 2, 2, 2.0, 0.1, 0.001,
                                       random_offset=11, max_bounce=mb)
        outs[0].mean().backward()
        torch.cuda.synchronize()
        t2 = time.time()
        print("iteration %d: F=%d N=%d  LBVH %.2f ms, render+backward (spp %d, %d indirect vertices) %.1f ms, coverage %.3f, "
              "finite %s, |grad_env| %.3e, peak memory %.1f GB" %
              (it, f.shape[0], n, 1e3 * (t1 - t0), spp, mb, 1e3 * (t2 - t1), float(occ.mean()),
               bool(torch.isfinite(outs[0]).all()), float(env.grad.abs().sum()), torch.cuda.max_memory_allocated() / 2 ** 30))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
