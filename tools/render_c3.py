"""Manual check of BASELINE config 3 (novel-view eval render, forward only): the C2 scene at 800 x 800 with --spp N through
run_restir_di_with_pt under torch.no_grad(); prints device time and path samples/s.  Not part of the test-suite.

    python tools/render_c3.py [spp]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R, slangpy_shim  # noqa: E402


def main(spp):
    cfg = synth.CONFIGS["C3"]
    dev = torch.device("cuda", 0)
    W, H, mb = cfg["W"], cfg["H"], cfg["max_bounce"]
    n = W * H
    v, f = synth.make_mesh(cfg)
    env = torch.from_numpy(synth.envmap(*cfg["env"])).to(dev)
    vert, tri = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    pose = torch.from_numpy(synth.camera_pose(view=0)).to(dev)
    worker = R.restirbvhWorker(vert, tri)
    k = slangpy_shim.get_kernels()
    mods = R.load_m_for_restir(W, H, device=dev, max_bounce=mb)
    mat = synth.ProceduralMaterial(0.0)
    with torch.no_grad():
        for it in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            worker.update_mesh(vert, tri)
            ro, rd = synth.camera_rays_torch(W, H, pose)
            occ, depth = torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
            pos, nrm = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
            k.gbuffer_primary(worker.packed, ro, rd, occ, pos, nrm, depth, ws=slangpy_shim.workspace(dev, n))
            kd, rs = mat.gbuffer_materials(pos, occ)
            outs = R.run_restir_di_with_pt(False, 1, 1, 1, mat, None, worker, *mods, env, occ, nrm, depth, kd, rs, rd, pos,
                                           None, None, None, None, W, H, spp, 2, 2, 2.0, 0.1, 0.001, random_offset=11,
                                           max_bounce=mb)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("render %d: %dx%d spp %d forward only: %.1f ms, %.3e path samples/s, finite %s, mean %.4f, peak memory %.1f GB"
                  % (it, W, H, spp, ms, n * spp / (ms * 1e-3), bool(torch.isfinite(outs[0]).all()), float(outs[0].mean()),
                     torch.cuda.max_memory_allocated() / 2 ** 30))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 512)
