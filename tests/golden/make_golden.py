"""Generates the committed golden fixtures (run in the BUILD container only; /root/reference does not travel).

  refdriver_<cfg>.npz  the REFERENCE's own, unmodified host code -- nerf/renderer_restir.py, nerf/ScreenSpaceReSTIR/
                       {Resampling,GenerateLightTiles,Denoising}.py imported in place from /root/reference -- driving the
                       kernels through the slangpy-protocol shim (host-check flavour of the product kernels, CPU tensors).
                       Pins everything the reference's Python decides: frame-index schedule, reservoir / bounce buffer
                       ping-pong, accumulation, denoise + composite, and the autograd wiring including the stale-alias
                       semantics of the saved tensors.  The kernels themselves are pinned by the oracle.
  oracle_<cfg>.npz     oracle outputs on the same inputs (regression pin of the oracle; the reference ships no vectors).

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

import hostcheck as H  # noqa: E402
import parity as P  # noqa: E402
from mirres_restir_nerf_mesh_b200 import slangpy_shim, synth, renderer_restir as MINE  # noqa: E402

SPP, DENOISE_ITER, STEP, PHI = 3, 2, 2, (2.0, 0.1, 0.001)  # nerf/renderer.py:1103-1108
SEED = 0


def import_reference_driver():
    """Import the reference's renderer_restir.py unmodified, with its unavailable third-party imports stubbed."""
    for name in ("pyexr", "torchvision", "torchvision.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["slangpy"] = slangpy_shim
    sys.path.insert(0, REF)
    # the reference hard-codes device='cuda'; redirect allocations to the CPU for this harness
    for fn in ("zeros", "ones", "empty"):
        orig = getattr(torch, fn)

        def patched(*a, _orig=orig, **k):
            if k.get("device") == "cuda":
                k["device"] = "cpu"
            return _orig(*a, **k)
        setattr(torch, fn, patched)
    ref = importlib.import_module("nerf.renderer_restir")
    # Two documented contract deviations are injected so that the fixture is reproducible on any device; everything
    # else is the reference's own code:
    #  * F.normalize -> the same formula with a defined rounding order (device-dependent reduction order otherwise);
    #  * make_sampleable's torch.sum / torch.cumsum (parallel, device-dependent order) -> sequential fp32 prefix sums
    #    (include/mirres_b200.h: mirres_env_build_distribution).  tests/test_gpu.py::test_env_distribution_and_tiles
    #    checks that the reference's own two-kernel + torch-scan protocol stays within 1e-4 of it.
    ref.safe_l2_normalize = lambda x, dim=-1: MINE._normalize_rows(x)
    ref.make_sampleable = MINE.make_sampleable
    return ref


def run_reference_driver(ref, sc):
    W, Hh = sc["W"], sc["H"]
    worker = ref.restirbvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    info, aabb = H.t(sc["bvh"].info.copy()), H.t(sc["bvh"].aabb.copy())
    worker.update_bvh = lambda: (info, aabb)  # LBVH from the oracle (bit-identical to the CUDA builder, tested on GPU)
    worker.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    cwd = os.getcwd()
    os.chdir(REF)  # the reference loads its .slang files by relative path
    try:
        mods = ref.load_m_for_restir(W, Hh)
    finally:
        os.chdir(cwd)
    g = {k: H.t(v) for k, v in sc["gbuffer"].items()}
    env = H.t(sc["env"]).requires_grad_(True)
    normal = g["normal_map"].clone().requires_grad_(True)
    kd = g["diffuse_map"].clone().requires_grad_(True)
    rs = g["roughness_specular"].clone().requires_grad_(True)
    np.random.seed(SEED)
    outs = ref.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, worker, *mods, env,
                                     g["occ_map"], normal, g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"], None,
                                     None, None, None, W, Hh, SPP, DENOISE_ITER, STEP, *PHI)
    w = torch.linspace(0.5, 1.5, W * Hh * 3).reshape(W * Hh, 3)
    (outs[0] * w).sum().backward()
    names = ("final_color", "denoised_diffuse", "denoised_spec", "denoised_indirect", "denoised_indirect_diff",
             "denoised_indirect_spec")
    out = {n: o.detach().numpy() for n, o in zip(names, outs)}
    out.update(grad_env=env.grad.numpy(), grad_normal=normal.grad.numpy(), grad_kd=kd.grad.numpy(), grad_rs=rs.grad.numpy())
    np.random.seed(SEED)
    out["random_offset"] = np.int64(np.random.randint(2 ** 20))
    return out


def main():
    H.activate()
    ref = import_reference_driver()
    for name, metallic in (("T0", 0.25),):
        sc = P.scene(name, metallic)
        out = run_reference_driver(ref, sc)
        np.savez_compressed(os.path.join(HERE, "refdriver_%s.npz" % name), metallic=np.float32(metallic), **out)
        o = P.oracle_run(sc, random_offset=int(out["random_offset"]), spp=SPP)
        last = o["snapshots"][-1]
        np.savez_compressed(os.path.join(HERE, "oracle_%s.npz" % name), metallic=np.float32(metallic),
                            random_offset=out["random_offset"], bvh_info=sc["bvh"].info, bvh_aabb=sc["bvh"].aabb,
                            sorted_codes=sc["bvh"].sorted_codes, res_ld=last["res"][0], res_M=last["res"][2],
                            res_w=last["res"][3], vis=last["vis"], color=o["color"], diff=o["diff"], spec=o["spec"],
                            color_1=o["color_1"], hit=sc["hit"], prim=sc["prim"], t=sc["t"])
        print(name, "final mean", out["final_color"].mean(), "grad_env abs sum", np.abs(out["grad_env"]).sum(),
              "random_offset", int(out["random_offset"]))


if __name__ == "__main__":
    main()
