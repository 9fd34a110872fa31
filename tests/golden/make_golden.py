"""Generates the committed golden fixtures (run in the BUILD container only; /root/reference does not travel).

  refdriver_<cfg>.npz  the REFERENCE's own, unmodified host code -- nerf/renderer_restir.py, nerf/ScreenSpaceReSTIR/
                       {Resampling,GenerateLightTiles,Denoising}.py imported in place from /root/reference -- driving the
                       kernels through the slangpy-protocol shim (host-check flavour of the product kernels, CPU tensors).
                       Pins everything the reference's Python decides: frame-index schedule, reservoir / bounce buffer
                       ping-pong, accumulation, denoise + composite, and the autograd wiring including the stale-alias
                       semantics of the saved tensors.  The kernels themselves are pinned by the oracle.
  oracle_<cfg>.npz     oracle outputs on the same inputs (regression pin of the oracle; the reference ships no vectors).

    python tests/golden/make_golden.py

  --real-slangpy       (for a maintainer whose machine has the real `slangpy` + slangc + a GPU; NOT runnable in the build
                       container or on the GPU box of this project, hence untested here)  Runs the same unmodified
                       reference driver on the same inputs with the REAL Slang kernels on CUDA and diffs every output
                       against the committed fixture -> profiles/slangpy_pin_<cfg>.json.  This is the one command that
                       turns "parity unpinned"  into a pin: integer-like outputs are expected to agree up to
                       the decision flips profiles/numerics_sensitivity.json bounds, floats to ~1e-4.
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


def import_reference_driver(real_slangpy=False):
    """Import the reference's renderer_restir.py unmodified, with its unavailable third-party imports stubbed."""
    for name in ("pyexr", "torchvision", "torchvision.utils"):
        try:
            importlib.import_module(name)
        except ImportError:
            sys.modules.setdefault(name, types.ModuleType(name))
    if not real_slangpy:
        sys.modules["slangpy"] = slangpy_shim
    sys.path.insert(0, REF)
    # the reference hard-codes device='cuda'; redirect allocations to the CPU for this harness
    for fn in (() if real_slangpy else ("zeros", "ones", "empty")):
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
    if not real_slangpy:  # (with the real kernels the reference runs entirely as it is)
        ref.safe_l2_normalize = lambda x, dim=-1: MINE._normalize_rows(x)
        ref.make_sampleable = MINE.make_sampleable
    return ref


def run_reference_driver(ref, sc, device="cpu", own_lbvh=False):
    W, Hh = sc["W"], sc["H"]
    t = lambda a: H.t(a).to(device)
    worker = ref.restirbvhWorker(t(sc["vert"]), t(sc["tri"]))
    if not own_lbvh:
        info, aabb = t(sc["bvh"].info.copy()), t(sc["bvh"].aabb.copy())
        worker.update_bvh = lambda: (info, aabb)  # LBVH from the oracle (bit-identical to the CUDA builder, tested on GPU)
    worker.update_mesh(t(sc["vert"]), t(sc["tri"]))
    cwd = os.getcwd()
    os.chdir(REF)  # the reference loads its .slang files by relative path
    try:
        mods = ref.load_m_for_restir(W, Hh)
    finally:
        os.chdir(cwd)
    g = {k: t(v) for k, v in sc["gbuffer"].items()}
    env = t(sc["env"]).requires_grad_(True)
    normal = g["normal_map"].clone().requires_grad_(True)
    kd = g["diffuse_map"].clone().requires_grad_(True)
    rs = g["roughness_specular"].clone().requires_grad_(True)
    np.random.seed(SEED)
    outs = ref.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, worker, *mods, env,
                                     g["occ_map"], normal, g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"], None,
                                     None, None, None, W, Hh, SPP, DENOISE_ITER, STEP, *PHI)
    w = torch.linspace(0.5, 1.5, W * Hh * 3).reshape(W * Hh, 3).to(device)
    (outs[0] * w).sum().backward()
    names = ("final_color", "denoised_diffuse", "denoised_spec", "denoised_indirect", "denoised_indirect_diff",
             "denoised_indirect_spec")
    out = {n: o.detach().cpu().numpy() for n, o in zip(names, outs)}
    out.update(grad_env=env.grad.cpu().numpy(), grad_normal=normal.grad.cpu().numpy(), grad_kd=kd.grad.cpu().numpy(),
               grad_rs=rs.grad.cpu().numpy())
    np.random.seed(SEED)
    out["random_offset"] = np.int64(np.random.randint(2 ** 20))
    if own_lbvh:
        out["bvh_info"] = worker.LBVHNode_info.cpu().numpy()
        out["bvh_aabb"] = worker.LBVHNode_aabb.cpu().numpy()
    return out


def pin_against_real_slangpy():
    """--real-slangpy: the reference's driver AND the reference's Slang kernels (LBVH included) against the fixture."""
    import json
    ref = import_reference_driver(real_slangpy=True)
    report = {}
    for name in ("T0",):
        fx = np.load(os.path.join(HERE, "refdriver_%s.npz" % name))
        sc = P.scene(name, float(fx["metallic"]))
        out = run_reference_driver(ref, sc, device="cuda", own_lbvh=True)
        rep = {"lbvh_topology_equal": bool(np.array_equal(out["bvh_info"].reshape(-1), sc["bvh"].info.reshape(-1))),
               "lbvh_boxes_equal": bool(np.array_equal(out["bvh_aabb"].reshape(-1), sc["bvh"].aabb.reshape(-1)))}
        for k in fx.files:
            if k in ("metallic", "random_offset") or k not in out:
                continue
            a, b = out[k].astype(np.float64).reshape(-1), fx[k].astype(np.float64).reshape(-1)
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
            rep[k] = {"max_rel_err": float(err.max()), "p999_rel_err": float(np.quantile(err, 0.999)),
                      "fraction_bit_equal": float((out[k].reshape(-1) == fx[k].reshape(-1)).mean()),
                      "rel_diff_of_mean": float(abs(a.mean() - b.mean()) / max(abs(b.mean()), 1e-12))}
            print(name, k, rep[k])
        report[name] = rep
    path = os.path.join(ROOT, "profiles", "slangpy_pin.json")
    json.dump(report, open(path, "w"), indent=1)
    print("->", path)


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
    if "--real-slangpy" in sys.argv:
        pin_against_real_slangpy()
    else:
        main()
