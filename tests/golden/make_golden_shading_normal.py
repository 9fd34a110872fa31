"""Golden vectors for prepare_shading_normal from the REFERENCE's own pure-torch restatement of its CUDA kernel.

nerf/renderutils/ops.py ships `bsdf_prepare_shading_normal` (ops.py:82-112, reached with use_python=True) "for
validation" of normal.cu; this script imports that module in place from /root/reference (BUILD container only; the
reference does not travel), evaluates it in float32 on seeded inputs and differentiates it with torch.autograd.
Output: tests/golden/shading_normal_ref.npz (inputs, outputs and the six input gradients for the four flag settings).

    python tests/golden/make_golden_shading_normal.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OPS = "/root/reference/nerf/renderutils/ops.py"
FX, FY = 23, 17  # ragged against 8 x 8 blocks


def reference_ops():
    spec = importlib.util.spec_from_file_location("ref_renderutils_ops", REF_OPS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # imports torch / numpy only; the CUDA plugin is built lazily and never touched here
    return mod


def inputs(seed=0):
    rng = np.random.default_rng(seed)
    n = FX * FY
    pos = rng.uniform(-0.8, 0.8, (n, 3))
    view_pos = np.array([[0.3, 2.9, 1.1]])
    smooth = rng.standard_normal((n, 3)) * rng.uniform(0.2, 3.0, (n, 1))        # un-normalised on purpose
    tng = np.cross(smooth, rng.standard_normal((n, 3))) * rng.uniform(0.5, 2.0, (n, 1))
    geom = smooth / np.linalg.norm(smooth, axis=1, keepdims=True) + 0.3 * rng.standard_normal((n, 3))
    geom /= np.linalg.norm(geom, axis=1, keepdims=True)
    pert = rng.standard_normal((n, 3)) * np.array([[0.3, 0.3, 1.0]])             # some rows with z < 0
    # random frames: ~1 pixel in 20 lands in the bend branch (0 < dot(V, S) < 0.1), half are back-facing
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(pos=f(pos), view_pos=f(view_pos), perturbed_nrm=f(pert), smooth_nrm=f(smooth), smooth_tng=f(tng),
                geom_nrm=f(geom), grad_out=f(rng.standard_normal((n, 3))))


NAMES = ("pos", "view_pos", "perturbed_nrm", "smooth_nrm", "smooth_tng", "geom_nrm")


def main():
    ops = reference_ops()
    data = inputs()
    out = {k: v for k, v in data.items()}
    out["fx"], out["fy"] = np.int32(FX), np.int32(FY)
    for two_sided in (0, 1):
        for opengl in (0, 1):
            ts = []
            for k in NAMES:
                t = torch.from_numpy(data[k]).clone()
                t = t.view(1, FY, FX, 3) if t.shape[0] != 1 else t.view(1, 1, 1, 3)
                ts.append(t.requires_grad_(True))
            y = ops.bsdf_prepare_shading_normal(*ts, bool(two_sided), bool(opengl))
            g = torch.autograd.grad(y, ts, torch.from_numpy(data["grad_out"]).view(1, FY, FX, 3))
            tag = "ts%d_gl%d" % (two_sided, opengl)
            out["out_" + tag] = y.detach().reshape(-1, 3).numpy()
            for k, gk in zip(NAMES, g):
                out["g_%s_%s" % (k, tag)] = gk.reshape(-1, 3).numpy()
    # the call of nerf/renderer.py:1013: no normal map (perturbed_nrm None -> (0,0,1)), zero tangents, two-sided, OpenGL
    ts = [torch.from_numpy(data[k]).clone() for k in NAMES]
    ts[2] = torch.tensor([[0.0, 0.0, 1.0]])
    ts[4] = torch.zeros_like(ts[4])
    ts = [(t.view(1, FY, FX, 3) if t.shape[0] != 1 else t.view(1, 1, 1, 3)).requires_grad_(k in (0, 3, 5)) for k, t in enumerate(ts)]
    y = ops.bsdf_prepare_shading_normal(*ts, True, True)
    g = torch.autograd.grad(y, [ts[0], ts[3], ts[5]], torch.from_numpy(data["grad_out"]).view(1, FY, FX, 3))
    out["out_callsite"] = y.detach().reshape(-1, 3).numpy()
    for k, gk in zip(("pos", "smooth_nrm", "geom_nrm"), g):
        out["g_%s_callsite" % k] = gk.reshape(-1, 3).numpy()
    np.savez_compressed(os.path.join(HERE, "shading_normal_ref.npz"), **out)
    print("wrote", os.path.join(HERE, "shading_normal_ref.npz"), "%d arrays" % len(out))


if __name__ == "__main__":
    main()
