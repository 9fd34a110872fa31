"""Golden vectors for auto_normals from the REFERENCE's own function (meshutils.py:14-39, top level of the reference).

Imports the reference's meshutils.py in place from /root/reference (BUILD container only) with its one unavailable
third-party import (pymeshlab, used by other functions of that file) stubbed and the hard-coded device='cuda' of the
fallback constant redirected to the CPU; evaluates auto_normals in float32 on a small seeded mesh that includes an
unreferenced vertex and a zero-area triangle, and differentiates it with torch.autograd.
Output: tests/golden/auto_normals_ref.npz.

    python tests/golden/make_golden_auto_normals.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/meshutils.py"


def reference_meshutils():
    sys.modules.setdefault("pymeshlab", types.ModuleType("pymeshlab"))
    orig = torch.tensor

    def tensor(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return orig(*a, **k)
    torch.tensor = tensor
    spec = importlib.util.spec_from_file_location("ref_meshutils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def mesh(nu=12, nv=9, seed=0):
    """bumpy torus grid (closed, every vertex shared by six triangles) + one unreferenced vertex + one zero-area triangle"""
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(nu) * 2 * np.pi / nu, np.arange(nv) * 2 * np.pi / nv, indexing="ij")
    r = 0.25 + 0.03 * rng.standard_normal(u.shape)
    x = (0.6 + r * np.cos(v)) * np.cos(u)
    y = (0.6 + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v)
    vert = np.stack([x, y, z], -1).reshape(-1, 3)
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    tri = []
    for i in range(nu):
        for j in range(nv):
            tri.append([idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)])
            tri.append([idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)])
    vert = np.concatenate([vert, [[0.1, 0.2, 0.3]]], 0)       # unreferenced vertex -> fallback (0, 0, 1)
    tri.append([3, 3, 17])                                     # zero-area triangle -> contributes nothing
    return vert.astype(np.float32), np.asarray(tri, np.int32)


def main():
    mu = reference_meshutils()
    vert, tri = mesh()
    v = torch.from_numpy(vert).clone().requires_grad_(True)
    vn, tri_out = mu.auto_normals(v, torch.from_numpy(tri))
    assert tri_out.shape == tri.shape
    g = np.random.default_rng(1).standard_normal(vert.shape).astype(np.float32)
    (gv,) = torch.autograd.grad(vn, v, torch.from_numpy(g))
    np.savez_compressed(os.path.join(HERE, "auto_normals_ref.npz"), vert=vert, tri=tri, vnrm=vn.detach().numpy(),
                        grad_vnrm=g, grad_vert=gv.numpy())
    print("wrote auto_normals_ref.npz: V=%d F=%d" % (vert.shape[0], tri.shape[0]))


if __name__ == "__main__":
    main()
