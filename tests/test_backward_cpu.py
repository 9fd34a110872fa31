"""Hand-derived backward kernels (host-check flavour) against float64 autograd restatements (oracle/backward.py)."""
import numpy as np
import pytest
import torch

import hostcheck as H
import parity as P
from oracle import backward as B


def assert_grad_close(got, want, name, rtol=1e-3):
    """Gradient tolerance of the north star (1e-3): relative to the tensor's scale everywhere, and per row (pixel)
    relative for all but a handful of ill-conditioned rows (fp32 cancellation at specular peaks)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    scale = np.abs(want).max()
    assert np.isfinite(got).all(), name
    assert np.abs(got - want).max() <= rtol * scale + 1e-12, (name, np.abs(got - want).max(), scale)
    rows = np.abs(want).max(axis=1) > 1e-6 * scale
    if rows.any():
        rel = (np.abs(got - want).max(axis=1) / (np.abs(want).max(axis=1) + 1e-30))[rows]
        assert (rel > rtol).mean() <= 0.01, (name, (rel > rtol).mean())
        assert rel.max() <= 5e-2, (name, rel.max())


@pytest.fixture(scope="module", autouse=True)
def _bind_hostcheck():
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    H.activate()
    yield
    slangpy_shim.set_kernels(None)


@pytest.fixture(scope="module")
def case():
    sc = P.scene("T1", 0.3)
    ref = P.oracle_run(sc, spp=1)
    return sc, ref


def test_final_shading_backward(case):
    sc, ref = case
    s, g = ref["snapshots"][0], ref["prepared"]
    W, Hh = sc["W"], sc["H"]
    n = W * Hh
    rng = np.random.default_rng(0)
    gc, gd, gs = [rng.standard_normal((n, 3)).astype(np.float32) for _ in range(3)]
    gN, gK, gR, gL = torch.zeros(n, 3), torch.zeros(n, 3), torch.zeros(n, 2), torch.zeros(n, 3)
    H.kernels().final_shading_bwd(H.t(s["fs_dir"]), H.t(s["fs_dist"]), H.t(s["fs_Li"]), W, Hh, H.t(g["occ_map"]),
                                  H.t(g["normal_map"]), H.t(g["ray_dir_map"]), H.t(g["diffuse_map"]),
                                  H.t(g["roughness_specular"]), H.t(gc), H.t(gd), H.t(gs), gN, gK, gR, gL)
    rN, rK, rR, rL = B.final_shading_grads(s["fs_dir"], s["fs_dist"], s["fs_Li"], g["occ_map"], g["normal_map"],
                                           g["ray_dir_map"], g["diffuse_map"], g["roughness_specular"], gc, gd, gs)
    assert (np.abs(rN).max(axis=1) > 0).sum() > 100
    for a, b, nm in ((gN, rN, "normal"), (gK, rK, "kd"), (gR, rR, "rough_metal"), (gL, rL, "Li")):
        assert_grad_close(a.numpy(), b, nm)


def test_eval_final_backward(case):
    sc, ref = case
    s = ref["snapshots"][0]
    W, Hh = sc["W"], sc["H"]
    He, We = sc["env"].shape[:2]
    rng = np.random.default_rng(1)
    gLi = rng.standard_normal((W * Hh, 3)).astype(np.float32)
    ge = torch.zeros(He * We, 3)
    H.kernels().eval_final_bwd([H.t(a) for a in s["res"]], We, He, W, Hh, H.t(s["vis"]), H.t(gLi), ge)
    want = B.eval_final_grad_env(s["res"][0], s["res"][3], s["vis"], gLi, We, He)
    assert (want != 0).any(axis=1).sum() > 50
    assert np.abs(ge.numpy() - want).max() <= 1e-4 * np.abs(want).max()


@pytest.mark.parametrize("step", [2, 1])
def test_eaw_backward_gather_equals_autograd(case, step):
    sc, ref = case
    g = ref["prepared"]
    W, Hh = sc["W"], sc["H"]
    n = W * Hh
    rng = np.random.default_rng(2)
    color = (rng.random((n, 3)) * g["occ_map"]).astype(np.float32)
    go = rng.standard_normal((n, 3)).astype(np.float32)
    k = H.kernels()
    out = torch.zeros(n, 3)
    k.eaw_fwd(2.0, 0.1, 0.001, W, Hh, step, H.t(g["occ_map"]), H.t(color), H.t(g["normal_map"]), H.t(g["pos_map"]), out)
    gC, gN, gP, scr = torch.zeros(n, 3), torch.zeros(n, 3), torch.zeros(n, 3), torch.zeros(n)
    k.eaw_bwd(2.0, 0.1, 0.001, W, Hh, step, H.t(g["occ_map"]), H.t(color), H.t(g["normal_map"]), H.t(g["pos_map"]), out,
              H.t(go), gC, gN, gP, scr)
    o64, rC, rN, rP = B.eaw_grads(2.0, 0.1, 0.001, W, Hh, step, g["occ_map"], color, g["normal_map"], g["pos_map"], go)
    assert np.abs(out.numpy() - o64).max() < 1e-5
    for a, b, nm in ((gC, rC, "color"), (gN, rN, "normal"), (gP, rP, "pos")):
        assert_grad_close(a.numpy(), b, "eaw " + nm, rtol=1e-4)


def test_autograd_functions_route_gradients(case):
    """End-to-end through the reference-shaped driver: run_restir_di_with_pt -> loss -> backward reaches env, normal,
    kd, roughness/metallic with finite values (the reference's only gradient outputs)."""
    from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth
    sc, _ = case
    W, Hh = sc["W"], sc["H"]
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    mods = R.load_m_for_restir(W, Hh, device="cpu")
    g = {k: H.t(v) for k, v in sc["gbuffer"].items()}
    env = H.t(sc["env"]).requires_grad_(True)
    normal = g["normal_map"].clone().requires_grad_(True)
    kd = g["diffuse_map"].clone().requires_grad_(True)
    rs = g["roughness_specular"].clone().requires_grad_(True)
    outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(0.3), None, w, *mods, env, g["occ_map"], normal,
                                   g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"], None, None, None, None, W, Hh,
                                   2, 2, 2, 2.0, 0.1, 0.001, random_offset=99)
    final = outs[0]
    assert final.shape == (W * Hh, 3) and torch.isfinite(final).all()
    final.sum().backward()
    for t in (env, normal, kd, rs):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
