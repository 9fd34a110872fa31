"""prepare_shading_normal (nerf/renderutils/ops.py:129-163 -> c_src/normal.cu:95-178).

CPU: the product's kernel bodies (host-check flavour) and the Python mirror of the reference operator against golden
vectors produced by the REFERENCE's own pure-torch restatement of its kernel (ops.py:82-112, imported from the
reference tree by tests/golden/make_golden_shading_normal.py) and its torch.autograd gradients.
GPU: the shipped library against the same vectors and against the reference's own CUDA kernels
(oracle/_ref/libref_renderutils.so, compiled from normal.cu where it lies).
Tolerances are the north-star ones: 1e-4 forward, 1e-3 gradients."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "shading_normal_ref.npz"))
NAMES = ("pos", "view_pos", "perturbed_nrm", "smooth_nrm", "smooth_tng", "geom_nrm")
FWD_RTOL, GRAD_RTOL = 1e-4, 1e-3
FX, FY = int(GOLD["fx"]), int(GOLD["fy"])
N = FX * FY


def close(got, want, rtol):
    # absolute floor 1e-6: with no normal map d(perturbed_nrm).z = dot(dS, N) is identically zero in exact arithmetic
    # (dS is orthogonal to S = N) and both sides return fp32 rounding noise of O(1e-7) for O(1) upstream gradients
    return np.allclose(got, want, rtol=rtol, atol=max(rtol * 1e-2 * float(np.abs(want).max()), 1e-6))


def run_all_flags(k, dev):
    """the four flag settings through the C ABI: outputs and the six gradients against the reference vectors"""
    ins = [torch.from_numpy(GOLD[n]).to(dev) for n in NAMES]
    go = torch.from_numpy(GOLD["grad_out"]).to(dev)
    for two_sided in (0, 1):
        for opengl in (0, 1):
            tag = "ts%d_gl%d" % (two_sided, opengl)
            out = torch.zeros(N, 3, device=dev)
            k.shading_normal_fwd(N, ins, two_sided, opengl, out)
            assert close(out.cpu().numpy(), GOLD["out_" + tag], FWD_RTOL), tag
            grads = [torch.zeros(N, 3, device=dev) for _ in range(6)]
            k.shading_normal_bwd(N, ins, two_sided, opengl, go, grads)
            for n, g in zip(NAMES, grads):
                want = GOLD["g_%s_%s" % (n, tag)]
                got = g.cpu().numpy()
                if want.shape[0] == 1:  # broadcast input: the kernel hands back per-pixel gradients (as the reference's does)
                    got = got.sum(0, keepdims=True)
                assert close(got, want, GRAD_RTOL), (n, tag, np.abs(got - want).max())


def run_operator(dev):
    """the Python mirror with the arguments of nerf/renderer.py:1013 (no normal map, zero tangents)"""
    from mirres_restir_nerf_mesh_b200 import renderutils_ops as OPS
    v = lambda n, g: torch.from_numpy(GOLD[n]).to(dev).view(1, FY, FX, 3).requires_grad_(g)
    pos, smooth, geom = v("pos", True), v("smooth_nrm", True), v("geom_nrm", True)
    view = torch.from_numpy(GOLD["view_pos"]).to(dev).view(1, 1, 1, 3)
    out = OPS.prepare_shading_normal(pos, view, None, smooth, torch.zeros_like(smooth), geom, two_sided_shading=True, opengl=True)
    assert out.shape == (1, FY, FX, 3)
    assert close(out.detach().cpu().numpy().reshape(-1, 3), GOLD["out_callsite"], FWD_RTOL)
    out.backward(torch.from_numpy(GOLD["grad_out"]).to(dev).view(1, FY, FX, 3))
    for n, t in (("pos", pos), ("smooth_nrm", smooth), ("geom_nrm", geom)):
        assert close(t.grad.cpu().numpy().reshape(-1, 3), GOLD["g_%s_callsite" % n], GRAD_RTOL), n
    # a broadcast operand that needs a gradient gets it reduced to its own shape
    view.requires_grad_(True)
    OPS.prepare_shading_normal(pos, view, None, smooth, torch.zeros_like(smooth), geom).sum().backward()
    assert view.grad.shape == (1, 1, 1, 3)


def test_host_flavour_against_reference_torch_restatement():
    import hostcheck as H
    run_all_flags(H.kernels(), "cpu")


def test_operator_mirror_against_reference_torch_restatement():
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    H.activate()
    try:
        run_operator("cpu")
    finally:
        slangpy_shim.set_kernels(None)


def test_strided_and_degenerate_rows():
    """row strides (a [n,6] tensor viewed as two [n,3] operands, SURVEY.md 8b stride caveat), zero vectors, bad arguments"""
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import kernels as K
    k = H.kernels()
    wide = torch.from_numpy(np.concatenate([GOLD["smooth_nrm"], GOLD["geom_nrm"]], 1).copy())
    ins = [torch.from_numpy(GOLD[n]) for n in NAMES]
    ins[3], ins[5] = wide[:, 0:3], wide[:, 3:6]
    out = torch.zeros(N, 3)
    k.shading_normal_fwd(N, ins, 1, 1, out)
    assert close(out.numpy(), GOLD["out_ts1_gl1"], FWD_RTOL)
    # all-zero operands: nrm(0) = 0 everywhere, finite outputs and gradients
    z = [torch.zeros(4, 3) for _ in range(6)]
    out = torch.full((4, 3), 7.0)
    k.shading_normal_fwd(4, z, 1, 1, out)
    assert bool((out == 0).all())
    grads = [torch.full((4, 3), 7.0) for _ in range(6)]
    k.shading_normal_bwd(4, z, 1, 1, torch.ones(4, 3), grads)
    assert all(bool(torch.isfinite(g).all()) for g in grads)
    with pytest.raises(K.AbiError):
        k.shading_normal_fwd(N, ins[:5] + [torch.zeros(N - 1, 3)], 1, 1, torch.zeros(N, 3))
    assert k.lib.mirres_shading_normal_fwd(4, *([None, 3] * 6), 1, 1, None, None) == -1
    assert k.lib.mirres_shading_normal_fwd(0, *([None, 3] * 6), 1, 1, None, None) == 0  # empty frame: nothing to do


@pytest.mark.gpu
def test_gpu_against_reference_vectors_and_reference_kernel():
    from mirres_restir_nerf_mesh_b200.slangpy_shim import get_kernels, set_kernels
    set_kernels(None)
    k = get_kernels()
    run_all_flags(k, "cuda")
    run_operator("cuda")
    from oracle import ref as REF
    if not REF.available():
        pytest.skip("oracle/_ref/libref_renderutils.so not built (needs the reference tree at build time)")
    ins = [torch.from_numpy(GOLD[n]).cuda() for n in NAMES]
    go = torch.from_numpy(GOLD["grad_out"]).cuda()
    variants = [ins, ins[:2] + [torch.tensor([[0.0, 0.0, 1.0]]).cuda(), ins[3], torch.zeros_like(ins[4]), ins[5]]]
    for operands in variants:
        for two_sided in (0, 1):
            for opengl in (0, 1):
                want = REF.prepare_shading_normal_fwd(FX, FY, operands, two_sided, opengl).cpu().numpy()
                out = torch.zeros(N, 3, device="cuda")
                k.shading_normal_fwd(N, operands, two_sided, opengl, out)
                assert close(out.cpu().numpy(), want, FWD_RTOL), (two_sided, opengl)
                wantg = REF.prepare_shading_normal_bwd(FX, FY, operands, two_sided, opengl, go)
                grads = [torch.zeros(N, 3, device="cuda") for _ in range(6)]
                k.shading_normal_bwd(N, operands, two_sided, opengl, go, grads)
                for n, g, w in zip(NAMES, grads, wantg):
                    assert close(g.cpu().numpy(), w.cpu().numpy(), GRAD_RTOL), (n, two_sided, opengl)
