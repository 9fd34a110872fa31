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


# ---- auto_normals (meshutils.py:14-39) ---------------------------------------------------------------------------------
AN = np.load(os.path.join(HERE, "golden", "auto_normals_ref.npz"))


def run_auto_normals(k, dev):
    vert, tri = torch.from_numpy(AN["vert"]).to(dev), torch.from_numpy(AN["tri"]).to(dev)
    vsum, vnrm = torch.full_like(vert, 9.0), torch.zeros_like(vert)  # the forward zero-fills vsum itself
    k.vertex_normals_fwd(vert, tri, vsum, vnrm)
    assert close(vnrm.cpu().numpy(), AN["vnrm"], FWD_RTOL)
    assert (vnrm[-1].cpu().numpy() == np.array([0, 0, 1], np.float32)).all()  # unreferenced vertex: the fallback
    gv = torch.zeros_like(vert)
    k.vertex_normals_bwd(vert, tri, vsum, torch.from_numpy(AN["grad_vnrm"]).to(dev), gv)
    assert close(gv.cpu().numpy(), AN["grad_vert"], GRAD_RTOL), np.abs(gv.cpu().numpy() - AN["grad_vert"]).max()
    assert bool((gv[-1] == 0).all())
    # accumulation contract and out-of-range indices
    k.vertex_normals_bwd(vert, tri, vsum, torch.from_numpy(AN["grad_vnrm"]).to(dev), gv)
    assert close(gv.cpu().numpy(), 2 * AN["grad_vert"], GRAD_RTOL)
    bad = torch.cat([tri, torch.tensor([[0, 1, 10 ** 6], [-1, 2, 3]], dtype=torch.int32, device=dev)])
    vn2 = torch.zeros_like(vert)
    k.vertex_normals_fwd(vert, bad, vsum, vn2)
    assert close(vn2.cpu().numpy(), AN["vnrm"], FWD_RTOL)


def run_auto_normals_operator(dev):
    from mirres_restir_nerf_mesh_b200 import meshutils as MU
    v = torch.from_numpy(AN["vert"]).to(dev).requires_grad_(True)
    tri = torch.from_numpy(AN["tri"]).to(dev)
    vn, t2 = MU.auto_normals(v, tri.long())  # the reference passes whatever integer dtype the mesh has
    assert t2.dtype == torch.int64 and close(vn.detach().cpu().numpy(), AN["vnrm"], FWD_RTOL)
    vn.backward(torch.from_numpy(AN["grad_vnrm"]).to(dev))
    assert close(v.grad.cpu().numpy(), AN["grad_vert"], GRAD_RTOL)


def test_auto_normals_host_flavour_against_reference_function():
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    run_auto_normals(H.kernels(), "cpu")
    H.activate()
    try:
        run_auto_normals_operator("cpu")
    finally:
        slangpy_shim.set_kernels(None)


@pytest.mark.gpu
def test_auto_normals_gpu_against_reference_function():
    from mirres_restir_nerf_mesh_b200.slangpy_shim import get_kernels, set_kernels
    set_kernels(None)
    run_auto_normals(get_kernels(), "cuda")
    run_auto_normals_operator("cuda")


@pytest.mark.gpu
def test_vertex_gradient_chain_gpu():
    """grad_normal of the path back to the mesh vertices: shading_normal_bwd -> interpolate_bwd -> vertex_normals_bwd
    against torch.autograd over the same chain written in torch (C2-sized mesh, one pixel per triangle)."""
    from mirres_restir_nerf_mesh_b200 import synth, meshutils as MU, renderutils_ops as OPS
    from mirres_restir_nerf_mesh_b200.slangpy_shim import get_kernels, set_kernels
    set_kernels(None)
    k = get_kernels()
    vert_np, tri_np = synth.torus_knot(200, 50) if hasattr(synth, "torus_knot") else (AN["vert"], AN["tri"])
    vert = torch.from_numpy(np.ascontiguousarray(vert_np, dtype=np.float32)).cuda().requires_grad_(True)
    tri = torch.from_numpy(np.ascontiguousarray(tri_np, dtype=np.int32)).cuda()
    F = tri.shape[0]
    rng = np.random.default_rng(5)
    prim = torch.from_numpy(rng.integers(0, F, F).astype(np.int32)).cuda()
    bary = torch.from_numpy(rng.dirichlet([1, 1, 1], F)[:, 1:].astype(np.float32)).cuda()
    eye = torch.tensor([[0.3, 2.9, 1.1]], device="cuda")
    go = torch.from_numpy(rng.standard_normal((F, 3)).astype(np.float32)).cuda()

    def chain(auto_normals, shading_normal, interpolate):
        v = vert.detach().clone().requires_grad_(True)
        vn, _ = auto_normals(v, tri)
        w = torch.cat([1 - bary.sum(1, keepdim=True), bary], 1)
        corners = tri.long()[prim.long()]
        smooth = interpolate(vn, corners, w)
        pos = (v.detach()[corners] * w[..., None]).sum(1)
        e1, e2 = v.detach()[corners[:, 1]] - v.detach()[corners[:, 0]], v.detach()[corners[:, 2]] - v.detach()[corners[:, 0]]
        geom = torch.nn.functional.normalize(torch.linalg.cross(e1, e2), dim=-1)
        out = shading_normal(pos.view(1, 1, F, 3), eye.view(1, 1, 1, 3), None, smooth.view(1, 1, F, 3),
                             torch.zeros(1, 1, F, 3, device="cuda"), geom.view(1, 1, F, 3))
        out.backward(go.view(1, 1, F, 3))
        return out.detach(), v.grad

    class _Interp(torch.autograd.Function):  # product: barycentric gather forward, mirres_interpolate_bwd backward
        @staticmethod
        def forward(ctx, vn, corners, w):
            ctx.V = vn.shape[0]
            return (vn[corners] * w[..., None]).sum(1)

        @staticmethod
        def backward(ctx, g):
            out = torch.zeros(ctx.V, 3, device=g.device)
            k.interpolate_bwd(g.contiguous(), prim, bary, tri, out)
            return out, None, None

    def ref_auto_normals(v, t):  # the torch expression of meshutils.py:14-39
        i = t.long()
        fn = torch.linalg.cross(v[i[:, 1]] - v[i[:, 0]], v[i[:, 2]] - v[i[:, 0]])
        s = torch.zeros_like(v).index_add_(0, i[:, 0], fn).index_add_(0, i[:, 1], fn).index_add_(0, i[:, 2], fn)
        d = (s * s).sum(-1, keepdim=True)
        s = torch.where(d > 1e-20, s, torch.tensor([0.0, 0.0, 1.0], device=v.device))
        return s / torch.sqrt(torch.clamp((s * s).sum(-1, keepdim=True), min=1e-20)), t

    def ref_shading_normal(pos, view_pos, pert, smooth, tng, geom):  # ops.py:82-112 with pert = (0,0,1), tng = 0
        nrm = torch.nn.functional.normalize
        s = nrm(nrm(smooth, dim=-1), dim=-1)
        vv = nrm(view_pos - pos, dim=-1)
        front = (geom * vv).sum(-1, keepdim=True) > 0
        s, g = torch.where(front, s, -s), torch.where(front, geom, -geom)
        t = torch.clamp((vv * s).sum(-1, keepdim=True) / 0.1, min=0, max=1)
        return torch.lerp(g, s, t)

    out_p, gv_p = chain(MU.auto_normals, OPS.prepare_shading_normal, _Interp.apply)
    out_r, gv_r = chain(ref_auto_normals, ref_shading_normal, lambda vn, c, w: (vn[c] * w[..., None]).sum(1))
    assert close(out_p.cpu().numpy(), out_r.cpu().numpy(), FWD_RTOL)
    assert close(gv_p.cpu().numpy(), gv_r.cpu().numpy(), GRAD_RTOL), float((gv_p - gv_r).abs().max())


def test_auto_normals_c2_mesh_against_torch_expression():
    """config C2's mesh (250 000 vertices, 500 000 triangles): vertex normals and their gradient from the product kernels
    (host flavour) against the torch expression of meshutils.py:14-39 differentiated by autograd"""
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import synth
    v_np, t_np = synth.make_mesh(synth.CONFIGS["C2"])
    vert, tri = torch.from_numpy(v_np), torch.from_numpy(t_np)
    k = H.kernels()
    vsum, vnrm = torch.empty_like(vert), torch.empty_like(vert)
    k.vertex_normals_fwd(vert, tri, vsum, vnrm)
    v = vert.clone().requires_grad_(True)
    i = tri.long()
    fn = torch.linalg.cross(v[i[:, 1]] - v[i[:, 0]], v[i[:, 2]] - v[i[:, 0]])
    s = torch.zeros_like(v).index_add_(0, i[:, 0], fn).index_add_(0, i[:, 1], fn).index_add_(0, i[:, 2], fn)
    want = s / torch.sqrt(torch.clamp((s * s).sum(-1, keepdim=True), min=1e-20))
    assert close(vnrm.numpy(), want.detach().numpy(), FWD_RTOL)
    g = torch.from_numpy(np.random.default_rng(2).standard_normal(v_np.shape).astype(np.float32))
    want.backward(g)
    gv = torch.zeros_like(vert)
    k.vertex_normals_bwd(vert, tri, vsum, g, gv)
    # the gradient of a unit normal scales with 1 / |sum of face normals| (tiny faces: ~1e5 here), compare relative to its scale
    assert close(gv.numpy(), v.grad.numpy(), GRAD_RTOL), float((gv - v.grad).abs().max() / v.grad.abs().max())
