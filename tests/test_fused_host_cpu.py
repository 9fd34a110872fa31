"""CPU checks (host-check flavour of the kernels, tests/hostcheck.py) of the single-launch replacements of the driver's
torch chains (csrc/screen.cu, mirres_final_shading_bwd_multi) and of the concurrent schedule's host logic driven over
CPU tensors: every fused kernel against the torch expressions it replaces, the one-node backward of the spp loop
against the per-pass autograd Functions."""
import numpy as np
import pytest
import torch

import fused_checks as C
import hostcheck as H
import parity as P
from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth


@pytest.fixture(scope="module", autouse=True)
def _bind_hostcheck():
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    H.activate()
    yield
    slangpy_shim.set_kernels(None)


def test_material_kernel_equals_torch_and_numpy_expressions():
    C.material_kernel_equals_torch_and_numpy_expressions(H.kernels(), "cpu")


def test_sum_images_is_the_sequential_torch_sum():
    C.sum_images_is_the_sequential_torch_sum(H.kernels(), "cpu")


def test_composite_forward_exact_backward_matches_autograd():
    C.composite_forward_exact_backward_matches_autograd(H.kernels(), "cpu")


def test_final_shading_bwd_multi_equals_the_single_pass_kernels():
    C.final_shading_bwd_multi_equals_the_single_pass_kernels(H.kernels(), "cpu")


def _render(sc, w, prepared_lighting=False, spp=3, early_blocks=False, **kw):
    mods = R.load_m_for_restir(sc["W"], sc["H"], device="cpu")
    g = {k: H.t(v) for k, v in sc["gbuffer"].items()}
    env = H.t(sc["env"]).requires_grad_(True)
    if prepared_lighting:
        kw["lighting"] = R.prepare_lighting(mods[0], mods[1], mods[8], mods[9], mods[10], env, spp, 99,
                                            frame_pixels=sc["W"] * sc["H"] if early_blocks else None)
    normal = g["normal_map"].clone().requires_grad_(True)
    tex = torch.cat((g["diffuse_map"], torch.zeros_like(g["diffuse_map"])), dim=1).requires_grad_(True)
    kd = tex[:, 0:3]  # strided view, as render_stage1 passes it
    rs = g["roughness_specular"].clone().requires_grad_(True)
    outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, w, *mods, env,
                                   g["occ_map"].clone(), normal, g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"], None,
                                   None, None, None, sc["W"], sc["H"], spp, 2, 2, 2.0, 0.1, 0.001, random_offset=99, **kw)
    wgt = torch.linspace(0.5, 1.5, outs[0].numel()).reshape(outs[0].shape)
    (outs[0] * wgt).sum().backward()
    return [o.detach() for o in outs], [env.grad, normal.grad, tex.grad, rs.grad]


@pytest.mark.parametrize("strict", [True, False])
def test_concurrent_schedule_host_logic_equals_sequential_schedule(strict):
    """The concurrent schedule (one autograd node for all shading passes, fused sums / composite / material merge) driven
    over CPU tensors against the reference's sequential order: images bit-identical, gradients to 1e-5 of their scale (the
    accumulation order over passes is the engine's, the association with the denoiser's contributions differs)."""
    sc = P.scene("T0", 0.3)
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    old = R.STRICT_REFERENCE_ALIASING
    R.STRICT_REFERENCE_ALIASING = strict
    try:
        seq, gseq = _render(sc, w, overlap=False, batched_denoise=False, fused_prepare=False, fused_composite=False)
        par, gpar = _render(sc, w, overlap=True)
    finally:
        R.STRICT_REFERENCE_ALIASING = old
    for a, b in zip(seq, par):
        assert torch.equal(a, b)
    for a, b in zip(gseq, gpar):
        assert a.abs().sum() > 0
        assert (a - b).abs().max() <= 1e-5 * a.abs().max()


def test_prepared_lighting_changes_nothing():
    """prepare_lighting() + run_restir_di_with_pt(lighting=...) == the loop computing the same things itself."""
    sc = P.scene("T0", 0.0)
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    a, ga = _render(sc, w, overlap=True)
    b, gb = _render(sc, w, prepared_lighting=True, overlap=True)
    c, gc = _render(sc, w, prepared_lighting=True, early_blocks=True, overlap=True)  # + zero-filled blocks, flipped map
    for x, y, z in zip(a + ga, b + gb, c + gc):
        assert torch.equal(x, y) and torch.equal(x, z)


def test_long_loop_flushes_and_chunked_backward():
    """spp = 19: the concurrent schedule flushes its running sums every 16 images and the one-node backward runs in two
    chunks of passes (16 + 3); results must still equal the sequential schedule."""
    sc = P.scene("T0", 0.0)
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    seq, gseq = _render(sc, w, spp=19, overlap=False, batched_denoise=False, fused_prepare=False, fused_composite=False)
    par, gpar = _render(sc, w, spp=19, overlap=True)
    for a, b in zip(seq, par):
        assert torch.equal(a, b)
    for a, b in zip(gseq, gpar):
        assert a.abs().sum() > 0
        assert (a - b).abs().max() <= 1e-5 * a.abs().max()
