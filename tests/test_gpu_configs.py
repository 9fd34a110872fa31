"""GPU parity at the configurations round 1 had only covered on the host flavour: BASELINE configs[4] sizes, the
reference's real frame size (ssaa = 2 -> 1600 x 1600, /root/reference main.py:140, nerf/utils.py:770-777), MAX_Bounce
1 and 3 (the reference hard-codes 2, nerf/ScreenSpaceReSTIR/FinalShading.slang:7), long spp loops against the ORACLE
(not against the repo's own sequential schedule) and process_normal_ao (EAWDenoise.slang:591-647).

All through the C ABI (libmirres_b200.so) on cuda:0; bar as in test_gpu.py: integer outputs exact, floats to 1e-4 --
and, because oracle and kernels share the numerical contract of include/mirres_fpmath.h, bit equality.
"""
import numpy as np
import pytest
import torch

import parity as P
from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
FWD_RTOL = 1e-4


def tt(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def make_worker(sc):
    w = R.restirbvhWorker(tt(sc["vert"]), tt(sc["tri"]))
    w.LBVHNode_info, w.LBVHNode_aabb = w.update_bvh(want_sorted_codes=True)
    return w


@pytest.fixture(scope="module")
def kernels():
    from mirres_restir_nerf_mesh_b200.slangpy_shim import get_kernels, set_kernels
    set_kernels(None)
    return get_kernels()


def _assert_lbvh_and_gbuffer(kernels, sc, w):
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    assert (w.sorted_codes.cpu().numpy() == sc["bvh"].sorted_codes).all()
    assert (w.LBVHNode_info.cpu().numpy() == sc["bvh"].info).all()
    assert (w.LBVHNode_aabb.cpu().numpy() == sc["bvh"].aabb).all()
    n = sc["W"] * sc["H"]
    occ, depth = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pos, nrm = torch.zeros(n, 3, device=DEV), torch.zeros(n, 3, device=DEV)
    prim, bary = torch.zeros(n, dtype=torch.int32, device=DEV), torch.zeros(n, 2, device=DEV)
    kernels.gbuffer_primary(w.packed, tt(sc["rays_o"]), tt(sc["rays_d"]), occ, pos, nrm, depth, prim, bary,
                            ws=slangpy_shim.workspace(torch.device(DEV), n))
    m = sc["hit"] > 0
    assert (occ.cpu().numpy() == sc["hit"]).all() and (prim.cpu().numpy() == sc["prim"]).all()
    assert (pos.cpu().numpy()[m] == sc["pos"][m]).all() and (nrm.cpu().numpy()[m] == sc["nrm"][m]).all()


def test_c5_sizes_one_iteration_against_oracle(kernels, oracle):
    """BASELINE configs[4] sizes: 2 000 000 triangles, 2048 x 2048, 2k x 1k envmap, 4 path vertices (MAX_Bounce 3); ONE
    spp iteration (the oracle needs ~20 s for it): LBVH, primary G-buffer and every intermediate tensor."""
    sc = P.scene("C5")
    ref = P.oracle_run(sc, spp=1)
    w = make_worker(sc)
    _assert_lbvh_and_gbuffer(kernels, sc, w)
    got = P.product_run(sc, w, DEV, ref["prepared"], spp=1)
    assert got["mb"] == 3
    bad = P.compare(ref, got, rtol=FWD_RTOL)
    assert not bad, bad[:5]
    assert not P.compare(ref, got, rtol=0.0), "forward pass is expected to be bit-exact"


def test_ssaa2_frame_1600_against_oracle(kernels, oracle):
    """The frame stage 1 really renders for 800 x 800 data (ssaa = 2): 1600 x 1600 on the C2 mesh, two spp iterations (the
    second one runs the temporal pass): LBVH / G-buffer / every intermediate tensor against the oracle, then the default
    training call (concurrent schedule, denoiser, backward) on that frame: finite outputs and gradients."""
    sc = P.scene("C2S")
    ref = P.oracle_run(sc)
    w = make_worker(sc)
    _assert_lbvh_and_gbuffer(kernels, sc, w)
    got = P.product_run(sc, w, DEV, ref["prepared"])
    bad = P.compare(ref, got, rtol=FWD_RTOL)
    assert not bad, bad[:5]
    assert not P.compare(ref, got, rtol=0.0)
    W, Hh = sc["W"], sc["H"]
    mods = R.load_m_for_restir(W, Hh, device=DEV)
    g = {k: tt(v) for k, v in sc["gbuffer"].items()}
    env = tt(sc["env"]).requires_grad_(True)
    normal = g["normal_map"].clone().requires_grad_(True)
    kd = g["diffuse_map"].clone().requires_grad_(True)
    rs = g["roughness_specular"].clone().requires_grad_(True)
    outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(0.0), None, w, *mods, env, g["occ_map"], normal,
                                   g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"], None, None, None, None, W, Hh,
                                   2, 2, 2, 2.0, 0.1, 0.001, random_offset=4242)
    outs[0].mean().backward()
    torch.cuda.synchronize()
    assert all(o.shape == (W * Hh, 3) and torch.isfinite(o).all() for o in outs)
    for t in (env, normal, kd, rs):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
    # the un-denoised part of that call against the oracle's totals
    with torch.no_grad():
        tot = R.restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(0.0), w, 2, W, Hh, *mods[:7], *mods[8:],
                                  tt(sc["env"]), tt(ref["prepared"]["occ_map"]), g["pos_map"], g["normal_map"],
                                  g["depth_map"], g["diffuse_map"], g["roughness_specular"],
                                  tt(ref["prepared"]["ray_dir_map"]), None, None, None, None, None, None,
                                  random_offset=4242, max_bounce=2)
    m = np.float32(2)
    for j, k in ((0, "color"), (1, "color_1"), (2, "diff"), (3, "spec"), (4, "diff_1"), (5, "spec_1")):
        assert np.array_equal(tot[j].cpu().numpy() / m, ref[k]), k


@pytest.mark.parametrize("name,metallic", [("T0", 0.0), ("T2", 0.4)])
@pytest.mark.parametrize("mb", [1, 3])
def test_max_bounce_one_and_three(kernels, name, metallic, mb):
    """MAX_Bounce is a compile-time 2 in the reference (FinalShading.slang:7); the run-time generalisation at 1 and 3
    indirect vertices against the oracle's, on CUDA (round 1: host flavour only)."""
    sc = P.scene(name, metallic)
    ref = P.oracle_run(sc, max_bounce=mb)
    got = P.product_run(sc, make_worker(sc), DEV, ref["prepared"], max_bounce=mb)
    assert got["mb"] == mb
    assert P.compare(ref, got, rtol=FWD_RTOL) == []
    assert P.compare(ref, got, rtol=0.0) == []
    # default (concurrent) schedule of the same call: totals equal the oracle's
    _concurrent_totals_equal_oracle(sc, ref, spp=sc["cfg"]["spp"], mb=mb)


def _concurrent_totals_equal_oracle(sc, ref, spp, mb, random_offset=4242):
    W, Hh = sc["W"], sc["H"]
    mods = R.load_m_for_restir(W, Hh, device=DEV, max_bounce=mb)
    g = {k: tt(v) for k, v in sc["gbuffer"].items()}
    w = make_worker(sc)
    with torch.no_grad():
        tot = R.restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), w, spp, W, Hh, *mods[:7],
                                  *mods[8:], tt(sc["env"]), tt(ref["prepared"]["occ_map"]), g["pos_map"], g["normal_map"],
                                  g["depth_map"], g["diffuse_map"], g["roughness_specular"],
                                  tt(ref["prepared"]["ray_dir_map"]), None, None, None, None, None, None,
                                  random_offset=random_offset, max_bounce=mb)
    torch.cuda.synchronize()
    m = np.float32(spp)
    for j, k in ((0, "color"), (1, "color_1"), (2, "diff"), (3, "spec"), (4, "diff_1"), (5, "spec_1")):
        assert np.array_equal(tot[j].cpu().numpy() / m, ref[k]), k
    assert tot[7] == spp


def test_long_loop_against_oracle(kernels):
    """spp = 24 (> one flush of the concurrent schedule's running sums, > the 16-pass chunk of the one-node backward, > 20 =
    MAX_HISTORY_LENGTH of the temporal pass): every intermediate tensor of every iteration against the ORACLE on the
    sequential schedule, then the totals of the default concurrent schedule against the oracle's."""
    sc = P.scene("T1", 0.2)
    spp = 24
    ref = P.oracle_run(sc, spp=spp, random_offset=1717)
    got = P.product_run(sc, make_worker(sc), DEV, ref["prepared"], spp=spp, random_offset=1717)
    assert len(got["snapshots"]) == spp
    assert P.compare(ref, got, rtol=FWD_RTOL) == []
    assert P.compare(ref, got, rtol=0.0) == []
    assert max(int(s["res"][2].max()) for s in ref["snapshots"]) > 20  # history clamp of the temporal pass is exercised
    _concurrent_totals_equal_oracle(sc, ref, spp=spp, mb=sc["cfg"]["max_bounce"], random_offset=1717)


def test_normal_ao_gpu(kernels, oracle):
    """process_normal_ao (EAWDenoise.slang:591-647, called from nerf/renderer.py:1153) through the C ABI, bit-exact against
    the oracle; ragged frame and an all-background frame."""
    sc = P.scene("T2")
    g = sc["gbuffer"]
    for W, Hh in ((sc["W"], sc["H"]), (sc["W"] - 5, sc["H"] - 3), (1, 1)):
        n = W * Hh
        occ, nrm = g["occ_map"][:n].copy(), g["normal_map"][:n].copy()
        ao = torch.full((n, 3), 7.0, device=DEV)
        kernels.normal_ao(W, Hh, tt(occ), tt(nrm), ao)
        assert np.array_equal(ao.cpu().numpy(), oracle.normal_ao(W, Hh, occ, nrm)), (W, Hh)
    n = sc["W"] * sc["H"]
    ao = torch.full((n, 3), 7.0, device=DEV)
    kernels.normal_ao(sc["W"], sc["H"], torch.zeros(n, 1, device=DEV), tt(g["normal_map"]), ao)
    assert np.array_equal(ao.cpu().numpy(), oracle.normal_ao(sc["W"], sc["H"], np.zeros((n, 1), np.float32), g["normal_map"]))


@pytest.mark.parametrize("name,world,balanced,spp", [("T2", 2, False, 3), ("T2", 3, True, 5), ("C1", 4, True, 3)])
def test_row_bands_on_one_gpu_are_bit_identical(kernels, name, world, balanced, spp):
    """Row-band rendering with the ranks as host threads on cuda:0 (tests/band_threads.py): every virtual rank
    runs the spp loop on the slice of the maps that holds its band and halo, on its own stream, with the concurrent schedule;
    halo rows and the final gather are tensor copies behind a barrier.  All six outputs must equal the single render bit for
    bit -- row offset of the random streams, jittered row of temporal reuse, band words of the spatial pass, alive lists
    and split tracers on band-sized queues, and two host threads driving the library at once."""
    import band_threads as BT
    from mirres_restir_nerf_mesh_b200 import dist as D
    sc = P.scene(name, 0.3)
    w = make_worker(sc)
    want = BT.render(sc, w, DEV, spp=spp)
    torch.cuda.synchronize()
    occ = tt(sc["gbuffer"]["occ_map"])
    bounds = D.balanced_bounds(occ, sc["W"], sc["H"], world) if balanced else D.uniform_bounds(sc["H"], world)
    outs, errors = BT.render_in_threads(sc, w, DEV, world, bounds, spp=spp)
    torch.cuda.synchronize()
    assert not errors, errors
    for r in range(world):
        for a, b in zip(outs[r], want):
            assert torch.equal(a, b), (r, (a != b).sum().item())


@pytest.mark.parametrize("shift", [None, (3.0, -2.0)])
def test_visibility_tags_cast_only_unknown_rays(kernels, oracle, shift):
    """Visibility tags on the GPU (tests/test_vis_tags_cpu.py has the CPU side): every tensor of the loop equals the
    oracle's -- which traces every final-visibility ray -- while the pass queues exactly the rays the oracle's provenance
    count calls unknown; with tags off it queues one ray per valid sample."""
    sc = P.scene("C1", 0.4)
    motion = None
    if shift is not None:
        motion = np.empty((sc["W"] * sc["H"], 2), np.float32)
        motion[:, 0], motion[:, 1] = shift[0] / sc["W"], shift[1] / sc["H"]
    ref = P.oracle_run(sc, spp=4, motion=motion)
    assert ref["provenance_violations"] == 0
    w = make_worker(sc)
    valid = [int((s["res"][0][:, 0] > 0.1).sum()) for s in ref["snapshots"]]
    unknown = [v - k for v, k in zip(valid, ref["known_final_rays"])]
    old = R.USE_VIS_TAGS
    try:
        for tags, want in ((1, unknown), (0, valid)):
            R.USE_VIS_TAGS = tags
            got = P.product_run(sc, w, DEV, ref["prepared"], spp=4, motion=motion)
            assert P.compare(ref, got, rtol=FWD_RTOL) == []
            assert P.compare(ref, got) == []
            assert [s["final_rays"] for s in got["snapshots"]] == want
    finally:
        R.USE_VIS_TAGS = old
    if shift is None:
        assert sum(unknown) <= 1e-3 * sum(valid)      # history comes from the same pixel: (almost) nothing left to cast
    else:
        assert 0 < sum(unknown) < sum(valid)          # history from another pixel has never been tested from here
