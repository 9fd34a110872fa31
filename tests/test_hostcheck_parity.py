"""CPU parity: the product's kernel source (host-check flavour) + the product's Python driver against the
independently written oracle, bit for bit, on small synthetic scenes.  See tests/hostcheck.py."""
import numpy as np
import pytest
import torch

import hostcheck as H
import parity as P


@pytest.fixture(scope="module", autouse=True)
def _bind_hostcheck():
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    H.activate()
    yield
    slangpy_shim.set_kernels(None)


def _worker(sc):
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    return w


# C1 is BASELINE.json configs[0], the "CPU plumbing workload": icosphere level 5 (20 480 triangles), 256 x 256, 1 spp
@pytest.mark.parametrize("name,metallic", [("T0", 0.0), ("T1", 0.0), ("T2", 0.4), ("C1", 0.0)])
def test_pipeline_bit_exact(name, metallic):
    sc = P.scene(name, metallic)
    ref = P.oracle_run(sc)
    got = P.product_run(sc, _worker(sc), "cpu", ref["prepared"])
    assert P.compare(ref, got) == []


@pytest.mark.parametrize("name,metallic,view", [("T2", 0.4, 13), ("T2", 0.0, 37), ("T1", 0.4, 71), ("C1", 0.4, 50)])
def test_pipeline_bit_exact_other_views(name, metallic, view):
    """other cameras of the 100-view ring and the metallic material variant (SURVEY.md 8d synthetic inputs)"""
    sc = P.scene(name, metallic, view)
    ref = P.oracle_run(sc, random_offset=977 + view)
    got = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], random_offset=977 + view)
    assert P.compare(ref, got) == []


def test_c2_full_size_bit_exact():
    """BASELINE.json configs[1] at FULL size (500 000 triangles, 800 x 800, spp 4, 3 path vertices): every intermediate
    tensor of the forward spp loop, product kernel source (host flavour) against the oracle; ~10 s on 8 cores."""
    sc = P.scene("C2", 0.0)
    ref = P.oracle_run(sc)
    got = P.product_run(sc, _worker(sc), "cpu", ref["prepared"])
    assert P.compare(ref, got) == []


def test_c5_sizes_one_iteration_bit_exact():
    """BASELINE.json configs[4] sizes (2 000 000 triangles, 2048 x 2048, 2k x 1k envmap, 4 path vertices), ONE spp
    iteration: the largest shapes of the contract through the product kernel source against the oracle; ~20 s."""
    sc = P.scene("C5", 0.0)
    ref = P.oracle_run(sc, spp=1)
    got = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], spp=1)
    assert P.compare(ref, got) == []


def test_three_and_one_indirect_bounces():
    sc = P.scene("T0")
    for mb in (1, 3):
        ref = P.oracle_run(sc, max_bounce=mb)
        got = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], max_bounce=mb)
        assert P.compare(ref, got) == []


def test_rays_bit_exact_including_negative_t(oracle):
    sc = P.scene("T1")
    w = _worker(sc)
    k = H.kernels()
    rng = np.random.default_rng(3)
    hitm = sc["hit"] > 0
    d = rng.standard_normal((int(hitm.sum()), 3)).astype(np.float32)
    o = (sc["pos"][hitm] + 0.01 * d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    n = len(o)
    hit = torch.zeros(n, dtype=torch.int32)
    t, pos, nrm = torch.zeros(n), torch.zeros(n, 3), torch.zeros(n, 3)
    prim = torch.zeros(n, dtype=torch.int32)
    k.trace_closest(w.packed, H.t(o), H.t(d), hit, t, pos, nrm, prim)
    oh, ot, op, on, opr = oracle.trace(sc["bvh"], o, d)
    assert (ot[oh > 0] < 0).any()  # the quirk is exercised
    assert (hit.numpy() == oh).all() and (prim.numpy() == opr).all()
    assert (t.numpy() == ot).all() and (pos.numpy() == op).all() and (nrm.numpy()[oh > 0] == on[oh > 0]).all()
    anyh = torch.zeros(n, dtype=torch.int32)
    k.trace_any(w.packed, H.t(o), H.t(d), anyh)
    assert (anyh.numpy() == oh).all()


def test_env_and_offsets_bit_exact(oracle):
    from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R
    env = synth.envmap(32, 64)
    tex = np.ascontiguousarray(env[::-1].reshape(-1, 3))
    want = oracle.env_build_distribution(tex, 64, 32)
    got = R.make_sampleable(None, H.t(tex), 64, 32)
    for a, b in zip(got, want):
        assert (a.numpy() == b).all()
    out = torch.zeros(8192 * 2, 1)
    H.kernels().neighbor_offsets(8192, out)
    assert (out.numpy() == oracle.neighbor_offsets(8192)).all()


def test_eaw_and_ao_bit_exact(oracle):
    sc = P.scene("T1")
    g = sc["gbuffer"]
    rng = np.random.default_rng(0)
    color = (rng.random((sc["W"] * sc["H"], 3)) * g["occ_map"]).astype(np.float32)
    k = H.kernels()
    for step in (2, 1):
        out = torch.zeros(color.shape)
        k.eaw_fwd(2.0, 0.1, 0.001, sc["W"], sc["H"], step, H.t(g["occ_map"]), H.t(color), H.t(g["normal_map"]),
                  H.t(g["pos_map"]), out)
        want = oracle.eaw_fwd(2.0, 0.1, 0.001, sc["W"], sc["H"], step, g["occ_map"], color, g["normal_map"], g["pos_map"])
        assert (out.numpy() == want).all()
    ao = torch.zeros(color.shape)
    k.normal_ao(sc["W"], sc["H"], H.t(g["occ_map"]), H.t(g["normal_map"]), ao)
    assert (ao.numpy() == oracle.normal_ao(sc["W"], sc["H"], g["occ_map"], g["normal_map"])).all()


def _scatter_reference(grad, prim, bary, tri, V):
    out = np.zeros((V, grad.shape[1]), np.float64)
    for i in np.nonzero(prim >= 0)[0]:
        u, v = (bary[i] if bary is not None else (1 / 3, 1 / 3))
        w = (1 - u - v, u, v) if bary is not None else (1 / 3, 1 / 3, 1 / 3)
        for k in range(3):
            out[tri[prim[i], k]] += w[k] * grad[i].astype(np.float64)
    return out


def test_gbuffer_primary_and_interpolate_bwd(oracle):
    """primary-ray G-buffer (bit-exact against the oracle's closest-hit) and its reverse scatter."""
    sc = P.scene("T2")
    w = _worker(sc)
    k = H.kernels()
    n = len(sc["rays_o"])
    occ, pos, nrm, depth = torch.zeros(n), torch.zeros(n, 3), torch.zeros(n, 3), torch.zeros(n)
    prim, bary = torch.zeros(n, dtype=torch.int32), torch.zeros(n, 2)
    k.gbuffer_primary(w.packed, H.t(sc["rays_o"]), H.t(sc["rays_d"]), occ, pos, nrm, depth, prim, bary)
    oh, ot, op, on, opr = oracle.trace(sc["bvh"], sc["rays_o"], sc["rays_d"])
    m = oh > 0
    assert m.any() and (~m).any()
    assert (occ.numpy() == oh).all() and (prim.numpy() == opr).all()
    assert (pos.numpy()[m] == op[m]).all() and (nrm.numpy()[m] == on[m]).all()
    assert (pos.numpy()[~m] == 0).all() and (nrm.numpy()[~m] == 0).all() and (prim.numpy()[~m] == -1).all()
    np.testing.assert_allclose(depth.numpy()[m], np.linalg.norm(op[m] - sc["rays_o"][m], axis=1), rtol=1e-6)
    # the wavefront launch shape (persistent queue tracer) gives the same maps
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    o2, p2, n2, d2 = torch.zeros(n), torch.zeros(n, 3), torch.zeros(n, 3), torch.zeros(n)
    pr2, b2 = torch.zeros(n, dtype=torch.int32), torch.zeros(n, 2)
    k.gbuffer_primary(w.packed, H.t(sc["rays_o"]), H.t(sc["rays_d"]), o2, p2, n2, d2, pr2, b2, ws=slangpy_shim.workspace("cpu", n))
    for a_, b_ in ((occ, o2), (pos, p2), (nrm, n2), (depth, d2), (prim, pr2), (bary, b2)):
        assert (a_.numpy() == b_.numpy()).all()
    # barycentrics reproduce the hit point
    tri, vert = sc["tri"], sc["vert"]
    b = bary.numpy()[m]
    pv = vert[tri[opr[m]]]
    rec = (1 - b[:, 0:1] - b[:, 1:2]) * pv[:, 0] + b[:, 0:1] * pv[:, 1] + b[:, 1:2] * pv[:, 2]
    np.testing.assert_allclose(rec, op[m], atol=2e-5)
    # interpolated vertex normals
    vn = vert / np.linalg.norm(vert, axis=1, keepdims=True)
    nrm2 = torch.zeros(n, 3)
    face = torch.full((n, 3), 7.0)
    k.gbuffer_primary(w.packed, H.t(sc["rays_o"]), H.t(sc["rays_d"]), occ, pos, nrm2, depth, prim, bary, H.t(vn.astype(np.float32)), H.t(tri),
                      geom_normal=face)
    assert torch.equal(face, nrm)  # the face normal (zero on misses) survives beside the interpolated one
    vv = vn.astype(np.float32)[tri[opr[m]]]
    want = (1 - b[:, 0:1] - b[:, 1:2]) * vv[:, 0] + b[:, 0:1] * vv[:, 1] + b[:, 1:2] * vv[:, 2]
    np.testing.assert_allclose(nrm2.numpy()[m], want, atol=1e-6)
    # reverse scatter, with and without barycentric weights, C = 3 and C = 8
    rng = np.random.default_rng(0)
    for C, use_bary in ((3, True), (8, True), (5, False)):
        grad = rng.standard_normal((n, C)).astype(np.float32)
        grad[~m] = 7.0  # background gradients must be ignored
        out = torch.zeros(len(vert), C)
        k.interpolate_bwd(H.t(grad), prim, bary if use_bary else None, H.t(tri), out)
        want = _scatter_reference(grad, prim.numpy(), bary.numpy() if use_bary else None, tri, len(vert))
        np.testing.assert_allclose(out.numpy(), want, rtol=1e-4, atol=1e-5)


def test_frame_offset_word_shifts_every_random_stream():
    """The device-resident frame offset (include/mirres_b200.h) must be indistinguishable from a larger random_offset:
    it is what lets a captured CUDA graph be replayed with fresh random streams."""
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    sc = P.scene("T0")
    ref = P.oracle_run(sc)
    want = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], random_offset=4242 + 37)
    slangpy_shim.set_frame_offset("cpu", 37)
    try:
        got = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], random_offset=4242)
    finally:
        slangpy_shim.set_frame_offset("cpu", 0)
    for a, b in zip(want["totals"], got["totals"]):
        assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
    base = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], random_offset=4242)
    assert not np.array_equal(np.asarray(base["totals"][0]), np.asarray(got["totals"][0]))


def _bilateral_inputs(sc, seed=0):
    rng = np.random.default_rng(seed)
    g = sc["gbuffer"]
    n = sc["W"] * sc["H"]
    col = (rng.random((n, 3)) * g["occ_map"]).astype(np.float32)
    nrm = g["normal_map"] + 0.05 * rng.standard_normal((n, 3)).astype(np.float32)
    nrm = (nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-10)).astype(np.float32)
    z = g["depth_map"].astype(np.float32)
    zdz = np.concatenate([z, np.abs(rng.standard_normal((n, 1))).astype(np.float32) * 0.01 + 1e-3], 1).astype(np.float32)
    return col, nrm, zdz


def test_cross_bilateral_denoiser(oracle):
    """SURVEY.md 8f-3 (--use_bi_de): forward and transposed-gather backward bit-exact against the oracle's restatement of
    nerf/renderutils/c_src/denoising.cu; the backward is the adjoint of the forward (float64 autograd); the gb_depth
    branch of run_restir_di_with_pt runs on it."""
    from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth
    sc = P.scene("T0")
    W, Hh = sc["W"], sc["H"]
    col, nrm, zdz = _bilateral_inputs(sc)
    k = H.kernels()
    sigma = 1.0  # radius 7: 225 taps
    out = torch.zeros(W * Hh, 4)
    k.bilateral_fwd(W, Hh, sigma, H.t(col), H.t(nrm), H.t(zdz), out)
    want = oracle.bilateral_fwd(W, Hh, sigma, col, nrm, zdz)
    assert np.array_equal(out.numpy(), want)
    go = np.random.default_rng(1).standard_normal((W * Hh, 4)).astype(np.float32)
    g = torch.zeros(W * Hh, 3)
    k.bilateral_bwd(W, Hh, sigma, H.t(nrm), H.t(zdz), H.t(go), g)
    assert np.array_equal(g.numpy(), oracle.bilateral_bwd(W, Hh, sigma, nrm, zdz, go))
    # adjoint check in float64: out[:, :3] is linear in col with weights w(c, t)
    rad = 2 * int(np.ceil(sigma * 2.5)) + 1
    c64 = torch.tensor(col, dtype=torch.float64).view(Hh, W, 3).requires_grad_(True)
    n64, z64 = torch.tensor(nrm, dtype=torch.float64).view(Hh, W, 3), torch.tensor(zdz, dtype=torch.float64).view(Hh, W, 2)
    acc = torch.zeros(Hh, W, 3, dtype=torch.float64)
    ys, xs = torch.meshgrid(torch.arange(Hh), torch.arange(W), indexing="ij")
    for oy in range(-rad, rad + 1):
        for ox in range(-rad, rad + 1):
            y, x = ys + oy, xs + ox
            ok = (y >= 0) & (y < Hh) & (x >= 0) & (x < W)
            yc, xc = y.clamp(0, Hh - 1), x.clamp(0, W - 1)
            d2 = float(ox * ox + oy * oy)
            w = np.exp(-d2 / (2 * sigma * sigma)) * (n64[yc, xc] * n64).sum(-1).clamp(1e-4, 1.0) ** 128 * torch.exp(
                -(z64[yc, xc, 0] - z64[..., 0]).abs() / (z64[..., 1] * np.sqrt(d2)).clamp(min=1e-4))
            acc = acc + (w * ok)[..., None] * c64[yc, xc]
    (acc * torch.tensor(go[:, :3], dtype=torch.float64).view(Hh, W, 3)).sum().backward()
    ref = c64.grad.view(-1, 3).numpy()
    assert np.abs(g.numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
    np.testing.assert_allclose(out.numpy()[:, :3], acc.detach().view(-1, 3).numpy(), rtol=2e-4, atol=1e-6)
    # the driver's gb_depth branch (nerf/renderer_restir.py:529-541)
    w_ = _worker(sc)
    mods = R.load_m_for_restir(W, Hh, device="cpu")
    gb = {kk: H.t(v) for kk, v in sc["gbuffer"].items()}
    kd = gb["diffuse_map"].clone().requires_grad_(True)
    outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(0.0), H.t(zdz), w_, *mods, H.t(sc["env"]),
                                   gb["occ_map"], gb["normal_map"], gb["depth_map"], kd, gb["roughness_specular"],
                                   gb["ray_dir_map"], gb["pos_map"], None, None, None, None, W, Hh, 2, 2, 2, 2.0, 0.1, 0.001,
                                   random_offset=5)
    outs[0].sum().backward()
    assert all(torch.isfinite(o).all() for o in outs) and kd.grad.abs().sum() > 0


def test_path_kernels_without_alive_lists_give_the_same_images():
    """The path kernels walk the list of paths that are still alive when the workspace carries the signature of the
    call sequence, and all foreground pixels (with the reference's stop-flag test) otherwise.  Both must give the same
    result: the driver's normal run against a run whose kernels wipe the signature before every shaded vertex."""
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    sc = P.scene("T2", 0.4)
    ref = P.oracle_run(sc, max_bounce=3)
    got = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], max_bounce=3)
    assert P.compare(ref, got) == []

    class NoLists(H.HostKernels):
        wiped = 0

        def bounce_shade(self, *a, **kw):
            ws = kw["ws"] if "ws" in kw else a[-1]  # the workspace is the last argument of Kernels.bounce_shade
            ws[72:80].view(torch.int32).zero_()  # MR_CTR_ALIVE_SIG words: a foreign signature -> fallback to all foreground pixels
            NoLists.wiped += 1
            return super().bounce_shade(*a, **kw)

    slangpy_shim.set_kernels(NoLists())
    try:
        got2 = P.product_run(sc, _worker(sc), "cpu", ref["prepared"], max_bounce=3)
    finally:
        H.activate()
    assert NoLists.wiped == sc["cfg"]["spp"] * 3
    assert P.compare(ref, got2) == []
