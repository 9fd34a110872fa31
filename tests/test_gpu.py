"""GPU parity tests proper: libmirres_b200.so through the C ABI against the oracle on identical seeded inputs.

Bar (BASELINE.json north_star): BVH topology, hit ids and reservoir sample indices bit-exact; radiance and reservoir
weights within 1e-4 relative; gradients within 1e-3.  Because oracle and kernels share the numerical contract
(include/mirres_fpmath.h, no FMA contraction) the forward pass is in fact required to be bit-exact here.
"""
import numpy as np
import pytest
import torch

import parity as P
from oracle import backward as B
from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
FWD_RTOL = 1e-4   # stated tolerance for radiance / reservoir weights (integer outputs are always exact)
GRAD_RTOL = 1e-3  # stated tolerance for gradients


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


@pytest.mark.parametrize("name", ["T0", "T2", "C1"])
def test_lbvh_bit_exact(kernels, oracle, name):
    cfg = synth.CONFIGS[name]
    v, f = synth.make_mesh(cfg)
    b = oracle.Bvh(v, f)
    w = R.restirbvhWorker(tt(v), tt(f))
    info, aabb = w.update_bvh(want_sorted_codes=True)
    assert (w.sorted_codes.cpu().numpy() == b.sorted_codes).all()
    assert (info.cpu().numpy() == b.info).all()
    assert (aabb.cpu().numpy() == b.aabb).all()


def test_lbvh_edge_cases(kernels, oracle):
    v = np.array([[0, 0, 0], [1, 0, 0.25], [0, 1, 0.5], [1, 1, 0.5], [2, 2, 2]], np.float32)
    for tri in ([[0, 1, 2]], [[0, 1, 2], [1, 3, 2]], [[0, 1, 2], [0, 1, 2], [0, 1, 2], [1, 3, 2], [1, 3, 4]]):
        tri = np.array(tri, np.int32)
        b = oracle.Bvh(v, tri)
        w = R.restirbvhWorker(tt(v), tt(tri))
        info, aabb = w.update_bvh(want_sorted_codes=True)
        assert (info.cpu().numpy() == b.info).all() and (aabb.cpu().numpy() == b.aabb).all()
        o = np.array([[0.2, 0.2, 1.0], [0.9, 0.9, 3.0]], np.float32)
        d = np.array([[0, 0, -1], [0.01, 0.02, -1]], np.float32)
        hit = torch.zeros(2, dtype=torch.int32, device=DEV)
        t = torch.zeros(2, device=DEV)
        prim = torch.zeros(2, dtype=torch.int32, device=DEV)
        kernels.trace_closest(w.packed, tt(o), tt(d), hit, t, None, None, prim)
        oh, ot, _, _, opr = oracle.trace(b, o, d)
        assert (hit.cpu().numpy() == oh).all() and (t.cpu().numpy() == ot).all() and (prim.cpu().numpy() == opr).all()


def test_lbvh_granular_stages_match_fused(kernels, oracle):
    """The per-kernel entry points the slangpy-protocol shim serves (reference update_bvh, renderer_restir.py:25-89)."""
    from mirres_restir_nerf_mesh_b200 import slangpy_shim as slangpy
    v, f = synth.make_mesh(synth.CONFIGS["T2"])
    b = oracle.Bvh(v, f)
    vt, ft = tt(v), tt(f)
    F = f.shape[0]
    m_ele = slangpy.loadModule('nerf/bvhworkers/get_elements.slang')
    m_mc = slangpy.loadModule('nerf/bvhworkers/lbvh_morton_codes.slang')
    m_sort = slangpy.loadModule('nerf/bvhworkers/lbvh_single_radixsort.slang')
    m_h = slangpy.loadModule('nerf/bvhworkers/lbvh_hierarchy.slang')
    m_bb = slangpy.loadModule('nerf/bvhworkers/lbvh_bounding_boxes.slang')
    pidx = torch.zeros((F, 1), dtype=torch.int, device=DEV)
    eaabb = torch.zeros((F, 6), dtype=torch.float, device=DEV)
    m_ele.generateElements(vert=vt, v_indx=ft, ele_primitiveIdx=pidx, ele_aabb=eaabb).launchRaw(blockSize=(256, 1, 1), gridSize=((F + 255) // 256, 1, 1))
    pc = m_mc.pushConstantsMortonCodes(g_num_elements=F, g_min_x=eaabb[:, 0].min(), g_min_y=eaabb[:, 1].min(), g_min_z=eaabb[:, 2].min(),
                                       g_max_x=eaabb[:, 3].max(), g_max_y=eaabb[:, 4].max(), g_max_z=eaabb[:, 5].max())
    codes = torch.zeros((F, 2), dtype=torch.int, device=DEV)
    m_mc.morton_codes(pc=pc, ele_aabb=eaabb, morton_codes_ele=codes).launchRaw(blockSize=(256, 1, 1), gridSize=((F + 255) // 256, 1, 1))
    pingpong = torch.zeros((F, 2), dtype=torch.int, device=DEV)
    m_sort.radix_sort(g_num_elements=F, g_elements_in=codes, g_elements_out=pingpong).launchRaw(blockSize=(256, 1, 1), gridSize=(1, 1, 1))
    assert (codes.cpu().numpy() == b.sorted_codes).all()
    info = torch.zeros((2 * F - 1, 3), dtype=torch.int, device=DEV)
    aabb = torch.zeros((2 * F - 1, 6), dtype=torch.float, device=DEV)
    cinfo = torch.zeros((2 * F - 1, 2), dtype=torch.int, device=DEV)
    m_h.hierarchy(g_num_elements=F, ele_primitiveIdx=pidx, ele_aabb=eaabb, g_sorted_morton_codes=codes, g_lbvh_info=info,
                  g_lbvh_aabb=aabb, g_lbvh_construction_infos=cinfo).launchRaw(blockSize=(256, 1, 1), gridSize=((F + 255) // 256, 1, 1))
    heights = torch.zeros((F, 1), dtype=torch.int, device=DEV)
    m_bb.get_bvh_height(g_num_elements=F, g_lbvh_info=info, g_lbvh_aabb=aabb, g_lbvh_construction_infos=cinfo, tree_heights=heights).launchRaw()
    for i in range(int(heights.max())):
        m_bb.get_bbox(g_num_elements=F, expected_height=i + 1, g_lbvh_info=info, g_lbvh_aabb=aabb, g_lbvh_construction_infos=cinfo).launchRaw()
    m_bb.set_root(g_lbvh_info=info, g_lbvh_aabb=aabb).launchRaw(blockSize=(1, 1, 1), gridSize=(1, 1, 1))
    assert (info.cpu().numpy() == b.info).all() and (aabb.cpu().numpy() == b.aabb).all()


def test_rays_bit_exact(kernels, oracle):
    sc = P.scene("C1")
    w = make_worker(sc)
    rng = np.random.default_rng(5)
    hitm = sc["hit"] > 0
    d = rng.standard_normal((int(hitm.sum()), 3)).astype(np.float32)
    o = (sc["pos"][hitm] + 0.01 * d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    for org, dirs in ((sc["rays_o"], sc["rays_d"]), (o, d)):
        n = len(org)
        hit = torch.zeros(n, dtype=torch.int32, device=DEV)
        t, pos, nrm = torch.zeros(n, device=DEV), torch.zeros(n, 3, device=DEV), torch.zeros(n, 3, device=DEV)
        prim = torch.zeros(n, dtype=torch.int32, device=DEV)
        kernels.trace_closest(w.packed, tt(org), tt(dirs), hit, t, pos, nrm, prim)
        oh, ot, op, on, opr = oracle.trace(sc["bvh"], org, dirs)
        assert (hit.cpu().numpy() == oh).all() and (prim.cpu().numpy() == opr).all()
        assert (t.cpu().numpy() == ot).all() and (pos.cpu().numpy() == op).all()
        assert (nrm.cpu().numpy()[oh > 0] == on[oh > 0]).all()
        anyh = torch.zeros(n, dtype=torch.int32, device=DEV)
        kernels.trace_any(w.packed, tt(org), tt(dirs), anyh)
        assert (anyh.cpu().numpy() == oh).all()
    assert (ot[oh > 0] < 0).any()  # negative-t quirk exercised


def test_gbuffer_primary_and_gradient_scatter(kernels, oracle):
    """SURVEY.md 8f-2 through the C ABI: primary G-buffer bit-exact against the oracle's closest-hit; the warp-aggregated
    reverse scatter against a float64 index_add (tolerance: float atomics are order-nondeterministic)."""
    sc = P.scene("C1")
    w = make_worker(sc)
    n = len(sc["rays_o"])
    occ, depth = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pos, nrm = torch.zeros(n, 3, device=DEV), torch.zeros(n, 3, device=DEV)
    prim, bary = torch.zeros(n, dtype=torch.int32, device=DEV), torch.zeros(n, 2, device=DEV)
    kernels.gbuffer_primary(w.packed, tt(sc["rays_o"]), tt(sc["rays_d"]), occ, pos, nrm, depth, prim, bary)
    oh, ot, op, on, opr = oracle.trace(sc["bvh"], sc["rays_o"], sc["rays_d"])
    m = oh > 0
    assert (occ.cpu().numpy() == oh).all() and (prim.cpu().numpy() == opr).all()
    assert (pos.cpu().numpy()[m] == op[m]).all() and (nrm.cpu().numpy()[m] == on[m]).all()
    # wavefront launch shape (persistent queue tracer): identical maps
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    o2, d2 = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p2, n2 = torch.zeros(n, 3, device=DEV), torch.zeros(n, 3, device=DEV)
    pr2, b2 = torch.zeros(n, dtype=torch.int32, device=DEV), torch.zeros(n, 2, device=DEV)
    kernels.gbuffer_primary(w.packed, tt(sc["rays_o"]), tt(sc["rays_d"]), o2, p2, n2, d2, pr2, b2,
                            ws=slangpy_shim.workspace(torch.device(DEV), n))
    for a_, b_ in ((occ, o2), (pos, p2), (nrm, n2), (depth, d2), (prim, pr2), (bary, b2)):
        assert torch.equal(a_, b_)
    tri = tt(sc["tri"])
    V = len(sc["vert"])
    # interpolated vertex normals with the face normal kept beside them (both launch shapes)
    vn = torch.nn.functional.normalize(tt(sc["vert"]), dim=-1)
    for ws in (None, slangpy_shim.workspace(torch.device(DEV), n)):
        sm, face = torch.zeros(n, 3, device=DEV), torch.full((n, 3), 7.0, device=DEV)
        kernels.gbuffer_primary(w.packed, tt(sc["rays_o"]), tt(sc["rays_d"]), o2, p2, sm, d2, pr2, b2, vnormal=vn, tri=tri,
                                ws=ws, geom_normal=face)
        assert torch.equal(face, nrm)
        fg = prim >= 0
        c = vn[tri.long()[prim[fg].long()]]
        want = (1 - bary[fg, 0:1] - bary[fg, 1:2]) * c[:, 0] + bary[fg, 0:1] * c[:, 1] + bary[fg, 1:2] * c[:, 2]
        assert torch.allclose(sm[fg], want, atol=1e-6) and bool((sm[~fg] == 0).all())
    g = torch.Generator(device="cpu").manual_seed(0)
    for C, use_bary, contended in ((3, True, False), (8, True, False), (5, False, False), (8, True, True)):
        grad = torch.randn(n, C, generator=g).to(DEV)
        pr = prim.clone()
        if contended:  # every foreground pixel sees one of four triangles: exercises the aggregation path
            pr = torch.where(prim >= 0, prim % 4, prim)
        out = torch.zeros(V, C, device=DEV)
        kernels.interpolate_bwd(grad, pr, bary if use_bary else None, tri, out)
        b = bary.double() if use_bary else torch.full((n, 2), 1 / 3, device=DEV, dtype=torch.float64)
        wts = torch.stack((1 - b[:, 0] - b[:, 1], b[:, 0], b[:, 1]), 1)
        fg = pr >= 0
        want = torch.zeros(V, C, device=DEV, dtype=torch.float64)
        for kk in range(3):
            want.index_add_(0, tri[pr[fg].long(), kk].long(), wts[fg, kk:kk + 1] * grad[fg].double())
        torch.testing.assert_close(out.double(), want, rtol=GRAD_RTOL, atol=1e-3 if contended else 1e-5)


@pytest.mark.parametrize("name,metallic", [("T0", 0.0), ("T1", 0.0), ("T2", 0.4), ("C1", 0.0)])
def test_pipeline_parity(kernels, name, metallic):
    sc = P.scene(name, metallic)
    spp = 2 if name == "C1" else None
    mb = 2 if name == "C1" else None
    ref = P.oracle_run(sc, spp=spp, max_bounce=mb)
    got = P.product_run(sc, make_worker(sc), DEV, ref["prepared"], spp=spp, max_bounce=mb)
    # stated bar: integer outputs exact, floats to 1e-4 ...
    assert P.compare(ref, got, rtol=FWD_RTOL) == []
    # ... and, given the shared numerical contract, bit equality of every tensor
    assert P.compare(ref, got, rtol=0.0) == []


def test_env_distribution_and_tiles(kernels, oracle):
    env = synth.envmap(256, 512)
    tex = np.ascontiguousarray(env[::-1].reshape(-1, 3))
    want = oracle.env_build_distribution(tex, 512, 256)
    got = R.make_sampleable(None, tt(tex), 512, 256)
    for a, b in zip(got, want):
        assert (a.cpu().numpy() == b).all()
    ld, uv, pdf = oracle.light_tiles(tex, 512, 256, want, 12345)
    gld = torch.zeros((131072, 3), device=DEV)
    guv = torch.zeros((131072, 2), dtype=torch.int32, device=DEV)
    gpdf = torch.zeros((131072, 1), device=DEV)
    kernels.light_tiles(tt(tex), 512, 256, got, 12345, 128, 1024, gld, guv, gpdf)
    assert (guv.cpu().numpy() == uv).all() and (gld.cpu().numpy() == ld).all() and (gpdf.cpu().numpy() == pdf).all()
    # granular protocol (reference GenerateLightTiles.py:4-29 with torch scans in between) stays close to the fused path
    from mirres_restir_nerf_mesh_b200 import slangpy_shim as slangpy
    m = slangpy.loadModule('nerf/ScreenSpaceReSTIR/make_sampleable.slang')
    weight = torch.zeros([512 * 256, 1], device=DEV)
    m.make_sampleable(env_tex=tt(tex), weight=weight, width=512, height=256).launchRaw()
    pdf_ = weight.reshape(256, 512, 1)
    cdf_ = torch.cat([torch.zeros([256, 1], device=DEV), pdf_.cumsum(1).reshape(256, 512)], dim=-1).reshape(-1, 1)
    m.Distribution2D(w=512, h=256, pdf_=weight, cdf_=cdf_).launchRaw()
    assert torch.allclose(weight, got[0], rtol=1e-4, atol=1e-9)


def test_backward_kernels(kernels):
    from test_backward_cpu import assert_grad_close
    sc = P.scene("T1", 0.3)
    ref = P.oracle_run(sc, spp=1)
    s, g = ref["snapshots"][0], ref["prepared"]
    W, Hh = sc["W"], sc["H"]
    n = W * Hh
    rng = np.random.default_rng(0)
    gc, gd, gs = [rng.standard_normal((n, 3)).astype(np.float32) for _ in range(3)]
    z = lambda *s_: torch.zeros(*s_, device=DEV)
    gN, gK, gR, gL = z(n, 3), z(n, 3), z(n, 2), z(n, 3)
    kernels.final_shading_bwd(tt(s["fs_dir"]), tt(s["fs_dist"]), tt(s["fs_Li"]), W, Hh, tt(g["occ_map"]), tt(g["normal_map"]),
                              tt(g["ray_dir_map"]), tt(g["diffuse_map"]), tt(g["roughness_specular"]), tt(gc), tt(gd),
                              tt(gs), gN, gK, gR, gL)
    rN, rK, rR, rL = B.final_shading_grads(s["fs_dir"], s["fs_dist"], s["fs_Li"], g["occ_map"], g["normal_map"],
                                           g["ray_dir_map"], g["diffuse_map"], g["roughness_specular"], gc, gd, gs)
    for a, b, nm in ((gN, rN, "normal"), (gK, rK, "kd"), (gR, rR, "rough_metal"), (gL, rL, "Li")):
        assert_grad_close(a.cpu().numpy(), b, nm, GRAD_RTOL)
    He, We = sc["env"].shape[:2]
    gLi = rng.standard_normal((n, 3)).astype(np.float32)
    ge = z(He * We, 3)
    kernels.eval_final_bwd([tt(a) for a in s["res"]], We, He, W, Hh, tt(s["vis"]), tt(gLi), ge)
    want = B.eval_final_grad_env(s["res"][0], s["res"][3], s["vis"], gLi, We, He)
    assert np.abs(ge.cpu().numpy() - want).max() <= GRAD_RTOL * np.abs(want).max()
    color = (rng.random((n, 3)) * g["occ_map"]).astype(np.float32)
    go = rng.standard_normal((n, 3)).astype(np.float32)
    for step in (2, 1):
        out = z(n, 3)
        kernels.eaw_fwd(2.0, 0.1, 0.001, W, Hh, step, tt(g["occ_map"]), tt(color), tt(g["normal_map"]), tt(g["pos_map"]), out)
        gC, gNn, gP, scr = z(n, 3), z(n, 3), z(n, 3), z(n)
        kernels.eaw_bwd(2.0, 0.1, 0.001, W, Hh, step, tt(g["occ_map"]), tt(color), tt(g["normal_map"]), tt(g["pos_map"]),
                        out, tt(go), gC, gNn, gP, scr)
        o64, rC, rNn, rP = B.eaw_grads(2.0, 0.1, 0.001, W, Hh, step, g["occ_map"], color, g["normal_map"], g["pos_map"], go)
        assert np.abs(out.cpu().numpy() - o64).max() < 1e-5
        for a, b, nm in ((gC, rC, "color"), (gNn, rNn, "normal"), (gP, rP, "pos")):
            assert_grad_close(a.cpu().numpy(), b, "eaw " + nm, GRAD_RTOL)


def test_env_gradient_scatter_under_contention(kernels):
    """All pixels pick the same texel quad (worst case for atomics): the warp-aggregated scatter must still sum right."""
    n, We, He = 4096, 64, 32
    ld = np.zeros((n, 3), np.float32)
    ld[:, 0] = 1.0
    ld[:, 1] = 0.61
    ld[:, 2] = 0.37
    ld[::7, 1] = 0.2  # a second quad in some lanes
    res = (ld, np.ones((n, 1), np.float32), np.ones((n, 1), np.int32), np.full((n, 1), 0.5, np.float32))
    vis = np.ones((n, 1), np.float32)
    rng = np.random.default_rng(3)
    gLi = rng.standard_normal((n, 3)).astype(np.float32)
    ge = torch.zeros(He * We, 3, device=DEV)
    kernels.eval_final_bwd([tt(a) for a in res], We, He, 64, 64, tt(vis), tt(gLi), ge)
    want = B.eval_final_grad_env(ld, res[3], vis, gLi, We, He)
    assert np.abs(ge.cpu().numpy() - want).max() <= GRAD_RTOL * np.abs(want).max()


def test_training_step_through_reference_api(kernels):
    """run_restir_di_with_pt with the reference signature: fwd + bwd, gradients finite and deterministic in forward."""
    sc = P.scene("T2", 0.2)
    W, Hh = sc["W"], sc["H"]
    w = R.restirbvhWorker(tt(sc["vert"]), tt(sc["tri"]))
    w.update_mesh(tt(sc["vert"]), tt(sc["tri"]))
    mods = R.load_m_for_restir(W, Hh)
    finals = []
    for _ in range(2):
        g = {k: tt(v) for k, v in sc["gbuffer"].items()}
        env = tt(sc["env"]).requires_grad_(True)
        normal = g["normal_map"].clone().requires_grad_(True)
        kd = g["diffuse_map"].clone().requires_grad_(True)
        rs = g["roughness_specular"].clone().requires_grad_(True)
        outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(0.2), None, w, *mods, env, g["occ_map"], normal,
                                       g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"], None, None, None, None, W, Hh,
                                       2, 2, 2, 2.0, 0.1, 0.001, random_offset=7)
        assert len(outs) == 6 and all(o.shape == (W * Hh, 3) for o in outs)
        outs[0].sum().backward()
        for t in (env, normal, kd, rs):
            assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
        finals.append(outs[0].detach().clone())
    assert torch.equal(finals[0], finals[1])


def test_full_size_properties(kernels):
    """BASELINE config C2 sizes (500k triangles, 800x800): size-independent properties instead of an oracle run."""
    cfg = synth.CONFIGS["C2"]
    v, f = synth.make_mesh(cfg)
    F = f.shape[0]
    w = R.restirbvhWorker(tt(v), tt(f))
    info, aabb = w.update_bvh(want_sorted_codes=True)
    codes = w.sorted_codes.long()
    # sortedness + stability + permutation
    assert bool((codes[1:, 0] >= codes[:-1, 0]).all())
    same = codes[1:, 0] == codes[:-1, 0]
    assert bool((codes[1:, 1][same] > codes[:-1, 1][same]).all())
    assert torch.equal(torch.sort(codes[:, 1])[0], torch.arange(F, device=DEV))
    # tree: every node except the root is the child of exactly one internal node; boxes are exact unions
    leaf = F - 1
    il = info[:leaf].long()
    children = torch.cat([il[:, 0], il[:, 1]])
    assert torch.equal(torch.sort(children)[0], torch.arange(1, 2 * F - 1, device=DEV))
    assert torch.equal(aabb[:leaf, :3], torch.minimum(aabb[il[:, 0], :3], aabb[il[:, 1], :3]))
    assert torch.equal(aabb[:leaf, 3:], torch.maximum(aabb[il[:, 0], 3:], aabb[il[:, 1], 3:]))
    assert torch.equal(torch.sort(info[leaf:, 2].long())[0], torch.arange(F, device=DEV))
    # rebuild is idempotent (bit-identical)
    info2, aabb2 = w.update_bvh()
    assert torch.equal(info, info2) and torch.equal(aabb, aabb2)
    # rays: any-hit agrees with closest-hit on 640k primary rays; hits lie on the reported triangle
    ro, rd = synth.camera_rays(cfg["W"], cfg["H"])
    n = len(ro)
    hit = torch.zeros(n, dtype=torch.int32, device=DEV)
    t, pos = torch.zeros(n, device=DEV), torch.zeros(n, 3, device=DEV)
    prim = torch.zeros(n, dtype=torch.int32, device=DEV)
    kernels.trace_closest(w.packed, tt(ro), tt(rd), hit, t, pos, None, prim)
    anyh = torch.zeros(n, dtype=torch.int32, device=DEV)
    kernels.trace_any(w.packed, tt(ro), tt(rd), anyh)
    assert torch.equal(hit, anyh) and 0.1 < hit.float().mean() < 0.6
    m = hit > 0
    tri = tt(f).long()[prim[m].long()]
    vv = tt(v)
    a, b_, c = vv[tri[:, 0]], vv[tri[:, 1]], vv[tri[:, 2]]
    nrm = torch.cross(b_ - a, c - a, dim=1)
    nrm = nrm / nrm.norm(dim=1, keepdim=True)
    assert float(((pos[m] - a) * nrm).sum(1).abs().max()) < 1e-4


def test_captured_step_replays_with_fresh_random_streams(kernels):
    """graphed.CapturedStep: a whole render + backward recorded into a CUDA graph (side streams included) reproduces
    the eager call, and the device-resident frame offset gives replay k the random streams of random_offset + k."""
    from mirres_restir_nerf_mesh_b200.graphed import CapturedStep
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    sc = P.scene("T1", 0.2)
    W, Hh = sc["W"], sc["H"]
    worker = make_worker(sc)
    mods = R.load_m_for_restir(W, Hh, device=DEV)
    mat = synth.ProceduralMaterial(sc["metallic"])
    g = {k: tt(v) for k, v in sc["gbuffer"].items()}
    wgt = torch.linspace(0.5, 1.5, W * Hh * 3, device=DEV).reshape(W * Hh, 3)

    def step(env, random_offset=900):
        env_l = env.detach().clone().requires_grad_(True)
        normal = g["normal_map"].clone().requires_grad_(True)
        outs = R.run_restir_di_with_pt(False, 1, 1, 1, mat, None, worker, *mods, env_l, g["occ_map"].clone(), normal,
                                       g["depth_map"], g["diffuse_map"], g["roughness_specular"], g["ray_dir_map"],
                                       g["pos_map"], None, None, None, None, W, Hh, 3, 2, 2, 2.0, 0.1, 0.001,
                                       random_offset=random_offset)
        (outs[0] * wgt).sum().backward()
        return outs[0].detach(), env_l.grad, normal.grad

    env = tt(sc["env"])
    cap = CapturedStep(step, dict(env=env))
    try:
        for k in (0, 5):
            want = [x.clone() for x in step(env, random_offset=900 + k)]
            cap.set_frame_offset(k)
            got = cap(env=env)
            torch.cuda.synchronize()
            assert torch.equal(got[0], want[0]), (k, (got[0] - want[0]).abs().max().item(), (got[0] != want[0]).sum().item())
            torch.testing.assert_close(got[1], want[1], rtol=GRAD_RTOL, atol=1e-5)
            torch.testing.assert_close(got[2], want[2], rtol=GRAD_RTOL, atol=1e-5)
        assert not torch.equal(step(env, 900)[0], step(env, 905)[0])
    finally:
        slangpy_shim.set_frame_offset(torch.device(DEV, torch.cuda.current_device()), 0)


def _render_t2(sc, w, mods, **kw):
    g = {k: tt(v) for k, v in sc["gbuffer"].items()}
    env = tt(sc["env"]).requires_grad_(True)
    normal = g["normal_map"].clone().requires_grad_(True)
    tex = torch.cat((g["diffuse_map"], torch.zeros_like(g["diffuse_map"])), dim=1).requires_grad_(True)
    kd = tex[:, 0:3]  # strided view, row stride 6 floats, as render_stage1 passes it (nerf/renderer.py:1018-1020)
    rs = g["roughness_specular"].clone().requires_grad_(True)
    outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, w, *mods, env,
                                   g["occ_map"], normal, g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"], None, None,
                                   None, None, sc["W"], sc["H"], 3, 2, 2, 2.0, 0.1, 0.001, random_offset=99, **kw)
    wgt = torch.linspace(0.5, 1.5, outs[0].numel(), device=DEV).reshape(outs[0].shape)
    (outs[0] * wgt).sum().backward()
    return [o.detach() for o in outs], [env.grad, normal.grad, tex.grad, rs.grad]


def test_concurrent_schedule_equals_sequential_schedule(kernels):
    """The default schedule on CUDA tensors (indirect chains, initial candidates and shading on side streams, batched
    denoiser, fused map preparation) against the reference's sequential order: images bit-identical, gradients to the
    stated tolerance (accumulation order of the autograd engine differs)."""
    sc = P.scene("T2", 0.3)
    w = R.restirbvhWorker(tt(sc["vert"]), tt(sc["tri"]))
    w.update_mesh(tt(sc["vert"]), tt(sc["tri"]))
    mods = R.load_m_for_restir(sc["W"], sc["H"])
    seq, gseq = _render_t2(sc, w, mods, overlap=False, batched_denoise=False, fused_prepare=False)
    par, gpar = _render_t2(sc, w, mods)
    for a, b in zip(seq, par):
        assert torch.equal(a, b)
    for a, b in zip(gseq, gpar):
        assert a.abs().sum() > 0
        assert (a - b).abs().max() <= GRAD_RTOL * a.abs().max()


def test_empty_and_single_pixel_frames(kernels):
    """Edge cases of the wavefront machinery: no foreground pixel at all (empty ray queues), and a 1x1 frame."""
    sc = P.scene("T0")
    w = R.restirbvhWorker(tt(sc["vert"]), tt(sc["tri"]))
    w.update_mesh(tt(sc["vert"]), tt(sc["tri"]))
    for W, Hh, occ_value in ((sc["W"], sc["H"], 0.0), (1, 1, 1.0)):
        n = W * Hh
        mods = R.load_m_for_restir(W, Hh)
        g = {k: tt(v)[:n].clone() for k, v in sc["gbuffer"].items()}
        if occ_value > 0:  # one foreground pixel: take the first hit of the scene
            i = int(np.nonzero(sc["hit"] > 0)[0][0])
            g = {k: tt(v)[i:i + 1].clone() for k, v in sc["gbuffer"].items()}
        g["occ_map"].fill_(occ_value)
        env = tt(sc["env"]).requires_grad_(True)
        outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(0.0), None, w, *mods, env, g["occ_map"],
                                       g["normal_map"], g["depth_map"], g["diffuse_map"], g["roughness_specular"],
                                       g["ray_dir_map"], g["pos_map"], None, None, None, None, W, Hh, 2, 2, 2, 2.0, 0.1, 0.001,
                                       random_offset=3)
        torch.cuda.synchronize()
        assert all(o.shape == (n, 3) and torch.isfinite(o).all() for o in outs)
        if occ_value == 0:
            assert bool((outs[0] == 1.0).all())  # background pixels of the final image are forced to 1 (:546-547)


def test_cross_bilateral_denoiser_gpu(kernels, oracle):
    """SURVEY.md 8f-3 through the C ABI: bit-exact against the oracle (both directions are deterministic gathers)."""
    import test_hostcheck_parity as THP
    sc = P.scene("T1")
    W, Hh = sc["W"], sc["H"]
    col, nrm, zdz = THP._bilateral_inputs(sc, 3)
    for sigma in (1.0, 4.0):  # 4.0 = the reference's factor 2 (radius 21, 1849 taps)
        out = torch.zeros(W * Hh, 4, device=DEV)
        kernels.bilateral_fwd(W, Hh, sigma, tt(col), tt(nrm), tt(zdz), out)
        assert np.array_equal(out.cpu().numpy(), oracle.bilateral_fwd(W, Hh, sigma, col, nrm, zdz))
        go = np.random.default_rng(1).standard_normal((W * Hh, 4)).astype(np.float32)
        g = torch.zeros(W * Hh, 3, device=DEV)
        kernels.bilateral_bwd(W, Hh, sigma, tt(nrm), tt(zdz), tt(go), g)
        assert np.array_equal(g.cpu().numpy(), oracle.bilateral_bwd(W, Hh, sigma, nrm, zdz, go))


def test_cross_bilateral_against_reference_kernel(kernels, oracle):
    """The REFERENCE's own kernels (nerf/renderutils/c_src/denoising.cu:14-130, compiled unmodified from the reference
    tree into oracle/_ref/libref_renderutils.so and launched as torch_bindings.cpp:201-246 launches them) against the
    product (mirres_bilateral_fwd/_bwd) and against the oracle's restatement.  The reference build contracts to FMA and
    calls libdevice expf / powf, the product evaluates the same expression tree without contraction, so the bar is the
    north-star tolerance (1e-4 forward, 1e-3 gradients), not bit-exactness.  Frames with ragged 8 x 8 block edges."""
    from oracle import ref as REF
    if not REF.available():
        pytest.skip("oracle/_ref/libref_renderutils.so not built (needs the reference tree at build time)")
    import test_hostcheck_parity as THP
    sc = P.scene("T1")
    W, Hh = sc["W"], sc["H"]
    col, nrm, zdz = THP._bilateral_inputs(sc, 3)

    def close(got, want, rtol):
        # signed sums (backward) cancel: scale the absolute term by the magnitude of the data
        return np.allclose(got, want, rtol=rtol, atol=rtol * 1e-2 * float(np.abs(want).max()))

    for fx, fy in ((W, Hh), (W - 3, Hh - 5), (5, 3), (1, 1)):
        n = fx * fy
        c, nn, z = col[:n].copy(), nrm[:n].copy(), zdz[:n].copy()
        for sigma in (1.0, 4.0, 1e-4):  # 4.0 = the reference's factor 2 (radius 21); 1e-4 = its lower clamp (ops.py:195)
            want = REF.bilateral_fwd(fx, fy, sigma, tt(c), tt(nn), tt(z)).cpu().numpy()
            out = torch.zeros(n, 4, device=DEV)
            kernels.bilateral_fwd(fx, fy, sigma, tt(c), tt(nn), tt(z), out)
            assert close(out.cpu().numpy(), want, FWD_RTOL), (fx, fy, sigma)
            assert close(oracle.bilateral_fwd(fx, fy, sigma, c, nn, z), want, FWD_RTOL), (fx, fy, sigma)
            go = np.random.default_rng(1).standard_normal((n, 4)).astype(np.float32)
            wantg = REF.bilateral_bwd(fx, fy, sigma, tt(c), tt(nn), tt(z), tt(go)).cpu().numpy()
            g = torch.zeros(n, 3, device=DEV)
            kernels.bilateral_bwd(fx, fy, sigma, tt(nn), tt(z), tt(go), g)
            assert close(g.cpu().numpy(), wantg, GRAD_RTOL), (fx, fy, sigma)
            assert close(oracle.bilateral_bwd(fx, fy, sigma, nn, z, go), wantg, GRAD_RTOL), (fx, fy, sigma)
            # the composed op of ops.py:192-200 (normalised colour), on lit pixels
            lit = want[:, 3] > 1e-3
            assert np.allclose(out.cpu().numpy()[lit, :3] / out.cpu().numpy()[lit, 3:4], want[lit, :3] / want[lit, 3:4],
                               rtol=FWD_RTOL, atol=1e-6)


def test_c2_full_size_parity_against_oracle(kernels, oracle):
    """BASELINE config C2 at FULL size (500 000 triangles, 800 x 800, spp 4, 3 path vertices): LBVH, primary G-buffer and
    every intermediate tensor of the forward spp loop against the oracle (the oracle needs a few seconds on the host
    cores), then the concurrent schedule against the sequential one."""
    sc = P.scene("C2")
    ref = P.oracle_run(sc)
    w = make_worker(sc)
    assert (w.sorted_codes.cpu().numpy() == sc["bvh"].sorted_codes).all()
    assert (w.LBVHNode_info.cpu().numpy() == sc["bvh"].info).all()
    assert (w.LBVHNode_aabb.cpu().numpy() == sc["bvh"].aabb).all()
    n = sc["W"] * sc["H"]
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    occ, depth = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pos, nrm = torch.zeros(n, 3, device=DEV), torch.zeros(n, 3, device=DEV)
    prim, bary = torch.zeros(n, dtype=torch.int32, device=DEV), torch.zeros(n, 2, device=DEV)
    kernels.gbuffer_primary(w.packed, tt(sc["rays_o"]), tt(sc["rays_d"]), occ, pos, nrm, depth, prim, bary,
                            ws=slangpy_shim.workspace(torch.device(DEV), n))
    m = sc["hit"] > 0
    assert (occ.cpu().numpy() == sc["hit"]).all() and (prim.cpu().numpy() == sc["prim"]).all()
    assert (pos.cpu().numpy()[m] == sc["pos"][m]).all() and (nrm.cpu().numpy()[m] == sc["nrm"][m]).all()
    got = P.product_run(sc, w, DEV, ref["prepared"])
    bad = P.compare(ref, got, rtol=FWD_RTOL)
    assert not bad, bad[:5]
    assert not P.compare(ref, got, rtol=0.0), "forward pass is expected to be bit-exact"
    # default (concurrent) schedule == sequential schedule at full size
    mods = R.load_m_for_restir(sc["W"], sc["H"])
    g = {k: tt(v) for k, v in sc["gbuffer"].items()}
    outs = []
    for kw in (dict(overlap=False, batched_denoise=False, fused_prepare=False), dict()):
        with torch.no_grad():
            outs.append(R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(0.0), None, w, *mods, tt(sc["env"]),
                                                g["occ_map"].clone(), g["normal_map"], g["depth_map"], g["diffuse_map"],
                                                g["roughness_specular"], g["ray_dir_map"], g["pos_map"], None, None, None,
                                                None, sc["W"], sc["H"], 4, 2, 2, 2.0, 0.1, 0.001, random_offset=4242, **kw))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_fused_screen_kernels_gpu(kernels):
    """csrc/screen.cu and mirres_final_shading_bwd_multi on the GPU against the torch expressions / single-pass kernels
    they replace (the same checks the CPU suite runs on the host-check flavour)."""
    import fused_checks as C
    C.material_kernel_equals_torch_and_numpy_expressions(kernels, DEV)
    C.sum_images_is_the_sequential_torch_sum(kernels, DEV)
    C.composite_forward_exact_backward_matches_autograd(kernels, DEV)
    C.final_shading_bwd_multi_equals_the_single_pass_kernels(kernels, DEV)


def test_long_loop_and_prepared_lighting_gpu(kernels):
    """spp = 19 on the GPU: the concurrent schedule flushes its running sums every 16 images, releases the per-iteration
    images early and runs the one-node backward in two chunks; with and without prepare_lighting() the images equal the
    sequential schedule's bit for bit, the gradients to the stated tolerance."""
    sc = P.scene("T0", 0.3)
    w = R.restirbvhWorker(tt(sc["vert"]), tt(sc["tri"]))
    w.update_mesh(tt(sc["vert"]), tt(sc["tri"]))
    W, Hh, spp = sc["W"], sc["H"], 19

    def render(prepared, **kw):
        mods = R.load_m_for_restir(W, Hh)
        g = {k: tt(v) for k, v in sc["gbuffer"].items()}
        env = tt(sc["env"]).requires_grad_(True)
        normal = g["normal_map"].clone().requires_grad_(True)
        kd = g["diffuse_map"].clone().requires_grad_(True)
        rs = g["roughness_specular"].clone().requires_grad_(True)
        if prepared:
            kw["lighting"] = R.prepare_lighting(mods[0], mods[1], mods[8], mods[9], mods[10], env, spp, 321,
                                                frame_pixels=sc["W"] * sc["H"])
        outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, w, *mods, env,
                                       g["occ_map"].clone(), normal, g["depth_map"], kd, rs, g["ray_dir_map"], g["pos_map"],
                                       None, None, None, None, W, Hh, spp, 2, 2, 2.0, 0.1, 0.001, random_offset=321, **kw)
        wgt = torch.linspace(0.5, 1.5, outs[0].numel(), device=DEV).reshape(outs[0].shape)
        (outs[0] * wgt).sum().backward()
        torch.cuda.synchronize()
        return [o.detach() for o in outs], [env.grad, normal.grad, kd.grad, rs.grad]

    seq, gseq = render(False, overlap=False, batched_denoise=False, fused_prepare=False, fused_composite=False)
    for prepared in (False, True):
        par, gpar = render(prepared)
        for a, b in zip(seq, par):
            assert torch.equal(a, b)
        for a, b in zip(gseq, gpar):
            assert a.abs().sum() > 0
            assert (a - b).abs().max() <= GRAD_RTOL * a.abs().max()
