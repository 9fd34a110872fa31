"""Visibility tags (include/mirres_b200.h, mirres_set_visibility_tags): the final-visibility pass skips the rays whose
answer an earlier pass of the same spp loop has produced.  Three checks, all on the CPU:
  * the oracle records the same provenance beside its reservoirs and traces EVERY final ray anyway: a tagged ray that
    turns out occluded is a violation of the claim the product relies on (must be 0), with and without motion vectors;
  * the product (host flavour of the kernel source) with tags on, with tags off and the oracle agree bit for bit;
  * with tags on the pass queues exactly the rays the oracle counts as unknown."""
import numpy as np
import pytest

import hostcheck as H
import parity as P
from mirres_restir_nerf_mesh_b200 import renderer_restir as R


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


def _motion(sc, dx, dy):
    """motion vectors in frame fractions as TemporalResampling.slang:52-56 reads them: constant, or (dx = "random")
    independent per pixel within +-dy pixels, so that some pixels find their history at home and others do not"""
    n = sc["W"] * sc["H"]
    m = np.empty((n, 2), np.float32)
    if dx == "random":
        r = np.random.default_rng(5).uniform(-dy, dy, size=(n, 2))
        r[::3] = 0.0                                   # a third of the pixels keep their place exactly
        m[:, 0], m[:, 1] = r[:, 0] / sc["W"], r[:, 1] / sc["H"]
    else:
        m[:, 0], m[:, 1] = dx / sc["W"], dy / sc["H"]
    return m


def _run(sc, tags, ref, **kw):
    old = R.USE_VIS_TAGS
    R.USE_VIS_TAGS = int(tags)
    try:
        return P.product_run(sc, _worker(sc), "cpu", ref["prepared"], **kw)
    finally:
        R.USE_VIS_TAGS = old


@pytest.mark.parametrize("name,metallic,spp,shift", [("T1", 0.0, 4, None), ("T2", 0.4, 4, None), ("C1", 0.0, 3, None),
                                                     ("T1", 0.0, 4, (3, -2)), ("C1", 0.4, 3, (-5, 1)), ("T2", 0.0, 5, (0.6, 0.6)),
                                                     ("T2", 0.4, 6, ("random", 1.5)), ("C1", 0.0, 4, ("random", 0.7))])
def test_tags_change_nothing_but_the_ray_count(name, metallic, spp, shift):
    sc = P.scene(name, metallic)
    motion = None if shift is None else _motion(sc, *shift)
    ref = P.oracle_run(sc, spp=spp, motion=motion)
    assert ref["provenance_violations"] == 0
    on = _run(sc, True, ref, spp=spp, motion=motion)
    off = _run(sc, False, ref, spp=spp, motion=motion)
    assert P.compare(ref, on) == []
    assert P.compare(ref, off) == []
    valid = [int((s["res"][0][:, 0] > 0.1).sum()) for s in ref["snapshots"]]
    assert [s["final_rays"] for s in off["snapshots"]] == valid                   # the reference casts one ray per valid sample
    unknown = [v - k for v, k in zip(valid, ref["known_final_rays"])]
    assert [s["final_rays"] for s in on["snapshots"]] == unknown                  # the product only the unknown ones
    assert unknown[0] == 0                                                        # first iteration: every sample has passed its ray
    if shift is None:
        assert sum(unknown) <= 1e-3 * sum(valid)                                  # history comes from the same pixel
    else:
        assert 0 < sum(unknown) < sum(valid)                                      # history from another pixel: not known


def test_history_from_another_pixel_can_be_occluded():
    """the case the tags must NOT cover: a history sample taken from another pixel has never been tested from here, and
    some of them are occluded (vis = 0) -- the scene / shift are chosen so that this happens"""
    sc = P.scene("T2", 0.0)
    ref = P.oracle_run(sc, spp=6, motion=_motion(sc, 7, 5))
    occluded = sum(int((s["vis"] == 0).sum()) for s in ref["snapshots"])
    assert occluded > 0 and ref["provenance_violations"] == 0
    on = _run(sc, True, ref, spp=6, motion=_motion(sc, 7, 5))
    assert P.compare(ref, on) == []


def test_no_tag_outlives_its_loop():
    """the caller's reservoir sets carry no tag after the loop, also when the loop dies half way"""
    import torch
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    sc = P.scene("T0", 0.0)
    ref = P.oracle_run(sc, spp=2)
    W, Hh = sc["W"], sc["H"]
    mods = R.load_m_for_restir(W, Hh, device=torch.device("cpu"), max_bounce=2)
    res, prev = mods[11], mods[12]  # reservoirs, prev_reservoirs of the 17-tuple
    assert slangpy_shim.vis_tag(res) is None

    def run(hook):
        g = {k: H.t(v) for k, v in sc["gbuffer"].items()}
        with torch.no_grad():
            R.restir_di_with_pt(False, 1, 1, 1, None, _worker(sc), 2, W, Hh, *mods[:7], *mods[8:], H.t(sc["env"]),
                                g["occ_map"], g["pos_map"], g["normal_map"], g["depth_map"], g["diffuse_map"],
                                g["roughness_specular"], H.t(ref["prepared"]["ray_dir_map"]), None, None, None, None, None,
                                None, random_offset=1, max_bounce=2, hooks=hook)

    seen = []

    def boom(kind, i, d):
        seen.append(slangpy_shim.vis_tag(d["reservoirs"]) is not None if kind == "direct" else None)
        raise RuntimeError("stop here")

    with pytest.raises(RuntimeError):
        run(boom)
    assert seen == [True]                                  # the tags were on inside the loop ...
    assert slangpy_shim.vis_tag(res) is None and slangpy_shim.vis_tag(prev) is None   # ... and are gone after it


@pytest.mark.parametrize("name,metallic", [("T2", 0.0), ("C1", 0.4)])
def test_roofline_charges_exactly_the_rays_the_spatial_pass_casts(name, metallic):
    """bench.py charges the spatial pass with the oracle's node / triangle counts WITHOUT the rays the oracle classifies as
    unable to reach the output (orc_kernels.cpp, counters 9..11).  That classification must be the product's: one spatial
    pass launched on the oracle's own inputs queues exactly (rays - dead) rays, and produces the oracle's reservoirs."""
    import numpy as np
    import torch
    from oracle import oracle as O, driver as D
    from mirres_restir_nerf_mesh_b200 import slangpy_shim, synth
    sc = P.scene(name, metallic)
    per, snaps = {}, []
    D.run_no_denoise(sc["bvh"], sc["env"], sc["gbuffer"], 1, sc["W"], sc["H"], 4242, lambda p: synth.material(p, metallic),
                     max_bounce=1, counters=per, snapshots=snaps)
    c = per["spatial_resampling"]
    rays, dead = int(c[7]), int(c[11])
    assert 0 < dead < rays
    g = D.prepare_gbuffer(sc["gbuffer"])
    W, Hh, n = sc["W"], sc["H"], sc["W"] * sc["H"]
    env = np.ascontiguousarray(sc["env"][::-1].reshape(-1, 3), np.float32)
    offs = (O.neighbor_offsets(8192).reshape(-1, 2) / np.float32(127)).astype(np.float32)
    k, w = H.kernels(), _worker(sc)
    ws = slangpy_shim.workspace("cpu", n)
    k.workspace_prepare(H.t(g["occ_map"]), ws)
    prev = [H.t(a.copy()) for a in snaps[0]["prev"]]
    res = [torch.zeros_like(a) for a in prev]
    k.spatial_resampling(w.packed, H.t(g["pos_map"]), res, prev, H.t(offs), H.t(env), sc["env"].shape[1], sc["env"].shape[0],
                         W, Hh, 4242 + 3, H.t(g["occ_map"]), H.t(g["normal_depth"]), H.t(g["brdf_map"]),
                         H.t(g["ray_dir_map"]), ws)
    assert int(ws[:16].view(torch.int32)[2]) == rays - dead
    for a, b in zip(res, snaps[0]["res"]):
        assert np.array_equal(a.numpy().reshape(-1), b.reshape(-1))
