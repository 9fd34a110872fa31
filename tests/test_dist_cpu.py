"""Multi-process tests of the multi-GPU host logic on CPU (gloo, world sizes 2, 3 and 4): row-band rendering of a frame
(every rank runs the spp loop on the slice of the maps that holds its band and halo, receives the halo rows of the
reservoirs point to point after every spatial pass, and the accumulated images are gathered before the denoiser) must
reproduce the single-process images bit for bit; the gradient all-reduce sums the flat buffer.  At world size 4 the
bands (24 rows) are narrower than the halo, so a rank talks to more than its two neighbours; at world size 3 the bands
are cut by foreground-pixel count (unequal heights)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
SPP, DENOISE_ITER, STEP, PHI = 3, 2, 2, (2.0, 0.1, 0.001)


def _render(sc, shard, overlap=None):
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth
    W, Hh = sc["W"], sc["H"]
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    mods = R.load_m_for_restir(W, Hh, device="cpu")
    g = {k: H.t(v) for k, v in sc["gbuffer"].items()}
    with torch.no_grad():
        return R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, w, *mods,
                                       H.t(sc["env"]), g["occ_map"], g["normal_map"], g["depth_map"], g["diffuse_map"],
                                       g["roughness_specular"], g["ray_dir_map"], g["pos_map"], None, None, None, None,
                                       W, Hh, SPP, DENOISE_ITER, STEP, *PHI, random_offset=777, shard=shard, overlap=overlap)


def _worker(rank, world, port, out_dir, overlap=None, balanced=False):
    for p in (os.path.dirname(HERE), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hostcheck as H
    import parity as P
    from mirres_restir_nerf_mesh_b200 import dist as D
    H.activate()
    sc = P.scene("T2", 0.4)
    bounds = D.balanced_bounds(torch.from_numpy(sc["gbuffer"]["occ_map"]), sc["W"], sc["H"], world) if balanced else None
    shard = D.RowBandShard(sc["W"], sc["H"], bounds=bounds)
    full = _render(sc, shard, overlap)  # full-frame on every rank
    flat = torch.full((5,), float(rank + 1))
    D.allreduce_gradients(flat)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), *[f.numpy() for f in full], flat=flat.numpy(),
                 active=np.array(shard.active), bounds=np.array(shard.bounds), halo_bytes=shard.halo_bytes())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,overlap,balanced", [(2, None, False), (4, None, False), (2, True, False), (3, True, True)])
def test_row_band_sharding_is_bit_identical(tmp_path, world, overlap, balanced):
    """overlap=True drives the concurrent schedule's host logic (exchange on the reuse chain) over CPU tensors."""
    import hostcheck as H
    import parity as P
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    port = 29500 + (os.getpid() + 7 * world + (3 if overlap else 0)) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), overlap, balanced), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    H.activate()
    try:
        sc = P.scene("T2", 0.4)
        want = _render(sc, None)
    finally:
        slangpy_shim.set_kernels(None)
    if balanced:
        b = got["bounds"]
        fg = sc["gbuffer"]["occ_map"].reshape(sc["H"], sc["W"]).sum(1)
        per_band = [fg[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(per_band) < 1.25 * fg.sum() / world and len(set(np.diff(b))) > 1  # equal work, unequal heights
    else:
        assert tuple(got["active"]) == (0, sc["H"] // world + 31)
    # what rank 0 receives per exchange: the rows below its band, 24 bytes per pixel, not whole bands
    assert int(got["halo_bytes"]) == (got["active"][1] - got["bounds"][1]) * sc["W"] * 24
    for i, w in enumerate(want):
        assert np.array_equal(got["arr_%d" % i], w.numpy(), equal_nan=True), i
    assert (got["flat"] == float(sum(range(1, world + 1)))).all()


def test_temporal_pass_on_a_row_slice_rounds_like_the_full_frame():
    """A band of a TALL frame handed to the temporal pass as a frame of its own (with the row offset) must pick the same
    previous pixel as the full-frame launch: `int(pixel + u)` is rounded at the magnitude of the row number (SURVEY.md
    quirk 8), so a pass that jittered band-local row numbers would differ in a few pixels per million -- too few for the
    small frames of the multi-process tests above, enough to break bit-identity at 2048 rows (seen on 2 B200s)."""
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import slangpy_shim, synth
    k = H.activate()
    try:
        fx, fy, y0, y1 = 32, 70000, 65500, 65600   # rows near 2^16: float spacing 2^-7, one jitter in ~128 rounds up
        n = fx * fy
        g = torch.Generator().manual_seed(1)
        env = H.t(np.ascontiguousarray(synth.envmap(16, 32)[::-1].reshape(-1, 3)))
        rnd = lambda *s: torch.rand(*s, generator=g)
        nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
        maps = dict(occ=torch.ones(n, 1), nd=torch.cat((nrm, 2.0 + 0.01 * rnd(n, 1)), 1).contiguous(), brdf=torch.cat((rnd(n, 1), torch.zeros(n, 1), 0.1 + rnd(n, 1)), 1).contiguous(),
                    ray=-nrm.clone())

        def reservoirs():
            ld = torch.cat((torch.ones(n, 1), rnd(n, 2)), 1).contiguous()
            return [ld, rnd(n, 1), torch.randint(1, 5, (n, 1), generator=g, dtype=torch.int32), rnd(n, 1)]
        cur, prev = reservoirs(), reservoirs()

        def run(rows):
            a, b = rows
            cut = lambda t: t[a * fx:b * fx].clone()
            res = [cut(t) for t in cur]
            m = {k_: cut(v) for k_, v in maps.items()}
            with slangpy_shim.row_offset(a), slangpy_shim.active_rows(y0 - a, y1 - a, fx):
                ws = slangpy_shim.workspace(torch.device("cpu"), (b - a) * fx)
                slangpy_shim.prepare_workspace(m["occ"])
                k.temporal_resampling(res, [cut(t) for t in prev], env, 32, 16, fx, b - a, 77, m["occ"], m["nd"], m["brdf"], m["ray"],
                                      m["occ"], m["nd"], m["brdf"], m["ray"], ws)
            return [t[(y0 - a) * fx:(y1 - a) * fx] for t in res]

        full, band = run((0, fy)), run((y0 - 31, y1 + 31))
        moved = sum(int((f != c[y0 * fx:y1 * fx]).any(dim=1).sum()) for f, c in zip(full[:1], cur[:1]))
        assert moved > 100  # the pass did something
        for f, b_ in zip(full, band):
            assert torch.equal(f, b_)
    finally:
        slangpy_shim.set_kernels(None)


@pytest.mark.parametrize("world,balanced", [(2, False), (3, True)])
def test_bands_as_threads_of_one_process(world, balanced):
    """The virtual ranks as host threads (tests/band_threads.py) over the host-check kernels: per-thread workspaces and launch
    context, slices, row offset and band words; same bit-identity as the multi-process runs."""
    import band_threads as BT
    import hostcheck as H
    import parity as P
    from mirres_restir_nerf_mesh_b200 import dist as D, slangpy_shim
    H.activate()
    try:
        sc = P.scene("T2", 0.4)
        w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
        w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
        want = BT.render(sc, w, "cpu")
        bounds = D.balanced_bounds(torch.from_numpy(sc["gbuffer"]["occ_map"]), sc["W"], sc["H"], world) if balanced \
            else D.uniform_bounds(sc["H"], world)
        outs, errors = BT.render_in_threads(sc, w, "cpu", world, bounds)
        assert not errors, errors
        for r in range(world):
            for a, b in zip(outs[r], want):
                assert torch.equal(a, b) or bool(((a == b) | (a.isnan() & b.isnan())).all())
    finally:
        slangpy_shim.set_kernels(None)


def test_band_boundaries():
    """balanced_bounds / rebalanced_bounds / RowBandShard plans on a synthetic occupancy: equal cost per band, rows of pure
    background outside every band, point-to-point plans symmetric (what a rank sends is what its peer expects)."""
    from mirres_restir_nerf_mesh_b200 import dist as D
    fx, fy = 64, 200
    occ = torch.zeros(fy, fx)
    occ[40:160, 10:50] = 1.0
    occ[90:110, :] = 1.0  # a dense stripe: bands must get thinner there
    b = D.balanced_bounds(occ.reshape(-1, 1), fx, fy, 4)
    assert b[0] == 40 and b[-1] == 160 and all(y1 > y0 for y0, y1 in zip(b, b[1:]))
    cost = lambda y0, y1: float(occ[y0:y1].sum()) + D.AREA_WEIGHT * fx * (y1 - y0)
    costs = [cost(y0, y1) for y0, y1 in zip(b, b[1:])]
    assert max(costs) < 1.15 * sum(costs) / 4
    assert min(y1 - y0 for y0, y1 in zip(b, b[1:])) < max(y1 - y0 for y0, y1 in zip(b, b[1:]))
    assert D.balanced_bounds(occ.reshape(-1, 1), fx, fy, 4, clip=False)[0] == 0
    assert D.balanced_bounds(torch.zeros(fy * fx, 1), fx, fy, 4) == D.uniform_bounds(fy, 4)
    # measured cost says band 0 is twice as expensive per row: it must shrink, the total range must stay
    r = D.rebalanced_bounds(b, [2.0, 1.0, 1.0, 1.0])
    assert r[0] == b[0] and r[-1] == b[-1] and r[1] - r[0] < b[1] - b[0]
    shards = [D.RowBandShard(fx, fy, rank=q, world=4, bounds=b) for q in range(4)]
    for s in shards:
        assert s.first_row <= s.active[0] <= s.wide[0] <= s.rows[0] and s.rows[1] <= s.wide[1] <= s.active[1] <= s.last_row
        for q, rows in s.send_plan:
            assert (s.rank, rows) in shards[q].recv_plan
        for q, rows in s.recv_plan:
            assert (s.rank, rows) in shards[q].send_plan
        assert s.halo_bytes() == sum(y1 - y0 for _, (y0, y1) in s.recv_plan) * fx * 24


# ---- view-parallel training (BASELINE config 4): one view per rank, the flat gradient buffer summed -------------
def _view_gradients(view):
    """forward + backward of one view of the small scene through the product driver (host flavour): the envmap gradient
    and the per-pixel gradients scattered to the vertices, as bench.py's step lays them out in ONE flat buffer"""
    import hostcheck as H
    import parity as P
    from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth
    sc = P.scene("T0", 0.3, view=view)
    W, Hh = sc["W"], sc["H"]
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    mods = R.load_m_for_restir(W, Hh, device="cpu")
    g = {k: H.t(v) for k, v in sc["gbuffer"].items()}
    env = H.t(sc["env"]).requires_grad_(True)
    normal = g["normal_map"].clone().requires_grad_(True)
    outs = R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, w, *mods, env,
                                   g["occ_map"], normal, g["depth_map"], g["diffuse_map"], g["roughness_specular"],
                                   g["ray_dir_map"], g["pos_map"], None, None, None, None, W, Hh, 2, DENOISE_ITER, STEP, *PHI,
                                   random_offset=31 + 17 * view)
    torch.nn.functional.mse_loss(outs[0], torch.full_like(outs[0], 0.5)).backward()
    # per-pixel normal gradients land on the vertices of the triangle the pixel sees (here: its first vertex)
    V = sc["vert"].shape[0]
    vgrad = torch.zeros(V, 3)
    hit = torch.from_numpy(sc["hit"] > 0)
    first_vertex = torch.from_numpy(sc["tri"][np.maximum(sc["prim"], 0), 0].astype(np.int64))
    vgrad.index_add_(0, first_vertex[hit], normal.grad[hit])
    return torch.cat((env.grad.reshape(-1), vgrad.reshape(-1)))


def _train_worker(rank, world, port, out_dir):
    for p in (os.path.dirname(HERE), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import dist as D
    H.activate()
    flat = _view_gradients(view=7 * rank)
    ne = flat.numel() // 2  # (any split point will do: the two parts are independent all-reduces)
    # the step's collective in the two parts bench.py issues: the envmap segment first, the vertex segments after it
    D.allreduce_gradients(flat[:ne])
    D.allreduce_gradients(flat[ne:])
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
    dist.destroy_process_group()


def test_view_parallel_training_step_sums_the_gradients_of_all_views(tmp_path):
    """world 2 over gloo: each rank renders ITS view forward + backward, the flat gradient buffer is all-reduced in two
    parts; the result equals the sum of the two views' buffers computed in one process (to fp32 addition order: two
    addends, so exactly)."""
    world, port = 2, 29731 + os.getpid() % 200
    mp.spawn(_train_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "flat.npy"))
    import hostcheck as H
    from mirres_restir_nerf_mesh_b200 import slangpy_shim
    H.activate()
    try:
        want = (_view_gradients(0) + _view_gradients(7)).numpy()
    finally:
        slangpy_shim.set_kernels(None)
    assert np.abs(want).sum() > 0
    assert np.array_equal(got, want)
