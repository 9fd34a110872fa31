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
