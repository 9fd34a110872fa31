"""Manual multi-GPU check of the rendering partition (BASELINE config 3 shape): the C2 scene at 800 x 800, forward only,
cut into row bands of equal foreground-pixel count over the ranks (dist.RowBandShard: every rank runs the spp loop on the
slice of the maps that holds its band and halo, receives 31 halo rows of the reservoirs point to point after every spatial
pass over NCCL; the accumulated images are gathered before the denoiser).  Rank 0 also renders the whole frame alone and
compares: all six output images must be bit-identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_rows_sharded.py [spp] [C3|C5]
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R, slangpy_shim, dist as D  # noqa: E402


def main(spp, cfg_name="C3"):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS[cfg_name]
    W, H, mb = cfg["W"], cfg["H"], cfg["max_bounce"]
    n = W * H
    v, f = synth.make_mesh(cfg)
    env = torch.from_numpy(synth.envmap(*cfg["env"])).to(dev)
    vert, tri = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    pose = torch.from_numpy(synth.camera_pose(view=3)).to(dev)
    worker = R.restirbvhWorker(vert, tri)
    k = slangpy_shim.get_kernels()
    mat = synth.ProceduralMaterial(0.0)

    def render(shard):
        mods = R.load_m_for_restir(W, H, device=dev, max_bounce=mb)
        worker.update_mesh(vert, tri)
        ro, rd = synth.camera_rays_torch(W, H, pose)
        occ, depth = torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
        pos, nrm = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
        k.gbuffer_primary(worker.packed, ro, rd, occ, pos, nrm, depth, ws=slangpy_shim.workspace(dev, n))
        kd, rs = mat.gbuffer_materials(pos, occ)
        return R.run_restir_di_with_pt(False, 1, 1, 1, mat, None, worker, *mods, env, occ, nrm, depth, kd, rs, rd, pos, None,
                                       None, None, None, W, H, spp, 2, 2, 2.0, 0.1, 0.001, random_offset=5, max_bounce=mb,
                                       shard=shard)

    with torch.no_grad():
        ro, rd = synth.camera_rays_torch(W, H, pose)
        worker.update_mesh(vert, tri)
        occ0 = torch.empty(n, 1, device=dev)
        k.gbuffer_primary(worker.packed, ro, rd, occ0, torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev),
                          torch.empty(n, 1, device=dev), ws=slangpy_shim.workspace(dev, n))
        shard = D.RowBandShard(W, H, bounds=D.balanced_bounds(occ0, W, H, world))
        for it in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            outs = render(shard)  # full-frame on every rank
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            torch.cuda.synchronize()
            e0.record()
            want = render(None)
            e1.record()
            torch.cuda.synchronize()
            same = all(bool(torch.equal(a, b)) for a, b in zip(outs, want))
            print(cfg_name + " row bands over %d GPUs: %dx%d spp %d forward: %.1f ms (max over ranks, incl. reservoir exchange), "
                  "%.3e samples/s; one GPU alone: %.1f ms; bands %s; all six images bit-identical: %s"
                  % (world, W, H, spp, float(ms), n * spp / (float(ms) * 1e-3), e0.elapsed_time(e1), shard.bounds, same))
            assert same
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16, sys.argv[2] if len(sys.argv) > 2 else "C3")
