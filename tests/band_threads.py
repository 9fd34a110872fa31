"""Row-band rendering with the ranks as THREADS of one process (one device): every virtual rank runs the ordinary
run_restir_di_with_pt(shard=...) on its own CUDA stream, the halo exchange and the gather are direct tensor copies between
the threads behind a barrier.  What it checks that the gloo multi-process tests cannot: the CUDA kernels' handling of row
slices (row offset of the random streams, band words of the spatial pass, jittered row of temporal reuse) on the GPU the
driver runs `pytest -m gpu` on -- no NCCL, no second GPU needed -- and that two host threads can drive two renderers at
once (per-thread workspaces, launch context and tuning)."""
import threading

import torch

from mirres_restir_nerf_mesh_b200 import dist as D, renderer_restir as R, synth


class ThreadShard(D.RowBandShard):
    def __init__(self, fx, fy, rank, world, bounds, barrier, board):
        super().__init__(fx, fy, rank=rank, world=world, bounds=bounds)
        self.barrier, self.board = barrier, board

    @staticmethod
    def _sync(t):
        if t.is_cuda:
            torch.cuda.synchronize(t.device)

    def exchange(self, tensors, row0=0):
        tensors = list(tensors)
        self._sync(tensors[0])
        self.board[self.rank] = (tensors, row0)
        self.barrier.wait()
        for q, (y0, y1) in self.recv_plan:
            src, r0 = self.board[q]
            for t, s in zip(tensors, src):
                tv, sv = t.view(t.shape[0] // self.fx, -1), s.view(s.shape[0] // self.fx, -1)
                tv[y0 - row0:y1 - row0].copy_(sv[y0 - r0:y1 - r0])
        self._sync(tensors[0])
        self.barrier.wait()

    def gather_bands(self, images, row0=0):
        images = [im.detach() for im in images]
        self._sync(images[0])
        self.board[self.rank] = (images, row0)
        self.barrier.wait()
        full = [torch.zeros((self.fy, self.fx * im.shape[1]), dtype=torch.float32, device=im.device) for im in images]
        for r in range(self.world):
            src, r0 = self.board[r]
            b0, b1 = self.band(r)
            for f, s in zip(full, src):
                f[b0:b1] = s.view(s.shape[0] // self.fx, -1)[b0 - r0:b1 - r0]
        self._sync(images[0])
        self.barrier.wait()
        return [f.view(-1, im.shape[1]) for f, im in zip(full, images)]


def render(sc, worker, device, shard=None, spp=3, random_offset=777, overlap=None):
    dev = torch.device(device)
    tt = lambda a: torch.from_numpy(a).to(dev)
    W, Hh = sc["W"], sc["H"]
    mods = R.load_m_for_restir(W, Hh, device=dev)
    g = {k: tt(v) for k, v in sc["gbuffer"].items()}
    with torch.no_grad():
        return R.run_restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), None, worker, *mods, tt(sc["env"]),
                                       g["occ_map"], g["normal_map"], g["depth_map"], g["diffuse_map"], g["roughness_specular"],
                                       g["ray_dir_map"], g["pos_map"], None, None, None, None, W, Hh, spp, 2, 2, 2.0, 0.1, 0.001,
                                       random_offset=random_offset, shard=shard, overlap=overlap)


def render_in_threads(sc, worker, device, world, bounds, spp=3, overlap=None):
    """Returns the outputs of virtual rank 0 (full-frame on every rank) and the list of exceptions of all threads."""
    barrier, board = threading.Barrier(world), {}
    outs, errors = [None] * world, []

    def run(r):
        try:
            shard = ThreadShard(sc["W"], sc["H"], r, world, bounds, barrier, board)
            if torch.device(device).type == "cuda":
                with torch.cuda.stream(torch.cuda.Stream(device=device)):
                    outs[r] = render(sc, worker, device, shard, spp=spp, overlap=overlap)
                    torch.cuda.current_stream().synchronize()
            else:
                outs[r] = render(sc, worker, device, shard, spp=spp, overlap=overlap)
        except BaseException as e:  # noqa: BLE001 -- a thread that dies must not leave the others in the barrier
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return outs, errors
