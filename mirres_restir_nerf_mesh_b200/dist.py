"""Multi-GPU partitioning of the path: one process per GPU, torch.distributed for the plumbing.

Training (BASELINE config 4): data parallel over views -- every rank renders its own view and the per-step gradients
(envmap, vertex normals, vertex texture) are summed with ONE all-reduce of a flat buffer (`allreduce_gradients`).

Rendering (configs 3 and 5): the frame is cut into contiguous ROW BANDS.  Pixels are independent except that spatial
reuse reads neighbours within 30 px and temporal reuse within 1 px (GATHER_RADIUS, nerf/renderer_restir.py:176), and the
a-trous filter reaches 6 px.  A rank therefore processes its band plus a 31-row halo, and after every spatial pass the
ranks exchange their own rows of the reservoirs, so that the halo rows a rank reads in the next iteration hold their
owners' values (recomputing the halo locally is not enough: temporal reuse makes the dependency cone grow by 30 px
per spp iteration).  RNG streams are keyed on global pixel coordinates (the maps stay full-frame; only the list of
processed pixels is restricted), so the assembled image is bit-identical to the single-GPU one.  The BVH, the envmap
distribution and the light tiles are rebuilt identically on every rank.
"""
import torch
import torch.distributed as dist

HALO_ROWS = 31  # 30 px gather radius + 1 px temporal jitter


class RowBandShard:
    def __init__(self, framedim_x, framedim_y, rank=None, world=None, group=None, halo=HALO_ROWS):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        if framedim_y % self.world:
            raise ValueError("frame height %d is not divisible by %d ranks" % (framedim_y, self.world))
        self.fx, self.fy = int(framedim_x), int(framedim_y)
        band = self.fy // self.world
        self.rows = (self.rank * band, (self.rank + 1) * band)          # rows this rank owns
        self.active = (max(self.rows[0] - halo, 0), min(self.rows[1] + halo, self.fy))  # rows it processes

    def _own(self, t):
        return t.view(self.fy, -1)[self.rows[0]:self.rows[1]].contiguous()

    def exchange(self, tensors):
        """In place: every [N, k] tensor ends up with each row band holding its owner's values.  All tensors travel in ONE
        collective: their own rows are packed side by side as 32-bit words (fp32 and int32 alike) and unpacked after the
        all-gather -- per spp iteration that is one NCCL call and five small copies instead of four collectives with
        their staging."""
        tensors = list(tensors)
        band = self.rows[1] - self.rows[0]
        own = [t.view(self.fy, -1)[self.rows[0]:self.rows[1]] for t in tensors]
        packed = torch.cat([o.view(torch.float32) for o in own], dim=1).contiguous()
        width = packed.shape[1]
        out = torch.empty((self.world, band, width), dtype=torch.float32, device=packed.device)
        try:
            dist.all_gather_into_tensor(out.view(-1), packed.view(-1), group=self.group)
        except (RuntimeError, NotImplementedError, AttributeError):
            parts = [torch.empty_like(packed) for _ in range(self.world)]
            dist.all_gather(parts, packed, group=self.group)
            out = torch.stack(parts)
        c0 = 0
        for t, o in zip(tensors, own):
            c1 = c0 + o.shape[1]
            t.view(self.world, band, o.shape[1]).view(torch.float32).copy_(out[:, :, c0:c1])
            c0 = c1

    def gather_image(self, img):
        """Full-frame [N, k] image assembled from the bands every rank owns."""
        out = img.clone()
        self.exchange([out])
        return out


def render_rows_sharded(run, shard, *args, **kw):
    """`run` = renderer_restir.run_restir_di_with_pt; returns its outputs assembled over all ranks."""
    outs = run(*args, shard=shard, **kw)
    return tuple(shard.gather_image(o.detach()) for o in outs)


def allreduce_gradients(flat, group=None):
    """The per-step collective of view-parallel training: sum of the flat gradient buffer over ranks."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    return flat
