"""Multi-GPU partitioning of the path: one process per GPU, torch.distributed for the plumbing.

Training (BASELINE config 4): data parallel over views -- every rank renders its own view and the per-step gradients
(envmap, vertex normals, vertex texture) are summed with ONE all-reduce of a flat buffer (`allreduce_gradients`).

Rendering (configs 3 and 5): the frame is cut into contiguous ROW BANDS, one per rank.  Pixels are independent except
that spatial reuse reads the post-temporal reservoirs of neighbours within 30 px (GATHER_RADIUS,
nerf/renderer_restir.py:176) and temporal reuse reads the previous iteration's reservoir within 1 px.  Per spp iteration
a rank therefore runs
    initial candidates + temporal reuse      on its band + 30 rows on either side   (`wide` rows)
    spatial reuse, visibility, shading,
    the indirect paths                       on its band only                       (`rows`)
and after every spatial pass it receives the 31 rows above and below its band from the ranks that own them -- 24 bytes
per pixel, 0.6 MB per boundary at 800 px width -- point to point, nothing else travels inside the loop.
(Recomputing the halo locally is not enough: temporal reuse makes the dependency cone grow by 30 px per iteration.)
The `[N, k]` maps are row-major, so the rows a rank touches are one contiguous slice of every tensor: the spp loop runs
on the slices as on a frame of their own, and the workspace's row-offset word keeps the random streams (keyed on pixel
coordinates) those of the full frame, so every reservoir a rank computes or receives is bit-identical to the single-GPU
one.  After the loop the
six accumulated images of the bands are all-gathered and denoised / composited at full frame on every rank, exactly as
on one GPU.  The BVH, the envmap distribution and the light tiles are rebuilt identically on every rank.

Band boundaries need not be uniform: `balanced_bounds` cuts the frame so that every band holds the same number of
FOREGROUND pixels (background pixels leave every kernel after the compaction, so rows of sky cost nothing).
"""
import torch
import torch.distributed as dist

REUSE_ROWS = 30                 # GATHER_RADIUS: rows of post-temporal reservoirs spatial reuse reads beyond the band
HALO_ROWS = REUSE_ROWS + 1      # + 1 px jitter of temporal reuse: rows of finished reservoirs a rank needs from its neighbours


def uniform_bounds(framedim_y, world):
    """world + 1 row indices, bands of (almost) equal height."""
    return [(framedim_y * r) // world for r in range(world + 1)]


AREA_WEIGHT = 0.026  # cost of one pixel of a band's rows relative to one foreground pixel (per-pixel streams: zero fills,
                     # path-state prologues, copies; measured on B200 at 2048 x 2048: 0.17 ns against 6.5 ns per iteration)


def balanced_bounds(occ_map, framedim_x, framedim_y, world, min_rows=1, area_weight=AREA_WEIGHT, clip=True):
    """world + 1 row indices that cut the frame into bands of about equal COST: foreground pixels (occ > 0.5) plus
    `area_weight` per pixel of the band's rows.  With `clip` the first band starts at the first row that holds foreground
    and the last band ends behind the last such row: rows of pure background belong to no band (nothing is computed for
    them; their accumulated images are zero, which is what the spp loop produces there).  One host synchronisation; every
    rank derives the same boundaries from the same full-frame occupancy."""
    per_row = (occ_map.reshape(framedim_y, framedim_x) > 0.5).sum(dim=1).to(torch.float64).cpu()
    total_fg = float(per_row.sum())
    if total_fg <= 0:
        return uniform_bounds(framedim_y, world)
    rows = torch.nonzero(per_row > 0).reshape(-1)
    y_first, y_last = (int(rows[0]), int(rows[-1]) + 1) if clip else (0, framedim_y)
    if y_last - y_first < world * min_rows:
        y_first, y_last = 0, framedim_y
    cost = per_row[y_first:y_last] + area_weight * framedim_x
    cum = torch.cumsum(cost, 0)
    total = float(cum[-1])
    bounds = [y_first]
    for r in range(1, world):
        y = y_first + int(torch.searchsorted(cum, torch.tensor(total * r / world, dtype=torch.float64)).item()) + 1
        y = max(y, bounds[-1] + min_rows)
        y = min(y, y_last - (world - r) * min_rows)
        bounds.append(y)
    bounds.append(y_last)
    return bounds


def rebalanced_bounds(bounds, seconds, min_rows=1):
    """New boundaries from the time every rank needed for ITS band (any unit): the cost of a row is taken as constant
    inside a band, and the rows [bounds[0], bounds[-1]) are cut again into pieces of equal cost.  Foreground counts do
    not see that some regions cast longer rays than others; one or two rounds of this do."""
    world = len(bounds) - 1
    dens = []
    for r in range(world):
        rows = max(bounds[r + 1] - bounds[r], 1)
        dens += [float(seconds[r]) / rows] * (bounds[r + 1] - bounds[r])
    cum = torch.cumsum(torch.tensor(dens, dtype=torch.float64), 0)
    total = float(cum[-1])
    out = [bounds[0]]
    for r in range(1, world):
        y = bounds[0] + int(torch.searchsorted(cum, torch.tensor(total * r / world, dtype=torch.float64)).item()) + 1
        y = max(y, out[-1] + min_rows)
        y = min(y, bounds[-1] - (world - r) * min_rows)
        out.append(y)
    out.append(bounds[-1])
    return out


class RowBandShard:
    def __init__(self, framedim_x, framedim_y, rank=None, world=None, group=None, halo=HALO_ROWS, bounds=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.fx, self.fy = int(framedim_x), int(framedim_y)
        self.bounds = [int(b) for b in (bounds if bounds is not None else uniform_bounds(self.fy, self.world))]
        # the bands are consecutive; rows before the first and behind the last one belong to no rank (pure background,
        # see balanced_bounds): nothing is computed for them and gather_bands returns zeros there
        if len(self.bounds) != self.world + 1 or self.bounds[0] < 0 or self.bounds[-1] > self.fy or \
                any(b1 < b0 for b0, b1 in zip(self.bounds, self.bounds[1:])):
            raise ValueError("band boundaries %r do not fit %d rows over %d ranks" % (self.bounds, self.fy, self.world))
        self.first_row, self.last_row = self.bounds[0], self.bounds[-1]
        self.halo = int(halo)
        self.rows = self.band(self.rank)                                                   # rows this rank owns
        # (clipped to the rows that belong to some band: beyond them there is no foreground, hence nothing to read)
        self.wide = (max(self.rows[0] - (halo - 1), self.first_row), min(self.rows[1] + (halo - 1), self.last_row))  # initial + temporal
        self.active = (max(self.rows[0] - halo, self.first_row), min(self.rows[1] + halo, self.last_row))  # rows whose reservoirs it reads
        # point-to-point plan: (peer, rows) -- what this rank needs from a peer = the peer's band cut with its halo
        self.recv_plan = [(q, self._cut(self.band(q), self.active)) for q in range(self.world) if q != self.rank]
        self.recv_plan = [(q, c) for q, c in self.recv_plan if c is not None]
        self.send_plan = [(q, self._cut(self.rows, self._active_of(q))) for q in range(self.world) if q != self.rank]
        self.send_plan = [(q, c) for q, c in self.send_plan if c is not None]
        # the collective is chosen once, from the backend (gloo has no flat all-gather), never by catching a failed call
        self._flat_gather = self.world > 1 and dist.is_initialized() and dist.get_backend(group) != "gloo"

    def band(self, r):
        return (self.bounds[r], self.bounds[r + 1])

    def _active_of(self, r):
        y0, y1 = self.band(r)
        return (max(y0 - self.halo, self.first_row), min(y1 + self.halo, self.last_row))

    @staticmethod
    def _cut(a, b):
        y0, y1 = max(a[0], b[0]), min(a[1], b[1])
        return (y0, y1) if y1 > y0 else None

    def halo_bytes(self, row_bytes=24):
        """Bytes this rank receives per exchange (row_bytes per pixel: the four reservoir tensors)."""
        return sum((c[1] - c[0]) for _, c in self.recv_plan) * self.fx * row_bytes

    def _peer(self, q):
        return q if self.group is None else dist.get_global_rank(self.group, q)

    def _row_views(self, tensors, row0):
        # [rows, fx * k] views of [rows * fx, k] tensors that start at frame row `row0`, as 32-bit words
        views = [t.view(t.shape[0] // self.fx, -1) for t in tensors]
        return [v if v.dtype == torch.float32 else v.view(torch.float32) for v in views]

    def exchange(self, tensors, row0=0):
        """In place: the halo rows of every [rows * fx, k] tensor (rows of the frame from `row0` on; the whole frame by
        default) take their owners' values.  A band of rows is a contiguous piece of each tensor, so the rows travel
        straight from and into the tensors -- no staging copies: per exchange ONE group of point-to-point operations,
        and a rank talks only to the ranks whose bands touch its halo."""
        tensors = list(tensors)
        if self.world == 1 or not (self.recv_plan or self.send_plan):
            return
        views = [t.view(t.shape[0] // self.fx, -1) for t in tensors]  # [rows, fx * k], any 32-bit dtype
        ops = []
        for q, (y0, y1) in self.send_plan:
            ops += [dist.P2POp(dist.isend, v[y0 - row0:y1 - row0], self._peer(q), self.group) for v in views]
        for q, (y0, y1) in self.recv_plan:
            ops += [dist.P2POp(dist.irecv, v[y0 - row0:y1 - row0], self._peer(q), self.group) for v in views]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def gather_bands(self, images, row0=0):
        """Full-frame copies of fp32 images of which every rank holds its own band (tensors [rows * fx, k] that start at
        frame row `row0`): one all-gather for all of them, bands padded to the tallest."""
        images = list(images)
        views = self._row_views(images, row0)
        y0, y1 = self.rows
        if self.world == 1:
            full = [torch.zeros((self.fy, v.shape[1]), dtype=torch.float32, device=v.device) for v in views]
            for f, v in zip(full, views):
                f[y0:y1] = v[y0 - row0:y1 - row0]
            return [f.view(-1, im.shape[1]) for f, im in zip(full, images)]
        width = sum(v.shape[1] for v in views)
        tall = max(b1 - b0 for b0, b1 in zip(self.bounds, self.bounds[1:]))
        mine = torch.zeros((tall, width), dtype=torch.float32, device=views[0].device)
        mine[:y1 - y0] = torch.cat([v[y0 - row0:y1 - row0] for v in views], dim=1)
        out = torch.empty((self.world, tall, width), dtype=torch.float32, device=mine.device)
        if self._flat_gather:
            dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=self.group)
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            out = torch.stack(parts)
        full = [torch.zeros((self.fy, v.shape[1]), dtype=torch.float32, device=mine.device) for v in views]
        for r in range(self.world):
            b0, b1 = self.band(r)
            c0 = 0
            for f in full:
                c1 = c0 + f.shape[1]
                f[b0:b1] = out[r, :b1 - b0, c0:c1]
                c0 = c1
        return [f.view(-1, im.shape[1]) for f, im in zip(full, images)]

    def gather_image(self, img):
        """Full-frame [N, k] image assembled from the bands every rank owns."""
        return self.gather_bands([img])[0]

    def view(self):
        """The shard as the spp loop sees it when it is handed the rows `active` of every map as a frame of their own."""
        return RowBandView(self)


class RowBandView:
    """Row numbers relative to the first row of the slice [active[0], active[1]) of the frame."""

    def __init__(self, shard):
        self.shard, self.row0 = shard, shard.active[0]
        rel = lambda r: (r[0] - self.row0, r[1] - self.row0)
        self.rows, self.wide, self.active = rel(shard.rows), rel(shard.wide), rel(shard.active)

    def exchange(self, tensors):
        self.shard.exchange(tensors, self.row0)


def render_rows_sharded(run, shard, *args, **kw):
    """`run` = renderer_restir.run_restir_di_with_pt; its outputs are full-frame on every rank (the bands' accumulated
    images are gathered before the denoiser)."""
    outs = run(*args, shard=shard, **kw)
    return tuple(o.detach() for o in outs)


def allreduce_gradients(flat, group=None):
    """The per-step collective of view-parallel training: sum of the flat gradient buffer over ranks."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    return flat
