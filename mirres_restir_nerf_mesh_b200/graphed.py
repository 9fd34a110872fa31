"""CUDA-graph capture of a whole stage-1 step.

`run_restir_di_with_pt` issues several hundred small launches per step (the reference: more, plus host syncs); at 800x800
the GPU work of a step is ~10 ms, which is about what the Python/driver side needs to enqueue it, so the step is
launch-bound and the two-stream overlap of the direct and indirect chains (renderer_restir.restir_di_with_pt) cannot
materialise.  Every entry point of libmirres_b200.so is capturable (no allocation, no synchronisation, stream-ordered
memsets only), so the step -- LBVH rebuild, G-buffer, spp loop, denoise, loss, backward, gradient scatter -- can be
recorded once and replayed:

    step = CapturedStep(fn, dict(vert=vert, tri=tri, env=env, rays_o=o, rays_d=d))
    out = step(vert=new_vert, env=new_env)     # copies into the static inputs, replays, returns the static outputs

`fn(**inputs)` must be shape-static and free of host synchronisation (use `sample_no_di_dense` materials and pass
`random_offset`).  Frame indices are kernel arguments and therefore baked into the graph; `set_frame_offset(k)` writes
a device-resident word that every kernel adds to its frame index (include/mirres_b200.h, "frame offset"), so replay k
of a training run draws the random streams the eager call with `random_offset + k` would draw.
"""
import torch


class CapturedStep:
    def __init__(self, fn, inputs, warmup=2):
        self.fn = fn
        self.static_in = {k: v.clone() if isinstance(v, torch.Tensor) else v for k, v in inputs.items()}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                out = fn(**self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        del out
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(**self.static_in)
        torch.cuda.synchronize()

    def load(self, **inputs):
        """Stream-ordered copies into the static input buffers (host tensors should be pinned)."""
        for k, v in inputs.items():
            self.static_in[k].copy_(v, non_blocking=True)

    def set_frame_offset(self, value):
        """Stream-ordered: the next replay draws the random streams of `random_offset + value`."""
        from . import slangpy_shim
        device = next(v.device for v in self.static_in.values() if isinstance(v, torch.Tensor))
        slangpy_shim.set_frame_offset(device, value)

    def replay(self):
        self.graph.replay()
        return self.static_out

    def __call__(self, **inputs):
        self.load(**inputs)
        return self.replay()
