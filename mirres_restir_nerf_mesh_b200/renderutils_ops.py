"""Host-side mirror of nerf/renderutils/ops.py for the two operators of that plugin the path touches.

  prepare_shading_normal(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading=True,
                         opengl=True)                      ops.py:129-163 -> normal.cu:95-178, called at renderer.py:1013
  bilateral_denoiser / bilateral_denoiser_no_di            ops.py:173-212 (implemented in renderer_restir.py, re-exported)

Same names, argument order, broadcasting rules and autograd contract as the reference: tensors are
[minibatch, height, width, 3] "or broadcastable equivalent"; the backward hands full-resolution gradients back for
every input, as the plugin does.  There is no `use_python` branch: the product has no second code path.
"""
import torch

from .renderer_restir import bilateral_denoiser, bilateral_denoiser_no_di  # noqa: F401
from .slangpy_shim import get_kernels


def _out_shape(tensors):
    shp = [1, 1, 1]
    for t in tensors:
        if t.dim() != 4 or t.shape[-1] != 3:
            raise ValueError("prepare_shading_normal: tensors must be [minibatch, height, width, 3] or broadcastable")
        for d in range(3):
            if t.shape[d] != 1:
                if shp[d] not in (1, t.shape[d]):
                    raise ValueError("prepare_shading_normal: shapes do not broadcast")
                shp[d] = t.shape[d]
    return shp


def _rows(t, shp):
    """[n,3] rows of a tensor that is either full-size or a single broadcast row; partial broadcasts are expanded."""
    if t.numel() == 3:
        return t.reshape(1, 3)
    if list(t.shape[:3]) != shp:
        t = t.expand(*shp, 3)
    return t.reshape(-1, 3)


class _prepare_shading_normal_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading, opengl):
        ctx.two_sided_shading, ctx.opengl = two_sided_shading, opengl
        tensors = (pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm)
        shp = _out_shape(tensors)
        n = shp[0] * shp[1] * shp[2]
        out = torch.empty(n, 3, dtype=torch.float32, device=pos.device)
        get_kernels().shading_normal_fwd(n, [_rows(t.detach().float(), shp) for t in tensors], two_sided_shading, opengl, out)
        ctx.save_for_backward(*tensors)
        ctx.shp = shp
        return out.view(*shp, 3)

    @staticmethod
    def backward(ctx, dout):
        tensors = ctx.saved_tensors
        shp = ctx.shp
        n = shp[0] * shp[1] * shp[2]
        grads = [torch.empty(n, 3, dtype=torch.float32, device=dout.device) if ctx.needs_input_grad[k] else None
                 for k in range(6)]
        get_kernels().shading_normal_bwd(n, [_rows(t.detach().float(), shp) for t in tensors], ctx.two_sided_shading,
                                         ctx.opengl, dout.contiguous().view(n, 3), grads)
        out = []
        for k, g in enumerate(grads):
            if g is None:
                out.append(None)
                continue
            g = g.view(*shp, 3)
            # the plugin returns full-resolution gradients; autograd needs the input's shape, so reduce broadcast axes
            t = tensors[k]
            dims = [d for d in range(3) if t.shape[d] == 1 and shp[d] != 1]
            out.append(g.sum(dim=dims, keepdim=True) if dims else g)
        return tuple(out) + (None, None)


def prepare_shading_normal(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading=True,
                           opengl=True):
    """nerf/renderutils/ops.py:129-163 (final shading normal: tangent space, two-sided flip, normal-map perturbation,
    back-facing normals bent towards the camera)."""
    if perturbed_nrm is None:
        # (0, 0, 1) built with device-side fills: a host->device copy could not be captured into a CUDA graph
        perturbed_nrm = torch.zeros(1, 1, 1, 3, dtype=torch.float32, device=pos.device)
        perturbed_nrm[..., 2] = 1.0
    out = _prepare_shading_normal_func.apply(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm,
                                             two_sided_shading, opengl)
    if torch.is_anomaly_enabled():
        assert torch.all(torch.isfinite(out)), "Output of prepare_shading_normal contains inf or NaN"
    return out
