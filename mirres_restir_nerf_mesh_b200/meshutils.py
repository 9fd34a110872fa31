"""Host-side mirror of the one function of the reference's top-level meshutils.py the G-buffer stage calls.

  auto_normals(v_pos, t_pos_idx) -> (v_nrm, t_pos_idx)        meshutils.py:14-39, called from nerf/renderer.py:979-1030

Same name, arguments and return value; differentiable with respect to v_pos (the reference differentiates the torch
expression; here the backward is mirres_vertex_normals_bwd)."""
import torch

from .slangpy_shim import get_kernels


class _auto_normals_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_pos, tri):
        v = v_pos.detach().float().contiguous()
        vsum, vnrm = torch.empty_like(v), torch.empty_like(v)
        get_kernels().vertex_normals_fwd(v, tri, vsum, vnrm)
        ctx.save_for_backward(v, tri, vsum)
        return vnrm

    @staticmethod
    def backward(ctx, g):
        v, tri, vsum = ctx.saved_tensors
        gv = torch.zeros_like(v)
        get_kernels().vertex_normals_bwd(v, tri, vsum, g.contiguous().float(), gv)
        return gv, None


def auto_normals(v_pos, t_pos_idx):
    tri = t_pos_idx if t_pos_idx.dtype == torch.int32 and t_pos_idx.is_contiguous() else t_pos_idx.to(torch.int32).contiguous()
    return _auto_normals_func.apply(v_pos, tri), t_pos_idx
