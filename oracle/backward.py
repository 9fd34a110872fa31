"""Backward oracle -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

The reference obtains its two `.bwd` kernels from Slang's automatic differentiation
(nerf/ScreenSpaceReSTIR/Resampling.py:119-143, 179-214; Denoising.py:30-48), which cannot be run here.  This file
restates the differentiable FORWARD functions in float64 PyTorch, vectorised over pixels, and lets torch.autograd
produce the reference gradients the hand-written CUDA backward is compared with:
  final_shading(...)   nerf/ScreenSpaceReSTIR/FinalShading.slang:14-109 (+ utils/brdfDi.slang, utils/helperDi.slang:1-40)
  eval_final_li(...)   nerf/ScreenSpaceReSTIR/EvaluateFinalSamples.slang:129-188 (bilinear env fetch helper.slang:72-99)
  eaw(...)             nerf/ScreenSpaceReSTIR/EAWDenoise.slang:50-174
Branch predicates are evaluated on detached values (autodiff freezes control flow).
"""
import math

import numpy as np
import torch

F64 = torch.float64


def _lum(v):
    return v[:, 0] * 0.212671 + v[:, 1] * 0.715160 + v[:, 2] * 0.072169


def _frame(n):
    sign = torch.where(n[:, 2] > 0, torch.ones_like(n[:, 2]), -torch.ones_like(n[:, 2]))
    a = -1.0 / (sign + n[:, 2])
    b = n[:, 0] * n[:, 1] * a
    fx = torch.stack([1.0 + sign * n[:, 0] * n[:, 0] * a, sign * b, -sign * n[:, 0]], -1)
    fy = torch.stack([b, sign + n[:, 1] * n[:, 1] * a, -n[:, 1]], -1)
    return fx, fy, n


def _local(fr, v):
    return torch.stack([(fr[0] * v).sum(-1), (fr[1] * v).sum(-1), (fr[2] * v).sum(-1)], -1)


def _lambda_ggx(a2, c):
    c2 = c * c
    t = torch.clamp(1 - c2, min=0) / c2
    lam = 0.5 * (-1 + torch.sqrt(1 + a2 * t))
    return torch.where(c > 0, lam, torch.zeros_like(lam))


def final_shading(fs_dir, fs_dist, fs_Li, occ, normal, ray_dir, kd, rm):
    """All inputs float64 tensors; returns (color, diff_light, spec_light) for the occ > 0.1 branch; background pixels
    (env lookup, no gradient) are returned as zeros."""
    F0 = 0.04
    with torch.no_grad():
        active = (occ.reshape(-1) > 0.1) & (fs_dist.reshape(-1) > 0)
    # inactive pixels carry zeros (normal = 0, dir = 0): give them a harmless configuration so that the masked-out
    # branch produces no NaN for autograd to propagate through torch.where
    up = torch.tensor([0.0, 0.0, 1.0], dtype=F64)
    normal = torch.where(active[:, None], normal, up.expand_as(normal))
    fs_dir = torch.where(active[:, None], fs_dir, up.expand_as(fs_dir))
    ray_dir = torch.where(active[:, None], ray_dir, -up.expand_as(ray_dir))
    rough, metallic = rm[:, 0], rm[:, 1]
    spec = F0 * (1.0 - metallic)[:, None] + kd * metallic[:, None]
    fr = _frame(normal)
    wo = _local(fr, -ray_dir)
    wi = _local(fr, fs_dir)
    alpha = rough * rough
    alpha = torch.where(alpha.detach() < 1e-4, torch.zeros_like(alpha), alpha)
    with torch.no_grad():
        pD = _lum(kd) * (1 - metallic)
        cosv = (-(ray_dir) * normal).sum(-1)
        fres = spec + (1 - spec) * torch.clamp(1 - cosv, min=0)[:, None] ** 5
        pS = _lum(fres) * (metallic + (1 - metallic))
        gate = ~(torch.minimum(wo[:, 2], wi[:, 2]) < 1e-6)
    Dl = torch.clamp(0.31830988 * wi[:, 2], min=0.0)  # M_1_PI literal of the reference
    Dl = torch.where(gate & (pD > 0), Dl, torch.zeros_like(Dl))
    diffuse_val = Dl[:, None] * fs_Li
    safe = gate & (alpha.detach() != 0)
    wo_s = torch.where(safe[:, None], wo, up.expand_as(wo))
    wi_s = torch.where(safe[:, None], wi, up.expand_as(wi))
    alpha = torch.where(safe, alpha, torch.full_like(alpha, 0.5))
    wo, wi = wo_s, wi_s
    h = wo + wi
    h = h / torch.sqrt((h * h).sum(-1, keepdim=True))
    c = (wo * h).sum(-1)
    a2 = alpha * alpha
    d = (h[:, 2] * a2 - h[:, 2]) * h[:, 2] + 1
    D = a2 / (d * d * math.pi)
    G = 1 / (1 + _lambda_ggx(a2, wo[:, 2]) + _lambda_ggx(a2, wi[:, 2]))
    Fr = spec + (1 - spec) * torch.clamp(1 - c, min=0)[:, None] ** 5
    Fs = Fr * (D * G * 0.25 / wo[:, 2])[:, None]
    smask = gate & (pS > 0) & (alpha.detach() != 0)
    Fs = torch.where(smask[:, None], Fs, torch.zeros_like(Fs))
    specular_val = Fs * fs_Li
    color = kd * (1.0 - metallic)[:, None] * diffuse_val + specular_val
    z = torch.zeros_like(color)
    a = active[:, None]
    return torch.where(a, color, z), torch.where(a, diffuse_val, z), torch.where(a, specular_val, z)


def final_shading_grads(fs_dir, fs_dist, fs_Li, occ, normal, ray_dir, kd, rm, g_color, g_diff, g_spec):
    t = lambda a, rg=False: torch.tensor(np.asarray(a), dtype=F64, requires_grad=rg)
    N, KD, RM, LI = t(normal, True), t(kd, True), t(rm, True), t(fs_Li, True)
    c, d, s = final_shading(t(fs_dir), t(fs_dist), LI, t(occ), N, t(ray_dir), KD, RM)
    loss = (c * t(g_color)).sum() + (d * t(g_diff)).sum() + (s * t(g_spec)).sum()
    loss.backward()
    z = lambda x: torch.zeros_like(x) if x.grad is None else x.grad
    return z(N).numpy(), z(KD).numpy(), z(RM).numpy(), z(LI).numpy()


def eval_final_grad_env(res_ld, res_w, vis, grad_Li, W, H):
    """grad_env [H*W,3] in float64 for Li = W * bilinear(env): the fp32 addressing decisions (texel indices, lerp
    weights) come from the oracle (orc_eval_final_taps), the accumulation is done in float64."""
    from . import oracle as O
    taps, uv, valid = O.eval_final_taps(res_ld, W, H)
    g = np.zeros((W * H, 3), np.float64)
    act = (res_ld[:, 0] > 0.1) & (np.asarray(vis).reshape(-1) > 0) & (valid > 0)
    gl = np.asarray(res_w).reshape(-1, 1).astype(np.float64) * np.asarray(grad_Li).astype(np.float64)
    u = uv[:, 0].astype(np.float64)
    v = uv[:, 1].astype(np.float64)
    for k, w in enumerate(((1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v)):
        np.add.at(g, taps[act, k], (w[:, None] * gl)[act])
    return g


def eaw(c_phi, n_phi, p_phi, fx, fy, step, occ, color, normal, pos):
    """float64 torch forward of the a-trous filter (differentiable in color, normal, pos)."""
    k1 = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0], dtype=F64)
    C = color.reshape(fy, fx, 3)
    Nn = normal.reshape(fy, fx, 3)
    Pp = pos.reshape(fy, fx, 3)
    ssum = torch.zeros_like(C)
    wsum = torch.zeros(fy, fx, dtype=F64)
    ys, xs = torch.meshgrid(torch.arange(fy), torch.arange(fx), indexing="ij")
    for j in range(5):
        for i in range(5):
            dx, dy = (i - 2) * step, (j - 2) * step
            ux, uy = xs + dx, ys + dy
            valid = (ux >= 0) & (ux < fx) & (uy >= 0) & (uy < fy)
            uxc, uyc = ux.clamp(0, fx - 1), uy.clamp(0, fy - 1)
            ct, nt, pt = C[uyc, uxc], Nn[uyc, uxc], Pp[uyc, uxc]
            cw = torch.clamp(torch.exp(-((C - ct) ** 2).sum(-1) / c_phi), max=1.0)
            nw = torch.clamp(torch.exp(-torch.clamp(((Nn - nt) ** 2).sum(-1), min=0) / n_phi), max=1.0)
            pw = torch.clamp(torch.exp(-torch.clamp(((Pp - pt) ** 2).sum(-1), min=0) / p_phi), max=1.0)
            w = cw * nw * pw * (k1[i] * k1[j] / 256.0)
            w = torch.where(valid, w, torch.zeros_like(w))
            ssum = ssum + ct * w[..., None]
            wsum = wsum + w
    out = ssum / wsum[..., None]
    occ2 = occ.reshape(fy, fx, 1)
    return torch.where(occ2 < 0.1, C.detach(), out).reshape(-1, 3)


def eaw_grads(c_phi, n_phi, p_phi, fx, fy, step, occ, color, normal, pos, g_out):
    t = lambda a, rg=False: torch.tensor(np.asarray(a), dtype=F64, requires_grad=rg)
    C, Nn, Pp = t(color, True), t(normal, True), t(pos, True)
    out = eaw(c_phi, n_phi, p_phi, fx, fy, int(step), t(occ), C, Nn, Pp)
    (out * t(g_out)).sum().backward()
    return out.detach().numpy(), C.grad.numpy(), Nn.grad.numpy(), Pp.grad.numpy()
