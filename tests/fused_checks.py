"""Checks of the single-launch replacements of the driver's torch chains (csrc/screen.cu,
mirres_final_shading_bwd_multi) against the torch expressions / single-pass kernels they replace.  Shared by the CPU
suite (host-check flavour of the kernels) and the GPU suite (libmirres_b200.so): every function takes the bound kernels
and the device."""
import numpy as np
import pytest
import torch

import parity as P
from mirres_restir_nerf_mesh_b200 import renderer_restir as R, synth

def material_kernel_equals_torch_and_numpy_expressions(k, dev):
    g = torch.Generator().manual_seed(0)
    n = 5000
    pos = (torch.rand(n, 3, generator=g) * 4 - 2).float().to(dev)
    occ = (torch.rand(n, 1, generator=g) > 0.4).float().to(dev)
    mat = synth.ProceduralMaterial(0.4)
    # mode 0: the G-buffer materials
    kd, rs = mat.gbuffer_materials(pos, occ)
    ref = mat.sample_no_di_dense(pos) * occ
    assert torch.equal(kd, ref[:, 0:3]) and torch.equal(rs, ref[:, 4:6])
    kd_np, rough_np, met_np = synth.material(pos.cpu().numpy(), 0.4)
    assert np.array_equal(kd.cpu().numpy(), kd_np * occ.cpu().numpy())
    assert np.array_equal(rs.cpu().numpy()[:, 0:1], rough_np * occ.cpu().numpy())
    # mode 1: the merge between bounces, with and without the albedo scale
    for scale in (None, (0.5, 1.5, 2.0)):
        kd0, rs0 = (torch.rand(n, 3, generator=g) * 2 - 0.5).to(dev), torch.rand(n, 2, generator=g).to(dev)
        want = R._query_material(_DenseOnly(mat), occ, pos, kd0.clone(), rs0.clone(), scale is not None, scale or (1, 1, 1))
        got = R._query_material(mat, occ, pos, kd0.clone(), rs0.clone(), scale is not None, scale or (1, 1, 1))
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        legacy = R._query_material(_Legacy(mat), occ, pos, kd0.clone(), rs0.clone(), scale is not None, scale or (1, 1, 1))
        assert torch.equal(got[0], legacy[0]) and torch.equal(got[1], legacy[1])


class _DenseOnly:
    def __init__(self, m):
        self.sample_no_di_dense = m.sample_no_di_dense
        self.sample_no_di = m.sample_no_di


class _Legacy:
    def __init__(self, m):
        self.sample_no_di = m.sample_no_di


def sum_images_is_the_sequential_torch_sum(k, dev):
    g = torch.Generator().manual_seed(1)
    srcs = [(torch.randn(777, 3, generator=g) * 10 ** float(e)).to(dev) for e in range(-3, 4)]
    want = torch.zeros(777, 3, device=dev)
    for s_ in srcs:
        want += s_
    dst = torch.zeros(777, 3, device=dev)
    k.sum_images(srcs[:3], dst, accumulate=True)
    k.sum_images(srcs[3:], dst, divisor=4.0, accumulate=True)
    assert torch.equal(dst, want / 4)
    fresh = torch.full((777, 3), 5.0, device=dev)
    k.sum_images(srcs, fresh)
    assert torch.equal(fresh, want)
    k.sum_images([], fresh, divisor=3.0, accumulate=True)
    assert torch.equal(fresh, want / 3.0)
    with pytest.raises(Exception):
        k.sum_images([fresh], fresh)


def _composite_torch(occ, kd, rs, dd, ds, di):
    diffuse = kd * (1.0 - rs[..., 1:2])
    c = diffuse * dd + ds + di
    c = torch.where(occ <= 0.1, torch.ones_like(c), c)
    return torch.nan_to_num(c, 0.0)


def composite_forward_exact_backward_matches_autograd(k, dev):
    g = torch.Generator().manual_seed(2)
    n = 3000
    occ = (torch.rand(n, 1, generator=g) > 0.3).float()
    mk = lambda *s: torch.rand(*s, generator=g)
    kd, rs, dd, ds, di = mk(n, 3), mk(n, 2), mk(n, 3) * 3, mk(n, 3) * 2, mk(n, 3)
    dd[5, 1] = float("nan")
    ds[9, 0] = float("inf")
    di[11, 2] = -float("inf")
    occ[5] = occ[9] = occ[11] = 1.0
    occ, kd, rs, dd, ds, di = (t.to(dev) for t in (occ, kd, rs, dd, ds, di))
    leaves = [t.clone().requires_grad_(True) for t in (kd, rs, dd, ds)]
    want = _composite_torch(occ, leaves[0], leaves[1], leaves[2], leaves[3], di)
    w = torch.linspace(0.5, 1.5, n * 3).reshape(n, 3).to(dev)
    (want * w).sum().backward()
    mine = [t.clone().requires_grad_(True) for t in (kd, rs, dd, ds)]
    got = R.Composite.apply(occ, mine[0], mine[1], mine[2], mine[3], di)
    assert torch.equal(got, want.detach())
    (got * w).sum().backward()
    for a, b in zip(mine, leaves):
        assert torch.equal(torch.isnan(a.grad), torch.isnan(b.grad))  # 0 * nan at the poisoned pixels, as in torch
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-6, atol=1e-7, equal_nan=True)


def final_shading_bwd_multi_equals_the_single_pass_kernels(k, dev):
    sc = P.scene("T1", 0.3)
    gb = {kk: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for kk, v in sc["gbuffer"].items()}
    fx, fy = sc["W"], sc["H"]
    n = fx * fy
    g = torch.Generator().manual_seed(3)
    K = 5
    dirs = [torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1).to(dev) for _ in range(K)]
    dists = [((torch.rand(n, 1, generator=g) > 0.3).float() * 1e6).to(dev) for _ in range(K)]
    Lis = [(torch.rand(n, 3, generator=g) * 3).to(dev) for _ in range(K)]
    gC, gD, gS = (torch.randn(n, 3, generator=g).to(dev) for _ in range(3))
    occ, nrm, ray = gb["occ_map"], gb["normal_map"], gb["ray_dir_map"]
    kd, rs = gb["diffuse_map"], gb["roughness_specular"]
    e = lambda c: torch.empty(n, c, device=dev)
    singles = []
    for j in range(K):
        o = (e(3), e(3), e(2), e(3))
        k.final_shading_bwd(dirs[j], dists[j], Lis[j], fx, fy, occ, nrm, ray, kd, rs, gC / 4.0, gD / 4.0, gS / 4.0, *o)
        singles.append(o)
    acc = [singles[K - 1][c].clone() for c in range(3)]
    for j in range(K - 2, -1, -1):
        for c in range(3):
            acc[c] += singles[j][c]
    assert acc[0].abs().sum() > 0 and acc[2].abs().sum() > 0
    out = (e(3), e(3), e(2))
    gLi = [e(3) for _ in range(K)]
    k.final_shading_bwd_multi(dirs, dists, Lis, fx, fy, occ, nrm, ray, kd, rs, gC, gD, gS, *out, gLi, grad_divisor=4.0)
    for c in range(3):
        assert torch.equal(out[c], acc[c])
    for j in range(K):
        assert torch.equal(gLi[j], singles[j][3])
    # two chunks (last passes first) with accumulation, summed radiance gradient, no colour gradient
    zero = torch.zeros(n, 3, device=dev)
    singles0 = []
    for j in range(K):
        o = (e(3), e(3), e(2), e(3))
        k.final_shading_bwd(dirs[j], dists[j], Lis[j], fx, fy, occ, nrm, ray, kd, rs, zero, gD, gS, *o)
        singles0.append(o)
    out2, gsum = (e(3), e(3), e(2)), [e(3)]
    k.final_shading_bwd_multi(dirs[3:], dists[3:], Lis[3:], fx, fy, occ, nrm, ray, kd, rs, None, gD, gS, *out2, gsum,
                              sum_grad_Li=True)
    k.final_shading_bwd_multi(dirs[:3], dists[:3], Lis[:3], fx, fy, occ, nrm, ray, kd, rs, None, gD, gS, *out2, gsum,
                              sum_grad_Li=True, accumulate=True)
    want = [singles0[K - 1][c].clone() for c in range(4)]
    for j in range(K - 2, -1, -1):
        for c in range(4):
            want[c] += singles0[j][c]
    for c in range(3):
        assert torch.equal(out2[c], want[c])
    assert torch.equal(gsum[0], want[3])
    # passes that share direction and distance (the reference's saved aliases): one evaluation on the summed radiance
    singles1 = []
    for j in range(K):
        o = (e(3), e(3), e(2), e(3))
        k.final_shading_bwd(dirs[0], dists[0], Lis[j], fx, fy, occ, nrm, ray, kd, rs, gC, gD, gS, *o)
        singles1.append(o)
    want1 = [singles1[K - 1][c].clone() for c in range(4)]
    for j in range(K - 2, -1, -1):
        for c in range(4):
            want1[c] += singles1[j][c]
    out3, gsum3 = (e(3), e(3), e(2)), [e(3)]
    k.final_shading_bwd_multi([dirs[0]] * K, [dists[0]] * K, Lis, fx, fy, occ, nrm, ray, kd, rs, gC, gD, gS, *out3, gsum3,
                              sum_grad_Li=True)
    for got, want_ in zip(list(out3) + gsum3, want1):
        assert want_.abs().sum() > 0
        assert (got - want_).abs().max() <= 2e-6 * want_.abs().max()
