"""The oracle itself: numerical contract, BVH invariants, brute-force ray agreement (CPU only)."""
import numpy as np
import pytest

from mirres_restir_nerf_mesh_b200 import synth


def test_fp_contract_not_contracted(oracle):
    assert oracle.fp_contract_selftest() == 0


def test_expf_is_within_stated_error(oracle):
    """mr_expf is the one contract function evaluated in fp32 (include/mirres_fpmath.h): within 1.5 ulp, monotone on
    the range the a-trous weights use, exact limits."""
    x = np.linspace(-100, 80, 400001, dtype=np.float32)
    got = oracle.fpmath(6, x)
    want = np.exp(x.astype(np.float64))
    ulp = np.spacing(np.abs(want.astype(np.float32))).astype(np.float64)
    assert np.max(np.abs(got.astype(np.float64) - want) / np.maximum(ulp, 1e-45)) <= 1.5
    assert (np.diff(got) >= 0).all()
    sp = oracle.fpmath(6, np.array([0.0, -200.0, 100.0, -0.0], np.float32))
    assert sp[0] == 1.0 and sp[1] == 0.0 and np.isinf(sp[2]) and sp[3] == 1.0


@pytest.mark.parametrize("op,fn,lo,hi", [(0, np.sin, -7, 7), (1, np.cos, -7, 7), (2, np.arccos, -1, 1)])
def test_fpmath_is_correctly_rounded(oracle, op, fn, lo, hi):
    x = np.linspace(lo, hi, 400001, dtype=np.float32)
    got = oracle.fpmath(op, x)
    want = fn(x.astype(np.float64)).astype(np.float32)
    assert (got == want).mean() > 0.9999
    assert np.max(np.abs(got.astype(np.float64) - want) / np.maximum(np.spacing(np.abs(want)), 1e-45)) <= 1.0


def test_atan2_quadrants(oracle):
    rng = np.random.default_rng(0)
    y = rng.standard_normal(200000).astype(np.float32)
    x = rng.standard_normal(200000).astype(np.float32)
    got = oracle.fpmath(3, y, x)
    want = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    assert (got == want).mean() > 0.9999
    sp = oracle.fpmath(3, np.array([0, 0, 1, -1], np.float32), np.array([1, -1, 0, 0], np.float32))
    assert np.allclose(sp, [0, np.pi, np.pi / 2, -np.pi / 2])


@pytest.mark.parametrize("mesh", ["T0", "T2"])
def test_bvh_invariants(oracle, mesh):
    v, f = synth.make_mesh(synth.CONFIGS[mesh])
    b = oracle.Bvh(v, f)
    F = f.shape[0]
    leaf = F - 1
    info, aabb = b.info, b.aabb
    assert (np.diff(b.sorted_codes[:, 0].astype(np.int64)) >= 0).all()
    # stable: equal codes keep element order
    same = np.diff(b.sorted_codes[:, 0]) == 0
    assert (np.diff(b.sorted_codes[:, 1])[same] > 0).all()
    assert (info[leaf:, :2] == 0).all() and sorted(info[leaf:, 2].tolist()) == list(range(F))
    children = np.concatenate([info[:leaf, 0], info[:leaf, 1]])
    assert sorted(children.tolist()) == list(range(1, 2 * F - 1))
    l, r = info[:leaf, 0], info[:leaf, 1]
    assert (aabb[:leaf, :3] == np.minimum(aabb[l, :3], aabb[r, :3])).all()
    assert (aabb[:leaf, 3:] == np.maximum(aabb[l, 3:], aabb[r, 3:])).all()
    tv = v[f[info[leaf:, 2]]]
    assert (aabb[leaf:, :3] == tv.min(1)).all() and (aabb[leaf:, 3:] == tv.max(1)).all()


def test_single_and_two_triangle_meshes(oracle):
    v = np.array([[0, 0, 0], [1, 0, 0.25], [0, 1, 0.5], [1, 1, 0.5]], np.float32)
    b1 = oracle.Bvh(v, np.array([[0, 1, 2]], np.int32))
    assert b1.info.tolist() == [[0, 0, 0]]
    hit, t, pos, nrm, prim = oracle.trace(b1, np.array([[0.2, 0.2, 1]], np.float32), np.array([[0, 0, -1]], np.float32))
    assert hit[0] == 1 and abs(t[0] - 0.85) < 1e-6 and prim[0] == 0
    # quirk: a triangle whose box is flat along an axis can never be hit (aabb_hit rejects t_max <= t_min)
    flat = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    bf = oracle.Bvh(flat, np.array([[0, 1, 2]], np.int32))
    assert oracle.trace(bf, np.array([[0.2, 0.2, 1]], np.float32), np.array([[0, 0, -1]], np.float32))[0][0] == 0
    b2 = oracle.Bvh(v, np.array([[0, 1, 2], [1, 3, 2]], np.int32))
    assert b2.info[0].tolist()[:2] == [1, 2]


def test_primary_rays_match_brute_force(oracle):
    v, f = synth.icosphere(2)
    b = oracle.Bvh(v, f)
    ro, rd = synth.camera_rays(48, 40)
    hit, t, pos, nrm, prim = oracle.trace(b, ro, rd)
    v0, v1, v2 = (v[f[:, k]].astype(np.float64) for k in range(3))
    for i in range(0, len(ro), 7):
        o, d = ro[i].astype(np.float64), rd[i].astype(np.float64)
        d /= np.linalg.norm(d)
        e1, e2 = v1 - v0, v2 - v0
        P = np.cross(d, e2)
        det = (e1 * P).sum(1)
        inv = 1 / det
        T = o - v0
        u = (T * P).sum(1) * inv
        Q = np.cross(T, e1)
        w = (Q * d).sum(1) * inv
        tt = (e2 * Q).sum(1) * inv
        m = (u >= 0) & (w >= 0) & (u + w <= 1) & (tt > 0)
        assert m.any() == bool(hit[i])
        if m.any():
            k = int(np.argmin(np.where(m, tt, np.inf)))
            assert k == prim[i] and abs(tt[k] - t[i]) < 1e-4


def test_negative_t_self_hits_are_reproduced(oracle):
    # quirk 9.1: hits behind the origin are accepted; a shadow ray starting just above a coarse triangle reports a hit
    v, f = synth.icosphere(1)
    b = oracle.Bvh(v, f)
    c = v[f].mean(1)
    n = c / np.linalg.norm(c, axis=1, keepdims=True)
    o = (c + 0.01 * n).astype(np.float32)
    hit, t, *_ = oracle.trace(b, o, n.astype(np.float32))
    assert hit.all() and (t < 0).all()


def test_env_distribution_is_a_distribution(oracle):
    env = synth.envmap(32, 64)
    tex = np.ascontiguousarray(env[::-1].reshape(-1, 3))
    pdf_, cdf_, mpdf_, mcdf_ = oracle.env_build_distribution(tex, 64, 32)
    cdf = cdf_.reshape(32, 65)
    assert (cdf[:, 0] == 0).all() and (cdf[:, -1] == 1).all() and (np.diff(cdf, axis=1) >= -1e-7).all()
    assert mcdf_[0] == 0 and mcdf_[-1] == 1 and abs(mpdf_.sum() - 1) < 1e-4
    assert np.allclose(pdf_.reshape(32, 64).sum(1), 1, atol=1e-4)
    ld, uv, pdf = oracle.light_tiles(tex, 64, 32, (pdf_, cdf_, mpdf_, mcdf_), 7, 16, 1024)
    assert ((ld[:, 0] == 1) | (ld[:, 0] == 0)).all() and (uv >= 0).all() and (uv[:, 0] < 64).all() and (uv[:, 1] < 32).all()
    assert ld[:, 0].mean() > 0.99 and (pdf[ld[:, 0] == 1] > 0).all()
    # importance sampling: the solid-angle pdf integrates to one => E[1/pdf] = 4 pi
    est = (1.0 / pdf[ld[:, 0] == 1]).mean()
    assert abs(est / (4 * np.pi) - 1) < 0.1


def test_neighbor_offsets(oracle):
    o = oracle.neighbor_offsets(8192).reshape(-1, 2)
    assert o.min() >= -127 and o.max() <= 127 and (o == np.round(o)).all()
    assert (np.hypot(o[:, 0], o[:, 1]) <= 128).all()
