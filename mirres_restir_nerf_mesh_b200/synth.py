"""Synthetic inputs of the TensoIR / NeRF-synthetic shapes (no dataset download is possible).

Everything the hot path consumes is generated here, seeded and deterministic:
  * meshes   : bumped icosphere (config C1) and torus-knot tube (C2/C3/C5), `vert [V,3] f32`, `tri [F,3] i32`
               (layout of nerf/renderer.py:171-172);
  * cameras  : NeRF-synthetic pinhole rays (nerf/utils.py:350-421 convention: pixel centre + 0.5, -y, -z);
  * envmap   : HDR `[He,We,3]` = sky gradient + sun lobe + three area lights, clamped >= 0.01 (nerf/utils.py:1589);
  * materials: procedural kd / roughness / metallic of the surface position, also exposed through an object with the
               `sample_no_di(x[M,3]) -> [M,6]` protocol that nerf/renderer_restir.py:399-402 expects from `mlp_mat`.
Pure numpy; the G-buffer itself is produced by a closest-hit trace supplied by the caller (CUDA in the product,
the oracle in CPU tests).
"""
import numpy as np


def icosphere(level=5, radius=0.6, bump=0.05):
    """F = 20 * 4**level triangles; level 5 -> 20480 (config C1)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * (len(v) + 1) + es[:, 1]
        uk, inv = np.unique(key, return_inverse=True)
        a = uk // (len(v) + 1)
        b = uk % (len(v) + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        n = len(f)
        m01 = base + inv[0:n]
        m12 = base + inv[n:2 * n]
        m20 = base + inv[2 * n:3 * n]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    r = radius * (1.0 + (bump / radius) * np.sin(9 * v[:, 0:1]) * np.sin(7 * v[:, 1:2]) * np.sin(5 * v[:, 2:3]))
    v = v * r
    return v.astype(np.float32), f.astype(np.int32)


def torus_knot(nu=1000, nv=250, p=2, q=3, tube=0.5, scale=0.8):
    """Closed tube around a (p,q) torus knot on a periodic (nu,nv) grid; F = 2*nu*nv (C2: 500k, C5: 2M)."""
    u = np.linspace(0.0, 2 * np.pi, nu, endpoint=False)
    def curve(t):
        r = np.cos(q * t) + 2.0
        return np.stack([r * np.cos(p * t), r * np.sin(p * t), -np.sin(q * t)], -1)
    c = curve(u)
    d = curve(u + 1e-4) - c
    tng = d / np.linalg.norm(d, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    nrm = np.cross(tng, up)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    bnm = np.cross(tng, nrm)
    w = np.linspace(0.0, 2 * np.pi, nv, endpoint=False)
    rad = tube * (1.0 + 0.15 * np.sin(8 * u)[:, None] * np.cos(3 * w)[None, :])
    pts = c[:, None, :] + rad[..., None] * (np.cos(w)[None, :, None] * nrm[:, None, :] + np.sin(w)[None, :, None] * bnm[:, None, :])
    pts = pts.reshape(-1, 3)
    pts -= 0.5 * (pts.max(0) + pts.min(0))
    pts *= scale / np.abs(pts).max()
    iu = np.arange(nu)[:, None]
    iv = np.arange(nv)[None, :]
    a = iu * nv + iv
    b = ((iu + 1) % nu) * nv + iv
    cidx = ((iu + 1) % nu) * nv + (iv + 1) % nv
    didx = iu * nv + (iv + 1) % nv
    tri = np.concatenate([np.stack([a, b, cidx], -1).reshape(-1, 3), np.stack([a, cidx, didx], -1).reshape(-1, 3)], 0)
    return pts.astype(np.float32), tri.astype(np.int32)


def camera_rays(W, H, view=0, n_views=100, radius=3.2, elevation_deg=30.0, fov_x=0.6911112):
    """Returns (rays_o [N,3], rays_d [N,3] unit) for view `view`, pixelIndex = y*W + x."""
    az = 2 * np.pi * view / n_views
    el = np.deg2rad(elevation_deg)
    eye = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    fl = W / (2.0 * np.tan(fov_x / 2.0))
    xs = (np.arange(W) + 0.5 - W / 2.0) / fl
    ys = -(np.arange(H) + 0.5 - H / 2.0) / fl
    d = xs[None, :, None] * right[None, None, :] + ys[:, None, None] * upv[None, None, :] + fwd[None, None, :]
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(eye, d.shape)
    return o.reshape(-1, 3).astype(np.float32).copy(), d.reshape(-1, 3).astype(np.float32)


def camera_pose(view=0, n_views=100, radius=3.2, elevation_deg=30.0):
    """[4,3] f32 rows (right, up, forward, eye) of the camera `camera_rays` uses for `view`."""
    az = 2 * np.pi * view / n_views
    el = np.deg2rad(elevation_deg)
    eye = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    return np.stack([right, upv, fwd, eye]).astype(np.float32)


def camera_rays_torch(W, H, pose, fov_x=0.6911112):
    """Device-side twin of `camera_rays` (the reference generates rays on the GPU from the pose as well, nerf/utils.py
    get_rays): pose [4,3] tensor -> (rays_o [N,3], rays_d [N,3] unit), pixelIndex = y*W + x."""
    import torch
    fl = W / (2.0 * np.tan(fov_x / 2.0))
    dev = pose.device
    xs = (torch.arange(W, device=dev, dtype=torch.float32) + 0.5 - W / 2.0) / fl
    ys = -(torch.arange(H, device=dev, dtype=torch.float32) + 0.5 - H / 2.0) / fl
    d = xs[None, :, None] * pose[0][None, None, :] + ys[:, None, None] * pose[1][None, None, :] + pose[2][None, None, :]
    d = d / torch.sqrt((d * d).sum(-1, keepdim=True))
    return pose[3].expand(H * W, 3).contiguous(), d.reshape(-1, 3).contiguous()


def envmap(He=256, We=512, seed=0):
    """HDR `[He,We,3]` f32, the layout of `lgt.base` (nerf/render_helper.py)."""
    rng = np.random.default_rng(seed)
    v = (np.arange(He) + 0.5) / He
    u = (np.arange(We) + 0.5) / We
    theta = v[:, None] * np.pi
    phi = u[None, :] * 2 * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), np.sin(theta) * np.sin(phi)], -1)
    sky = 0.2 + 0.8 * (0.5 + 0.5 * d[..., 1:2]) * np.array([0.6, 0.75, 1.0])
    img = sky.copy()
    def lobe(direction, sigma_deg, peak, col):
        direction = np.asarray(direction, np.float64)
        direction /= np.linalg.norm(direction)
        cosang = np.clip((d * direction).sum(-1), -1, 1)
        ang = np.arccos(cosang)
        s = np.deg2rad(sigma_deg)
        return peak * np.exp(-0.5 * (ang / s) ** 2)[..., None] * np.asarray(col)
    img += lobe([0.5, 0.7, 0.3], 2.0, 5.0e3, [1.0, 0.95, 0.85])
    for _ in range(3):
        dirv = rng.standard_normal(3)
        img += lobe(dirv, 12.0, 50.0, 0.5 + 0.5 * rng.random(3))
    return np.maximum(img, 0.01).astype(np.float32)


def _tri_np(x):
    # triangle wave in [0,1] from exactly-rounded elementwise ops only (identical in numpy, torch CPU and torch CUDA,
    # unlike sin, whose last ulp differs between libraries and would break bit-exact parity of the bounce materials)
    f = x - np.floor(x)
    return np.abs(np.float32(2.0) * f - np.float32(1.0))


def material(pos, metallic=0.0):
    """Procedural kd [M,3], roughness [M,1], metallic [M,1] of surface positions pos [M,3]."""
    pos = np.asarray(pos, np.float32)
    f = np.float32
    kd = (f(0.1) + f(0.8) * _tri_np(pos * f(1.7))).astype(np.float32)
    rough = (f(0.08) + f(0.92) * _tri_np(pos[:, 0:1] * f(0.8) + pos[:, 1:2] * f(0.5))).astype(np.float32)
    met = np.full_like(rough, metallic)
    return kd, rough, met


class ProceduralMaterial:
    """Stand-in for the tiny-cuda-nn `mlp_mat` (out of scope): `.sample_no_di(x) -> [M,6]` with kd in 0:3,
    roughness in 4, metallic in 5 (nerf/renderer_restir.py:399-402).  Works on torch tensors of any device and
    returns bit-identical values to `material()` (every op is a single exactly-rounded elementwise kernel)."""

    def __init__(self, metallic=0.0):
        self.metallic = float(metallic)

    @staticmethod
    def _tri(x):
        import torch
        f = x - torch.floor(x)
        return torch.abs(2.0 * f - 1.0)

    def sample_no_di(self, x):
        import torch
        kd = 0.1 + 0.8 * self._tri(x * 1.7)
        rough = 0.08 + 0.92 * self._tri(x[:, 0:1] * 0.8 + x[:, 1:2] * 0.5)
        met = torch.full_like(rough, self.metallic)
        return torch.cat([kd, torch.zeros_like(rough), rough, met], dim=-1)

    # evaluated on every pixel and merged with a mask by the driver: no host synchronisation in the spp loop
    sample_no_di_dense = sample_no_di

    def sample_no_di_masked_(self, occ, pos, kd_out, rs_out, scale=None):
        """The lookup and the driver's torch.where merge in ONE launch (mirres_material_procedural, CUDA tensors only):
        pixels with occ >= 0.5 get the material at `pos`, the others keep theirs; same values as sample_no_di."""
        from .slangpy_shim import get_kernels, _c
        get_kernels().material_procedural(_c(pos), _c(occ), 1, self.metallic, kd_out, rs_out, scale)
        return kd_out, rs_out

    def gbuffer_materials(self, pos, occ):
        """kd [n,3] and (roughness, metallic) [n,2] of the primary hits, zero where occ is zero: the columns 0:3 and 4:6 of
        `sample_no_di_dense(pos) * occ`, one launch on CUDA tensors."""
        import torch
        from .slangpy_shim import get_kernels, _c
        n = pos.shape[0]
        kd = torch.empty((n, 3), dtype=torch.float32, device=pos.device)
        rs = torch.empty((n, 2), dtype=torch.float32, device=pos.device)
        get_kernels().material_procedural(_c(pos), _c(occ), 0, self.metallic, kd, rs)
        return kd, rs


def gbuffer_from_hits(rays_o, rays_d, hit, t, pos, normal, metallic=0.0):
    """Assemble the G-buffer maps render_stage1 hands to run_restir_di_with_pt (nerf/renderer.py:1092-1096,1121)."""
    hit = np.asarray(hit).reshape(-1) > 0
    occ = hit.astype(np.float32)[:, None]
    pos_map = np.where(hit[:, None], pos, 0.0).astype(np.float32)
    normal_map = np.where(hit[:, None], normal, 0.0).astype(np.float32)
    depth = np.where(hit, np.linalg.norm(pos_map - rays_o, axis=1), 0.0).astype(np.float32)[:, None]
    kd, rough, met = material(pos_map, metallic)
    kd = np.where(hit[:, None], kd, 0.0).astype(np.float32)
    rs = np.where(hit[:, None], np.concatenate([rough, met], 1), 0.0).astype(np.float32)
    return dict(occ_map=occ, pos_map=pos_map, normal_map=normal_map, depth_map=depth, diffuse_map=kd,
                roughness_specular=rs, ray_dir_map=np.ascontiguousarray(rays_d, np.float32))


CONFIGS = {
    # name: mesh generator, frame, env (He, We), spp, indirect bounces (reference MAX_Bounce semantics)
    "C1": dict(mesh=("icosphere", dict(level=5)), W=256, H=256, env=(256, 512), spp=1, max_bounce=1),
    "C2": dict(mesh=("torus_knot", dict(nu=1000, nv=250)), W=800, H=800, env=(256, 512), spp=4, max_bounce=2),
    "C3": dict(mesh=("torus_knot", dict(nu=1000, nv=250)), W=800, H=800, env=(256, 512), spp=512, max_bounce=2),
    # the frame stage 1 really renders for 800 x 800 data: ssaa = 2 (reference main.py:140, nerf/utils.py:770-777)
    "C2S": dict(mesh=("torus_knot", dict(nu=1000, nv=250)), W=1600, H=1600, env=(256, 512), spp=2, max_bounce=2),
    "C5": dict(mesh=("torus_knot", dict(nu=2000, nv=500)), W=2048, H=2048, env=(1024, 2048), spp=128, max_bounce=3),
    # small cases for parity tests
    "T0": dict(mesh=("icosphere", dict(level=2)), W=48, H=40, env=(16, 32), spp=2, max_bounce=2),
    "T1": dict(mesh=("icosphere", dict(level=3)), W=96, H=64, env=(32, 64), spp=3, max_bounce=2),
    "T2": dict(mesh=("torus_knot", dict(nu=96, nv=24)), W=128, H=96, env=(64, 128), spp=2, max_bounce=2),
}


def make_mesh(cfg):
    kind, kw = cfg["mesh"]
    return icosphere(**kw) if kind == "icosphere" else torus_knot(**kw)
