"""ctypes front-end of the CPU oracle (oracle/libmirres_oracle.so) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for the Slang kernels (see oracle/orc_common.h); the bilateral denoiser restatement (bilateral_fwd /
bilateral_bwd) IS pinned: tests/test_gpu.py checks it against the reference's own denoising.cu compiled into
oracle/_ref (oracle/ref.py).  All arrays are C-contiguous numpy arrays; outputs are
allocated here.  Function names follow the reference kernels they restate:
  bvh_build          nerf/renderer_restir.py:25-89 (+ nerf/bvhworkers/*.slang)
  trace              nerf/ScreenSpaceReSTIR/utils/helperDi.slang:313-395
  env_build_distribution   nerf/ScreenSpaceReSTIR/GenerateLightTiles.py:4-29
  light_tiles        nerf/ScreenSpaceReSTIR/GenerateLightTiles.slang:16-62
  initial/temporal/spatial_resampling, final_visibility, eval_final_fwd, final_shading_fwd,
  bounce_first, bounce_shade   nerf/ScreenSpaceReSTIR/*.slang (see orc_kernels.cpp)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libmirres_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "mirres_fpmath.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


class flavour:
    """`with flavour("fast"):` runs the oracle built with FMA contraction and C-library transcendentals (make fast; see
    include/mirres_fpmath.h) inside the block -- the sensitivity study's stand-in for the reference binary's numerics.
    The contract flavour is restored on exit."""

    def __init__(self, name):
        assert name in ("contract", "fast")
        self.name = name

    def __enter__(self):
        global _LIB
        self.saved = _LIB
        if self.name == "fast":
            so = os.path.join(_HERE, "libmirres_oracle_fast.so")
            subprocess.check_call(["make", "-C", _HERE, "-s", "fast"])
            _LIB = ctypes.CDLL(so)
        else:
            _LIB = None
            lib()
        return self

    def __exit__(self, *a):
        global _LIB
        _LIB = self.saved


def _p(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "oracle needs contiguous numpy arrays"
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def set_threads(n):
    """OpenMP threads of every oracle entry point from now on (overrides an inherited OMP_NUM_THREADS)."""
    lib().orc_set_threads(int(n))


def threads_in_use():
    """Threads an OpenMP region of the oracle runs with right now (measured inside a parallel region)."""
    return int(lib().orc_threads_in_use())


def new_counters():
    return np.zeros(12, dtype=np.int64)


def fpmath(op, x, y=None):
    x = _f32(x)
    y = _f32(y) if y is not None else x
    out = np.empty_like(x)
    rc = lib().orc_fpmath_eval(int(op), _p(x), _p(y), _p(out), int(x.size))
    assert rc == 0
    return out


def fp_contract_selftest():
    out = np.zeros(1, np.float32)
    return lib().orc_fpmath_eval(100, _p(out), _p(out), _p(out), 1)


class Bvh:
    def __init__(self, vert, tri):
        self.vert = _f32(vert)
        self.tri = _i32(tri)
        F = self.tri.shape[0]
        self.F = F
        self.info = np.zeros((2 * F - 1, 3), np.int32)
        self.aabb = np.zeros((2 * F - 1, 6), np.float32)
        self.sorted_codes = np.zeros((F, 2), np.int32)
        self.parent = np.zeros(2 * F - 1, np.int32)
        self.extent = np.zeros(6, np.float32)
        rc = lib().orc_bvh_build(_p(self.vert), int(self.vert.shape[0]), _p(self.tri), int(F), _p(self.info),
                                 _p(self.aabb), _p(self.sorted_codes), _p(self.parent), _p(self.extent))
        assert rc == 0

    def args(self):
        return (_p(self.info), _p(self.aabb), _p(self.vert), _p(self.tri))


def trace(bvh, org, dirs, counters=None):
    org = _f32(org)
    dirs = _f32(dirs)
    n = org.shape[0]
    hit = np.zeros(n, np.int32)
    t = np.zeros(n, np.float32)
    pos = np.zeros((n, 3), np.float32)
    nrm = np.zeros((n, 3), np.float32)
    prim = np.zeros(n, np.int32)
    rc = lib().orc_trace(*bvh.args(), _p(org), _p(dirs), int(n), _p(hit), _p(t), _p(pos), _p(nrm), _p(prim),
                         _p(counters))
    assert rc == 0
    return hit, t, pos, nrm, prim


def env_build_distribution(env_tex, W, H):
    env_tex = _f32(env_tex)
    pdf_ = np.zeros((W * H, 1), np.float32)
    cdf_ = np.zeros((H * (W + 1), 1), np.float32)
    mpdf_ = np.zeros((H, 1), np.float32)
    mcdf_ = np.zeros((H + 1, 1), np.float32)
    rc = lib().orc_env_build_distribution(_p(env_tex), int(W), int(H), _p(pdf_), _p(cdf_), _p(mpdf_), _p(mcdf_))
    assert rc == 0
    return pdf_, cdf_, mpdf_, mcdf_


def neighbor_offsets(count):
    out = np.zeros((count * 2, 1), np.float32)
    assert lib().orc_neighbor_offsets(int(count), _p(out)) == 0
    return out


def light_tiles(env_tex, W, H, dist, frame_index, tile_count=128, tile_size=1024):
    pdf_, cdf_, mpdf_, mcdf_ = dist
    n = tile_count * tile_size
    light_data = np.zeros((n, 3), np.float32)
    light_uv = np.zeros((n, 2), np.int32)
    light_pdf = np.zeros((n, 1), np.float32)
    rc = lib().orc_light_tiles(_p(_f32(env_tex)), int(W), int(H), _p(pdf_), _p(cdf_), _p(mpdf_), _p(mcdf_),
                               ctypes.c_uint32(frame_index & 0xFFFFFFFF), int(tile_count), int(tile_size),
                               _p(light_data), _p(light_uv), _p(light_pdf))
    assert rc == 0
    return light_data, light_uv, light_pdf


def new_reservoirs(n):
    return (np.zeros((n, 3), np.float32), np.zeros((n, 1), np.float32), np.zeros((n, 1), np.int32),
            np.zeros((n, 1), np.float32))


def _res(r):
    return (_p(r[0]), _p(r[1]), _p(r[2]), _p(r[3]))


def set_provenance(res_tag=None, prev_tag=None):
    """uint8 [n] arrays or None: the sample-provenance tags the next pass records / reads (orc_kernels.cpp)."""
    lib().orc_set_provenance(_p(res_tag), _p(prev_tag))


def provenance_violations():
    f = lib().orc_provenance_violations
    f.restype = ctypes.c_longlong
    return int(f())


def initial_resampling(bvh, pos_map, res, env_tex, W, H, fx, fy, frame_index, occ, normal_depth, brdf_map, ray_dir,
                       dist, tiles, tile_count=128, tile_size=1024, screen_tile=8, n_light=32, n_brdf=1,
                       counters=None):
    pdf_, cdf_, mpdf_, mcdf_ = dist
    rc = lib().orc_initial_resampling(*bvh.args(), _p(pos_map), *_res(res), _p(env_tex), int(W), int(H), int(fx),
                                      int(fy), ctypes.c_uint32(frame_index & 0xFFFFFFFF), _p(occ), _p(normal_depth),
                                      _p(brdf_map), _p(ray_dir), _p(pdf_), _p(cdf_), _p(mpdf_), _p(mcdf_),
                                      _p(tiles[0]), _p(tiles[1]), _p(tiles[2]), int(tile_count), int(tile_size),
                                      int(screen_tile), int(n_light), int(n_brdf), _p(counters))
    assert rc == 0


def temporal_resampling(res, prev, env_tex, W, H, fx, fy, frame_index, occ, normal_depth, brdf_map, ray_dir,
                        prev_occ, prev_normal_depth, prev_brdf_map, prev_ray_dir, motion=None, max_history=20):
    rc = lib().orc_temporal_resampling(*_res(res), *_res(prev), _p(env_tex), int(W), int(H), int(fx), int(fy),
                                       ctypes.c_uint32(frame_index & 0xFFFFFFFF), _p(occ), _p(normal_depth),
                                       _p(brdf_map), _p(ray_dir), _p(prev_occ), _p(prev_normal_depth),
                                       _p(prev_brdf_map), _p(prev_ray_dir), _p(motion), int(max_history))
    assert rc == 0


def spatial_resampling(bvh, pos_map, res, prev, neighbor_offsets_, env_tex, W, H, fx, fy, frame_index, occ,
                       normal_depth, brdf_map, ray_dir, offset_count=8192, neighbor_count=5, gather_radius=30.0,
                       counters=None):
    rc = lib().orc_spatial_resampling(*bvh.args(), _p(pos_map), *_res(res), *_res(prev), _p(neighbor_offsets_),
                                      _p(env_tex), int(W), int(H), int(fx), int(fy),
                                      ctypes.c_uint32(frame_index & 0xFFFFFFFF), _p(occ), _p(normal_depth),
                                      _p(brdf_map), _p(ray_dir), int(offset_count), int(neighbor_count),
                                      ctypes.c_float(gather_radius), _p(counters))
    assert rc == 0


def final_visibility(bvh, res, fx, fy, pos_map, vis_map, counters=None):
    rc = lib().orc_final_visibility(*bvh.args(), _p(res[0]), int(fx), int(fy), _p(pos_map), _p(vis_map),
                                    _p(counters))
    assert rc == 0


def eval_final_fwd(res, env_tex, W, H, fx, fy, fs_dir, fs_dist, fs_Li, vis_map):
    rc = lib().orc_eval_final_fwd(*_res(res), _p(env_tex), int(W), int(H), int(fx), int(fy), _p(fs_dir), _p(fs_dist),
                                  _p(fs_Li), _p(vis_map))
    assert rc == 0


def final_shading_fwd(fs_dir, fs_dist, fs_Li, env_tex, W, H, fx, fy, occ, normal, ray_dir, diffuse, rough_metal):
    n = fx * fy
    color = np.zeros((n, 3), np.float32)
    diff = np.zeros((n, 3), np.float32)
    spec = np.zeros((n, 3), np.float32)
    rc = lib().orc_final_shading_fwd(_p(fs_dir), _p(fs_dist), _p(fs_Li), _p(env_tex), int(W), int(H), int(fx), int(fy),
                                     _p(occ), _p(normal), _p(ray_dir), _p(diffuse), _p(rough_metal), _p(color),
                                     _p(diff), _p(spec))
    assert rc == 0
    return color, diff, spec


def bounce_first(bvh, frame_index, bounce_count, max_bounce, fx, fy, occ, pos_map, normal, ray_dir, prd, diffuse,
                 rough_metal, new_pos, new_ray_d, new_occ, new_normal, counters=None):
    rc = lib().orc_bounce_first(*bvh.args(), ctypes.c_uint32(frame_index & 0xFFFFFFFF), ctypes.c_uint32(bounce_count),
                                int(max_bounce), int(fx), int(fy), _p(occ), _p(pos_map), _p(normal), _p(ray_dir),
                                _p(prd), _p(diffuse), _p(rough_metal), _p(new_pos), _p(new_ray_d), _p(new_occ),
                                _p(new_normal), _p(counters))
    assert rc == 0


def bounce_shade(bvh, frame_index, bounce_count, max_bounce, fx, fy, env_tex, W, H, dist, occ, pos_map, normal,
                 ray_dir, prd, diffuse, rough_metal, color, diff_color, spec_color, new_pos, new_ray_d, new_occ,
                 new_normal, counters=None):
    pdf_, cdf_, mpdf_, mcdf_ = dist
    rc = lib().orc_bounce_shade(*bvh.args(), ctypes.c_uint32(frame_index & 0xFFFFFFFF), ctypes.c_uint32(bounce_count),
                                int(max_bounce), int(fx), int(fy), _p(env_tex), int(W), int(H), _p(pdf_), _p(cdf_),
                                _p(mpdf_), _p(mcdf_), _p(occ), _p(pos_map), _p(normal), _p(ray_dir), _p(prd),
                                _p(diffuse), _p(rough_metal), _p(color), _p(diff_color), _p(spec_color), _p(new_pos),
                                _p(new_ray_d), _p(new_occ), _p(new_normal), _p(counters))
    assert rc == 0


def eaw_fwd(c_phi, n_phi, p_phi, fx, fy, step_width, occ, color, normal, pos):
    out = np.zeros((fx * fy, 3), np.float32)
    rc = lib().orc_eaw_fwd(ctypes.c_float(c_phi), ctypes.c_float(n_phi), ctypes.c_float(p_phi), int(fx), int(fy),
                           int(step_width), _p(occ), _p(color), _p(normal), _p(pos), _p(out))
    assert rc == 0
    return out


def normal_ao(fx, fy, occ, normal):
    out = np.zeros((fx * fy, 3), np.float32)
    assert lib().orc_normal_ao(int(fx), int(fy), _p(occ), _p(normal), _p(out)) == 0
    return out


def eval_final_taps(res_ld, W, H):
    res_ld = _f32(res_ld)
    n = res_ld.shape[0]
    taps = np.zeros((n, 4), np.int32)
    uv = np.zeros((n, 2), np.float32)
    valid = np.zeros(n, np.int32)
    assert lib().orc_eval_final_taps(_p(res_ld), int(W), int(H), int(n), _p(taps), _p(uv), _p(valid)) == 0
    return taps, uv, valid


def bilateral_fwd(fx, fy, sigma, col, nrm, zdz):
    """nerf/renderutils/c_src/denoising.cu:14-70 -> out [N,4]"""
    out = np.zeros((fx * fy, 4), np.float32)
    rc = lib().orc_bilateral_fwd(int(fx), int(fy), ctypes.c_float(sigma), _p(_f32(col)), _p(_f32(nrm)), _p(_f32(zdz)), _p(out))
    assert rc == 0
    return out


def bilateral_bwd(fx, fy, sigma, nrm, zdz, out_grad):
    """nerf/renderutils/c_src/denoising.cu:72-130 -> col_grad [N,3]"""
    g = np.zeros((fx * fy, 3), np.float32)
    rc = lib().orc_bilateral_bwd(int(fx), int(fy), ctypes.c_float(sigma), _p(_f32(nrm)), _p(_f32(zdz)), _p(_f32(out_grad)), _p(g))
    assert rc == 0
    return g
