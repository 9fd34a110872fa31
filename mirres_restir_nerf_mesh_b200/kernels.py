"""Typed Python front-end of the C ABI: torch tensors in, stream-ordered launches out.

One method per entry point of include/mirres_b200.h.  Tensors must be dense, of the documented dtype, and live on
the GPU; the launch goes to torch's current CUDA stream (the reference launches on the same stream through
slangpy's launchRaw).  Nothing here computes anything: it validates and forwards.
"""
import ctypes

import torch

from . import _lib


class AbiError(RuntimeError):
    pass


_ERR = {-1: "null pointer", -2: "bad shape/size argument", -3: "misaligned pointer", -4: "scratch too small",
        -5: "input/output aliasing"}


class Kernels:
    """Binds a loaded library.  `require_cuda=False` exists only for the test-only host-check flavour."""

    def __init__(self, lib=None, require_cuda=True):
        self.lib = lib if lib is not None else _lib.load()
        self.require_cuda = require_cuda

    # -- helpers ---------------------------------------------------------------------------------------------
    def _p(self, t, dtype=None, optional=False):
        if t is None:
            if optional:
                return None
            raise AbiError("required tensor is None")
        if not isinstance(t, torch.Tensor):
            raise AbiError("expected a torch.Tensor, got %r" % type(t))
        if self.require_cuda and not t.is_cuda:
            raise AbiError("mirres-b200 kernels need CUDA tensors (no CPU fallback)")
        if dtype is not None and t.dtype != dtype:
            raise AbiError("expected dtype %s, got %s" % (dtype, t.dtype))
        if not t.is_contiguous():
            raise AbiError("tensor must be contiguous (call .contiguous() at the boundary)")
        return ctypes.c_void_p(t.data_ptr())

    def _f(self, t, optional=False):
        return self._p(t, torch.float32, optional)

    def _i(self, t, optional=False):
        return self._p(t, torch.int32, optional)

    def _stream(self):
        if not self.require_cuda:
            return None
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _check(rc, name):
        if rc != 0:
            if rc <= -100:
                raise AbiError("%s: CUDA launch error %d" % (name, -rc - 100))
            raise AbiError("%s: %s (%d)" % (name, _ERR.get(rc, "error"), rc))

    @staticmethod
    def _u32(x):
        return ctypes.c_uint(int(x) & 0xFFFFFFFF)

    # -- tuning ------------------------------------------------------------------------------------------------
    TUNE_ANY_BLOCKS, TUNE_CLOSEST_BLOCKS, TUNE_MIXED_BLOCKS, TUNE_CLOSEST_SPLIT, TUNE_ANY_TOP = 0, 1, 2, 3, 4

    def get_tuning(self, key):
        v = self.lib.mirres_get_tuning(int(key))
        self._check(min(v, 0), "mirres_get_tuning")
        return v

    def set_tuning(self, key, value):
        self._check(self.lib.mirres_set_tuning(int(key), int(value)), "mirres_set_tuning")

    # -- BVH ---------------------------------------------------------------------------------------------------
    def bvh_sizes(self, F):
        return (self.lib.mirres_bvh_scratch_bytes(F), self.lib.mirres_bvh_packed_node_bytes(F),
                self.lib.mirres_bvh_packed_tri_bytes(F))

    def bvh_build(self, vert, tri, info, aabb, packed_nodes, packed_tris, scratch, sorted_codes=None):
        F = tri.shape[0]
        rc = self.lib.mirres_bvh_build(self._f(vert), vert.shape[0], self._i(tri), F, self._i(info), self._f(aabb),
                                       self._p(packed_nodes, torch.uint8, True), self._p(packed_tris, torch.uint8, True),
                                       self._i(sorted_codes, True), self._p(scratch, torch.uint8), scratch.numel(),
                                       self._stream())
        self._check(rc, "mirres_bvh_build")

    def bvh_elements(self, vert, tri, ele_primitiveIdx, ele_aabb):
        rc = self.lib.mirres_bvh_elements(self._f(vert), self._i(tri), tri.shape[0], self._i(ele_primitiveIdx, True),
                                          self._f(ele_aabb), self._stream())
        self._check(rc, "mirres_bvh_elements")

    def bvh_morton(self, ele_aabb, extent, morton_codes_ele):
        rc = self.lib.mirres_bvh_morton(self._f(ele_aabb), ele_aabb.shape[0], *[float(x) for x in extent],
                                        self._i(morton_codes_ele), self._stream())
        self._check(rc, "mirres_bvh_morton")

    def bvh_sort(self, pairs, scratch):
        rc = self.lib.mirres_bvh_sort(self._i(pairs), pairs.shape[0], self._p(scratch, torch.uint8), scratch.numel(),
                                      self._stream())
        self._check(rc, "mirres_bvh_sort")

    def bvh_hierarchy_refit(self, sorted_pairs, ele_aabb, info, aabb, scratch):
        rc = self.lib.mirres_bvh_hierarchy_refit(self._i(sorted_pairs), self._f(ele_aabb), sorted_pairs.shape[0],
                                                 self._i(info), self._f(aabb), self._p(scratch, torch.uint8),
                                                 scratch.numel(), self._stream())
        self._check(rc, "mirres_bvh_hierarchy_refit")

    def bvh_pack(self, info, aabb, vert, tri, packed_nodes, packed_tris):
        rc = self.lib.mirres_bvh_pack(self._i(info), self._f(aabb), self._f(vert), self._i(tri), tri.shape[0],
                                      self._p(packed_nodes, torch.uint8), self._p(packed_tris, torch.uint8),
                                      self._stream())
        self._check(rc, "mirres_bvh_pack")

    # -- rays --------------------------------------------------------------------------------------------------
    def trace_closest(self, packed, org, dirs, hit, t=None, pos=None, normal=None, prim=None, visits=None):
        rc = self.lib.mirres_trace_closest(self._p(packed[0]), self._p(packed[1]), self._f(org), self._f(dirs),
                                           org.shape[0], self._i(hit), self._f(t, True), self._f(pos, True),
                                           self._f(normal, True), self._i(prim, True), self._i(visits, True),
                                           self._stream())
        self._check(rc, "mirres_trace_closest")

    def trace_any(self, packed, org, dirs, hit, visits=None):
        rc = self.lib.mirres_trace_any(self._p(packed[0]), self._p(packed[1]), self._f(org), self._f(dirs), org.shape[0],
                                       self._i(hit), self._i(visits, True), self._stream())
        self._check(rc, "mirres_trace_any")

    # -- environment -------------------------------------------------------------------------------------------
    def env_build_distribution(self, env_tex, W, H, pdf_, cdf_, mpdf_, mcdf_, row_scratch):
        rc = self.lib.mirres_env_build_distribution(self._f(env_tex), W, H, self._f(pdf_), self._f(cdf_), self._f(mpdf_),
                                                    self._f(mcdf_), self._f(row_scratch), self._stream())
        self._check(rc, "mirres_env_build_distribution")

    def env_weights(self, env_tex, W, H, weight):
        self._check(self.lib.mirres_env_weights(self._f(env_tex), W, H, self._f(weight), self._stream()),
                    "mirres_env_weights")

    def env_distribution2d(self, W, H, pdf_, cdf_):
        self._check(self.lib.mirres_env_distribution2d(W, H, self._f(pdf_), self._f(cdf_), self._stream()),
                    "mirres_env_distribution2d")

    def neighbor_offsets(self, count, out):
        self._check(self.lib.mirres_neighbor_offsets(count, self._f(out), self._stream()), "mirres_neighbor_offsets")

    def light_tiles(self, env_tex, W, H, dist, frame_index, tile_count, tile_size, light_data, light_uv, light_pdf,
                    light_cache=None, frame_offset=None):
        rc = self.lib.mirres_light_tiles(self._f(env_tex), W, H, self._f(dist[0]), self._f(dist[1]), self._f(dist[2]),
                                         self._f(dist[3]), self._u32(frame_index), tile_count, tile_size,
                                         self._f(light_data), self._i(light_uv), self._f(light_pdf), self._f(light_cache, True),
                                         self._i(frame_offset, True), self._stream())
        self._check(rc, "mirres_light_tiles")

    # -- wavefront workspace -----------------------------------------------------------------------------------
    def workspace_bytes(self, n_pixels):
        return self.lib.mirres_workspace_bytes(int(n_pixels))

    def workspace_prepare(self, occ, ws):
        rc = self.lib.mirres_workspace_prepare(self._f(occ), occ.shape[0], self._p(ws, torch.uint8), ws.numel(),
                                               self._stream())
        self._check(rc, "mirres_workspace_prepare")

    def _ws(self, ws):
        return (self._p(ws, torch.uint8), ws.numel())

    # -- ReSTIR ------------------------------------------------------------------------------------------------
    def _res(self, r):
        return (self._f(r[0]), self._f(r[1]), self._i(r[2]), self._f(r[3]))

    class _Tags:
        """mirres_set_visibility_tags around ONE pass (include/mirres_b200.h): the tags are per-call arguments here, the
        library's thread-local setting never outlives the call."""

        def __init__(self, k, n, res_tag, prev_tag):
            self.k, self.on = k, (res_tag is not None or prev_tag is not None)
            for t in (res_tag, prev_tag):
                if t is not None and t.numel() != n:
                    raise AbiError("visibility tag: expected %d bytes, got %d" % (n, t.numel()))
            self.args = (k._p(res_tag, torch.uint8, True), k._p(prev_tag, torch.uint8, True))

        def __enter__(self):
            if self.on:
                self.k._check(self.k.lib.mirres_set_visibility_tags(*self.args), "mirres_set_visibility_tags")

        def __exit__(self, *a):
            if self.on:
                self.k.lib.mirres_set_visibility_tags(None, None)

    def initial_resampling(self, packed, pos_map, res, env_tex, W, H, fx, fy, frame_index, occ, normal_depth, brdf_map,
                           ray_dir, pdf_, mpdf_, light_data, light_pdf, ws, tile_count=128, tile_size=1024, screen_tile=8,
                           n_light=32, n_brdf=1, light_cache=None, vis_tag=None):
        with self._Tags(self, fx * fy, vis_tag, None):
            rc = self._initial_resampling(packed, pos_map, res, env_tex, W, H, fx, fy, frame_index, occ, normal_depth,
                                          brdf_map, ray_dir, pdf_, mpdf_, light_data, light_pdf, ws, tile_count, tile_size,
                                          screen_tile, n_light, n_brdf, light_cache)
        self._check(rc, "mirres_initial_resampling")

    def _initial_resampling(self, packed, pos_map, res, env_tex, W, H, fx, fy, frame_index, occ, normal_depth, brdf_map,
                            ray_dir, pdf_, mpdf_, light_data, light_pdf, ws, tile_count, tile_size, screen_tile, n_light,
                            n_brdf, light_cache):
        return self.lib.mirres_initial_resampling(self._p(packed[0]), self._p(packed[1]), self._f(pos_map), *self._res(res),
                                                self._f(env_tex), W, H, fx, fy, self._u32(frame_index), self._f(occ),
                                                self._f(normal_depth), self._f(brdf_map), self._f(ray_dir), self._f(pdf_),
                                                self._f(mpdf_), self._f(light_data), self._f(light_pdf),
                                                self._f(light_cache, True), tile_count,
                                                tile_size, screen_tile, n_light, n_brdf, *self._ws(ws), self._stream())

    def temporal_resampling(self, res, prev, env_tex, W, H, fx, fy, frame_index, occ, normal_depth, brdf_map, ray_dir,
                            prev_occ, prev_normal_depth, prev_brdf_map, prev_ray_dir, ws, motion=None, max_history=20,
                            vis_tag=None, prev_vis_tag=None):
        with self._Tags(self, fx * fy, vis_tag, prev_vis_tag):
            rc = self.lib.mirres_temporal_resampling(*self._res(res), *self._res(prev), self._f(env_tex), W, H, fx, fy,
                                                     self._u32(frame_index), self._f(occ), self._f(normal_depth),
                                                     self._f(brdf_map), self._f(ray_dir), self._f(prev_occ),
                                                     self._f(prev_normal_depth), self._f(prev_brdf_map),
                                                     self._f(prev_ray_dir), self._f(motion, True), max_history,
                                                     *self._ws(ws), self._stream())
        self._check(rc, "mirres_temporal_resampling")

    def spatial_resampling(self, packed, pos_map, res, prev, neighbor_offsets, env_tex, W, H, fx, fy, frame_index, occ,
                           normal_depth, brdf_map, ray_dir, ws, offset_count=8192, neighbor_count=5, gather_radius=30.0,
                           vis_tag=None, prev_vis_tag=None):
        with self._Tags(self, fx * fy, vis_tag, prev_vis_tag):
            rc = self.lib.mirres_spatial_resampling(self._p(packed[0]), self._p(packed[1]), self._f(pos_map),
                                                    *self._res(res), *self._res(prev), self._f(neighbor_offsets),
                                                    self._f(env_tex), W, H, fx, fy, self._u32(frame_index), self._f(occ),
                                                    self._f(normal_depth), self._f(brdf_map), self._f(ray_dir), offset_count,
                                                    neighbor_count, float(gather_radius), *self._ws(ws), self._stream())
        self._check(rc, "mirres_spatial_resampling")

    def final_visibility(self, packed, res_ld, fx, fy, pos_map, vis_map, ws, vis_tag=None):
        with self._Tags(self, fx * fy, vis_tag, None):
            rc = self.lib.mirres_final_visibility(self._p(packed[0]), self._p(packed[1]), self._f(res_ld), fx, fy,
                                                  self._f(pos_map), self._f(vis_map), *self._ws(ws), self._stream())
        self._check(rc, "mirres_final_visibility")

    def eval_final_fwd(self, res, env_tex, W, H, fx, fy, fs_dir, fs_dist, fs_Li, vis_map):
        rc = self.lib.mirres_eval_final_fwd(*self._res(res), self._f(env_tex), W, H, fx, fy, self._f(fs_dir),
                                            self._f(fs_dist), self._f(fs_Li), self._f(vis_map), self._stream())
        self._check(rc, "mirres_eval_final_fwd")

    def eval_final_bwd(self, res, W, H, fx, fy, vis_map, grad_Li, grad_env):
        rc = self.lib.mirres_eval_final_bwd(*self._res(res), W, H, fx, fy, self._f(vis_map), self._f(grad_Li),
                                            self._f(grad_env), self._stream())
        self._check(rc, "mirres_eval_final_bwd")

    # -- shading -----------------------------------------------------------------------------------------------
    def final_shading_fwd(self, fs_dir, fs_dist, fs_Li, env_tex, W, H, fx, fy, occ, normal, ray_dir, diffuse, rough_metal,
                          color, diff_light, spec_light):
        rc = self.lib.mirres_final_shading_fwd(self._f(fs_dir), self._f(fs_dist), self._f(fs_Li), self._f(env_tex), W, H,
                                               fx, fy, self._f(occ), self._f(normal), self._f(ray_dir), self._f(diffuse),
                                               self._f(rough_metal), self._f(color), self._f(diff_light),
                                               self._f(spec_light), self._stream())
        self._check(rc, "mirres_final_shading_fwd")

    def final_shading_bwd(self, fs_dir, fs_dist, fs_Li, fx, fy, occ, normal, ray_dir, diffuse, rough_metal, g_color,
                          g_diff, g_spec, g_normal, g_diffuse, g_rough_metal, g_Li):
        rc = self.lib.mirres_final_shading_bwd(self._f(fs_dir), self._f(fs_dist), self._f(fs_Li), fx, fy, self._f(occ),
                                               self._f(normal), self._f(ray_dir), self._f(diffuse), self._f(rough_metal),
                                               self._f(g_color), self._f(g_diff), self._f(g_spec), self._f(g_normal),
                                               self._f(g_diffuse), self._f(g_rough_metal), self._f(g_Li), self._stream())
        self._check(rc, "mirres_final_shading_bwd")

    def bounce_first(self, packed, frame_index, bounce_count, max_bounce, fx, fy, occ, pos_map, normal, ray_dir, prd,
                     diffuse, rough_metal, new_pos, new_ray_d, new_occ, new_normal, ws):
        rc = self.lib.mirres_bounce_first(self._p(packed[0]), self._p(packed[1]), self._u32(frame_index),
                                          self._u32(bounce_count), max_bounce, fx, fy, self._f(occ), self._f(pos_map),
                                          self._f(normal), self._f(ray_dir), self._f(prd), self._f(diffuse),
                                          self._f(rough_metal), self._f(new_pos), self._f(new_ray_d), self._f(new_occ),
                                          self._f(new_normal), *self._ws(ws), self._stream())
        self._check(rc, "mirres_bounce_first")

    def bounce_shade(self, packed, frame_index, bounce_count, max_bounce, fx, fy, env_tex, W, H, dist, occ, pos_map,
                     normal, ray_dir, prd, diffuse, rough_metal, color, diff_color, spec_color, new_pos, new_ray_d,
                     new_occ, new_normal, ws):
        rc = self.lib.mirres_bounce_shade(self._p(packed[0]), self._p(packed[1]), self._u32(frame_index),
                                          self._u32(bounce_count), max_bounce, fx, fy, self._f(env_tex), W, H,
                                          self._f(dist[0]), self._f(dist[1]), self._f(dist[2]), self._f(dist[3]),
                                          self._f(occ), self._f(pos_map), self._f(normal), self._f(ray_dir), self._f(prd),
                                          self._f(diffuse), self._f(rough_metal), self._f(color), self._f(diff_color),
                                          self._f(spec_color), self._f(new_pos), self._f(new_ray_d), self._f(new_occ),
                                          self._f(new_normal), *self._ws(ws), self._stream())
        self._check(rc, "mirres_bounce_shade")

    # -- denoiser ----------------------------------------------------------------------------------------------
    def eaw_fwd(self, c_phi, n_phi, p_phi, fx, fy, step_width, occ, color, normal, pos, out_color):
        rc = self.lib.mirres_eaw_fwd(float(c_phi), float(n_phi), float(p_phi), fx, fy, float(step_width), self._f(occ),
                                     self._f(color), self._f(normal), self._f(pos), self._f(out_color), self._stream())
        self._check(rc, "mirres_eaw_fwd")

    def eaw_bwd(self, c_phi, n_phi, p_phi, fx, fy, step_width, occ, color, normal, pos, out_color, g_out, g_color,
                g_normal, g_pos, cum_w_scratch):
        rc = self.lib.mirres_eaw_bwd(float(c_phi), float(n_phi), float(p_phi), fx, fy, float(step_width), self._f(occ),
                                     self._f(color), self._f(normal), self._f(pos), self._f(out_color), self._f(g_out),
                                     self._f(g_color), self._f(g_normal), self._f(g_pos), self._f(cum_w_scratch),
                                     self._stream())
        self._check(rc, "mirres_eaw_bwd")

    def _ptr_array(self, tensors, optional=False):
        arr = (ctypes.c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            p = self._f(t, optional)
            arr[i] = None if p is None else p.value
        return arr

    def eaw_fwd_multi(self, c_phi, n_phi, p_phi, fx, fy, step_width, occ, normal, pos, colors, outs, cum_w=None):
        rc = self.lib.mirres_eaw_fwd_multi(float(c_phi), float(n_phi), float(p_phi), fx, fy, float(step_width),
                                           self._f(occ), self._f(normal), self._f(pos), len(colors),
                                           self._ptr_array(colors), self._ptr_array(outs),
                                           None if cum_w is None else self._ptr_array(cum_w, True), self._stream())
        self._check(rc, "mirres_eaw_fwd_multi")

    def eaw_bwd_multi(self, c_phi, n_phi, p_phi, fx, fy, step_width, occ, normal, pos, colors, outs, cum_w, g_outs,
                      g_colors, g_normal_sum, g_pos_sum):
        rc = self.lib.mirres_eaw_bwd_multi(float(c_phi), float(n_phi), float(p_phi), fx, fy, float(step_width),
                                           self._f(occ), self._f(normal), self._f(pos), len(colors),
                                           self._ptr_array(colors), self._ptr_array(outs), self._ptr_array(cum_w),
                                           self._ptr_array(g_outs), self._ptr_array(g_colors),
                                           self._f(g_normal_sum, True), self._f(g_pos_sum, True), self._stream())
        self._check(rc, "mirres_eaw_bwd_multi")

    def bilateral_fwd(self, fx, fy, sigma, col, nrm, zdz, out):
        rc = self.lib.mirres_bilateral_fwd(int(fx), int(fy), float(sigma), self._f(col), self._f(nrm), self._f(zdz),
                                           self._f(out), self._stream())
        self._check(rc, "mirres_bilateral_fwd")

    def bilateral_bwd(self, fx, fy, sigma, nrm, zdz, out_grad, col_grad):
        rc = self.lib.mirres_bilateral_bwd(int(fx), int(fy), float(sigma), self._f(nrm), self._f(zdz), self._f(out_grad),
                                           self._f(col_grad), self._stream())
        self._check(rc, "mirres_bilateral_bwd")

    def normal_ao(self, fx, fy, occ, normal, out_ao):
        rc = self.lib.mirres_normal_ao(fx, fy, self._f(occ), self._f(normal), self._f(out_ao), self._stream())
        self._check(rc, "mirres_normal_ao")

    # -- G-buffer producer / gradient scatter ------------------------------------------------------------------------
    def gbuffer_primary(self, packed, org, dirs, occ, pos, normal, depth, prim=None, bary=None, vnormal=None, tri=None,
                        ws=None, geom_normal=None):
        wsp = (None, 0) if ws is None else self._ws(ws)
        rc = self.lib.mirres_gbuffer_primary(self._p(packed[0]), self._p(packed[1]), self._f(org), self._f(dirs),
                                             org.shape[0], self._f(vnormal, True), self._i(tri, True), self._f(occ),
                                             self._f(pos), self._f(normal), self._f(depth), self._i(prim, True),
                                             self._f(bary, True), self._f(geom_normal, True), wsp[0], wsp[1],
                                             self._stream())
        self._check(rc, "mirres_gbuffer_primary")

    def prepare_maps(self, occ, normal, depth, diffuse, rough_metal, ray_dir, normal_depth, brdf_map, ray_out):
        rc = self.lib.mirres_prepare_maps(occ.shape[0], self._f(occ), self._f(normal), self._f(depth), self._f(diffuse),
                                          self._f(rough_metal), self._f(ray_dir), self._f(normal_depth),
                                          self._f(brdf_map), self._f(ray_out), self._stream())
        self._check(rc, "mirres_prepare_maps")

    def vertex_normals_fwd(self, vert, tri, vsum, vnrm):
        rc = self.lib.mirres_vertex_normals_fwd(self._f(vert), vert.shape[0], self._i(tri), tri.shape[0], self._f(vsum),
                                                self._f(vnrm), self._stream())
        self._check(rc, "mirres_vertex_normals_fwd")

    def vertex_normals_bwd(self, vert, tri, vsum, grad_vnrm, grad_vert):
        rc = self.lib.mirres_vertex_normals_bwd(self._f(vert), vert.shape[0], self._i(tri), tri.shape[0], self._f(vsum),
                                                self._f(grad_vnrm), self._f(grad_vert), self._stream())
        self._check(rc, "mirres_vertex_normals_bwd")

    def _rows3(self, t, n):
        """[n,3] (any row stride) or [1,3] fp32 rows -> (pointer, row stride in floats, tensor kept alive)."""
        if not isinstance(t, torch.Tensor) or t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != 3:
            raise AbiError("expected a float32 [rows, 3] tensor")
        if self.require_cuda and not t.is_cuda:
            raise AbiError("mirres-b200 kernels need CUDA tensors (no CPU fallback)")
        if t.shape[0] not in (1, n):
            raise AbiError("expected %d rows (or 1 broadcast row), got %d" % (n, t.shape[0]))
        if t.stride(1) != 1 or t.stride(0) < 0:
            t = t.contiguous()
        return ctypes.c_void_p(t.data_ptr()), (0 if t.shape[0] == 1 and n != 1 else int(t.stride(0))), t

    def shading_normal_fwd(self, n, inputs, two_sided, opengl, out):
        """inputs = (pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm), each n rows or one broadcast row"""
        args, keep = [], []
        for t in inputs:
            ptr, rs, t2 = self._rows3(t, n)
            args += [ptr, rs]
            keep.append(t2)
        rc = self.lib.mirres_shading_normal_fwd(int(n), *args, int(bool(two_sided)), int(bool(opengl)), self._f(out),
                                                self._stream())
        self._check(rc, "mirres_shading_normal_fwd")

    def shading_normal_bwd(self, n, inputs, two_sided, opengl, grad_out, grads):
        """grads: six [n,3] tensors or None (skipped)"""
        args, keep = [], []
        for t in inputs:
            ptr, rs, t2 = self._rows3(t, n)
            args += [ptr, rs]
            keep.append(t2)
        rc = self.lib.mirres_shading_normal_bwd(int(n), *args, int(bool(two_sided)), int(bool(opengl)),
                                                self._f(grad_out), *[self._f(g, True) for g in grads], self._stream())
        self._check(rc, "mirres_shading_normal_bwd")

    def interpolate_bwd(self, grad, prim, bary, tri, out):
        rc = self.lib.mirres_interpolate_bwd(self._f(grad), grad.shape[0], grad.shape[1], self._i(prim),
                                             self._f(bary, True), self._i(tri), tri.shape[0], self._f(out),
                                             self._stream())
        self._check(rc, "mirres_interpolate_bwd")

    # -- host-side chains of the spp loop as single launches -----------------------------------------
    def material_procedural(self, pos, occ, mode, metallic, kd, rough_metal, scale=None):
        sc = None if scale is None else (ctypes.c_float * 3)(*[float(x) for x in scale])
        rc = self.lib.mirres_material_procedural(pos.shape[0], self._f(pos), self._f(occ, True), int(mode), float(metallic),
                                                 sc, self._f(kd), self._f(rough_metal), self._stream())
        self._check(rc, "mirres_material_procedural")

    def sum_images(self, srcs, dst, divisor=0.0, accumulate=False):
        n = dst.numel()
        for s in srcs:
            if s.numel() != n:
                raise AbiError("mirres_sum_images: size mismatch")
        rc = self.lib.mirres_sum_images(n, len(srcs), self._ptr_array(srcs) if srcs else None, float(divisor),
                                        1 if accumulate else 0, self._f(dst), self._stream())
        self._check(rc, "mirres_sum_images")

    def composite_fwd(self, occ, kd, rough_metal, dd, ds, di, out):
        rc = self.lib.mirres_composite_fwd(occ.shape[0], self._f(occ), self._f(kd), self._f(rough_metal), self._f(dd),
                                           self._f(ds), self._f(di), self._f(out), self._stream())
        self._check(rc, "mirres_composite_fwd")

    def composite_bwd(self, occ, kd, rough_metal, dd, ds, di, g_out, g_kd, g_rm, g_dd, g_ds):
        rc = self.lib.mirres_composite_bwd(occ.shape[0], self._f(occ), self._f(kd), self._f(rough_metal), self._f(dd),
                                           self._f(ds), self._f(di), self._f(g_out), self._f(g_kd), self._f(g_rm),
                                           self._f(g_dd), self._f(g_ds), self._stream())
        self._check(rc, "mirres_composite_bwd")

    def final_shading_bwd_multi(self, fs_dirs, fs_dists, fs_Lis, fx, fy, occ, normal, ray_dir, diffuse, rough_metal, g_color,
                                g_diff, g_spec, g_normal, g_diffuse, g_rough_metal, g_Lis, sum_grad_Li=False,
                                accumulate=False, grad_divisor=0.0):
        rc = self.lib.mirres_final_shading_bwd_multi(
            len(fs_dirs), self._ptr_array(fs_dirs), self._ptr_array(fs_dists), self._ptr_array(fs_Lis), fx, fy,
            self._f(occ), self._f(normal), self._f(ray_dir), self._f(diffuse), self._f(rough_metal), self._f(g_color, True),
            self._f(g_diff), self._f(g_spec), float(grad_divisor), 1 if accumulate else 0, self._f(g_normal), self._f(g_diffuse),
            self._f(g_rough_metal), 1 if sum_grad_Li else 0, self._ptr_array(g_Lis), self._stream())
        self._check(rc, "mirres_final_shading_bwd_multi")
