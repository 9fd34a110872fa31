"""Host-side mirror of the reference's ReSTIR / path-tracing driver, running on libmirres_b200.so.

Same names, argument order, tensor layouts and return arity as the reference, so stage-1 training and `--test`
rendering call it unchanged:

    restirbvhWorker(vt, vt_ind) / .update_mesh / .InitialResampling_ / .SpatialResampling_ /
        .EvaluateFinalSamples_get_vis                                  nerf/renderer_restir.py:13-146
    load_m_for_restir(framedim_x, framedim_y) -> 17-tuple              nerf/renderer_restir.py:148-228
    restir_di_with_pt(...) / run_restir_di_with_pt(...)                nerf/renderer_restir.py:230-550
    make_sampleable, GenerateLightTiles                                nerf/ScreenSpaceReSTIR/GenerateLightTiles.py
    TemporalResampling, EvaluateFinalSamples_di, FinalShading, process_new_dir_for_pt,
        indirect_one_hit_divided_no_grad                               nerf/ScreenSpaceReSTIR/Resampling.py
    EAWDenoise_run, EAWDenoise_use_phi, EAWDenoise_use_phi_no_di       nerf/ScreenSpaceReSTIR/Denoising.py

What is different underneath (none of it visible in results): the LBVH is built by one stream-ordered call with no
host synchronisation (the reference: ~tree-height launches + 2 syncs); make_sampleable is two launches instead of two
kernels + five torch scans; light tiles use a dense grid; the spp loop performs no host synchronisation when the
material object offers `sample_no_di_dense`; backward of the denoiser is a deterministic gather.
Extensions are keyword-only with reference defaults: `random_offset` (reference: np.random.randint(2**20)),
`max_bounce` (reference: MAX_Bounce = 2), `strict_reference_aliasing` (reference behaviour).
"""
import contextlib

import numpy as np
import torch

from . import slangpy_shim as slangpy
from .slangpy_shim import get_kernels

TOTAL_RIS_PASSES = 5 + 15  # frame-index stride per spp iteration (nerf/renderer_restir.py:242)

import os as _os
MAX_INITIAL_STREAMS = int(_os.environ.get("MIRRES_INITIAL_STREAMS", 3))   # concurrent initial-candidate stages (light tiles + initial RIS of different spp iterations)
MAX_INDIRECT_CHAINS = int(_os.environ.get("MIRRES_INDIRECT_CHAINS", 4))  # concurrent indirect-path chains (one CUDA stream + path state + ray-queue workspace each)
CRITICAL_ANY_BLOCKS = int(_os.environ.get("MIRRES_CRITICAL_ANY_BLOCKS", 0))  # persistent grid (blocks / SM) of the reuse chain's boolean-ray launches; 0 = library default
USE_PRIORITIES = int(_os.environ.get("MIRRES_PRIORITIES", 1))  # stream priorities for the critical reuse chain
# persistent grids (blocks / SM) of the indirect chains' tracers.  Since their closest-hit rays are split over the lanes of
# a warp the chains have slack, and a small grid leaves the SMs to the reuse chain (C2 step 4.35 -> 4.26 ms)
BACKGROUND_CLOSEST_BLOCKS = int(_os.environ.get("MIRRES_BACKGROUND_CLOSEST_BLOCKS", 1))
BACKGROUND_MIXED_BLOCKS = int(_os.environ.get("MIRRES_BACKGROUND_MIXED_BLOCKS", 2))
# visibility tags beside the loop's reservoir sets (include/mirres_b200.h, mirres_set_visibility_tags): the final
# visibility pass then only casts the rays whose answer no earlier pass of the same spp loop has already produced
USE_VIS_TAGS = int(_os.environ.get("MIRRES_VIS_TAGS", 1))
_SIDE_STREAMS = {}


class _NullStream:
    """Stand-in for a CUDA stream when the concurrent schedule is driven over CPU tensors (the host-check flavour of the
    kernels in the CPU test-suite): everything runs in program order, so waits and records are no-ops."""

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass


class _NullEvent:
    def record(self, stream=None):
        pass


def _on(stream):
    return torch.cuda.stream(stream) if isinstance(stream, torch.cuda.Stream) else contextlib.nullcontext()


def _host_kernels_bound():
    return not get_kernels().require_cuda


def _side_stream(device, k=0, priority=0):
    """Extra CUDA streams per device for the concurrent schedule (see restir_di_with_pt).  `priority`: 0 = default, more
    negative = served first by the block scheduler when several kernels have blocks pending."""
    if torch.device(device).type != "cuda":
        return _NullStream()
    index = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if not USE_PRIORITIES:
        priority = 0
    st = _SIDE_STREAMS.get((index, k, priority))
    if st is None:
        st = _SIDE_STREAMS[(index, k, priority)] = torch.cuda.Stream(device=index, priority=priority)
    return st


# =====================================================================================================================
# LBVH worker
# =====================================================================================================================
class restirbvhWorker:
    def __init__(self, vt, vt_ind):
        self.vrt = vt
        self.v_ind = vt_ind
        self._scratch = None
        self.LBVHNode_info = None
        self.LBVHNode_aabb = None
        self.packed = None

    def update_bvh(self, want_sorted_codes=False):
        k = get_kernels()
        vert = self.vrt if self.vrt.is_contiguous() else self.vrt.contiguous()
        tri = self.v_ind if self.v_ind.is_contiguous() else self.v_ind.contiguous()
        F = tri.shape[0]
        dev = vert.device
        sb, nb, tb = k.bvh_sizes(F)
        if self._scratch is None or self._scratch.numel() < sb or self._scratch.device != dev:
            self._scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
        info = torch.empty((2 * F - 1, 3), dtype=torch.int32, device=dev)
        aabb = torch.empty((2 * F - 1, 6), dtype=torch.float32, device=dev)
        nodes = torch.empty(nb, dtype=torch.uint8, device=dev)
        tris = torch.empty(tb, dtype=torch.uint8, device=dev)
        codes = torch.empty((F, 2), dtype=torch.int32, device=dev) if want_sorted_codes else None
        k.bvh_build(vert, tri, info, aabb, nodes, tris, self._scratch, codes)
        info._mirres_packed = (nodes, tris)
        self.packed = (nodes, tris)
        self.sorted_codes = codes
        return info, aabb

    def update_mesh(self, vt, vt_ind):
        self.vrt = vt
        self.v_ind = vt_ind
        self.LBVHNode_info, self.LBVHNode_aabb = self.update_bvh()

    def _bvh_kw(self):
        return dict(g_lbvh_info=self.LBVHNode_info, g_lbvh_aabb=self.LBVHNode_aabb, vert=self.vrt, v_indx=self.v_ind)

    def InitialResampling_(self, m, pos_map, reservoirs, env_tex, env_width, env_height, framedim_x, framedim_y,
                           frameIndex, occ_map, normal_depth, brdf_map, ray_dir, pdf_, cdf_, mpdf_, mcdf_, light_data,
                           light_uv, light_inv_pdf, prepare=True):
        if not prepare:
            with slangpy.workspace_prepared():
                return self.InitialResampling_(m, pos_map, reservoirs, env_tex, env_width, env_height, framedim_x,
                                               framedim_y, frameIndex, occ_map, normal_depth, brdf_map, ray_dir, pdf_,
                                               cdf_, mpdf_, mcdf_, light_data, light_uv, light_inv_pdf)
        m.process_InitialResampling_(pos_map=pos_map, reservoirs=reservoirs, env_tex=env_tex, env_width=env_width,
                                     env_height=env_height, framedim_x=framedim_x, framedim_y=framedim_y,
                                     frameIndex=frameIndex, occ_map=occ_map, normal_depth=normal_depth,
                                     brdf_map=brdf_map, ray_dir=ray_dir, pdf_=pdf_, cdf_=cdf_, mpdf_=mpdf_, mcdf_=mcdf_,
                                     light_data=light_data, light_uv=light_uv, light_inv_pdf=light_inv_pdf,
                                     **self._bvh_kw()).launchRaw()
        return 'hello'

    def SpatialResampling_(self, m, pos_map, reservoirs, prev_reservoirs, neighborOffsets, env_tex, env_width,
                           env_height, framedim_x, framedim_y, frameIndex, occ_map, normal_depth, brdf_map, ray_dir):
        m.process_SpatialResampling_(pos_map=pos_map, reservoirs=reservoirs, prevReservoirs=prev_reservoirs,
                                     neighborOffsets=neighborOffsets, env_tex=env_tex, env_width=env_width,
                                     env_height=env_height, framedim_x=framedim_x, framedim_y=framedim_y,
                                     frameIndex=frameIndex, occ_map=occ_map, normal_depth=normal_depth,
                                     brdf_map=brdf_map, ray_dir=ray_dir, **self._bvh_kw()).launchRaw()
        return 'hello'

    def EvaluateFinalSamples_get_vis(self, m, pos_map, reservoirs, framedim_x, framedim_y, vis_map):
        m.process_EvaluateFinalSamples_get_vis(reservoirs=reservoirs, framedim_x=framedim_x, framedim_y=framedim_y,
                                               pos_map=pos_map, vis_map=vis_map, **self._bvh_kw()).launchRaw()
        return 'hello'


# =====================================================================================================================
# module loading and persistent buffers
# =====================================================================================================================
def _reservoir_set(n, device, zero=True):
    make = torch.zeros if zero else torch.empty
    return (make((n, 3), dtype=torch.float, device=device), make((n, 1), dtype=torch.float, device=device),
            make((n, 1), dtype=torch.int, device=device), make((n, 1), dtype=torch.float, device=device))


def attach_vis_tags(*sets):
    """Puts a zeroed visibility tag beside each reservoir set (slangpy.vis_tag finds it).  Only a driver that owns the
    whole spp loop may do this: the tags are valid as long as every write to the sets goes through the four passes and
    pos_map / the tree do not change -- restir_di_with_pt attaches them at its start and drops them at its end."""
    if not USE_VIS_TAGS:
        return
    for s in sets:
        ld = slangpy._reservoir(s)[0]  # (a set may also be the reference's dict form)
        setattr(ld, slangpy.VIS_TAG, torch.zeros(ld.shape[0], dtype=torch.uint8, device=ld.device))


def drop_vis_tags(*sets):
    for s in sets:
        ld = slangpy._reservoir(s)[0]
        if hasattr(ld, slangpy.VIS_TAG):
            delattr(ld, slangpy.VIS_TAG)


def _no_tags_left_behind(fn):
    """The caller's reservoir sets never leave the spp loop with a tag on them, whichever way the loop ends: a tag that
    outlived its loop would vouch for samples it knows nothing about."""
    import functools
    import inspect
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(*a, **k):
        try:
            return fn(*a, **k)
        finally:
            args = sig.bind(*a, **k).arguments
            drop_vis_tags(*[args[n] for n in ("reservoirs", "prev_reservoirs") if args.get(n) is not None])
    return wrapper


def load_m_for_restir(framedim_x, framedim_y, device='cuda', max_bounce=2):
    light_tile_count, light_tile_size = 128, 1024
    tile_defs = {"LIGHT_TILE_COUNT": light_tile_count, "LIGHT_TILE_SIZE": light_tile_size}
    make_sampleable_m = slangpy.loadModule('nerf/ScreenSpaceReSTIR/make_sampleable.slang')
    generateLightTiles_m = slangpy.loadModule('nerf/ScreenSpaceReSTIR/GenerateLightTiles.slang', defines=tile_defs)
    InitialResampling_m = slangpy.loadModule(
        'nerf/ScreenSpaceReSTIR/InitialResampling.slang',
        defines=dict(tile_defs, SCREEN_TILE_SIZE=8, INITIAL_LIGHT_SAMPLE_COUNT=32, INITIAL_BRDF_SAMPLE_COUNT=1))
    TemporalResampling_m = slangpy.loadModule('nerf/ScreenSpaceReSTIR/TemporalResampling.slang',
                                              defines={"MAX_HISTORY_LENGTH": 20})
    n_offsets = 8192
    SpatialResampling_m = slangpy.loadModule(
        'nerf/ScreenSpaceReSTIR/SpatialResampling.slang',
        defines={"NEIGHBOR_OFFSET_COUNT": n_offsets, "NEIGHBOR_COUNT": 5, "GATHER_RADIUS": 30})
    EvaluateFinalSamples_m = slangpy.loadModule('nerf/ScreenSpaceReSTIR/EvaluateFinalSamples.slang')
    FinalShading_m = slangpy.loadModule('nerf/ScreenSpaceReSTIR/FinalShading.slang', defines={"MAX_Bounce": max_bounce})
    denoising_m = slangpy.loadModule('nerf/ScreenSpaceReSTIR/EAWDenoise.slang')

    n_tile = light_tile_count * light_tile_size
    light_data = torch.zeros((n_tile, 3), dtype=torch.float, device=device)
    light_uv = torch.zeros((n_tile, 2), dtype=torch.int, device=device)
    light_inv_pdf = torch.zeros((n_tile, 1), dtype=torch.float, device=device)
    n = framedim_x * framedim_y
    reservoirs = _reservoir_set(n, device)
    prev_reservoirs = _reservoir_set(n, device)
    final_samples = (torch.zeros((n, 3), dtype=torch.float, device=device),
                     torch.zeros((n, 1), dtype=torch.float, device=device),
                     torch.zeros((n, 3), dtype=torch.float, device=device))
    neighborOffsets = torch.zeros((n_offsets * 2, 1), dtype=torch.float, device=device)
    make_sampleable_m.createNeighborOffsetTexture(sampleCount=n_offsets, neighborOffsets=neighborOffsets).launchRaw()
    neighborOffsets = neighborOffsets.reshape(-1, 2) / 127
    return (make_sampleable_m, generateLightTiles_m, InitialResampling_m, TemporalResampling_m, SpatialResampling_m,
            EvaluateFinalSamples_m, FinalShading_m, denoising_m, light_data, light_uv, light_inv_pdf, reservoirs,
            prev_reservoirs, final_samples, neighborOffsets, light_tile_count, light_tile_size)


# =====================================================================================================================
# environment distribution and light tiles
# =====================================================================================================================
def make_sampleable(m, env_map, width, height):
    """Fused replacement of GenerateLightTiles.py:4-29; returns (pdf_, cdf_, mpdf_, mcdf_) with the reference shapes."""
    env_map = env_map.contiguous()
    dev = env_map.device
    pdf_ = torch.empty((width * height, 1), dtype=torch.float, device=dev)
    cdf_ = torch.empty((height * (width + 1), 1), dtype=torch.float, device=dev)
    mpdf_ = torch.empty((height, 1), dtype=torch.float, device=dev)
    mcdf_ = torch.empty((height + 1, 1), dtype=torch.float, device=dev)
    rows = torch.empty((height,), dtype=torch.float, device=dev)
    get_kernels().env_build_distribution(env_map, int(width), int(height), pdf_, cdf_, mpdf_, mcdf_, rows)
    return pdf_, cdf_, mpdf_, mcdf_


def GenerateLightTiles(m, debug_out, env_tex, pdf_, cdf_, mpdf_, mcdf_, width, height, frameIndex, light_data, light_uv,
                       light_inv_pdf, light_tile_count=128, light_tile_size=1024):
    m.process_GenerateLightTiles(env_tex=env_tex, pdf_=pdf_, cdf_=cdf_, mpdf_=mpdf_, mcdf_=mcdf_, width=int(width),
                                 height=int(height), frameIndex=int(frameIndex), light_data=light_data,
                                 light_uv=light_uv, light_inv_pdf=light_inv_pdf, debug_out=debug_out).launchRaw()
    return 'hello'


def TemporalResampling(m, reservoirs, prev_reservoirs, env_tex, env_width, env_height, framedim_x, framedim_y,
                       frameIndex, occ_map, normal_depth, brdf_map, ray_dir, prev_occ_map, prev_normal_depth,
                       prev_brdf_map, prev_ray_dir, motionVectors):
    m.process_TemporalResampling(reservoirs=reservoirs, prevReservoirs=prev_reservoirs, env_tex=env_tex,
                                 env_width=env_width, env_height=env_height, framedim_x=framedim_x,
                                 framedim_y=framedim_y, frameIndex=frameIndex, occ_map=occ_map,
                                 normal_depth=normal_depth, brdf_map=brdf_map, ray_dir=ray_dir,
                                 prev_occ_map=prev_occ_map, prev_normal_depth=prev_normal_depth,
                                 prev_brdf_map=prev_brdf_map, prev_ray_dir=prev_ray_dir,
                                 motionVectors=motionVectors).launchRaw()
    return 'hello'


# =====================================================================================================================
# autograd wrappers
# =====================================================================================================================
STRICT_REFERENCE_ALIASING = True


def _keep(t):
    """The reference saves aliases of buffers that later spp iterations overwrite through raw pointers, so the
    backward of iteration i < spp-1 sees the last iteration's samples.  Strict mode reproduces
    that; relaxed mode snapshots the tensors."""
    return t if STRICT_REFERENCE_ALIASING else t.clone()


class EvaluateFinalSamples_di(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, res_light_data, res_light_pdf, res_M, res_weight, env_tex, env_width, env_height, framedim_x,
                framedim_y, finalSamples_dir, finalSamples_distance, eva_vis_map):
        final_Li = torch.empty((framedim_x * framedim_y, 3), dtype=torch.float, device=res_light_data.device)  # every row is written
        m.process_EvaluateFinalSamples_di_(
            reservoirs=(res_light_data, res_light_pdf, res_M, res_weight), env_tex=env_tex, env_width=env_width,
            env_height=env_height, framedim_x=framedim_x, framedim_y=framedim_y,
            finalSample=(finalSamples_dir, finalSamples_distance, final_Li), vis_map=eva_vis_map).launchRaw()
        ctx.save_for_backward(_keep(res_light_data), _keep(res_light_pdf), _keep(res_M), _keep(res_weight), env_tex,
                              _keep(finalSamples_dir), _keep(finalSamples_distance), final_Li, _keep(eva_vis_map))
        ctx.nums = (env_width, env_height, framedim_x, framedim_y)
        ctx.slang_m = m
        return final_Li

    @staticmethod
    def backward(ctx, grad_final_Li):
        (res_light_data, res_light_pdf, res_M, res_weight, env_tex, fs_dir, fs_dist, final_Li, vis) = ctx.saved_tensors
        env_width, env_height, framedim_x, framedim_y = ctx.nums
        m = ctx.slang_m
        grad_env = torch.zeros_like(env_tex, memory_format=torch.contiguous_format)
        m.process_EvaluateFinalSamples_di_.bwd(
            reservoirs=m.Reservoir(light_data=res_light_data, light_pdf=res_light_pdf, M=res_M, weight=res_weight),
            env_tex=(env_tex, grad_env), env_width=env_width, env_height=env_height, framedim_x=framedim_x,
            framedim_y=framedim_y,
            finalSample=m.FinalSample(dir=fs_dir, distance=fs_dist, Li=(final_Li, grad_final_Li.contiguous())),
            vis_map=vis).launchRaw()
        return (None, None, None, None, None, grad_env, None, None, None, None, None, None, None)


class FinalShading(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, finalSamples_dir, finalSamples_distance, finalSamples_Li, env_tex, env_width, env_height,
                framedim_x, framedim_y, occ_map, normal, ray_dir, diffuse_map, linearRoughness_specular_map):
        n, dev = framedim_x * framedim_y, occ_map.device
        color = torch.empty((n, 3), dtype=torch.float, device=dev)  # the kernel writes every pixel of all three
        color_diff = torch.empty((n, 3), dtype=torch.float, device=dev)
        color_spec = torch.empty((n, 3), dtype=torch.float, device=dev)
        m.process_FinalShading(finalSample=(finalSamples_dir, finalSamples_distance, finalSamples_Li), env_tex=env_tex,
                               env_width=env_width, env_height=env_height, framedim_x=framedim_x, framedim_y=framedim_y,
                               occ_map=occ_map, normal=normal, ray_dir=ray_dir, diffuse_map=diffuse_map,
                               linearRoughness_specular_map=linearRoughness_specular_map, color=color,
                               diff_light=color_diff, spec_light=color_spec).launchRaw()
        ctx.save_for_backward(_keep(finalSamples_dir), _keep(finalSamples_distance), finalSamples_Li, env_tex, occ_map,
                              normal, ray_dir, diffuse_map, linearRoughness_specular_map, color, color_diff, color_spec)
        ctx.nums = (env_width, env_height, framedim_x, framedim_y)
        ctx.slang_m = m
        return color, color_diff, color_spec

    @staticmethod
    def backward(ctx, grad_color, grad_color_diff, grad_color_spec):
        (fs_dir, fs_dist, fs_Li, env_tex, occ_map, normal, ray_dir, diffuse_map, rs_map, color, color_diff,
         color_spec) = ctx.saved_tensors
        env_width, env_height, framedim_x, framedim_y = ctx.nums
        m = ctx.slang_m
        cf = torch.contiguous_format
        grad_normal = torch.empty_like(normal, memory_format=cf)  # mirres_final_shading_bwd overwrites all four
        grad_diffuse = torch.empty_like(diffuse_map, memory_format=cf)
        grad_rs = torch.empty_like(rs_map, memory_format=cf)
        grad_Li = torch.empty_like(fs_Li, memory_format=cf)
        m.process_FinalShading.bwd(
            finalSample=m.FinalSample(dir=fs_dir, distance=fs_dist, Li=(fs_Li, grad_Li)), env_tex=env_tex,
            env_width=env_width, env_height=env_height, framedim_x=framedim_x, framedim_y=framedim_y, occ_map=occ_map,
            normal=(normal, grad_normal), ray_dir=ray_dir, diffuse_map=(diffuse_map, grad_diffuse),
            linearRoughness_specular_map=(rs_map, grad_rs), color=(color, grad_color.contiguous()),
            diff_light=(color_diff, grad_color_diff.contiguous()),
            spec_light=(color_spec, grad_color_spec.contiguous())).launchRaw()
        return (None, None, None, grad_Li, None, None, None, None, None, None, grad_normal, None, grad_diffuse, grad_rs)


class DirectLightSum(torch.autograd.Function):
    """All K (evaluate final samples -> final shading) passes of an spp loop as ONE autograd node (concurrent schedule).

    The reference builds K x (EvaluateFinalSamples_di, FinalShading) nodes and adds their outputs (:443-459); backward then
    runs 2 K kernels plus the engine's accumulations, one after the other.  The passes share their surface inputs and --
    because the loop only sums them -- their upstream gradients, so here the forward kernels run outside autograd, the
    sums and `total / mFrameIndex` are one launch each, and backward is one mirres_final_shading_bwd_multi launch plus one
    env-gradient scatter when the passes saved aliases of one reservoir buffer (the reference's behaviour, SURVEY.md
    7.3-3; K scatters otherwise).  Per-pass arithmetic is that of the two Functions above; gradients of the passes are
    added in the engine's order (last pass first)."""

    @staticmethod
    def forward(ctx, env_tex, normal, diffuse_map, rs_map, pack):
        ctx.pack = pack
        ctx.save_for_backward(env_tex, normal, diffuse_map, rs_map)
        ctx.set_materialize_grads(False)
        return pack["color"], pack["diff"], pack["spec"]

    @staticmethod
    def backward(ctx, g_color, g_diff, g_spec):
        env_tex, normal, diffuse_map, rs_map = ctx.saved_tensors
        pack = ctx.pack
        k = get_kernels()
        fx, fy, W, H = pack["dims"]
        n, dev = fx * fy, normal.device
        passes = pack["passes"]  # (res4, fs_dir, fs_dist, fs_Li, vis) per pass
        cf = torch.contiguous_format
        zeros3 = lambda: torch.zeros((n, 3), dtype=torch.float, device=dev)
        g_diff = zeros3() if g_diff is None else g_diff.contiguous()
        g_spec = zeros3() if g_spec is None else g_spec.contiguous()
        g_color = None if g_color is None else g_color.contiguous()
        g_normal = torch.empty((n, 3), dtype=torch.float, device=dev)
        g_kd = torch.empty((n, 3), dtype=torch.float, device=dev)
        g_rs = torch.empty((n, 2), dtype=torch.float, device=dev)
        key = lambda ps: tuple(t.data_ptr() for t in ps[0]) + (ps[1].data_ptr(), ps[2].data_ptr(), ps[4].data_ptr())
        aliased = all(key(ps) == key(passes[0]) for ps in passes)
        g_Li = [torch.empty((n, 3), dtype=torch.float, device=dev) for _ in range(1 if aliased else len(passes))]
        occ, ray = pack["occ"], pack["ray"]
        nrm_c, kd_c, rs_c = normal.contiguous(), diffuse_map.contiguous(), rs_map.contiguous()
        CH = 16
        first = True
        for hi in range(len(passes), 0, -CH):  # last pass first
            lo = max(0, hi - CH)
            chunk = passes[lo:hi]
            k.final_shading_bwd_multi([c[1] for c in chunk], [c[2] for c in chunk], [c[3] for c in chunk], fx, fy, occ, nrm_c,
                                      ray, kd_c, rs_c, g_color, g_diff, g_spec, g_normal, g_kd, g_rs,
                                      g_Li if aliased else g_Li[lo:hi], sum_grad_Li=aliased, accumulate=not first,
                                      grad_divisor=float(pack["frame"]))
            first = False
        grad_env = None
        if ctx.needs_input_grad[0]:
            grad_env = torch.zeros_like(env_tex, memory_format=cf)
            if aliased:
                ps = passes[-1]
                k.eval_final_bwd(ps[0], W, H, fx, fy, ps[4], g_Li[0], grad_env)
            else:
                for j in range(len(passes) - 1, -1, -1):
                    k.eval_final_bwd(passes[j][0], W, H, fx, fy, passes[j][4], g_Li[j], grad_env)
        return grad_env, g_normal, g_kd, g_rs, None


class Composite(torch.autograd.Function):
    """final_color = nan_to_num(where(occ <= 0.1, 1, kd (1 - metallic) dd + ds + di)) (nerf/renderer_restir.py:543-549) as
    one launch per direction instead of ~10 + ~15 elementwise torch kernels; same operations in the same order."""

    @staticmethod
    def forward(ctx, occ_map, diffuse_map, rs_map, dd, ds, di):
        args = [t.detach().contiguous() for t in (occ_map, diffuse_map, rs_map, dd, ds, di)]
        out = torch.empty_like(args[3])
        get_kernels().composite_fwd(*args, out)
        ctx.save_for_backward(*args)
        return out

    @staticmethod
    def backward(ctx, g_out):
        occ, kd, rs, dd, ds, di = ctx.saved_tensors
        g_kd, g_rs, g_dd, g_ds = torch.empty_like(kd), torch.empty_like(rs), torch.empty_like(dd), torch.empty_like(ds)
        get_kernels().composite_bwd(occ, kd, rs, dd, ds, di, g_out.contiguous(), g_kd, g_rs, g_dd, g_ds)
        return None, g_kd, g_rs, g_dd, g_ds, None


class EAWDenoise_run(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth, occ_map, color, normal_map, pos_map):
        out_color = torch.zeros((framedim_x * framedim_y, 3), dtype=torch.float, device=occ_map.device)
        m.process_EAWDenoise(PHI=(c_phi, n_phi, p_phi), framedim_x=int(framedim_x), framedim_y=int(framedim_y),
                             stepWidth=int(stepWidth), occ_map=occ_map, color=color, normal_map=normal_map,
                             pos_map=pos_map, out_color=out_color).launchRaw()
        ctx.save_for_backward(occ_map, color, normal_map, pos_map, out_color)
        ctx.nums = (c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth)
        ctx.slang_m = m
        return out_color

    @staticmethod
    def backward(ctx, grad_out_color):
        occ_map, color, normal_map, pos_map, out_color = ctx.saved_tensors
        c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth = ctx.nums
        m = ctx.slang_m
        cf = torch.contiguous_format
        grad_color = torch.zeros_like(color, memory_format=cf)
        grad_normal = torch.zeros_like(normal_map, memory_format=cf)
        grad_pos = torch.zeros_like(pos_map, memory_format=cf)
        m.process_EAWDenoise.bwd(PHI=(c_phi, n_phi, p_phi), framedim_x=int(framedim_x), framedim_y=int(framedim_y),
                                 stepWidth=int(stepWidth), occ_map=occ_map, color=(color, grad_color),
                                 normal_map=(normal_map, grad_normal), pos_map=(pos_map, grad_pos),
                                 out_color=(out_color, grad_out_color.contiguous())).launchRaw()
        return (None, None, None, None, None, None, None, None, grad_color, grad_normal, grad_pos)


class EAWDenoiseMulti(torch.autograd.Function):
    """One a-trous level over several images that share occ / normal / pos (mirres_eaw_fwd_multi): per image the same
    arithmetic as EAWDenoise_run, one launch for all.  Gradients are produced only for the images that need them."""

    @staticmethod
    def forward(ctx, c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth, occ_map, normal_map, pos_map, *colors):
        k = get_kernels()
        n, dev = framedim_x * framedim_y, occ_map.device
        cols = [c.contiguous() for c in colors]
        outs = [torch.empty((n, 3), dtype=torch.float, device=dev) for _ in cols]
        needs = [bool(c.requires_grad) for c in colors]
        cum = [torch.empty((n,), dtype=torch.float, device=dev) if needs[i] else None for i in range(len(cols))]
        occ, nrm, pos = occ_map.contiguous(), normal_map.contiguous(), pos_map.contiguous()
        k.eaw_fwd_multi(c_phi, n_phi, p_phi, int(framedim_x), int(framedim_y), int(stepWidth), occ, nrm, pos, cols, outs,
                        cum if any(c is not None for c in cum) else None)
        ctx.nums = (c_phi, n_phi, p_phi, int(framedim_x), int(framedim_y), int(stepWidth))
        ctx.needs = needs
        ctx.n_img = len(cols)
        ctx.save_for_backward(occ, nrm, pos, *cols, *outs, *[c for c in cum if c is not None])
        ctx.has_cum = [c is not None for c in cum]
        # outputs of images that carry no gradient stay out of the graph (they correspond to the reference's no_grad
        # passes), so the next a-trous level sees the same differentiable / non-differentiable split
        ctx.mark_non_differentiable(*[o for o, nd in zip(outs, needs) if not nd])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        c_phi, n_phi, p_phi, fx, fy, step = ctx.nums
        saved = ctx.saved_tensors
        occ, nrm, pos = saved[0:3]
        ni = ctx.n_img
        cols, outs = saved[3:3 + ni], saved[3 + ni:3 + 2 * ni]
        cum_iter = iter(saved[3 + 2 * ni:])
        cum = [next(cum_iter) if h else None for h in ctx.has_cum]
        # only images whose colour is differentiable take part, exactly the passes the reference runs through
        # EAWDenoise_run (the others go through the no_grad twin and never reach autograd)
        active = [i for i in range(ni) if ctx.needs[i] and grad_outs[i] is not None]
        g_colors = [None] * ni
        g_normal = g_pos = None
        if active:
            k = get_kernels()
            mk = lambda: torch.empty_like(cols[0], memory_format=torch.contiguous_format)
            gc = [mk() for _ in active]
            g_normal = mk() if ctx.needs_input_grad[7] else None
            g_pos = mk() if ctx.needs_input_grad[8] else None
            k.eaw_bwd_multi(c_phi, n_phi, p_phi, fx, fy, step, occ, nrm, pos, [cols[i] for i in active],
                            [outs[i] for i in active], [cum[i] for i in active],
                            [grad_outs[i].contiguous() for i in active], gc, g_normal, g_pos)
            for j, i in enumerate(active):
                if ctx.needs[i]:
                    g_colors[i] = gc[j]
        return (None, None, None, None, None, None, None, g_normal, g_pos, *g_colors)


def EAWDenoise_multi_use_phi(c_phi, n_phi, p_phi, stepWidth, iter_time, framedim_x, framedim_y, occ_map, colors,
                             normal_map, pos_map):
    """Denoising.py:151-197 for several images at once: `iter_time` a-trous levels with step widths stepWidth,
    stepWidth/2, ...; images that do not require grad behave like EAWDenoise_use_phi_no_di."""
    outs = tuple(colors)
    for _ in range(iter_time):
        outs = EAWDenoiseMulti.apply(c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth, occ_map, normal_map, pos_map,
                                     *outs)
        stepWidth /= 2
    return outs


# ---- cross-bilateral denoiser of --use_bi_de (nerf/renderutils/ops.py:173-212) ---------------------------------------------
class _bilateral_denoiser_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, col, nrm, zdz, sigma):
        _, h, w, _ = col.shape
        col_c, nrm_c, zdz_c = col.reshape(-1, 3).contiguous(), nrm.reshape(-1, 3).contiguous(), zdz.reshape(-1, 2).contiguous()
        out = torch.empty((h * w, 4), dtype=torch.float, device=col.device)
        get_kernels().bilateral_fwd(w, h, sigma, col_c, nrm_c, zdz_c, out)
        ctx.save_for_backward(nrm_c, zdz_c)
        ctx.dims = (h, w, sigma)
        return out.view(1, h, w, 4)

    @staticmethod
    def backward(ctx, out_grad):
        nrm_c, zdz_c = ctx.saved_tensors
        h, w, sigma = ctx.dims
        col_grad = torch.empty((h * w, 3), dtype=torch.float, device=nrm_c.device)
        get_kernels().bilateral_bwd(w, h, sigma, nrm_c, zdz_c, out_grad.reshape(-1, 4).contiguous(), col_grad)
        return col_grad.view(1, h, w, 3), None, None, None


def _safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


def bilateral_denoiser(h, w, input, factor=1.0):
    """nerf/renderutils/ops.py:193-201: input [N, 8] = (colour, normal, depth, depth gradient)."""
    input = input.reshape(1, h, w, input.shape[-1])
    sigma = max(factor * 2, 0.0001)
    col_w = _bilateral_denoiser_func.apply(input[..., 0:3], _safe_normalize(input[..., 3:6]), input[..., 6:8], sigma)
    out_val = col_w[..., 0:3] / col_w[..., 3:4]
    return out_val.view(-1, out_val.shape[-1])


@torch.no_grad()
def bilateral_denoiser_no_di(h, w, input, factor=1.0):
    """nerf/renderutils/ops.py:203-212."""
    return bilateral_denoiser(h, w, input, factor)


def EAWDenoise_run_no_di(m, c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth, occ_map, color, normal_map,
                         pos_map):
    out_color = torch.zeros((framedim_x * framedim_y, 3), dtype=torch.float, device=occ_map.device)
    m.process_EAWDenoise_no_di(PHI=(c_phi, n_phi, p_phi), framedim_x=int(framedim_x), framedim_y=int(framedim_y),
                               stepWidth=int(stepWidth), occ_map=occ_map, color=color, normal_map=normal_map,
                               pos_map=pos_map, out_color=out_color).launchRaw()
    return out_color


def EAWDenoise_use_phi(m, c_phi, n_phi, p_phi, stepWidth, iter_time, framedim_x, framedim_y, occ_map, color,
                       normal_map, pos_map):
    """Denoising.py:151-197: `iter_time` a-trous passes with step widths stepWidth, stepWidth/2, ..."""
    out_color = color
    for _ in range(iter_time):
        out_color = EAWDenoise_run.apply(m, c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth, occ_map, out_color,
                                         normal_map, pos_map)
        stepWidth /= 2
    return out_color


@torch.no_grad()
def EAWDenoise_use_phi_no_di(m, c_phi, n_phi, p_phi, stepWidth, iter_time, framedim_x, framedim_y, occ_map, color,
                             normal_map, pos_map):
    out_color = color
    for _ in range(iter_time):
        out_color = EAWDenoise_run_no_di(m, c_phi, n_phi, p_phi, framedim_x, framedim_y, stepWidth, occ_map,
                                         out_color.detach(), normal_map.detach(), pos_map.detach())
        stepWidth /= 2
    return out_color


# =====================================================================================================================
# bounce launchers
# =====================================================================================================================
def process_new_dir_for_pt(m, LBVHNode_info, LBVHNode_aabb, vert, vert_ind, frameIndex, bounce_count, framedim_x,
                           framedim_y, occ_map, pos_map, normal, ray_dir, prd, diffuse_map,
                           linearRoughness_specular_map, new_pos_map, new_ray_d, new_occ_map, new_normal):
    m.process_new_dir_for_pt(g_lbvh_info=LBVHNode_info, g_lbvh_aabb=LBVHNode_aabb, vert=vert, v_indx=vert_ind,
                             frameIndex=frameIndex, bounce_count=bounce_count, framedim_x=framedim_x,
                             framedim_y=framedim_y, occ_map=occ_map, pos_map=pos_map, normal=normal, ray_dir=ray_dir,
                             prd=prd, diffuse_map=diffuse_map,
                             linearRoughness_specular_map=linearRoughness_specular_map, new_pos_map=new_pos_map,
                             new_ray_d=new_ray_d, new_occ_map=new_occ_map, new_normal=new_normal).launchRaw()
    return 'hello'


def indirect_one_hit_divided_no_grad(m, LBVHNode_info, LBVHNode_aabb, vert, vert_ind, frameIndex, bounce_count,
                                     framedim_x, framedim_y, env_tex, env_width, env_height, pdf_, cdf_, mpdf_, mcdf_,
                                     occ_map, pos_map, normal, ray_dir, prd, diffuse_map,
                                     linearRoughness_specular_map, color, diff_color, spec_color, new_pos_map,
                                     new_ray_d, new_occ_map, new_normal):
    m.process_path_tracing_divided_no_grad(
        g_lbvh_info=LBVHNode_info, g_lbvh_aabb=LBVHNode_aabb, vert=vert, v_indx=vert_ind, frameIndex=frameIndex,
        bounce_count=bounce_count, framedim_x=framedim_x, framedim_y=framedim_y, env_tex=env_tex, env_width=env_width,
        env_height=env_height, pdf_=pdf_, cdf_=cdf_, mpdf_=mpdf_, mcdf_=mcdf_, occ_map=occ_map, pos_map=pos_map,
        normal=normal, ray_dir=ray_dir, prd=prd, diffuse_map=diffuse_map,
        linearRoughness_specular_map=linearRoughness_specular_map, color=color, diff_color=diff_color,
        spec_color=spec_color, new_pos_map=new_pos_map, new_ray_d=new_ray_d, new_occ_map=new_occ_map,
        new_normal=new_normal).launchRaw()
    return 'hello'


def _query_material(mlp_mat, occ, pos, kd_out, rs_out, use_scale, scale):
    """Material lookup at the indirect vertices (nerf/renderer_restir.py:398-408).  Objects that implement
    `sample_no_di_masked_(occ, pos, kd_out, rs_out, scale)` update the maps in place where occ >= 0.5 (one launch);
    objects with `sample_no_di_dense(pos[N,3]) -> [N,6]` are evaluated on every pixel and merged with a mask (no host
    sync); anything else gets the reference protocol: compact with torch.where, call `sample_no_di`, scatter."""
    masked = getattr(mlp_mat, "sample_no_di_masked_", None)
    if masked is not None and (pos.is_cuda or _host_kernels_bound()):
        # the lookup and the merge under the occupancy mask in one launch, in place (stream order protects the readers)
        masked(occ, pos, kd_out, rs_out, scale if use_scale else None)
        return kd_out, rs_out
    hit = occ >= 0.5
    dense = getattr(mlp_mat, "sample_no_di_dense", None)
    if dense is not None:
        kd_ks = dense(pos)
        kd = kd_ks[..., 0:3]
        if use_scale:
            kd = kd * kd.new_tensor(scale)
        kd_out = torch.where(hit, kd, kd_out)
        rs_out = torch.where(hit, torch.cat((kd_ks[..., 4:5], kd_ks[..., 5:6]), dim=-1), rs_out)
    else:
        idx = torch.where(hit)[0]
        kd_ks = mlp_mat.sample_no_di(pos[idx])
        kd_out[idx] = kd_ks[..., 0:3]
        rs_out[idx] = torch.cat((kd_ks[..., 4:5], kd_ks[..., 5:6]), dim=-1)
        if use_scale:
            kd_out[idx] = kd_out[idx] * kd_out.new_tensor(scale)
    if use_scale:
        kd_out = torch.clamp(kd_out, min=0.0, max=1.0)
    return kd_out, rs_out


def _normalize_rows(x, eps=1e-6):
    # F.normalize(p=2, eps) written as explicit elementwise ops so the rounding order is defined
    n = torch.sqrt(x[:, 0:1] * x[:, 0:1] + x[:, 1:2] * x[:, 1:2] + x[:, 2:3] * x[:, 2:3])
    return x / torch.clamp(n, min=eps)


def prepare_lighting(make_sampleable_m, generateLightTiles_m, light_data, light_uv, light_inv_pdf, env_map_init, spp,
                     random_offset, light_tile_count=128, light_tile_size=1024, frame_pixels=None):
    """Everything restir_di_with_pt derives from the environment map alone: the flipped map, its sampling distribution
    (nerf/renderer_restir.py:305-312) and the light tiles of the first min(spp, MAX_INITIAL_STREAMS) iterations (:320-325).
    None of it needs the G-buffer, so a caller that produces the G-buffer itself can enqueue this first (or on another
    stream) and hand the result to run_restir_di_with_pt(lighting=...): the values are the ones the loop would compute.
    With frame_pixels = fx * fy the zero-filled blocks of the concurrent schedule (running sums, chain material maps) and
    the flipped differentiable map are produced here as well, i.e. off the serial front of the loop (~25 us at 800 x 800)."""
    height, width = env_map_init.shape[0], env_map_init.shape[1]
    extra = {}
    if frame_pixels is not None:
        dev, n = env_map_init.device, int(frame_pixels)
        n_chains = min(int(spp), MAX_INDIRECT_CHAINS)
        extra = dict(frame_pixels=n, sum_block=torch.zeros((7, n, 3), dtype=torch.float, device=dev),
                     chain_kd=torch.zeros((n_chains, n, 3), dtype=torch.float, device=dev),
                     chain_rs=torch.zeros((n_chains, n, 2), dtype=torch.float, device=dev),
                     env_flipped=torch.flip(env_map_init, dims=[0]).reshape(-1, env_map_init.shape[2]))
    env_map = torch.flip(env_map_init.detach(), dims=[0]).reshape(-1, env_map_init.shape[2])
    dist = make_sampleable(make_sampleable_m, env_map, width, height)
    R = min(int(spp), MAX_INITIAL_STREAMS)
    tiles = [(light_data, light_uv, light_inv_pdf)] + [
        (torch.empty_like(light_data), torch.empty_like(light_uv), torch.empty_like(light_inv_pdf)) for _ in range(R - 1)]
    for i in range(R):
        GenerateLightTiles(generateLightTiles_m, None, env_map, *dist, width, height,
                           random_offset + TOTAL_RIS_PASSES * i, *tiles[i], light_tile_count, light_tile_size)
    # the dict is valid for ONE loop over exactly this state of the envmap: an in-place optimiser step keeps data_ptr but
    # bumps the tensor's version counter
    return dict(env_map=env_map, dist=dist, tiles=tiles, ready=R, random_offset=random_offset, spp=int(spp),
                env_ptr=env_map_init.data_ptr(), env_version=env_map_init._version, consumed=False, **extra)


# =====================================================================================================================
# the spp loop
# =====================================================================================================================
@_no_tags_left_behind
def restir_di_with_pt(use_scale, scale_x, scale_y, scale_z, mlp_mat, bvh_restir_worker, spp, framedim_x, framedim_y,
                      make_sampleable_m, generateLightTiles_m, InitialResampling_m, TemporalResampling_m,
                      SpatialResampling_m, EvaluateFinalSamples_m, FinalShading_m, light_data, light_uv, light_inv_pdf,
                      reservoirs, prev_reservoirs, final_samples, neighborOffsets, light_tile_count, light_tile_size,
                      env_map_init, occ_map, pos_map, normal_map, depth_map, diffuse_map, roughness_specular,
                      ray_dir_map, prev_occ_map, prev_normal_depth, prev_brdf_map, prev_ray_dir, motionVectors, color,
                      *, random_offset=None, max_bounce=None, hooks=None, overlap=None, shard=None, prepared=None,
                      normalize=False, indirect_done=None, lighting=None):
    # lighting: result of prepare_lighting() for this env map / spp / random_offset (environment distribution and the light
    # tiles of the first iterations, computed earlier by the caller, e.g. while the G-buffer is still being traced)
    # indirect_done(color_1, diff_1, spec_1): called once the indirect sums are final (normalised when `normalize`), in the
    # concurrent schedule on the stream that produced them, so the caller can post-process them while the direct-light
    # chain is still running
    # normalize=True (run_restir_di_with_pt): the six sums come back already divided by the frame count (:505-515)
    n = framedim_x * framedim_y
    dev = pos_map.device
    if random_offset is None:
        random_offset = int(np.random.randint(2 ** 20))
    if max_bounce is None:
        max_bounce = FinalShading_m.define("MAX_Bounce", 2)
    worker = bvh_restir_worker
    bvh = (worker.LBVHNode_info, worker.LBVHNode_aabb, worker.vrt, worker.v_ind)

    # `shard` (dist.RowBandView): the frame handed in is a band of a larger frame plus its halo rows.  Light tiles and
    # temporal reuse cover `shard.wide` (the band + 30 rows on either side: what spatial reuse reads), everything else
    # -- spatial reuse, visibility, shading, the indirect paths -- the band's own rows, which is what run_restir_di_with_pt
    # has made the ambient row restriction of the pixel lists.
    def rows_wide():
        return slangpy.active_rows(shard.wide[0], shard.wide[1], framedim_x) if shard is not None else contextlib.nullcontext()

    def band_only():
        # the spatial pass runs on the wide list (the pixels around the band publish the samples the band reuses) and
        # resamples the band's rows
        return slangpy.spatial_band(shard.rows[0], shard.rows[1]) if shard is not None else contextlib.nullcontext()

    def zeros(*shape):
        return torch.zeros(shape, dtype=torch.float, device=dev)

    if overlap is None:
        overlap = hooks is None and pos_map.is_cuda
    # the six running sums and total_indirect_light (:291-297) as slices of ONE zero-filled block: one fill on the serial
    # front of the step instead of seven
    # (concurrent schedule only: there the sums are written by raw launches; the sequential schedule accumulates through
    # autograd, in place, which wants tensors of their own)
    names = ("color", "diff", "spec", "color_1", "diff_1", "spec_1")
    if lighting is not None and not (lighting["random_offset"] == random_offset and lighting["spp"] == spp and
                                     lighting["env_ptr"] == env_map_init.data_ptr() and
                                     lighting.get("env_version") == env_map_init._version and not lighting.get("consumed")):
        lighting = None  # prepared for another call, for an older state of the envmap, or already used by a loop
    if lighting is not None:
        lighting["consumed"] = True
    # blocks zero-filled by prepare_lighting(frame_pixels=n); they are consumed (popped): a second loop on the same
    # `lighting` allocates its own
    early = lighting if (overlap and lighting is not None and lighting.get("frame_pixels") == n and
                         lighting.get("sum_block") is not None and lighting["sum_block"].device == dev) else None
    if overlap:
        block = early.pop("sum_block") if early is not None else zeros(7, n, 3)
        sums = {k: block[j] for j, k in enumerate(names)}
        total_indirect_light = block[6]
    else:
        sums = {k: zeros(n, 3) for k in names}
        total_indirect_light = zeros(n, 3)
    prd = ping = pong = color_1 = color_diff_1 = color_spec_1 = new_diffuse_map = new_roughness_specular = None
    if not overlap:  # path state of the sequential schedule (the concurrent chains own theirs)
        color_1, color_diff_1, color_spec_1 = zeros(n, 3), zeros(n, 3), zeros(n, 3)
        prd = zeros(n, 5)
        ping = dict(pos=zeros(n, 3), ray=zeros(n, 3), occ=zeros(n, 1), nrm=zeros(n, 3))
        pong = dict(pos=zeros(n, 3), ray=zeros(n, 3), occ=zeros(n, 1), nrm=zeros(n, 3))
        new_diffuse_map = torch.zeros((n, 3), dtype=torch.float, device=dev)
        new_roughness_specular = torch.zeros((n, 2), dtype=torch.float, device=dev)

    kd, rs = diffuse_map.detach(), roughness_specular.detach()
    if prepared is not None:
        normal_depth, brdf_map = prepared  # run_restir_di_with_pt has produced both with mirres_prepare_maps
    else:
        normal_depth = torch.cat((normal_map, depth_map), dim=-1).detach()
        brdf_map = torch.cat((kd[:, 0:1] * 0.2126 + kd[:, 1:2] * 0.7152 + kd[:, 2:3] * 0.0722,
                              rs[:, 1:2] * 0.2126 + rs[:, 1:2] * 0.7152 + rs[:, 1:2] * 0.0722,
                              rs[:, 0:1]), dim=-1)
        brdf_map[:, 2].clamp_(min=0.01, max=1)
        brdf_map[:, 2] = brdf_map[:, 2] * brdf_map[:, 2]
    eva_vis_map = torch.empty((n, 1), dtype=torch.float, device=dev)  # get_vis sets every pixel (to 1 first, :103)

    # the reference ignores the reservoirs/prev_* it is handed for these and starts from zeros (:291-302)
    if not overlap:
        prev_reservoirs = _reservoir_set(n, dev)
    # the zero-filled prev_* maps of the reference (:299-302) are never read: temporal reuse starts at the second
    # iteration, by which time they have been replaced by the current maps (:455-458)
    prev_occ_map = prev_normal_depth = prev_brdf_map = prev_ray_dir = None

    height, width = env_map_init.shape[0], env_map_init.shape[1]
    env_map = lighting["env_map"] if lighting is not None else \
        torch.flip(env_map_init.detach(), dims=[0]).reshape(-1, env_map_init.shape[2])
    env_map_init = early.pop("env_flipped") if early is not None else \
        torch.flip(env_map_init, dims=[0]).reshape(-1, env_map_init.shape[2])
    pdf_, cdf_, mpdf_, mcdf_ = lighting["dist"] if lighting is not None else \
        make_sampleable(make_sampleable_m, env_map, width, height)

    frame = 0
    scale = (scale_x, scale_y, scale_z)

    # The indirect path of an spp iteration (continuation ray + `max_bounce` shaded vertices) reads only the G-buffer
    # and its own path state, never the reservoirs, so it is independent of the direct-light chain AND of the indirect
    # paths of the other iterations.  With `overlap`, up to MAX_INDIRECT_CHAINS of them run on their own CUDA streams
    # with their own path state and ray-queue workspace; the chains fill each other's latency tails (a traversal
    # launch ends with a few long rays on an otherwise idle GPU).  Their per-vertex outputs are added to the running
    # sums on the main stream in the reference's order (iteration-major, bounce-minor), so the sums stay bit-identical
    # to the sequential schedule.
    main_stream = None
    keepalive = []
    chains = []
    normal_detached = normal_map.detach()

    def make_chain(tag, stream):
        c = dict(tag=tag, stream=stream, prd=prd, ping=ping, pong=pong, kd=new_diffuse_map, rs=new_roughness_specular)
        if tag != "main":
            # path state of a concurrent chain: the bounce kernels initialise prd / new_occ for every pixel and write
            # pos / ray / normal wherever a later kernel reads them, so only the material maps need a defined start
            e = lambda *shape: torch.empty(shape, dtype=torch.float, device=dev)
            c.update(prd=e(n, 5),
                     ping=dict(pos=e(n, 3), ray=e(n, 3), occ=e(n, 1), nrm=e(n, 3)),
                     pong=dict(pos=e(n, 3), ray=e(n, 3), occ=e(n, 1), nrm=e(n, 3)),
                     kd=chain_kd[len(chains)], rs=chain_rs[len(chains)])
        return c

    caller_stream = None
    if overlap:
        caller_stream = torch.cuda.current_stream() if pos_map.is_cuda else _NullStream()
        # the reuse chain (temporal -> spatial of consecutive iterations) is the critical path of the loop: it gets a stream
        # of its own with the highest priority, so its blocks are placed before the pending blocks of everything else
        main_stream = _side_stream(dev, "reuse", -3) if (pos_map.is_cuda and USE_PRIORITIES) else caller_stream
        n_chains = min(spp, MAX_INDIRECT_CHAINS)
        # material maps of all chains: two fills on the serial front instead of two per chain
        if early is not None and early["chain_kd"].shape[0] == n_chains:
            chain_kd, chain_rs = early.pop("chain_kd"), early.pop("chain_rs")
        else:
            chain_kd = torch.zeros((n_chains, n, 3), dtype=torch.float, device=dev)
            chain_rs = torch.zeros((n_chains, n, 2), dtype=torch.float, device=dev)
        for c in range(n_chains):
            chains.append(make_chain("indirect%d" % c, _side_stream(dev, c)))
        for c in chains:
            c["stream"].wait_stream(caller_stream)
    else:
        chains.append(make_chain("main", None))
    pending = []  # (iteration, bounce, completion event, (color, diff, spec)) not yet added to the running sums

    def accumulate(upto_iteration, stream=None, divisor=0.0):
        # runs on `stream` (the caller has made it current); always in (iteration, bounce) order: one launch per output
        # adds up to 32 per-vertex images to the running sum, the last call also divides by the frame count
        stream = stream or main_stream
        batch = []
        while pending and pending[0][0] <= upto_iteration:
            batch.append(pending.pop(0))
        if not batch and divisor == 0.0:
            return
        for _, _, done, _ in batch:
            stream.wait_event(done)
        k = get_kernels()
        for lo in range(0, max(len(batch), 1), 32):
            part = batch[lo:lo + 32]
            last = lo + 32 >= len(batch)
            for j, name in enumerate(("color_1", "diff_1", "spec_1")):
                k.sum_images([b[3][j] for b in part], sums[name], divisor if last else 0.0, accumulate=True)
        if batch and pos_map.is_cuda:
            # the summed images were allocated on the chains' streams: those streams wait for this read before the blocks
            # can be handed out again, so the images can be released now instead of living until the end of the loop
            consumed = torch.cuda.Event()
            consumed.record(stream)
            for c in chains:
                c["stream"].wait_event(consumed)

    def indirect_chain(i, first_pass, c):
        base = random_offset + TOTAL_RIS_PASSES * i
        ris_pass = first_pass
        ping_, pong_, prd_ = c["ping"], c["pong"], c["prd"]
        process_new_dir_for_pt(FinalShading_m, *bvh, base + ris_pass, 0, framedim_x, framedim_y, occ_map, pos_map,
                               normal_detached, ray_dir_map, prd_, kd, rs, ping_["pos"], ping_["ray"], ping_["occ"],
                               ping_["nrm"])
        ris_pass += 5
        src, dst = ping_, pong_
        for bounce in range(1, max_bounce + 1):
            keepalive.extend((c["kd"], c["rs"]))
            c["kd"], c["rs"] = _query_material(mlp_mat, src["occ"], src["pos"], c["kd"], c["rs"], use_scale, scale)
            if c["stream"] is None:
                outs3 = (color_1, color_diff_1, color_spec_1)
            else:
                # allocated on the chain's stream, kept until the join; the kernel's prologue zero-fills all three
                outs3 = tuple(torch.empty((n, 3), dtype=torch.float, device=dev) for _ in range(3))
            indirect_one_hit_divided_no_grad(FinalShading_m, *bvh, base + ris_pass, bounce, framedim_x, framedim_y,
                                             env_map, width, height, pdf_, cdf_, mpdf_, mcdf_, src["occ"], src["pos"],
                                             src["nrm"], src["ray"], prd_, c["kd"], c["rs"],
                                             outs3[0], outs3[1], outs3[2], dst["pos"], dst["ray"], dst["occ"],
                                             dst["nrm"])
            if c["stream"] is None:
                sums["color_1"] += outs3[0]
                sums["diff_1"] += outs3[1]
                sums["spec_1"] += outs3[2]
            else:
                done = torch.cuda.Event() if pos_map.is_cuda else _NullEvent()
                done.record(c["stream"])
                pending.append((i, bounce, done, outs3))
            if hooks is not None:
                hooks("bounce", (i, bounce), dict(color=outs3[0], diff=outs3[1], spec=outs3[2], prd=prd_,
                                                  occ=dst["occ"], pos=dst["pos"]))
            ris_pass += 5
            src, dst = dst, src

    if overlap:
        # ---- pipelined schedule -------------------------------------------------------------------------------------------
        # Only temporal -> spatial reuse carries a dependency from one spp iteration to the next.  Light tiles + initial
        # candidates depend on the G-buffer and the envmap alone, and final visibility / evaluation / shading only
        # consume the finished spatial reservoirs, so the three stages run on three streams:
        #     I: tiles(i) -> initial(i) -> X[i % 2]
        #     M: temporal(X[i % R], prev = S[(i-1) % 3]) -> spatial(X[i % R] -> S[i % 3])          (the critical path)
        #     S: B <- S[i % 3]; visibility(B); evaluate(B); shade; running sums
        # B is ONE buffer shared by all iterations, so the autograd Functions keep saving aliases of buffers that later
        # iterations overwrite, exactly like the reference's two-buffer ping-pong.  Arithmetic, frame
        # indices and accumulation order are those of the sequential schedule; only the enqueue order differs.
        st_s = _side_stream(dev, MAX_INDIRECT_CHAINS, -1)
        st_i = _side_stream(dev, "indirect_sum")
        # initial candidates of different iterations are independent of each other as well: a ring of R streams, each
        # with its own light-tile set, reservoir buffer and workspace, lets them all start as soon as the G-buffer exists
        R = min(spp, MAX_INITIAL_STREAMS)
        # the first iteration's candidates gate the whole reuse chain, the later ones have slack
        st_init = [_side_stream(dev, MAX_INDIRECT_CHAINS + 1 + r, -2 if r == 0 else -1) for r in range(R)]
        X = tuple(_reservoir_set(n, dev, False) for _ in range(R))  # both passes write every pixel
        # a ring of three finished-reservoir sets: with two, spatial(i) had to wait until the shading stream had copied
        # S(i - 2) away, and a shading stream that lags (its visibility trace shares the SMs with everything else) stalled
        # the reuse chain for ~70 us per iteration (timeline of one band of an 8-way C5 render, profiles/README.md)
        NS = 3
        S = tuple(_reservoir_set(n, dev, False) for _ in range(NS))
        attach_vis_tags(reservoirs, *(X + S))
        tiles = lighting["tiles"] if lighting is not None else [(light_data, light_uv, light_inv_pdf)] + [
            (torch.empty_like(light_data), torch.empty_like(light_uv), torch.empty_like(light_inv_pdf))
            for _ in range(R - 1)]
        tiles_ready = lighting["ready"] if lighting is not None else 0
        B = reservoirs
        # fork first: the chains' own foreground lists are then built beside the reuse chain's, not after it
        for st in [st_s, st_i] + st_init + ([main_stream] if main_stream is not caller_stream else []):
            st.wait_stream(caller_stream)
        with _on(main_stream), rows_wide():  # the reuse chain's list: temporal reuse covers it, spatial reuse its band rows
            slangpy.prepare_workspace(occ_map)
        for r in range(R):
            with _on(st_init[r]), slangpy.workspace_tag("initial%d" % r), rows_wide():
                slangpy.prepare_workspace(occ_map)
        with _on(st_s), slangpy.workspace_tag("shade"):
            slangpy.prepare_workspace(occ_map)
        # ... and so are the indirect chains' lists: once per loop, not once per iteration (three launches each)
        for c in chains:
            with _on(c["stream"]), slangpy.workspace_tag(c["tag"]):
                slangpy.prepare_workspace(occ_map)
        ev = lambda stream: (lambda e: (e.record(stream), e)[1])(torch.cuda.Event() if pos_map.is_cuda else _NullEvent())
        init_done, spatial_done, copy_done = {}, {}, {}
        passes, direct_outs = [], []
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (env_map_init, normal_map, diffuse_map,
                                                                               roughness_specular))

        def flush_direct(divisor=0.0):
            # running sums of the shading outputs in iteration order (the reference's `total += color`, :443-459)
            k = get_kernels()
            for j, name in enumerate(("color", "diff", "spec")):
                k.sum_images([o[j] for o in direct_outs], sums[name], divisor, accumulate=True)
            direct_outs.clear()  # allocated and read on the shading stream: stream order protects the blocks
        for i in range(spp):
            base = random_offset + TOTAL_RIS_PASSES * i
            first_indirect_pass = 4 if i == 0 else 5
            c = chains[i % len(chains)]
            with _on(c["stream"]), slangpy.workspace_tag(c["tag"]), slangpy.workspace_prepared(), slangpy.trace_blocks(
                    closest_blocks=BACKGROUND_CLOSEST_BLOCKS, mixed_blocks=BACKGROUND_MIXED_BLOCKS):
                indirect_chain(i, first_indirect_pass, c)
            r = i % R
            with _on(st_init[r]), slangpy.workspace_tag("initial%d" % r):
                if i >= R:
                    st_init[r].wait_event(spatial_done[i - R])  # X[r] was last read by spatial(i - R)
                if i >= tiles_ready:
                    GenerateLightTiles(generateLightTiles_m, None, env_map, pdf_, cdf_, mpdf_, mcdf_, width, height, base,
                                       *tiles[r], light_tile_count, light_tile_size)
                worker.InitialResampling_(InitialResampling_m, pos_map, X[r], env_map, width, height, framedim_x,
                                          framedim_y, base + 2, occ_map, normal_depth, brdf_map, ray_dir_map, pdf_, cdf_,
                                          mpdf_, mcdf_, *tiles[r], prepare=False)
                init_done[i] = ev(st_init[r])
            main_stream.wait_event(init_done[i])
            ris_pass = 3
            with _on(main_stream):
                if i > 0:
                    TemporalResampling(TemporalResampling_m, X[r], S[(i - 1) % NS], env_map, width, height, framedim_x,
                                       framedim_y, base + ris_pass, occ_map, normal_depth, brdf_map, ray_dir_map,
                                       prev_occ_map, prev_normal_depth, prev_brdf_map, prev_ray_dir, motionVectors)
                    ris_pass += 1
                if i >= NS:
                    main_stream.wait_event(copy_done[i - NS])  # S[i % NS] was last read by the copy of iteration i - NS
                with slangpy.trace_blocks(any_blocks=CRITICAL_ANY_BLOCKS), band_only():
                    worker.SpatialResampling_(SpatialResampling_m, pos_map, S[i % NS], X[r], neighborOffsets, env_map,
                                              width, height, framedim_x, framedim_y, base + ris_pass, occ_map, normal_depth,
                                              brdf_map, ray_dir_map)
                if shard is not None:
                    # row bands: the halo rows this rank reads in the next iteration (through its temporal pass) take their
                    # owners' values; the collective is ordered on the reuse chain's stream
                    shard.exchange(S[i % NS])
            ris_pass += 1
            assert ris_pass == first_indirect_pass
            spatial_done[i] = ev(main_stream)
            with _on(st_s), slangpy.workspace_tag("shade"):
                st_s.wait_event(spatial_done[i])
                for dst_t, src_t in zip(B, S[i % NS]):
                    dst_t.data.copy_(src_t)  # raw overwrite, invisible to autograd like the reference's kernels
                if slangpy.vis_tag(B) is not None:
                    slangpy.vis_tag(B).copy_(slangpy.vis_tag(S[i % NS]))
                copy_done[i] = ev(st_s)
                worker.EvaluateFinalSamples_get_vis(EvaluateFinalSamples_m, pos_map, B, framedim_x, framedim_y,
                                                    eva_vis_map)
                # evaluation + shading outside autograd; DirectLightSum (below) is the one node that stands for all passes
                with torch.no_grad():
                    final_Li = torch.empty((n, 3), dtype=torch.float, device=dev)
                    EvaluateFinalSamples_m.process_EvaluateFinalSamples_di_(
                        reservoirs=B, env_tex=env_map, env_width=width, env_height=height, framedim_x=framedim_x,
                        framedim_y=framedim_y, finalSample=(final_samples[0], final_samples[1], final_Li),
                        vis_map=eva_vis_map).launchRaw()
                    outs_d = tuple(torch.empty((n, 3), dtype=torch.float, device=dev) for _ in range(3))
                    FinalShading_m.process_FinalShading(
                        finalSample=(final_samples[0], final_samples[1], final_Li), env_tex=env_map, env_width=width,
                        env_height=height, framedim_x=framedim_x, framedim_y=framedim_y, occ_map=occ_map,
                        normal=normal_detached, ray_dir=ray_dir_map, diffuse_map=kd, linearRoughness_specular_map=rs,
                        color=outs_d[0], diff_light=outs_d[1], spec_light=outs_d[2]).launchRaw()
                if needs_grad:
                    passes.append((tuple(_keep(t) for t in B), _keep(final_samples[0]), _keep(final_samples[1]), final_Li,
                                   _keep(eva_vis_map)))
                direct_outs.append(outs_d)
                if len(direct_outs) >= 16:
                    flush_direct()
            if len(pending) >= 16:  # bounds the memory held by long loops (--spp 512); normally one flush at the end
                with _on(st_i):
                    accumulate(i - 2, st_i)
            frame += 1
            prev_occ_map, prev_normal_depth, prev_brdf_map, prev_ray_dir = occ_map, normal_depth, brdf_map, ray_dir_map
        div = float(frame) if normalize else 0.0
        with _on(st_s):
            flush_direct(div)
            if needs_grad:
                # the node lives on the shading stream, like the per-pass Functions it stands for: its backward runs there
                pack = dict(color=sums["color"], diff=sums["diff"], spec=sums["spec"], passes=passes,
                            dims=(framedim_x, framedim_y, width, height), occ=slangpy._c(occ_map),
                            ray=slangpy._c(ray_dir_map), frame=div)
                sums["color"], sums["diff"], sums["spec"] = DirectLightSum.apply(env_map_init, normal_map, diffuse_map,
                                                                                 roughness_specular, pack)
        with _on(st_i):
            # the indirect sums never meet the direct-light chain: they are finished (and handed to the caller's
            # post-processing) on their own stream, typically while the last reuse passes are still running
            for c in chains:
                st_i.wait_stream(c["stream"])
            accumulate(spp, st_i, div)
            if indirect_done is not None:
                indirect_done(sums["color_1"], sums["diff_1"], sums["spec_1"])
                indirect_done = None
        for st in [st_s, st_i] + st_init + ([main_stream] if main_stream is not caller_stream else []):
            caller_stream.wait_stream(st)
        keepalive.extend(X + S)
        keepalive.extend(tiles)
        drop_vis_tags(reservoirs)
    else:
        attach_vis_tags(reservoirs, prev_reservoirs)
        for i in range(spp):
            base = random_offset + TOTAL_RIS_PASSES * frame
            # frame-index schedule of the reference (nerf/renderer_restir.py:314-459): tiles +0 (+1 inside), initial +2,
            # temporal +3 (i > 0), spatial next, new_dir = spatial + 1, shaded vertices +5 each
            first_indirect_pass = 4 if i == 0 else 5
            ris_pass = 0
            GenerateLightTiles(generateLightTiles_m, None, env_map, pdf_, cdf_, mpdf_, mcdf_, width, height,
                               base + ris_pass, light_data, light_uv, light_inv_pdf, light_tile_count, light_tile_size)
            ris_pass += 2
            with rows_wide():  # (rebuilds the pixel list of the one workspace this schedule uses)
                worker.InitialResampling_(InitialResampling_m, pos_map, reservoirs, env_map, width, height, framedim_x,
                                          framedim_y, base + ris_pass, occ_map, normal_depth, brdf_map, ray_dir_map, pdf_,
                                          cdf_, mpdf_, mcdf_, light_data, light_uv, light_inv_pdf)
            ris_pass += 1
            if i > 0:
                TemporalResampling(TemporalResampling_m, reservoirs, prev_reservoirs, env_map, width, height, framedim_x,
                                   framedim_y, base + ris_pass, occ_map, normal_depth, brdf_map, ray_dir_map, prev_occ_map,
                                   prev_normal_depth, prev_brdf_map, prev_ray_dir, motionVectors)
                ris_pass += 1
            reservoirs, prev_reservoirs = prev_reservoirs, reservoirs
            with band_only():
                worker.SpatialResampling_(SpatialResampling_m, pos_map, reservoirs, prev_reservoirs, neighborOffsets,
                                          env_map, width, height, framedim_x, framedim_y, base + ris_pass, occ_map,
                                          normal_depth, brdf_map, ray_dir_map)
            ris_pass += 1
            if shard is not None:
                slangpy.prepare_workspace(occ_map)  # from here on the band's own rows only
                # the temporal pass of the next iteration reads finished reservoirs up to 31 rows outside this rank's
                # band: those rows take their owners' values
                shard.exchange(reservoirs)
            worker.EvaluateFinalSamples_get_vis(EvaluateFinalSamples_m, pos_map, reservoirs, framedim_x, framedim_y,
                                                eva_vis_map)
            final_Li = EvaluateFinalSamples_di.apply(EvaluateFinalSamples_m, reservoirs[0], reservoirs[1], reservoirs[2],
                                                     reservoirs[3], env_map_init, width, height, framedim_x, framedim_y,
                                                     final_samples[0], final_samples[1], eva_vis_map)
            color, color_diff, color_spec = FinalShading.apply(FinalShading_m, final_samples[0], final_samples[1], final_Li,
                                                               env_map, width, height, framedim_x, framedim_y, occ_map,
                                                               normal_map, ray_dir_map, diffuse_map, roughness_specular)
            if hooks is not None:
                hooks("direct", i, dict(reservoirs=reservoirs, prev_reservoirs=prev_reservoirs, vis=eva_vis_map,
                                        final_samples=final_samples, final_Li=final_Li, color=color, diff=color_diff,
                                        spec=color_spec, light_data=light_data, light_uv=light_uv,
                                        light_pdf=light_inv_pdf))
            assert ris_pass == first_indirect_pass
            indirect_chain(i, first_indirect_pass, chains[0])
            frame += 1
            reservoirs, prev_reservoirs = prev_reservoirs, reservoirs
            prev_occ_map, prev_normal_depth, prev_brdf_map, prev_ray_dir = occ_map, normal_depth, brdf_map, ray_dir_map
            sums["color"] += color
            sums["diff"] += color_diff
            sums["spec"] += color_spec
    if overlap:
        for c in chains:
            caller_stream.wait_stream(c["stream"])
    else:
        drop_vis_tags(reservoirs, prev_reservoirs)
        if normalize:
            for name in sums:
                sums[name] = sums[name] / frame
    if indirect_done is not None:
        indirect_done(sums["color_1"], sums["diff_1"], sums["spec_1"])
    keepalive.clear()
    return (sums["color"], sums["color_1"], sums["diff"], sums["spec"], sums["diff_1"], sums["spec_1"],
            total_indirect_light, frame)


def run_restir_di_with_pt(use_scale, scale_x, scale_y, scale_z, mlp_mat, gb_depth, bvh_restir_worker,
                          make_sampleable_m, generateLightTiles_m, InitialResampling_m, TemporalResampling_m,
                          SpatialResampling_m, EvaluateFinalSamples_m, FinalShading_m, denoising_m, light_data, light_uv,
                          light_inv_pdf, reservoirs, prev_reservoirs, final_samples, neighborOffsets, light_tile_count,
                          light_tile_size, env_map, occ_map, normal_map, depth_map, diffuse_map, roughness_specular,
                          ray_dir_map, pos_map, prev_occ_map, prev_normal_depth, prev_brdf_map, prev_ray_dir,
                          framedim_x, framedim_y, spp, denoise_iter, stepWidth, c_phi_scale=1.0, n_phi_scale=0.1,
                          p_phi_scale=0.1, *, random_offset=None, max_bounce=None, hooks=None, bilateral=None,
                          overlap=None, batched_denoise=True, shard=None, fused_prepare=True, fused_composite=True,
                          lighting=None):
    n, dev = framedim_x * framedim_y, pos_map.device
    prepared = None
    if fused_prepare and occ_map.is_contiguous():
        # occupancy threshold (in place, as the reference does, :484-485), ray normalisation, normal_depth and brdf_map
        # (:279-287) in one launch; the operations and their order are those of the torch expressions in the else branch
        normal_depth = torch.empty((n, 4), dtype=torch.float, device=dev)
        brdf_map = torch.empty((n, 3), dtype=torch.float, device=dev)
        ray_out = torch.empty((n, 3), dtype=torch.float, device=dev)
        get_kernels().prepare_maps(occ_map, normal_map.detach().contiguous(), depth_map.contiguous(),
                                   diffuse_map.detach().contiguous(), roughness_specular.detach().contiguous(),
                                   ray_dir_map.contiguous(), normal_depth, brdf_map, ray_out)
        ray_dir_map = ray_out
        prepared = (normal_depth, brdf_map)
    else:
        occ_map.masked_fill_(occ_map <= 0.5, 0)  # in place, as the reference does (:484-485), but without a host sync
        ray_dir_map = _normalize_rows(ray_dir_map)
    motionVectors = None  # the reference passes zeros (:487); NULL means the same to the kernel
    color = None
    early = {}

    def denoise_indirect(color_1, diff_1, spec_1):
        # the three indirect images carry no gradient and do not depend on the direct-light chain: they are denoised as soon
        # as their sums exist (concurrent schedule: on the stream that finished them, overlapping the last reuse passes)
        with torch.no_grad():
            combined = diff_1 + spec_1
            early["combined"] = combined
            early["outs"] = EAWDenoise_multi_use_phi(c_phi_scale, n_phi_scale, p_phi_scale, stepWidth, denoise_iter,
                                                     framedim_x, framedim_y, occ_map, (combined, diff_1, spec_1),
                                                     normal_map.detach(), pos_map.detach())

    split_denoise = gb_depth is None and batched_denoise
    loop_args = (make_sampleable_m, generateLightTiles_m, InitialResampling_m, TemporalResampling_m, SpatialResampling_m,
                 EvaluateFinalSamples_m, FinalShading_m, light_data, light_uv, light_inv_pdf)
    loop_kw = dict(random_offset=random_offset, max_bounce=max_bounce, hooks=hooks, overlap=overlap, normalize=True,
                   lighting=lighting)
    if shard is None:
        (total_color, total_color_1, total_diff_light, total_spec_light, total_diff_light_1, total_spec_light_1,
         total_indirect_light, mFrameIndex) = restir_di_with_pt(
            use_scale, scale_x, scale_y, scale_z, mlp_mat, bvh_restir_worker, spp, framedim_x, framedim_y, *loop_args,
            reservoirs, prev_reservoirs, final_samples, neighborOffsets, light_tile_count, light_tile_size, env_map, occ_map,
            pos_map, normal_map, depth_map, diffuse_map, roughness_specular, ray_dir_map, prev_occ_map, prev_normal_depth,
            prev_brdf_map, prev_ray_dir, motionVectors, color, prepared=prepared,
            indirect_done=denoise_indirect if split_denoise else None, **loop_kw)
    else:
        # Row-band rendering (dist.RowBandShard).  The [N, k] maps are row-major, so the rows this rank reads
        # -- its band and the halo around it -- are one contiguous slice of every tensor: the spp loop runs on the slices
        # as on a frame of their own (every per-pixel stream shrinks with the band), the row offset keeps the random
        # streams those of the full frame, the pixel lists are restricted to the band (light tiles and temporal reuse:
        # band + 30 rows, see restir_di_with_pt) and after every spatial pass the halo rows are received from their owners.
        # The six accumulated images of all bands are then gathered and denoised at full frame, exactly as on one GPU.
        view = shard.view()
        a0, a1 = shard.active
        cut = lambda t: None if t is None else (tuple(cut(x) for x in t) if isinstance(t, (tuple, list))
                                                else t[a0 * framedim_x:a1 * framedim_x])
        with slangpy.row_offset(a0), slangpy.active_rows(view.rows[0], view.rows[1], framedim_x):
            band = restir_di_with_pt(
                use_scale, scale_x, scale_y, scale_z, mlp_mat, bvh_restir_worker, spp, framedim_x, a1 - a0, *loop_args,
                cut(reservoirs), cut(prev_reservoirs), cut(final_samples), neighborOffsets, light_tile_count,
                light_tile_size, env_map, cut(occ_map), cut(pos_map), cut(normal_map), cut(depth_map), cut(diffuse_map),
                cut(roughness_specular), cut(ray_dir_map), None, None, None, None, motionVectors, color,
                prepared=cut(prepared), shard=view, **loop_kw)
        mFrameIndex = band[7]
        (total_color, total_color_1, total_diff_light, total_spec_light, total_diff_light_1,
         total_spec_light_1) = shard.gather_bands([t.detach() for t in band[0:6]], row0=a0)
        if split_denoise:
            denoise_indirect(total_color_1, total_diff_light_1, total_spec_light_1)  # at full frame, after the gather
    # `total / mFrameIndex` of all six sums (:505-515) has happened inside (normalize=True)
    combined_color_indirect = early["combined"] if split_denoise else total_diff_light_1 + total_spec_light_1

    if split_denoise:
        # images that share occ / normal / pos go through one launch per a-trous level (per-image arithmetic unchanged):
        # the two differentiable ones here, the three indirect ones in denoise_indirect above
        denoised_diffuse, denoised_spec = EAWDenoise_multi_use_phi(
            c_phi_scale, n_phi_scale, p_phi_scale, stepWidth, denoise_iter, framedim_x, framedim_y, occ_map,
            (total_diff_light, total_spec_light), normal_map, pos_map)
        denoised_indirect, denoised_indirect_diff, denoised_indirect_spec = early["outs"]
    elif gb_depth is None:
        args = (denoising_m, c_phi_scale, n_phi_scale, p_phi_scale, stepWidth, denoise_iter, framedim_x, framedim_y,
                occ_map)
        denoised_diffuse = EAWDenoise_use_phi(*args, total_diff_light, normal_map, pos_map)
        denoised_spec = EAWDenoise_use_phi(*args, total_spec_light, normal_map, pos_map)
        denoised_indirect = EAWDenoise_use_phi_no_di(*args, combined_color_indirect, normal_map, pos_map)
        denoised_indirect_diff = EAWDenoise_use_phi_no_di(*args, total_diff_light_1, normal_map, pos_map)
        denoised_indirect_spec = EAWDenoise_use_phi_no_di(*args, total_spec_light_1, normal_map, pos_map)
    else:
        # --use_bi_de: the cross-bilateral denoiser of nerf/renderutils; `bilateral=` swaps in another pair
        bi, bi_no_di = bilateral if bilateral is not None else (bilateral_denoiser, bilateral_denoiser_no_di)
        factor = 2.0
        cat = lambda c: torch.cat((c, normal_map, gb_depth), dim=-1)
        denoised_diffuse = bi(framedim_y, framedim_x, cat(total_diff_light), factor)
        denoised_spec = bi(framedim_y, framedim_x, cat(total_spec_light), factor)
        denoised_indirect = bi_no_di(framedim_y, framedim_x, cat(combined_color_indirect), factor)
        denoised_indirect_diff = bi_no_di(framedim_y, framedim_x, cat(total_diff_light_1), factor)
        denoised_indirect_spec = bi_no_di(framedim_y, framedim_x, cat(total_spec_light_1), factor)

    if fused_composite and (occ_map.is_cuda or _host_kernels_bound()):
        final_color = Composite.apply(occ_map, diffuse_map, roughness_specular, denoised_diffuse, denoised_spec,
                                      denoised_indirect)
    else:
        diffuse = diffuse_map * (1.0 - roughness_specular[..., 1:2])
        final_color = diffuse * denoised_diffuse + denoised_spec + denoised_indirect
        final_color = torch.where(occ_map <= 0.1, torch.ones_like(final_color), final_color)
        final_color = torch.nan_to_num(final_color, 0.0)
    return (final_color, denoised_diffuse, denoised_spec, denoised_indirect, denoised_indirect_diff,
            denoised_indirect_spec)
