"""A stand-in for the `slangpy` module the reference imports (nerf/renderer_restir.py:5,19-23,150-187).

`loadModule(path, defines=...)` returns an object that speaks the call protocol the reference's Python uses:

    m.kernel(**kwargs).launchRaw(blockSize=(..), gridSize=(..))       forward launch
    m.kernel.bwd(**kwargs).launchRaw(...)                              reverse-mode launch; differentiable tensors are
                                                                       passed as (primal, grad) tuples
    m.Reservoir(...), m.FinalSample(...), m.pushConstantsMortonCodes(...)   struct constructors (tuples also accepted)

Every kernel name of SURVEY.md section 2.1 is served by the C ABI of libmirres_b200.so; block / grid sizes are ignored
(the library picks its own launch shapes).  A maintainer of the reference makes it run on this library with

    import mirres_restir_nerf_mesh_b200.slangpy_shim as slangpy        # instead of `import slangpy`

(see INTEGRATION.md).  Kernels the reference never calls (SURVEY.md 2.1 "dead Slang kernels") raise AttributeError.
"""
import os
import threading

import torch

from . import kernels as _kernels

_KERNELS = None


def get_kernels():
    global _KERNELS
    if _KERNELS is None:
        _KERNELS = _kernels.Kernels()
    return _KERNELS


def set_kernels(k):
    """Test hook: lets the CPU test-suite bind the host-check flavour of the same kernels."""
    global _KERNELS
    _KERNELS = k
    _WORKSPACES.clear()


def _c(t):
    """Dense view of an input tensor (the reference hands over strided views)."""
    return t if t.is_contiguous() else t.contiguous()


def _primal(x):
    return x[0] if isinstance(x, (tuple, list)) else x


def _grad(x):
    return x[1]


class _Struct(tuple):
    pass


def _reservoir(x):
    if isinstance(x, dict):
        return (x["light_data"], x["light_pdf"], x["M"], x["weight"])
    return tuple(x)


def _final_sample(x):
    if isinstance(x, dict):
        return (x["dir"], x["distance"], x["Li"])
    return tuple(x)


VIS_TAG = "_mirres_vis_tag"


def vis_tag(reservoirs):
    """The visibility tag a driver has put beside a reservoir set (include/mirres_b200.h, mirres_set_visibility_tags), or
    None.  It rides on the light_data tensor of the set as an attribute: it exists only where a driver that owns the
    whole spp loop has attached it (renderer_restir.attach_vis_tags) and goes away with the tensor."""
    return getattr(_reservoir(reservoirs)[0], VIS_TAG, None)


def packed_bvh(info, aabb, vert, tri):
    """Traversal records for reference-layout BVH tensors; cached on the `info` tensor object, which
    restirbvhWorker.update_mesh replaces on every rebuild."""
    cached = getattr(info, "_mirres_packed", None)
    if cached is not None:
        return cached
    k = get_kernels()
    F = tri.shape[0]
    _, nb, tb = k.bvh_sizes(F)
    nodes = torch.empty(nb, dtype=torch.uint8, device=info.device)
    tris = torch.empty(tb, dtype=torch.uint8, device=info.device)
    k.bvh_pack(_c(info), _c(aabb), _c(vert), _c(tri), nodes, tris)
    info._mirres_packed = (nodes, tris)
    return info._mirres_packed


_WORKSPACES = {}


class _Context(threading.local):
    """Launch context of the calling host thread (the `with` blocks below push and pop here): each thread that drives a
    renderer has its own, so two threads on different streams do not see each other's workspace tags or row restrictions."""

    def __init__(self):
        self.ws_tag = ["main"]
        self.skip_prepare = [False]
        self.active_rows = [None]
        self.row_offset = [0]
        self.band = [(0, 0)]


_CTX = _Context()


class workspace_tag:
    """Selects which wavefront workspace the ray-casting kernels launched inside the `with` block use.  Kernels that
    run concurrently on different CUDA streams (the direct-light chain and the indirect path of one spp iteration, see
    renderer_restir.restir_di_with_pt) must not share ray queues, so each chain gets its own workspace."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        _CTX.ws_tag.append(self.tag)

    def __exit__(self, *a):
        _CTX.ws_tag.pop()




class active_rows:
    """Restricts the foreground-pixel lists built inside the block to image rows [y0, y1) of a frame `fx` pixels wide
    (row-band sharding of a frame across GPUs, see dist.py).  The kernels still see the full-frame maps -- neighbour
    and occupancy tests are unchanged -- but only pixels of the band are processed."""

    def __init__(self, y0, y1, fx):
        self.rows = (int(y0), int(y1), int(fx))

    def __enter__(self):
        _CTX.active_rows.append(self.rows)

    def __exit__(self, *a):
        _CTX.active_rows.pop()


def _band_occ(occ):
    rows = _CTX.active_rows[-1]
    if rows is None:
        return occ
    y0, y1, fx = rows
    m = occ.clone()
    flat = m.view(-1)
    flat[:max(y0, 0) * fx] = 0
    flat[max(y1, 0) * fx:] = 0
    return m


class workspace_prepared:
    """Inside the block process_InitialResampling_ does not rebuild the foreground-pixel list of its workspace (the
    driver has already done it with prepare_workspace for the same occupancy map)."""

    def __enter__(self):
        _CTX.skip_prepare.append(True)

    def __exit__(self, *a):
        _CTX.skip_prepare.pop()


class trace_blocks:
    """Blocks per SM of the persistent queue tracers for the ray-casting calls inside the block (mirres_set_tuning).  The
    driver wraps the launches of its critical chain, which it issues on a high-priority stream, so that they get a full
    grid while the background chains keep the small default."""

    def __init__(self, any_blocks=0, closest_blocks=0, mixed_blocks=0):
        self.values = (int(any_blocks), int(closest_blocks), int(mixed_blocks))
        self.saved = None

    def __enter__(self):
        k = get_kernels()
        keys = (k.TUNE_ANY_BLOCKS, k.TUNE_CLOSEST_BLOCKS, k.TUNE_MIXED_BLOCKS)
        # 0 = "no opinion": the caller's own tuning (set through the public set_tuning) stays in force
        self.saved = tuple(k.get_tuning(key) for key in keys)
        for key, v, old in zip(keys, self.values, self.saved):
            if v > 0 and old == 0:
                k.set_tuning(key, v)

    def __exit__(self, *a):
        k = get_kernels()
        for key, old in zip((k.TUNE_ANY_BLOCKS, k.TUNE_CLOSEST_BLOCKS, k.TUNE_MIXED_BLOCKS), self.saved):
            k.set_tuning(key, old)


def prepare_workspace(occ_map):
    """Builds the foreground-pixel list of the current workspace (see workspace_tag) from the primary occupancy."""
    occ = _c(occ_map)
    get_kernels().workspace_prepare(_band_occ(occ), workspace(occ.device, occ.shape[0]))


def workspace(device, n_pixels):
    """Wavefront workspace (include/mirres_b200.h) for frames of n_pixels on `device`, allocated once and reused."""
    # one set of workspaces per host thread: two threads that drive renderers of the same frame size on different
    # streams must not share ray queues and pixel lists (the library's tuning is per thread as well)
    key = (_device_key(device), int(n_pixels), _CTX.ws_tag[-1], threading.get_ident())
    ws = _WORKSPACES.get(key)
    if ws is None:
        nbytes = get_kernels().workspace_bytes(n_pixels)
        buf = torch.zeros(nbytes + 256, dtype=torch.uint8, device=device)
        off = (-buf.data_ptr()) % 256  # CUDA allocations are already 512-byte aligned; host ones are not
        ws = buf[off:off + nbytes]
        _WORKSPACES[key] = ws
        if _FRAME_OFFSET.get(_device_key(device)) is not None:
            _frame_word(ws).copy_(_FRAME_OFFSET[_device_key(device)])
    if _CTX.band[-1] != getattr(ws, "_mirres_band", (0, 0)):
        ws[BAND_BYTES:BAND_BYTES + 8].view(torch.int32).copy_(_band_tensor(ws.device, _CTX.band[-1]))
        ws._mirres_band = _CTX.band[-1]
    if _CTX.row_offset[-1] != getattr(ws, "_mirres_row_offset", 0):
        # stream-ordered, like the frame offset: the launches that follow see the new value
        _row_word(ws).fill_(_CTX.row_offset[-1])
        ws._mirres_row_offset = _CTX.row_offset[-1]
    return ws


FRAME_OFFSET_BYTES = 32  # MIRRES_WORKSPACE_FRAME_OFFSET_BYTES
ROW_OFFSET_BYTES = 36    # MIRRES_WORKSPACE_CTX.row_offset_BYTES
_FRAME_OFFSET = {}


def _row_word(ws):
    return ws[ROW_OFFSET_BYTES:ROW_OFFSET_BYTES + 4].view(torch.int32)


ERROR_BYTES = 48  # MIRRES_WORKSPACE_ERROR_BYTES
BAND_BYTES = 40   # MIRRES_WORKSPACE_BAND_BYTES: two words, rows [lo, hi)
_BAND_TENSORS = {}


def _band_tensor(device, band):
    """Device-resident copy of a (lo, hi) pair, made once per value: writing the words of a workspace is then a
    device-to-device copy, which a CUDA graph capture records like any other launch."""
    key = (str(device), band)
    t = _BAND_TENSORS.get(key)
    if t is None:
        t = _BAND_TENSORS[key] = torch.tensor(band, dtype=torch.int32, device=device)
    return t


class spatial_band:
    """Inside the block the spatial pass resamples rows [lo, hi) of the frame it is handed only; the other pixels of its
    list just publish their samples for those rows to reuse (include/mirres_b200.h, MIRRES_WORKSPACE_BAND_BYTES)."""

    def __init__(self, lo, hi):
        self.band = (int(lo), int(hi))

    def __enter__(self):
        _CTX.band.append(self.band)

    def __exit__(self, *a):
        _CTX.band.pop()



def check_workspaces(device=None, clear=True):
    """Synchronises and raises if a ray-casting launch since the last check had to drop a traversal-stack entry (a mesh
    whose LBVH is deeper than the reference's 64-entry stack: the reference itself has undefined behaviour there)."""
    bad = []
    for key, ws in _WORKSPACES.items():
        if device is not None and key[0] != _device_key(device):
            continue
        word = ws[ERROR_BYTES:ERROR_BYTES + 4].view(torch.int32)
        if int(word.item()) != 0:
            bad.append(key)
            if clear:
                word.zero_()
    if bad:
        raise RuntimeError("traversal stack overflow (more than 64 deferred nodes on one ray) in workspaces %r" % (bad,))


class row_offset:
    """Inside the block the frames handed to the kernels are row bands of a larger frame that start at row `y0`: every
    workspace used inside gets y0 as its row offset (include/mirres_b200.h), so pixels draw the random streams of their
    position in the full frame.  Workspaces are keyed by frame size; one that is used again outside the block is reset."""

    def __init__(self, y0):
        self.y0 = int(y0)

    def __enter__(self):
        _CTX.row_offset.append(self.y0)

    def __exit__(self, *a):
        _CTX.row_offset.pop()


def _frame_word(ws):
    return ws[FRAME_OFFSET_BYTES:FRAME_OFFSET_BYTES + 4].view(torch.int32)


def _device_key(device):
    """'cuda', 'cuda:0' and torch.device('cuda', 0) name the same device: indexed form, current device when no index."""
    d = torch.device(device)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return str(d)


def _frame_offset_tensor(device):
    dev = _device_key(device)
    cur = _FRAME_OFFSET.get(dev)
    if cur is None:
        cur = _FRAME_OFFSET[dev] = torch.zeros(1, dtype=torch.int32, device=device)
    return cur


def set_frame_offset(device, value):
    """Adds `value` to the frame index of every subsequent launch on `device` (all workspaces), stream-ordered and
    without touching any launch argument -- which is what lets a captured CUDA graph be replayed with fresh random
    streams.  0 restores the reference schedule."""
    dev = _device_key(device)
    cur = _frame_offset_tensor(device)
    cur.fill_(int(value) & 0x7FFFFFFF)
    for key, ws in list(_WORKSPACES.items()):
        if key[0] == dev:
            _frame_word(ws).copy_(cur)


class _Launch:
    def __init__(self, fn, kw):
        self._fn, self._kw = fn, kw

    def launchRaw(self, blockSize=None, gridSize=None):
        self._fn(**self._kw)


class _Kernel:
    def __init__(self, module, fwd, bwd=None):
        self._module, self._fwd, self._bwd = module, fwd, bwd

    def __call__(self, **kw):
        return _Launch(lambda **k: self._fwd(self._module, **k), kw)

    def bwd(self, **kw):
        if self._bwd is None:
            raise NotImplementedError("this kernel is not differentiable in the reference either")
        return _Launch(lambda **k: self._bwd(self._module, **k), kw)


# ---- kernel bodies: reference keyword names -> C ABI ------------------------------------------------------------------
def _generateElements(m, vert, v_indx, ele_primitiveIdx, ele_aabb):
    get_kernels().bvh_elements(_c(vert), _c(v_indx), ele_primitiveIdx, ele_aabb)


def _morton_codes(m, pc, ele_aabb, morton_codes_ele):
    ext = [float(pc[k]) for k in ("g_min_x", "g_min_y", "g_min_z", "g_max_x", "g_max_y", "g_max_z")]
    get_kernels().bvh_morton(ele_aabb, ext, morton_codes_ele)


def _scratch(F, device):
    k = get_kernels()
    return torch.empty(k.bvh_sizes(F)[0], dtype=torch.uint8, device=device)


def _radix_sort(m, g_num_elements, g_elements_in, g_elements_out):
    get_kernels().bvh_sort(g_elements_in, _scratch(int(g_num_elements), g_elements_in.device))


def _hierarchy(m, g_num_elements, ele_primitiveIdx, ele_aabb, g_sorted_morton_codes, g_lbvh_info, g_lbvh_aabb,
               g_lbvh_construction_infos):
    # builds the hierarchy AND the final bounding boxes (one bottom-up pass); the level-by-level kernels below
    # therefore have nothing left to do
    get_kernels().bvh_hierarchy_refit(g_sorted_morton_codes, ele_aabb, g_lbvh_info, g_lbvh_aabb,
                                      _scratch(int(g_num_elements), g_lbvh_info.device))


def _get_bvh_height(m, g_num_elements, g_lbvh_info, g_lbvh_aabb, g_lbvh_construction_infos, tree_heights):
    tree_heights.zero_()  # => the reference's `for i in range(tree_height_max)` loop runs zero times


def _noop(m, **kw):
    pass


def _make_sampleable(m, env_tex, weight, width, height):
    get_kernels().env_weights(_c(env_tex), int(width), int(height), weight)


def _Distribution2D(m, w, h, pdf_, cdf_):
    get_kernels().env_distribution2d(int(w), int(h), pdf_, cdf_)


def _createNeighborOffsetTexture(m, sampleCount, neighborOffsets):
    get_kernels().neighbor_offsets(int(sampleCount), neighborOffsets)


def _GenerateLightTiles(m, env_tex, pdf_, cdf_, mpdf_, mcdf_, width, height, frameIndex, light_data, light_uv,
                        light_inv_pdf, debug_out=None):
    # per-slot (direction, radiance) cache, kept on the light_data tensor it describes; rewritten on every call
    cache = getattr(light_data, "_mirres_cache", None)
    if cache is None or cache.shape[0] != light_data.shape[0] or cache.device != light_data.device:
        cache = torch.empty((light_data.shape[0], 8), dtype=torch.float32, device=light_data.device)
        light_data._mirres_cache = cache
    get_kernels().light_tiles(_c(env_tex), int(width), int(height), (pdf_, cdf_, mpdf_, mcdf_), int(frameIndex),
                              m.define("LIGHT_TILE_COUNT", 128), m.define("LIGHT_TILE_SIZE", 1024), light_data, light_uv,
                              light_inv_pdf, cache, _frame_offset_tensor(light_data.device))


def _InitialResampling(m, g_lbvh_info, g_lbvh_aabb, vert, v_indx, pos_map, reservoirs, env_tex, env_width, env_height,
                       framedim_x, framedim_y, frameIndex, occ_map, normal_depth, brdf_map, ray_dir, pdf_, cdf_, mpdf_,
                       mcdf_, light_data, light_uv, light_inv_pdf):
    # first kernel of every spp iteration that sees the primary occupancy: (re)build the foreground-pixel list here
    occ = _c(occ_map)
    ws = workspace(occ.device, occ.shape[0])
    if not _CTX.skip_prepare[-1]:
        get_kernels().workspace_prepare(_band_occ(occ), ws)
    get_kernels().initial_resampling(packed_bvh(g_lbvh_info, g_lbvh_aabb, vert, v_indx), _c(pos_map),
                                     _reservoir(reservoirs), _c(env_tex), int(env_width), int(env_height),
                                     int(framedim_x), int(framedim_y), int(frameIndex), occ, _c(normal_depth),
                                     _c(brdf_map), _c(ray_dir), pdf_, mpdf_, light_data, light_inv_pdf, ws,
                                     m.define("LIGHT_TILE_COUNT", 128), m.define("LIGHT_TILE_SIZE", 1024),
                                     m.define("SCREEN_TILE_SIZE", 8), m.define("INITIAL_LIGHT_SAMPLE_COUNT", 32),
                                     m.define("INITIAL_BRDF_SAMPLE_COUNT", 1),
                                     light_cache=getattr(light_data, "_mirres_cache", None), vis_tag=vis_tag(reservoirs))


def _TemporalResampling(m, reservoirs, prevReservoirs, env_tex, env_width, env_height, framedim_x, framedim_y,
                        frameIndex, occ_map, normal_depth, brdf_map, ray_dir, prev_occ_map, prev_normal_depth,
                        prev_brdf_map, prev_ray_dir, motionVectors=None):
    get_kernels().temporal_resampling(_reservoir(reservoirs), _reservoir(prevReservoirs), _c(env_tex), int(env_width),
                                      int(env_height), int(framedim_x), int(framedim_y), int(frameIndex), _c(occ_map),
                                      _c(normal_depth), _c(brdf_map), _c(ray_dir), _c(prev_occ_map),
                                      _c(prev_normal_depth), _c(prev_brdf_map), _c(prev_ray_dir),
                                      workspace(occ_map.device, occ_map.shape[0]),
                                      None if motionVectors is None else _c(motionVectors),
                                      m.define("MAX_HISTORY_LENGTH", 20), vis_tag=vis_tag(reservoirs),
                                      prev_vis_tag=vis_tag(prevReservoirs))


def _SpatialResampling(m, g_lbvh_info, g_lbvh_aabb, vert, v_indx, pos_map, reservoirs, prevReservoirs, neighborOffsets,
                       env_tex, env_width, env_height, framedim_x, framedim_y, frameIndex, occ_map, normal_depth,
                       brdf_map, ray_dir):
    get_kernels().spatial_resampling(packed_bvh(g_lbvh_info, g_lbvh_aabb, vert, v_indx), _c(pos_map),
                                     _reservoir(reservoirs), _reservoir(prevReservoirs), _c(neighborOffsets),
                                     _c(env_tex), int(env_width), int(env_height), int(framedim_x), int(framedim_y),
                                     int(frameIndex), _c(occ_map), _c(normal_depth), _c(brdf_map), _c(ray_dir),
                                     workspace(occ_map.device, occ_map.shape[0]), m.define("NEIGHBOR_OFFSET_COUNT", 8192), m.define("NEIGHBOR_COUNT", 5),
                                     float(m.define("GATHER_RADIUS", 30)), vis_tag=vis_tag(reservoirs),
                                     prev_vis_tag=vis_tag(prevReservoirs))


def _get_vis(m, g_lbvh_info, g_lbvh_aabb, vert, v_indx, reservoirs, framedim_x, framedim_y, pos_map, vis_map):
    get_kernels().final_visibility(packed_bvh(g_lbvh_info, g_lbvh_aabb, vert, v_indx), _reservoir(reservoirs)[0],
                                   int(framedim_x), int(framedim_y), _c(pos_map), vis_map,
                                   workspace(vis_map.device, vis_map.shape[0]), vis_tag=vis_tag(reservoirs))


def _eval_final_fwd(m, reservoirs, env_tex, env_width, env_height, framedim_x, framedim_y, finalSample, vis_map):
    fs = _final_sample(finalSample)
    get_kernels().eval_final_fwd(_reservoir(reservoirs), _c(_primal(env_tex)), int(env_width), int(env_height),
                                 int(framedim_x), int(framedim_y), fs[0], fs[1], _primal(fs[2]), vis_map)


def _eval_final_bwd(m, reservoirs, env_tex, env_width, env_height, framedim_x, framedim_y, finalSample, vis_map):
    fs = _final_sample(finalSample)
    get_kernels().eval_final_bwd(_reservoir(reservoirs), int(env_width), int(env_height), int(framedim_x),
                                 int(framedim_y), vis_map, _c(_grad(fs[2])), _grad(env_tex))


def _final_shading_fwd(m, finalSample, env_tex, env_width, env_height, framedim_x, framedim_y, occ_map, normal, ray_dir,
                       diffuse_map, linearRoughness_specular_map, color, diff_light, spec_light):
    fs = _final_sample(finalSample)
    get_kernels().final_shading_fwd(fs[0], fs[1], _c(_primal(fs[2])), _c(env_tex), int(env_width), int(env_height),
                                    int(framedim_x), int(framedim_y), _c(occ_map), _c(_primal(normal)), _c(ray_dir),
                                    _c(_primal(diffuse_map)), _c(_primal(linearRoughness_specular_map)),
                                    _primal(color), _primal(diff_light), _primal(spec_light))


def _final_shading_bwd(m, finalSample, env_tex, env_width, env_height, framedim_x, framedim_y, occ_map, normal, ray_dir,
                       diffuse_map, linearRoughness_specular_map, color, diff_light, spec_light):
    fs = _final_sample(finalSample)
    get_kernels().final_shading_bwd(fs[0], fs[1], _c(_primal(fs[2])), int(framedim_x), int(framedim_y), _c(occ_map),
                                    _c(_primal(normal)), _c(ray_dir), _c(_primal(diffuse_map)),
                                    _c(_primal(linearRoughness_specular_map)), _c(_grad(color)), _c(_grad(diff_light)),
                                    _c(_grad(spec_light)), _grad(normal), _grad(diffuse_map),
                                    _grad(linearRoughness_specular_map), _grad(fs[2]))


def _new_dir(m, g_lbvh_info, g_lbvh_aabb, vert, v_indx, frameIndex, bounce_count, framedim_x, framedim_y, occ_map,
             pos_map, normal, ray_dir, prd, diffuse_map, linearRoughness_specular_map, new_pos_map, new_ray_d,
             new_occ_map, new_normal):
    if _CTX.ws_tag[-1] != "main" and int(bounce_count) == 0 and not _CTX.skip_prepare[-1]:
        # a chain with its own workspace builds its own foreground-pixel list from the primary occupancy (unless the
        # driver has done so once for the whole loop: workspace_prepared)
        get_kernels().workspace_prepare(_band_occ(_c(occ_map)), workspace(prd.device, prd.shape[0]))
    get_kernels().bounce_first(packed_bvh(g_lbvh_info, g_lbvh_aabb, vert, v_indx), int(frameIndex), int(bounce_count),
                               m.define("MAX_Bounce", 2), int(framedim_x), int(framedim_y), _c(occ_map), _c(pos_map),
                               _c(normal), _c(ray_dir), prd, _c(diffuse_map), _c(linearRoughness_specular_map),
                               new_pos_map, new_ray_d, new_occ_map, new_normal, workspace(prd.device, prd.shape[0]))


def _path_tracing(m, g_lbvh_info, g_lbvh_aabb, vert, v_indx, frameIndex, bounce_count, framedim_x, framedim_y, env_tex,
                  env_width, env_height, pdf_, cdf_, mpdf_, mcdf_, occ_map, pos_map, normal, ray_dir, prd, diffuse_map,
                  linearRoughness_specular_map, color, diff_color, spec_color, new_pos_map, new_ray_d, new_occ_map,
                  new_normal):
    get_kernels().bounce_shade(packed_bvh(g_lbvh_info, g_lbvh_aabb, vert, v_indx), int(frameIndex), int(bounce_count),
                               m.define("MAX_Bounce", 2), int(framedim_x), int(framedim_y), _c(env_tex), int(env_width),
                               int(env_height), (pdf_, cdf_, mpdf_, mcdf_), _c(occ_map), _c(pos_map), _c(normal),
                               _c(ray_dir), prd, _c(diffuse_map), _c(linearRoughness_specular_map), color, diff_color,
                               spec_color, new_pos_map, new_ray_d, new_occ_map, new_normal,
                               workspace(prd.device, prd.shape[0]))


def _phi(PHI):
    if isinstance(PHI, dict):
        PHI = (PHI["c_phi"], PHI["n_phi"], PHI["p_phi"])
    return tuple(float(x) for x in PHI)


def _eaw_fwd(m, PHI, framedim_x, framedim_y, stepWidth, occ_map, color, normal_map, pos_map, out_color):
    c, n, p = _phi(PHI)
    get_kernels().eaw_fwd(c, n, p, int(framedim_x), int(framedim_y), stepWidth, _c(occ_map), _c(_primal(color)),
                          _c(_primal(normal_map)), _c(_primal(pos_map)), _primal(out_color))


def _eaw_bwd(m, PHI, framedim_x, framedim_y, stepWidth, occ_map, color, normal_map, pos_map, out_color):
    c, n, p = _phi(PHI)
    occ = _c(occ_map)
    scratch = torch.empty(occ.shape[0], dtype=torch.float32, device=occ.device)
    get_kernels().eaw_bwd(c, n, p, int(framedim_x), int(framedim_y), stepWidth, occ, _c(_primal(color)),
                          _c(_primal(normal_map)), _c(_primal(pos_map)), _c(_primal(out_color)), _c(_grad(out_color)),
                          _grad(color), _grad(normal_map), _grad(pos_map), scratch)


def _normal_ao(m, framedim_x, framedim_y, occ_map, normal_map, ray_dir, out_ao):
    get_kernels().normal_ao(int(framedim_x), int(framedim_y), _c(occ_map), _c(normal_map), out_ao)


_TABLE = {
    "get_elements": {"generateElements": (_generateElements, None)},
    "lbvh_morton_codes": {"morton_codes": (_morton_codes, None)},
    "lbvh_single_radixsort": {"radix_sort": (_radix_sort, None)},
    "lbvh_hierarchy": {"hierarchy": (_hierarchy, None)},
    "lbvh_bounding_boxes": {"get_bvh_height": (_get_bvh_height, None), "get_bbox": (_noop, None),
                            "set_root": (_noop, None)},
    "make_sampleable": {"make_sampleable": (_make_sampleable, None), "Distribution2D": (_Distribution2D, None),
                        "createNeighborOffsetTexture": (_createNeighborOffsetTexture, None)},
    "GenerateLightTiles": {"process_GenerateLightTiles": (_GenerateLightTiles, None)},
    "InitialResampling": {"process_InitialResampling_": (_InitialResampling, None)},
    "TemporalResampling": {"process_TemporalResampling": (_TemporalResampling, None)},
    "SpatialResampling": {"process_SpatialResampling_": (_SpatialResampling, None)},
    "EvaluateFinalSamples": {"process_EvaluateFinalSamples_get_vis": (_get_vis, None),
                             "process_EvaluateFinalSamples_di_": (_eval_final_fwd, _eval_final_bwd)},
    "FinalShading": {"process_FinalShading": (_final_shading_fwd, _final_shading_bwd),
                     "process_new_dir_for_pt": (_new_dir, None),
                     "process_path_tracing_divided_no_grad": (_path_tracing, None)},
    "EAWDenoise": {"process_EAWDenoise": (_eaw_fwd, _eaw_bwd), "process_EAWDenoise_no_di": (_eaw_fwd, None),
                   "process_normal_ao": (_normal_ao, None)},
}


class Module:
    def __init__(self, path, defines=None):
        self.path = path
        self.name = os.path.splitext(os.path.basename(path))[0]
        self.defines = dict(defines or {})
        if self.name not in _TABLE:
            raise FileNotFoundError("mirres-b200 serves no kernels for %r" % path)
        for kname, (fwd, bwd) in _TABLE[self.name].items():
            setattr(self, kname, _Kernel(self, fwd, bwd))

    def define(self, key, default):
        return int(self.defines.get(key, default))

    # struct constructors used by the reference wrappers (Resampling.py:132-137,198-205; renderer_restir.py:44-47)
    @staticmethod
    def Reservoir(light_data, light_pdf, M, weight):
        return _Struct((light_data, light_pdf, M, weight))

    @staticmethod
    def FinalSample(dir, distance, Li):
        return _Struct((dir, distance, Li))

    @staticmethod
    def pushConstantsMortonCodes(**kw):
        return dict(kw)

    @staticmethod
    def phi(c_phi, n_phi, p_phi):
        return (c_phi, n_phi, p_phi)


def loadModule(path, defines=None, **unused):
    return Module(path, defines)
