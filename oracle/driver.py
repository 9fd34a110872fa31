"""Oracle restatement of the host loop -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

Follows nerf/renderer_restir.py:230-471 (restir_di_with_pt) and :473-515 (the pre/post-processing of
run_restir_di_with_pt up to, not including, the denoiser) on numpy arrays, calling the oracle kernels.
Differences from the reference, all explicit arguments:
  * `random_offset` is passed in (reference: np.random.randint(2**20), renderer_restir.py:245);
  * `max_bounce` generalises the hard-coded MAX_Bounce = 2 (FinalShading.slang:7) and the two hand-unrolled
    shade calls (renderer_restir.py:396-454);
  * the material callback is `material(pos[M,3]) -> (kd[M,3], rough[M,1], metal[M,1])` (reference: mlp_mat.sample_no_di).
"""
import numpy as np

from . import oracle as O

TOTAL_RIS_PASSES = 20  # renderer_restir.py:242


def l2_normalize(x, eps=1e-6):
    # safe_l2_normalize (nerf/render_dump.py:5-6) with an explicit, order-defined norm
    n = np.sqrt(x[:, 0:1] * x[:, 0:1] + x[:, 1:2] * x[:, 1:2] + x[:, 2:3] * x[:, 2:3])
    return (x / np.maximum(n, np.float32(eps))).astype(np.float32)


def prepare_gbuffer(g):
    """renderer_restir.py:484-486 and :279-287."""
    g = {k: np.array(v, dtype=np.float32, copy=True) for k, v in g.items()}
    g["occ_map"][g["occ_map"][:, 0] <= 0.5, :] = 0
    g["ray_dir_map"] = l2_normalize(g["ray_dir_map"])
    kd, rs = g["diffuse_map"], g["roughness_specular"]
    f = np.float32
    lum = kd[:, 0:1] * f(0.2126) + kd[:, 1:2] * f(0.7152) + kd[:, 2:3] * f(0.0722)
    met = rs[:, 1:2] * f(0.2126) + rs[:, 1:2] * f(0.7152) + rs[:, 1:2] * f(0.0722)
    a = np.clip(rs[:, 0:1], f(0.01), f(1.0))
    g["brdf_map"] = np.ascontiguousarray(np.concatenate([lum, met, a * a], axis=1), np.float32)
    g["normal_depth"] = np.ascontiguousarray(np.concatenate([g["normal_map"], g["depth_map"]], axis=1), np.float32)
    return g


def restir_di_with_pt(bvh, env_map, g, spp, fx, fy, random_offset, material, max_bounce=2, tile_count=128,
                      tile_size=1024, counters=None, snapshots=None, motion=None):
    """g: prepared G-buffer dict (prepare_gbuffer).  Returns dict of per-call sums and mFrameIndex.
    motion: optional [n,2] motion vectors of the temporal pass (the reference passes zeros, renderer_restir.py:487).
    The sample-provenance tags of orc_kernels.cpp ride along (they change no result): tot["known_final_rays"] counts the
    final-visibility rays whose answer an earlier pass had produced, tot["provenance_violations"] those of them that
    were occluded after all (must be 0)."""
    n = fx * fy
    He, We = env_map.shape[0], env_map.shape[1]
    # `counters` may be one array (everything accumulated) or a dict with one array per kernel
    ctr = (lambda name: counters.setdefault(name, O.new_counters())) if isinstance(counters, dict) else (lambda name: counters)
    env = np.ascontiguousarray(env_map[::-1].reshape(-1, 3), np.float32)  # torch.flip(dims=[0]).reshape(-1,3)
    dist = O.env_build_distribution(env, We, He)
    offs = (O.neighbor_offsets(8192).reshape(-1, 2) / np.float32(127)).astype(np.float32)
    z3 = lambda: np.zeros((n, 3), np.float32)
    z1 = lambda: np.zeros((n, 1), np.float32)
    tot = dict(color=z3(), diff=z3(), spec=z3(), color_1=z3(), diff_1=z3(), spec_1=z3())
    color_1, cdiff_1, cspec_1 = z3(), z3(), z3()
    prd = np.zeros((n, 5), np.float32)
    A = dict(pos=z3(), ray=z3(), occ=z1(), nrm=z3())
    B = dict(pos=z3(), ray=z3(), occ=z1(), nrm=z3())
    new_diffuse, new_rs = z3(), np.zeros((n, 2), np.float32)
    vis = np.ones((n, 1), np.float32)
    res, prev = O.new_reservoirs(n), O.new_reservoirs(n)
    tag_res, tag_prev = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    violations0, known_final = O.provenance_violations(), []
    fs_dir, fs_dist, fs_Li = z3(), z1(), z3()
    occ, nd, brdf, ray, pos = g["occ_map"], g["normal_depth"], g["brdf_map"], g["ray_dir_map"], g["pos_map"]
    normal, kd, rs = g["normal_map"], g["diffuse_map"], g["roughness_specular"]
    prev_occ, prev_nd, prev_brdf, prev_ray = z1(), np.zeros((n, 4), np.float32), z3(), z3()
    mFrame = 0
    for i in range(spp):
        cur = 0
        base = random_offset + TOTAL_RIS_PASSES * mFrame
        tiles = O.light_tiles(env, We, He, dist, base + cur, tile_count, tile_size)
        cur += 2
        O.set_provenance(tag_res, None)
        O.initial_resampling(bvh, pos, res, env, We, He, fx, fy, base + cur, occ, nd, brdf, ray, dist, tiles,
                             tile_count, tile_size, counters=ctr("initial_resampling"))
        cur += 1
        if i > 0:
            O.set_provenance(tag_res, tag_prev)  # prev: the finished set of the previous iteration, same pos_map and tree
            O.temporal_resampling(res, prev, env, We, He, fx, fy, base + cur, occ, nd, brdf, ray, prev_occ, prev_nd,
                                  prev_brdf, prev_ray, motion=motion)
            cur += 1
        res, prev = prev, res
        tag_res, tag_prev = tag_prev, tag_res
        O.set_provenance(tag_res, tag_prev)
        O.spatial_resampling(bvh, pos, res, prev, offs, env, We, He, fx, fy, base + cur, occ, nd, brdf, ray,
                             counters=ctr("spatial_resampling"))
        cur += 1
        O.set_provenance(tag_res, None)
        O.final_visibility(bvh, res, fx, fy, pos, vis, counters=ctr("final_visibility"))
        O.set_provenance(None, None)
        known_final.append(int(((res[0][:, 0] > 0.1) & (tag_res != 0)).sum()))
        O.eval_final_fwd(res, env, We, He, fx, fy, fs_dir, fs_dist, fs_Li, vis)
        color, cdiff, cspec = O.final_shading_fwd(fs_dir, fs_dist, fs_Li, env, We, He, fx, fy, occ, normal, ray, kd, rs)
        if snapshots is not None:
            snapshots.append(dict(res=[a.copy() for a in res], prev=[a.copy() for a in prev], vis=vis.copy(),
                                  fs_dir=fs_dir.copy(), fs_dist=fs_dist.copy(), fs_Li=fs_Li.copy(),
                                  color=color.copy(), diff=cdiff.copy(), spec=cspec.copy(), tiles=tiles))
        O.bounce_first(bvh, base + cur, 0, max_bounce, fx, fy, occ, pos, normal, ray, prd, kd, rs, A["pos"], A["ray"],
                       A["occ"], A["nrm"], counters=ctr("bounce_first"))
        cur += 5
        src, dst = A, B
        for b in range(1, max_bounce + 1):
            idx = np.where(src["occ"][:, 0] >= 0.5)[0]
            mkd, mr, mm = material(src["pos"][idx])
            new_diffuse[idx] = mkd
            new_rs[idx] = np.concatenate([mr, mm], axis=1)
            O.bounce_shade(bvh, base + cur, b, max_bounce, fx, fy, env, We, He, dist, src["occ"], src["pos"],
                           src["nrm"], src["ray"], prd, new_diffuse, new_rs, color_1, cdiff_1, cspec_1, dst["pos"],
                           dst["ray"], dst["occ"], dst["nrm"], counters=ctr("bounce_shade"))
            tot["color_1"] += color_1
            tot["diff_1"] += cdiff_1
            tot["spec_1"] += cspec_1
            if snapshots is not None:
                snapshots[-1]["bounce%d" % b] = dict(color=color_1.copy(), diff=cdiff_1.copy(), spec=cspec_1.copy(),
                                                     occ=dst["occ"].copy(), pos=dst["pos"].copy(), prd=prd.copy())
            cur += 5
            src, dst = dst, src
        mFrame += 1
        res, prev = prev, res
        tag_res, tag_prev = tag_prev, tag_res
        prev_occ, prev_nd, prev_brdf, prev_ray = occ, nd, brdf, ray
        tot["color"] += color
        tot["diff"] += cdiff
        tot["spec"] += cspec
    tot["mFrameIndex"] = mFrame
    tot["known_final_rays"] = known_final
    tot["provenance_violations"] = O.provenance_violations() - violations0
    return tot


def run_no_denoise(bvh, env_map, gbuffer, spp, fx, fy, random_offset, material, max_bounce=2, counters=None,
                   snapshots=None, motion=None):
    """run_restir_di_with_pt (renderer_restir.py:473-549) with the denoiser replaced by identity."""
    g = prepare_gbuffer(gbuffer)
    tot = restir_di_with_pt(bvh, env_map, g, spp, fx, fy, random_offset, material, max_bounce, counters=counters,
                            snapshots=snapshots, motion=motion)
    m = np.float32(tot["mFrameIndex"])
    out = {k: (tot[k] / m).astype(np.float32) for k in ("color", "diff", "spec", "color_1", "diff_1", "spec_1")}
    indirect = out["diff_1"] + out["spec_1"]
    kd = g["diffuse_map"] * (np.float32(1.0) - g["roughness_specular"][:, 1:2])
    final = kd * out["diff"] + out["spec"] + indirect
    final[g["occ_map"][:, 0] <= 0.1, :] = 1.0
    out["final"] = np.nan_to_num(final, nan=0.0).astype(np.float32)
    out["indirect"] = indirect
    out["prepared"] = g
    out["known_final_rays"] = tot["known_final_rays"]
    out["provenance_violations"] = tot["provenance_violations"]
    return out
