"""How much can "parity unpinned" hide?  (SURVEY.md 8c residual risk.)

The reference's real binary (slangc -> nvcc) contracts a*b+c into FMA and calls libdevice sin / cos / acos / atan2 / pow;
neither can be reproduced here, and the oracle + kernels deliberately do NOT (include/mirres_fpmath.h) so that their
outputs can be compared bit for bit.  This tool runs the ORACLE in two flavours on identical inputs,

    contract   -ffp-contract=off, correctly rounded double-precision polynomials           (what the tests pin)
    fast       -ffp-contract=fast -mfma, C-library sinf/cosf/acosf/atan2f/expf, powf       (reference-like numerics)

and reports what moves: LBVH topology, primary hit ids, light-tile texels, reservoir selections, visibility flags (integer
decisions), and the relative error of radiance where the decisions agree, plus the image-level difference of the spp
averages.  A flipped decision replaces a sample by another valid sample of the same estimator (the images differ by
noise, not by bias); the float columns say how far outputs move when nothing flips.

    python tools/numerics_sensitivity.py [C1 C2 ...]  ->  profiles/numerics_sensitivity.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O, driver as D, backward as BW  # noqa: E402
from mirres_restir_nerf_mesh_b200 import synth  # noqa: E402


def run(name, spp=None, random_offset=4242, metallic=0.0):
    cfg = synth.CONFIGS[name]
    v, f = synth.make_mesh(cfg)
    W, H = cfg["W"], cfg["H"]
    ro, rd = synth.camera_rays(W, H)
    env = synth.envmap(*cfg["env"])
    b = O.Bvh(v, f)
    hit, t, pos, nrm, prim = O.trace(b, ro, rd)
    g = synth.gbuffer_from_hits(ro, rd, hit, t, pos, nrm, metallic=metallic)
    snaps = []
    out = D.run_no_denoise(b, env, g, spp or cfg["spp"], W, H, random_offset, lambda p: synth.material(p, metallic),
                           max_bounce=cfg["max_bounce"], snapshots=snaps)
    return dict(bvh=b, hit=hit, t=t, pos=pos, nrm=nrm, prim=prim, g=g, snaps=snaps, out=out)


def relerr(a, b, floor=1e-6):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def gradients(snap, g, env_shape, weights):
    """Gradients that leave the path for ONE shading pass, from the float64 backward oracle evaluated on a
    flavour's own forward state: d loss / d (normal, kd, roughness-metallic) per pixel and the envmap scatter, for
    loss = sum(color * weights)."""
    gN, gK, gR, gL = BW.final_shading_grads(snap["fs_dir"], snap["fs_dist"], snap["fs_Li"], g["occ_map"], g["normal_map"],
                                            g["ray_dir_map"], g["diffuse_map"], g["roughness_specular"], weights,
                                            np.zeros_like(weights), np.zeros_like(weights))
    gE = BW.eval_final_grad_env(snap["res"][0], snap["res"][3], snap["vis"], gL.astype(np.float32), env_shape[1], env_shape[0])
    return {"normal": gN, "kd": gK, "rough_metal": gR, "env": gE}


def compare(name, spp=None, metallic=0.0):
    """The same G-buffer goes into both flavours' spp loops (the contract flavour's), so that a flipped primary hit does not
    mask the sensitivity of the ReSTIR / path kernels; the primary-ray flips are reported on their own."""
    A = run(name, spp, metallic=metallic)
    with O.flavour("fast"):
        B = run(name, spp, metallic=metallic)
        cfg = synth.CONFIGS[name]
        snapsB = []
        outB = D.run_no_denoise(B["bvh"], synth.envmap(*cfg["env"]), A["g"], spp or cfg["spp"], cfg["W"], cfg["H"], 4242,
                                lambda p: synth.material(p, metallic), max_bounce=cfg["max_bounce"], snapshots=snapsB)
    rep = {"config": name, "spp": spp or synth.CONFIGS[name]["spp"], "pixels": int(len(A["hit"])),
           "foreground_pixels": int((A["hit"] > 0).sum())}
    rep["lbvh"] = {"sorted_codes_equal": bool(np.array_equal(A["bvh"].sorted_codes, B["bvh"].sorted_codes)),
                   "topology_equal": bool(np.array_equal(A["bvh"].info, B["bvh"].info)),
                   "node_boxes_equal": bool(np.array_equal(A["bvh"].aabb, B["bvh"].aabb)),
                   "leaf_order_differs": int((A["bvh"].sorted_codes[:, 1] != B["bvh"].sorted_codes[:, 1]).sum())}
    both = (A["hit"] > 0) & (B["hit"] > 0)
    rep["primary_rays"] = {"hit_flag_flips": int((A["hit"] != B["hit"]).sum()), "primitive_id_flips": int((A["prim"][both] != B["prim"][both]).sum()),
                           "t_max_rel_err": float(relerr(A["t"][both], B["t"][both]).max()),
                           "normal_max_abs_err": float(np.abs(A["nrm"][both] - B["nrm"][both]).max())}
    fg = A["g"]["occ_map"][:, 0] > 0.5
    its = []
    for i, (sa, sb) in enumerate(zip(A["snaps"], snapsB)):
        uv_flip = int((sa["tiles"][1] != sb["tiles"][1]).any(axis=1).sum())
        # a reservoir selection is its stored light sample (validity flag + octahedral direction) and M
        same_sel = (np.abs(sa["res"][0] - sb["res"][0]).max(axis=1) <= 1e-6) & (sa["res"][2][:, 0] == sb["res"][2][:, 0])
        same_vis = sa["vis"][:, 0] == sb["vis"][:, 0]
        agree = fg & same_sel & same_vis
        w_err = relerr(sa["res"][3][agree, 0], sb["res"][3][agree, 0])
        li_err = relerr(sa["fs_Li"][agree], sb["fs_Li"][agree], 1e-4)
        col_err = relerr(sa["color"][agree], sb["color"][agree], 1e-4)
        its.append({"iteration": i, "light_tile_texel_flips": uv_flip, "light_tile_samples": int(len(sa["tiles"][1])),
                    "reservoir_selection_flips": int((fg & ~same_sel).sum()), "visibility_flips": int((fg & same_sel & ~same_vis).sum()),
                    "pixels_with_equal_decisions": int(agree.sum()),
                    "reservoir_weight_rel_err": {"max": float(w_err.max()) if w_err.size else 0.0, "p999": float(np.quantile(w_err, 0.999)) if w_err.size else 0.0},
                    "Li_rel_err": {"max": float(li_err.max()) if li_err.size else 0.0, "p999": float(np.quantile(li_err, 0.999)) if li_err.size else 0.0},
                    "direct_colour_rel_err": {"max": float(col_err.max()) if col_err.size else 0.0,
                                              "p999": float(np.quantile(col_err, 0.999)) if col_err.size else 0.0}})
    rep["iterations"] = its
    # gradients of the last shading pass under both flavours' forward states (same upstream weights).  Per-pixel
    # gradients are compared where the decisions agree; the envmap gradient is a scatter over texels, compared as a
    # whole (a flipped selection moves its contribution to another texel) relative to its largest entry.
    gA_in = D.prepare_gbuffer(A["g"])
    cfg = synth.CONFIGS[name]
    env_shape = synth.envmap(*cfg["env"]).shape
    rng = np.random.default_rng(3)
    wts = rng.uniform(0.5, 1.5, size=(len(fg), 3)).astype(np.float32)
    ga = gradients(A["snaps"][-1], gA_in, env_shape, wts)
    gb = gradients(snapsB[-1], gA_in, env_shape, wts)
    sa, sb = A["snaps"][-1], snapsB[-1]
    agree = fg & (np.abs(sa["res"][0] - sb["res"][0]).max(axis=1) <= 1e-6) & (sa["vis"][:, 0] == sb["vis"][:, 0])
    grads = {}
    for k in ("normal", "kd", "rough_metal"):
        scale = np.abs(ga[k]).max()
        d = np.abs(ga[k][agree] - gb[k][agree]).max(axis=1) / max(scale, 1e-30)
        grads[k] = {"max_err_over_scale": float(d.max()) if d.size else 0.0,
                    "p999_err_over_scale": float(np.quantile(d, 0.999)) if d.size else 0.0, "scale": float(scale)}
    scale = np.abs(ga["env"]).max()
    grads["env"] = {"max_err_over_scale": float(np.abs(ga["env"] - gb["env"]).max() / max(scale, 1e-30)),
                    "rel_diff_of_sum": float(abs(ga["env"].sum() - gb["env"].sum()) / max(abs(ga["env"].sum()), 1e-30)),
                    "scale": float(scale), "note": "includes the pixels whose decisions flipped"}
    rep["gradients_last_pass"] = grads
    img = {}
    for k in ("color", "color_1", "final"):
        a, b = A["out"][k][fg].astype(np.float64), outB[k][fg].astype(np.float64)
        img[k] = {"mean_abs_diff": float(np.abs(a - b).mean()), "mean_value": float(np.abs(a).mean()),
                  "rel_diff_of_mean": float(abs(a.mean() - b.mean()) / max(abs(a.mean()), 1e-12)),
                  "pixels_bit_equal_fraction": float((A["out"][k][fg] == outB[k][fg]).all(axis=1).mean())}
    rep["images"] = img
    return rep


def main(names):
    out = {"what": __doc__.split("\n\n")[1].replace("\n", " "), "configs": [compare(n) for n in names]}
    # the metallic material exercises the specular lobe (pow5 Fresnel, GGX) on the small scene
    out["configs"].append(dict(compare("C1", metallic=0.4), variant="metallic 0.4"))
    path = os.path.join(ROOT, "profiles", "numerics_sensitivity.json")
    json.dump(out, open(path, "w"), indent=1)
    for c in out["configs"]:
        it = c["iterations"][-1]
        print(c["config"], c.get("variant", ""), "| lbvh equal:", c["lbvh"]["topology_equal"], "| prim flips:", c["primary_rays"]["primitive_id_flips"],
              "| tile flips:", it["light_tile_texel_flips"], "| selection flips: %d of %d" % (it["reservoir_selection_flips"], c["foreground_pixels"]),
              "| vis flips:", it["visibility_flips"], "| Li max rel err %.2e" % it["Li_rel_err"]["max"],
              "| colour max rel err %.2e (p99.9 %.2e)" % (it["direct_colour_rel_err"]["max"], it["direct_colour_rel_err"]["p999"]),
              "| image mean rel diff %.2e" % c["images"]["final"]["rel_diff_of_mean"],
              "| grads (err / scale): " + ", ".join("%s %.1e" % (k, v["max_err_over_scale"]) for k, v in c["gradients_last_pass"].items()))


if __name__ == "__main__":
    main(sys.argv[1:] or ["C1", "C2"])
