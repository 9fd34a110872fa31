"""Shared parity harness: runs the product driver (any backend) next to the oracle driver on one synthetic config and
compares every intermediate tensor.  Used by the CPU host-check tests and by the GPU tests."""
import numpy as np
import torch

from oracle import oracle as O, driver as D
from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R

INT_EXACT = ("tiles.uv", "res.M", "prev.M", "vis", "occ")


def scene(name, metallic=0.0, view=0):
    cfg = synth.CONFIGS[name]
    v, f = synth.make_mesh(cfg)
    W, Hh = cfg["W"], cfg["H"]
    ro, rd = synth.camera_rays(W, Hh, view=view)
    bvh = O.Bvh(v, f)
    hit, t, pos, nrm, prim = O.trace(bvh, ro, rd)
    g = synth.gbuffer_from_hits(ro, rd, hit, t, pos, nrm, metallic=metallic)
    env = synth.envmap(*cfg["env"])
    return dict(cfg=cfg, vert=v, tri=f, W=W, H=Hh, rays_o=ro, rays_d=rd, bvh=bvh, gbuffer=g, env=env, metallic=metallic,
                hit=hit, t=t, pos=pos, nrm=nrm, prim=prim)


def oracle_run(sc, random_offset=4242, spp=None, max_bounce=None, motion=None):
    cfg = sc["cfg"]
    snaps = []
    counters = O.new_counters()
    ref = D.run_no_denoise(sc["bvh"], sc["env"], sc["gbuffer"], spp or cfg["spp"], sc["W"], sc["H"], random_offset,
                           lambda p: synth.material(p, sc["metallic"]), max_bounce=max_bounce or cfg["max_bounce"],
                           snapshots=snaps, counters=counters, motion=motion)
    ref["snapshots"] = snaps
    ref["counters"] = counters
    return ref


def product_run(sc, worker, device, ref_prepared, random_offset=4242, spp=None, max_bounce=None, motion=None):
    """worker: a restirbvhWorker (already holding the BVH) on `device`.  Runs the sequential schedule (hooks); each
    snapshot also records how many rays the final-visibility pass of its iteration queued ("final_rays")."""
    cfg = sc["cfg"]
    spp = spp or cfg["spp"]
    mb = max_bounce or cfg["max_bounce"]
    W, Hh = sc["W"], sc["H"]
    dev = torch.device(device)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    mods = R.load_m_for_restir(W, Hh, device=dev, max_bounce=mb)
    g = {k: tt(v) for k, v in sc["gbuffer"].items()}
    mine = []

    def snap(v):
        return tuple(x.detach().cpu().numpy().copy() for x in v) if isinstance(v, tuple) else v.detach().cpu().numpy().copy()

    def hook(kind, i, d):
        if kind == "direct":
            mine.append({k: snap(v) for k, v in d.items()})
            # the boolean-ray queue size of the workspace the sequential schedule uses: the final-visibility pass was the
            # last to fill it (word MR_CTR_ANY_SIZE = 2 of the counter block at the start of the workspace)
            from mirres_restir_nerf_mesh_b200 import slangpy_shim
            ws = slangpy_shim.workspace(dev, W * Hh)
            mine[-1]["final_rays"] = int(ws[:16].cpu().view(torch.int32)[2])
        else:
            mine[-1]["bounce%d" % i[1]] = {k: snap(v) for k, v in d.items()}

    with torch.no_grad():
        outs = R.restir_di_with_pt(False, 1, 1, 1, synth.ProceduralMaterial(sc["metallic"]), worker, spp, W, Hh,
                                   *mods[:7], *mods[8:], tt(sc["env"]), g["occ_map"], g["pos_map"], g["normal_map"],
                                   g["depth_map"], g["diffuse_map"], g["roughness_specular"],
                                   tt(ref_prepared["ray_dir_map"]), None, None, None, None,
                                   None if motion is None else tt(motion), None,
                                   random_offset=random_offset, max_bounce=mb, hooks=hook)
    return dict(snapshots=mine, totals=[o.cpu().numpy() if torch.is_tensor(o) else o for o in outs], spp=spp, mb=mb)


def compare(ref, got, rtol=0.0, report=None):
    """rtol = 0 demands bit equality everywhere.  With rtol > 0 integer-like outputs stay exact and floats are
    compared to the relative tolerance.  Returns the list of mismatch descriptions (empty = parity)."""
    bad = []

    def cmp(a, b, name):
        a = np.asarray(a)
        b = np.asarray(b)
        if a.shape != b.shape:
            b = b.reshape(a.shape)
        same = (a == b) | (np.isnan(a.astype(np.float64)) & np.isnan(b.astype(np.float64)))
        if same.all():
            return
        exact = rtol == 0.0 or a.dtype.kind in "iu" or any(name.startswith(p) for p in INT_EXACT)
        if not exact:
            err = np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b.astype(np.float64)), 1e-6)
            err = np.where(same, 0.0, err)
            if np.nanmax(err) <= rtol:
                return
        idx = tuple(np.argwhere(~same)[0])
        bad.append("%s: %.3g%% differ, first at %s: got %r want %r" % (name, 100 * (1 - same.mean()), idx, a[idx], b[idx]))

    for i, (s, m) in enumerate(zip(ref["snapshots"], got["snapshots"])):
        cmp(m["light_data"], s["tiles"][0], "tiles.ld it%d" % i)
        cmp(m["light_uv"], s["tiles"][1], "tiles.uv it%d" % i)
        cmp(m["light_pdf"], s["tiles"][2], "tiles.pdf it%d" % i)
        for j, nm in enumerate(("ld", "pdf", "M", "w")):
            cmp(m["reservoirs"][j], s["res"][j], "res.%s it%d" % (nm, i))
            cmp(m["prev_reservoirs"][j], s["prev"][j], "prev.%s it%d" % (nm, i))
        cmp(m["vis"], s["vis"], "vis it%d" % i)
        cmp(m["final_Li"], s["fs_Li"], "Li it%d" % i)
        for k in ("color", "diff", "spec"):
            cmp(m[k], s[k], "%s it%d" % (k, i))
        for b in range(1, got["mb"] + 1):
            for k in ("color", "diff", "spec", "prd", "occ", "pos"):
                cmp(m["bounce%d" % b][k], s["bounce%d" % b][k], "%s.bounce%d it%d" % (k, b, i))
    m = np.float32(got["spp"])
    cmp(got["totals"][0] / m, ref["color"], "total.color")
    cmp(got["totals"][1] / m, ref["color_1"], "total.color_1")
    cmp(got["totals"][2] / m, ref["diff"], "total.diff")
    cmp(got["totals"][3] / m, ref["spec"], "total.spec")
    if report is not None:
        report.extend(bad)
    return bad
