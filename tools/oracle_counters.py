"""Measures V_n / V_t of SURVEY.md 8d with the oracle: BVH nodes popped and triangles tested per kernel launch on a
BASELINE config, under (a) the contract schedule (closest-hit rays: reference DFS; boolean rays: same DFS, exit at the
first hit) and (b) the reference schedule (no early-out), for information.  Writes profiles/oracle_counters_<cfg>.json,
which bench.py reads for the roofline's algorithmic bytes.

    python tools/oracle_counters.py C2
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O, driver as D  # noqa: E402
from mirres_restir_nerf_mesh_b200 import synth  # noqa: E402


def main(name):
    cfg = synth.CONFIGS[name]
    v, f = synth.make_mesh(cfg)
    b = O.Bvh(v, f)
    W, H = cfg["W"], cfg["H"]
    ro, rd = synth.camera_rays(W, H)
    prim_ctr = O.new_counters()
    hit, t, pos, nrm, prim = O.trace(b, ro, rd, prim_ctr)
    g = synth.gbuffer_from_hits(ro, rd, hit, t, pos, nrm)
    env = synth.envmap(*cfg["env"])
    spp = min(cfg["spp"], 4)
    per = {}
    t0 = time.time()
    D.run_no_denoise(b, env, g, spp, W, H, 1234, lambda p: synth.material(p), max_bounce=cfg["max_bounce"], counters=per)
    dt = time.time() - t0
    launches = {"initial_resampling": spp, "spatial_resampling": spp, "final_visibility": spp, "bounce_first": spp,
                "bounce_shade": spp * cfg["max_bounce"]}
    out = {"config": name, "frame": [W, H], "spp_measured": spp, "triangles": int(f.shape[0]), "hit_fraction": float(hit.mean()),
           "oracle_seconds": dt, "max_stack_depth": int(max(c[6] for c in per.values())),
           "primary_rays": {"nodes_per_ray": prim_ctr[0] / len(ro), "tris_per_ray": prim_ctr[1] / len(ro)}}
    tot_n = tot_t = tot_rays = dead_n = dead_t = dead_rays = 0
    for k, c in per.items():
        L = launches[k]
        # contract: shadow rays counted with first-hit exit (c[2], c[3]); closest rays under the reference DFS (c[4], c[5])
        nodes, tris = int(c[2] + c[4]), int(c[3] + c[5])
        out[k] = {"launches": L, "shadow_rays_per_launch": c[7] / L, "closest_rays_per_launch": c[8] / L,
                  "nodes_per_launch": nodes / L, "tris_per_launch": tris / L,
                  "reference_schedule_nodes_per_launch": int(c[0] + c[4]) / L,
                  "reference_schedule_tris_per_launch": int(c[1] + c[5]) / L}
        # rays of the reference whose result cannot reach the output (oracle/orc_kernels.cpp, counters 9..11); the product
        # does not cast them, so the figures bench.py charges IT with leave them out
        out[k]["dead_product_rays_per_launch"] = c[11] / L
        out[k]["cast_nodes_per_launch"] = (nodes - int(c[9])) / L
        out[k]["cast_tris_per_launch"] = (tris - int(c[10])) / L
        tot_n += nodes
        tot_t += tris
        tot_rays += int(c[7] + c[8])
        dead_n += int(c[9])
        dead_t += int(c[10])
        dead_rays += int(c[11])
    samples = W * H * spp
    out["per_sample"] = {"rays": tot_rays / samples, "V_n": tot_n / samples, "V_t": tot_t / samples,
                         "B_alg_traversal_bytes": (36 * tot_n + 48 * tot_t) / samples,
                         "rays_cast": (tot_rays - dead_rays) / samples,
                         "B_alg_traversal_bytes_cast": (36 * (tot_n - dead_n) + 48 * (tot_t - dead_t)) / samples}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "profiles", "oracle_counters_%s.json" % name)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out["per_sample"]), "->", path, "(%.1f s)" % dt)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C2")
