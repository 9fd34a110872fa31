"""Bounds the risk that "parity unpinned" leaves open: the oracle built with
reference-like numerics (FMA contraction on, C-library sin / cos / acos / atan2 / exp, powf -- `make -C oracle fast`,
include/mirres_fpmath.h) against the contract flavour the CUDA kernels are pinned to, on identical inputs.

What is asserted (tools/numerics_sensitivity.py computes it, profiles/numerics_sensitivity.json holds C1 / C2):
  * integer decisions are almost never affected: LBVH identical, no primary hit flips on these scenes, reservoir
    selections / visibility flags flip on < 0.5 % of the foreground pixels;
  * where the decisions agree, radiance moves by < 1e-4 relative (the north-star forward tolerance) for the reference's
    default material (metallic 0), 99.9 % of the pixels by < 4e-5;
  * with a metallic material the GGX lobe of near-mirror pixels amplifies rounding (cancellation in (a^2 - 1) cos^2 + 1):
    99.9 % of the pixels stay within 5e-3, which is the honest size of the gap a 1e-4 claim against the real binary would
    have to survive there;
  * the gradients that leave the path (float64 backward oracle on each flavour's own forward state, same upstream weights)
    move by < 1e-3 of their scale -- the north-star gradient tolerance -- for the default material, 99.9 % of the pixels
    by < 1e-4; with the metallic material the same near-mirror pixels reach a few 1e-3.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("name,metallic", [("T1", 0.0), ("T2", 0.0), ("T2", 0.4)])
def test_reference_like_numerics_stay_within_tolerance(oracle, name, metallic):
    import numerics_sensitivity as NS
    c = NS.compare(name, metallic=metallic)
    assert c["lbvh"]["topology_equal"] and c["lbvh"]["node_boxes_equal"]
    assert c["primary_rays"]["hit_flag_flips"] == 0 and c["primary_rays"]["t_max_rel_err"] < 1e-4
    fg = c["foreground_pixels"]
    for it in c["iterations"]:
        assert it["reservoir_selection_flips"] + it["visibility_flips"] < 0.005 * fg
        assert it["light_tile_texel_flips"] < 0.001 * it["light_tile_samples"]
        assert it["pixels_with_equal_decisions"] > 0.99 * fg
        if metallic == 0.0:
            assert it["Li_rel_err"]["max"] < 1e-4 and it["direct_colour_rel_err"]["max"] < 1e-4
            assert it["direct_colour_rel_err"]["p999"] < 4e-5
        else:
            assert it["direct_colour_rel_err"]["p999"] < 5e-3
    # the spp average is the same estimator either way: its mean moves by rounding, not by bias
    assert c["images"]["final"]["rel_diff_of_mean"] < 1e-4
    g = c["gradients_last_pass"]
    for k in ("normal", "kd", "rough_metal"):
        assert g[k]["scale"] > 0
        assert g[k]["p999_err_over_scale"] < (1e-4 if metallic == 0.0 else 2e-3), (k, g[k])
        assert g[k]["max_err_over_scale"] < (1e-3 if metallic == 0.0 else 2e-2), (k, g[k])
    assert g["env"]["max_err_over_scale"] < 2e-3 and g["env"]["rel_diff_of_sum"] < 1e-4, g["env"]
