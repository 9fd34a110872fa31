"""ctypes binding of libmirres_b200.so (C ABI: include/mirres_b200.h).

There is no fallback: if the CUDA library is missing or fails to load, importing the kernels raises.  The
signature table below is the single place where Python meets the ABI; `check_exports` (used by the CPU test-suite)
verifies that every symbol the header declares is exported.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmirres_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "mirres_b200.h")

_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "u": ctypes.c_uint, "f": ctypes.c_float, "z": ctypes.c_size_t}

# name -> argument kinds (p pointer, i int, u unsigned, f float, z size_t); every function returns int unless noted
SIGNATURES = {
    "mirres_abi_version": "",
    "mirres_set_tuning": "ii",
    "mirres_get_tuning": "i",
    "mirres_bvh_build": "pipippppppzp",
    "mirres_bvh_elements": "ppippp",
    "mirres_bvh_morton": "piffffffpp",
    "mirres_bvh_sort": "pipzp",
    "mirres_bvh_hierarchy_refit": "ppipppzp",
    "mirres_bvh_pack": "ppppippp",
    "mirres_trace_closest": "ppppippppppp",
    "mirres_trace_any": "ppppippp",
    "mirres_env_build_distribution": "piippppp" + "p",
    "mirres_env_weights": "piipp",
    "mirres_env_distribution2d": "iippp",
    "mirres_neighbor_offsets": "ipp",
    "mirres_light_tiles": "piippppuiipppppp",
    "mirres_workspace_prepare": "pipzp",
    "mirres_initial_resampling": "ppp" + "pppp" + "pii" + "iiu" + "pppp" + "pp" + "ppp" + "iiiii" + "pz" + "p",
    "mirres_temporal_resampling": "pppp" + "pppp" + "pii" + "iiu" + "pppp" + "pppp" + "p" + "i" + "pz" + "p",
    "mirres_spatial_resampling": "ppp" + "pppp" + "pppp" + "p" + "pii" + "iiu" + "pppp" + "iif" + "pz" + "p",
    "mirres_final_visibility": "pppiipp" + "pz" + "p",
    "mirres_set_visibility_tags": "pp",
    "mirres_eval_final_fwd": "pppp" + "pii" + "ii" + "ppp" + "p" + "p",
    "mirres_eval_final_bwd": "pppp" + "ii" + "ii" + "ppp" + "p",
    "mirres_final_shading_fwd": "ppp" + "pii" + "ii" + "ppppp" + "ppp" + "p",
    "mirres_final_shading_bwd": "ppp" + "ii" + "ppppp" + "ppp" + "pppp" + "p",
    "mirres_bounce_first": "pp" + "uui" + "ii" + "pppp" + "p" + "pp" + "pppp" + "pz" + "p",
    "mirres_bounce_shade": "pp" + "uui" + "ii" + "pii" + "pppp" + "pppp" + "p" + "pp" + "ppp" + "pppp" + "pz" + "p",
    "mirres_eaw_fwd": "fffiif" + "ppppp" + "p",
    "mirres_eaw_bwd": "fffiif" + "ppppp" + "ppppp" + "p",
    "mirres_normal_ao": "iipppp",
    "mirres_bilateral_fwd": "iifppppp",
    "mirres_bilateral_bwd": "iifppppp",
    "mirres_eaw_fwd_multi": "fffiif" + "ppp" + "i" + "ppp" + "p",
    "mirres_eaw_bwd_multi": "fffiif" + "ppp" + "i" + "ppppppp" + "p",
    "mirres_gbuffer_primary": "ppppi" + "pp" + "pppppp" + "p" + "pz" + "p",
    "mirres_prepare_maps": "ippppppppp" + "p",
    "mirres_interpolate_bwd": "piipppipp",
    "mirres_vertex_normals_fwd": "pipippp",
    "mirres_vertex_normals_bwd": "pipipppp",
    "mirres_shading_normal_fwd": "i" + "pi" * 6 + "ii" + "p" + "p",
    "mirres_shading_normal_bwd": "i" + "pi" * 6 + "ii" + "p" + "pppppp" + "p",
    "mirres_material_procedural": "ippifpppp",
    "mirres_sum_images": "iipfipp",
    "mirres_composite_fwd": "ipppppppp",
    "mirres_composite_bwd": "ipppppp" + "p" + "pppp" + "p",
    "mirres_final_shading_bwd_multi": "ippp" + "ii" + "ppppp" + "ppp" + "fippp" + "ip" + "p",
}
SIZE_FUNCS = ("mirres_bvh_scratch_bytes", "mirres_bvh_packed_node_bytes", "mirres_bvh_packed_tri_bytes",
              "mirres_workspace_bytes")


def bind(lib, allow_missing=()):
    for name, sig in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in allow_missing:
                continue
            raise
        fn.argtypes = [_T[c] for c in sig]
        fn.restype = ctypes.c_int
    for name in SIZE_FUNCS:
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in allow_missing:
                continue
            raise
        fn.argtypes = [ctypes.c_int]
        fn.restype = ctypes.c_size_t
    return lib


_LIB = None


def load():
    """Load the CUDA library.  Fails loudly: the product has no CPU path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmirres_b200.so is not built (%s). Run `python -m mirres_restir_nerf_mesh_b200.build`; "
                "mirres-b200 has no CPU fallback." % LIB_PATH)
        _LIB = bind(ctypes.CDLL(LIB_PATH))
        if _LIB.mirres_abi_version() != 1:
            raise RuntimeError("libmirres_b200.so ABI version mismatch")
    return _LIB


def declared_symbols(header_path=HEADER_PATH):
    """Function names declared in include/mirres_b200.h."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mirres_[a-z0-9_]+)\s*\(", text)))


def check_exports(lib_path=LIB_PATH):
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    return missing
