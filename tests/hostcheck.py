"""Test-only harness: the product's per-pixel kernel bodies compiled for the host (-DMR_HOST_CHECK).

The shipped library never contains or loads this flavour; it exists so that the CPU test-suite (no GPU in the build
container) can run the *same* kernel source and the *same* Python driver against the oracle.  LBVH construction has
no host flavour (it is warp-level CUDA); the harness takes the hierarchy from the oracle and packs it with the
product's pack routine.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mirres_restir_nerf_mesh_b200 import _lib, build as _build, kernels as _kernels, renderer_restir as R  # noqa: E402
from mirres_restir_nerf_mesh_b200 import slangpy_shim  # noqa: E402

_MISSING = ("mirres_abi_version", "mirres_bvh_build", "mirres_bvh_elements", "mirres_bvh_morton", "mirres_bvh_sort",
            "mirres_bvh_hierarchy_refit", "mirres_bvh_scratch_bytes", "mirres_bvh_packed_node_bytes",
            "mirres_bvh_packed_tri_bytes")



class HostKernels(_kernels.Kernels):
    def __init__(self):
        path = _build.build_hostcheck()
        lib = _lib.bind(ctypes.CDLL(path), allow_missing=_MISSING)
        super().__init__(lib=lib, require_cuda=False)

    def bvh_sizes(self, F):
        return (0, 128 + 341 * 128 + 128 * max(F - 1, 1), 64 * F)  # top table + wide nodes, packed triangles


_K = None


def kernels():
    global _K
    if _K is None:
        _K = HostKernels()
    return _K


def activate():
    slangpy_shim.set_kernels(kernels())
    return kernels()


class OracleBvhWorker(R.restirbvhWorker):
    """restirbvhWorker whose hierarchy comes from the oracle (CPU tensors); traversal records from the product."""

    def update_bvh(self, want_sorted_codes=False):
        from oracle import oracle as O
        b = O.Bvh(self.vrt.numpy(), self.v_ind.numpy())
        info = torch.from_numpy(b.info.copy())
        aabb = torch.from_numpy(b.aabb.copy())
        self.packed = slangpy_shim.packed_bvh(info, aabb, self.vrt, self.v_ind)
        self.oracle_bvh = b
        return info, aabb


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))
