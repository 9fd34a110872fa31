"""ctypes front-end of oracle/_ref/libref_renderutils.so -- TEST INFRASTRUCTURE ONLY.

The library holds the REFERENCE's own cross-bilateral denoiser kernels (nerf/renderutils/c_src/denoising.cu:14-130) and
shading-normal kernels (nerf/renderutils/c_src/normal.cu:95-178), compiled for sm_100a from the reference tree where
it lies, behind the C launchers oracle/ref_renderutils_launch.cu and oracle/ref_normal_launch.cu
(recipe: `make -C oracle ref`, run by __graft_entry__.build() whenever /root/reference is present).  It is the one
piece of the path whose reference source is CUDA C++ rather than Slang, hence the one piece that can be pinned against
reference code that really ran: tests/test_gpu.py::test_cross_bilateral_against_reference_kernel compares the product
(mirres_bilateral_fwd/_bwd) and the oracle's restatement (orc_bilateral_*) with it.  The built .so is git-ignored and
travels to the GPU box with the snapshot; /root/reference itself is never read at test time.

Arguments are CUDA torch tensors (fp32, contiguous); the kernels run on torch's current stream."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = "/root/reference/nerf/renderutils/c_src"
SO = os.path.join(_HERE, "_ref", "libref_renderutils.so")
_LIB = None


def build(force=False):
    """Compile the reference kernels when the reference tree is here; returns the .so path or None."""
    if not os.path.isdir(REFERENCE_SRC):
        return SO if os.path.exists(SO) else None
    srcs = [os.path.join(_HERE, "ref_renderutils_launch.cu"), os.path.join(_HERE, "ref_normal_launch.cu"),
            os.path.join(REFERENCE_SRC, "denoising.cu"), os.path.join(REFERENCE_SRC, "normal.cu")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return SO


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(SO)
        f, p, i = ctypes.c_float, ctypes.c_void_p, ctypes.c_int
        _LIB.ref_bilateral_fwd.argtypes = [i, i, f, p, p, p, p, p]
        _LIB.ref_bilateral_bwd.argtypes = [i, i, f, p, p, p, p, p, p]
        _LIB.ref_bilateral_fwd.restype = _LIB.ref_bilateral_bwd.restype = i
        _LIB.ref_prepare_shading_normal_fwd.argtypes = [i, i, p, p, i, i, p, p]
        _LIB.ref_prepare_shading_normal_bwd.argtypes = [i, i, p, p, i, i, p, p, p]
        _LIB.ref_prepare_shading_normal_fwd.restype = _LIB.ref_prepare_shading_normal_bwd.restype = i
    return _LIB


def _ptr(t):
    import torch
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def bilateral_fwd(fx, fy, sigma, col, nrm, zdz):
    """bilateral_denoiser_fwd_kernel as torch_bindings.cpp:201-222 launches it -> out [N,4]"""
    import torch
    out = torch.zeros(fx * fy, 4, device=col.device)
    rc = lib().ref_bilateral_fwd(fx, fy, sigma, _ptr(col), _ptr(nrm), _ptr(zdz), _ptr(out), _stream())
    assert rc == 0, "cudaLaunchKernel failed: %d" % rc
    return out


def bilateral_bwd(fx, fy, sigma, col, nrm, zdz, out_grad):
    """bilateral_denoiser_bwd_kernel as torch_bindings.cpp:224-246 launches it -> col_grad [N,3]"""
    import torch
    g = torch.zeros(fx * fy, 3, device=col.device)
    rc = lib().ref_bilateral_bwd(fx, fy, sigma, _ptr(col), _ptr(nrm), _ptr(zdz), _ptr(out_grad), _ptr(g), _stream())
    assert rc == 0, "cudaLaunchKernel failed: %d" % rc
    return g


def _operands(fx, fy, inputs):
    ptrs = (ctypes.c_void_p * 6)()
    rows = (ctypes.c_int * 6)()
    for k, t in enumerate(inputs):
        assert t.dim() == 2 and t.shape[1] == 3 and t.shape[0] in (1, fx * fy)
        ptrs[k] = _ptr(t).value
        rows[k] = t.shape[0]
    return ptrs, rows


def prepare_shading_normal_fwd(fx, fy, inputs, two_sided, opengl):
    """PrepareShadingNormalFwdKernel (normal.cu:95-122).  inputs = (pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng,
    geom_nrm), each [fx*fy,3] or a broadcast [1,3] -> out [fx*fy,3]"""
    import torch
    ptrs, rows = _operands(fx, fy, inputs)
    out = torch.zeros(fx * fy, 3, device=inputs[0].device)
    rc = lib().ref_prepare_shading_normal_fwd(fx, fy, ptrs, rows, int(two_sided), int(opengl), _ptr(out), _stream())
    assert rc == 0, "cudaLaunchKernel failed: %d" % rc
    return out


def prepare_shading_normal_bwd(fx, fy, inputs, two_sided, opengl, grad_out):
    """PrepareShadingNormalBwdKernel (normal.cu:124-178) -> six full-resolution [fx*fy,3] gradients"""
    import torch
    ptrs, rows = _operands(fx, fy, inputs)
    grads = [torch.zeros(fx * fy, 3, device=grad_out.device) for _ in range(6)]
    gp = (ctypes.c_void_p * 6)(*[_ptr(g).value for g in grads])
    rc = lib().ref_prepare_shading_normal_bwd(fx, fy, ptrs, rows, int(two_sided), int(opengl), _ptr(grad_out), gp,
                                              _stream())
    assert rc == 0, "cudaLaunchKernel failed: %d" % rc
    return grads
