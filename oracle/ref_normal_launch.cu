// TEST INFRASTRUCTURE -- never linked into, loaded by or shipped with the product library.
//
// C launcher around the reference's OWN shading-normal kernels: #includes nerf/renderutils/c_src/normal.cu from the
// reference tree where it lies (-I on the nvcc command line, oracle/Makefile target `ref`) and launches
// PrepareShadingNormalFwdKernel / BwdKernel with the parameter block the reference's torch binding builds
// (nerf/renderutils/c_src/torch_bindings.cpp:77-197).  The binding's block-shape heuristic (common.cpp:15-49, at most
// 8 x 8) only chooses how pixels map to threads; every thread handles one pixel independently, so the launcher uses
// 8 x 8 blocks throughout.  Part of oracle/_ref/libref_renderutils.so (see ref_renderutils_launch.cu).
#include <cstring>
#include <cuda_runtime.h>

#include "normal.cu"  // the reference's kernels, unmodified

namespace {

// [1, fy, fx, 3] tensor with the broadcast rule of Tensor::nhwcIndex (a dimension of extent 1 is not advanced):
// rows == 1 describes a [1,1,1,3] operand, otherwise a dense frame.  grad (optional) is the dense [1,fy,fx,3] buffer
// store_grad() writes through _dims.
Tensor make(const float *val, int rows, int fx, int fy, float *grad)
{
    Tensor t;
    std::memset(&t, 0, sizeof(t));
    t.val = const_cast<float *>(val);
    t.d_val = grad;
    t.fp16 = false;
    const bool one = rows == 1;
    t.dims[0] = 1, t.dims[1] = one ? 1 : fy, t.dims[2] = one ? 1 : fx, t.dims[3] = 3;
    t.strides[0] = one ? 3 : fy * fx * 3, t.strides[1] = one ? 3 : fx * 3, t.strides[2] = 3, t.strides[3] = 1;
    t._dims[0] = 1, t._dims[1] = fy, t._dims[2] = fx, t._dims[3] = 3;
    return t;
}

int launch(const void *kernel, PrepareShadingNormalKernelParams &p, void *stream)
{
    dim3 block(8, 8, 1);
    dim3 grid((p.gridSize.x - 1) / block.x + 1, (p.gridSize.y - 1) / block.y + 1, p.gridSize.z);
    void *args[] = {&p};
    cudaError_t e = cudaLaunchKernel(kernel, grid, block, args, 0, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace

// in[k] / rows[k], k = pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm: device pointers to fx*fy rows of
// 3 floats, or to ONE row (rows[k] == 1) that is broadcast.  out [fx*fy,3].
extern "C" int ref_prepare_shading_normal_fwd(int fx, int fy, const float *const *in, const int *rows, int two_sided,
                                              int opengl, float *out, void *stream)
{
    PrepareShadingNormalKernelParams p;
    std::memset(&p, 0, sizeof(p));
    p.gridSize = dim3(fx, fy, 1);
    p.two_sided_shading = two_sided != 0;
    p.opengl = opengl != 0;
    Tensor *t[6] = {&p.pos, &p.view_pos, &p.perturbed_nrm, &p.smooth_nrm, &p.smooth_tng, &p.geom_nrm};
    for (int k = 0; k < 6; ++k) *t[k] = make(in[k], rows[k], fx, fy, nullptr);
    p.out = make(out, fx * fy, fx, fy, nullptr);
    return launch((const void *)PrepareShadingNormalFwdKernel, p, stream);
}

// grad_out [fx*fy,3]; grads[k] [fx*fy,3] dense, all six required (the reference binding allocates all six).
extern "C" int ref_prepare_shading_normal_bwd(int fx, int fy, const float *const *in, const int *rows, int two_sided,
                                              int opengl, const float *grad_out, float *const *grads, void *stream)
{
    PrepareShadingNormalKernelParams p;
    std::memset(&p, 0, sizeof(p));
    p.gridSize = dim3(fx, fy, 1);
    p.two_sided_shading = two_sided != 0;
    p.opengl = opengl != 0;
    Tensor *t[6] = {&p.pos, &p.view_pos, &p.perturbed_nrm, &p.smooth_nrm, &p.smooth_tng, &p.geom_nrm};
    for (int k = 0; k < 6; ++k) *t[k] = make(in[k], rows[k], fx, fy, grads[k]);
    p.out = make(grad_out, fx * fy, fx, fy, nullptr);
    return launch((const void *)PrepareShadingNormalBwdKernel, p, stream);
}
