// TEST INFRASTRUCTURE -- never linked into, loaded by or shipped with the product library.
//
// C launcher around the reference's OWN cross-bilateral denoiser kernels: this file holds no kernel code, it
// #includes nerf/renderutils/c_src/denoising.cu from the reference tree where it lies (-I on the nvcc command line, see
// oracle/Makefile target `ref`) and launches bilateral_denoiser_fwd_kernel / _bwd_kernel exactly as the reference's
// torch binding does (nerf/renderutils/c_src/torch_bindings.cpp:201-246: 8 x 8 x 1 blocks, one thread per pixel,
// cudaLaunchKernel with a single BilateralDenoiserParams argument).  The output, oracle/_ref/libref_renderutils.so,
// pins the oracle's restatement (oracle/orc_kernels.cpp, orc_bilateral_*) and the product's mirres_bilateral_fwd/_bwd
// against reference code that really ran -- the one kernel pair of the path whose reference source is CUDA C++ rather
// than Slang.
//
// The reference's accessor class (c_src/accessor.h:205-300) has host constructors only when it is NOT compiled by nvcc
// (its binding file is plain C++ built against torch headers).  Under nvcc it is an aggregate-like standard-layout
// class {T *data_; int32 sizes_[N]; int32 strides_[N];}, so the launcher fills a layout-identical POD and copies it in.
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "denoising.cu"  // the reference's kernels, unmodified (found through -I<reference>/nerf/renderutils/c_src)

namespace {

struct Acc4 {
    float *data;
    int32_t sizes[4];
    int32_t strides[4];
};
static_assert(sizeof(Acc4) == sizeof(PackedTensorAccessor32<float, 4>), "accessor layout changed");

// contiguous [1, fy, fx, c] view, as ops.py:190-211 hands the tensors over
void fill(PackedTensorAccessor32<float, 4> &dst, const float *p, int fy, int fx, int c) {
    Acc4 a;
    a.data = const_cast<float *>(p);
    a.sizes[0] = 1, a.sizes[1] = fy, a.sizes[2] = fx, a.sizes[3] = c;
    a.strides[0] = fy * fx * c, a.strides[1] = fx * c, a.strides[2] = c, a.strides[3] = 1;
    std::memcpy(&dst, &a, sizeof(a));
}

int launch(const void *kernel, BilateralDenoiserParams &params, int fx, int fy, void *stream) {
    dim3 block(8, 8, 1);
    dim3 grid((fx - 1) / block.x + 1, (fy - 1) / block.y + 1, 1);
    void *args[] = {&params};
    cudaError_t e = cudaLaunchKernel(kernel, grid, block, args, 0, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace

// col [fy*fx,3], nrm [fy*fx,3], zdz [fy*fx,2], out [fy*fx,4]: device pointers, fp32, contiguous.
extern "C" int ref_bilateral_fwd(int fx, int fy, float sigma, const float *col, const float *nrm, const float *zdz,
                                 float *out, void *stream) {
    if (fx <= 0 || fy <= 0) return 0;
    BilateralDenoiserParams params;
    std::memset(&params, 0, sizeof(params));
    fill(params.col, col, fy, fx, 3);
    fill(params.nrm, nrm, fy, fx, 3);
    fill(params.zdz, zdz, fy, fx, 2);
    fill(params.out, out, fy, fx, 4);
    params.sigma = sigma;
    return launch((const void *)bilateral_denoiser_fwd_kernel, params, fx, fy, stream);
}

// out_grad [fy*fx,4] (the kernel reads the first three channels), col_grad [fy*fx,3].  The backward kernel takes the
// frame size from params.col and fetches its taps without using them (denoising.cu:76-129), so col is passed as the
// reference binding passes it.
extern "C" int ref_bilateral_bwd(int fx, int fy, float sigma, const float *col, const float *nrm, const float *zdz,
                                 const float *out_grad, float *col_grad, void *stream) {
    if (fx <= 0 || fy <= 0) return 0;
    BilateralDenoiserParams params;
    std::memset(&params, 0, sizeof(params));
    fill(params.col, col, fy, fx, 3);
    fill(params.nrm, nrm, fy, fx, 3);
    fill(params.zdz, zdz, fy, fx, 2);
    fill(params.out_grad, out_grad, fy, fx, 4);
    fill(params.col_grad, col_grad, fy, fx, 3);
    params.sigma = sigma;
    return launch((const void *)bilateral_denoiser_bwd_kernel, params, fx, fy, stream);
}
