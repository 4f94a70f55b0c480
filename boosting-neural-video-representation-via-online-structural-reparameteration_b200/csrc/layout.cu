// layout.cu — layout converters at the module boundary: NeRVBlock.forward (reference model.py:518-567) takes and
// returns NCHW fp32 tensors, the kernels work on NHWC bf16 with channels padded to a multiple of 32.
#include "onr_common.cuh"

namespace onr {

static inline int layout_grid(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = (size_t)num_sms() * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ src, int B, int C, int H, int W, int Cp,
                                         __nv_bfloat16* __restrict__ dst) {
    const size_t total = (size_t)B * H * W * Cp;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Cp);
        size_t pix = idx / Cp;
        const int w = (int)(pix % W);
        pix /= W;
        const int h = (int)(pix % H);
        const int b = (int)(pix / H);
        const float v = c < C ? src[(((size_t)b * C + c) * H + h) * W + w] : 0.0f;
        dst[idx] = __float2bfloat16(v);
    }
}

__global__ void nhwc_bf16_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int B, int C, int H, int W,
                                         int Cp, float* __restrict__ dst) {
    const size_t total = (size_t)B * C * H * W;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(idx % W);
        size_t r = idx / W;
        const int h = (int)(r % H);
        r /= H;
        const int c = (int)(r % C);
        const int b = (int)(r / C);
        dst[idx] = __bfloat162float(src[(((size_t)b * H + h) * W + w) * Cp + c]);
    }
}


}  // namespace onr

extern "C" {

int onr_nchw_to_nhwc_bf16(const float* src, int B, int C, int H, int W, int Cp, void* dst, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp >= C && Cp % 32 == 0, "Cp must be a multiple of 32 >= C");
    const size_t total = (size_t)B * H * W * Cp;
    nchw_to_nhwc_bf16_kernel<<<layout_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        src, B, C, H, W, Cp, reinterpret_cast<__nv_bfloat16*>(dst));
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_nhwc_bf16_to_nchw(const void* src, int B, int C, int H, int W, int Cp, float* dst, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp >= C, "Cp must be >= C");
    const size_t total = (size_t)B * C * H * W;
    nhwc_bf16_to_nchw_kernel<<<layout_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(src), B, C, H, W, Cp, dst);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
