// layout.cu — layout converters at the module boundary: NeRVBlock.forward (reference model.py:518-567) takes and
// returns NCHW fp32 tensors, the kernels work on NHWC bf16 with channels padded to a multiple of 32.
#include "onr_common.cuh"
#include "act.cuh"

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

// Activation over an NHWC bf16 map for the activations that are not fused into the convolution epilogue (act.cuh):
// z[pixels][Cp] (the ONR_CONV_FPROP_Z output) -> y = act(z) IN PLACE and, when d != NULL, d = act'(z) (the map the
// dgrad epilogue and the head backward multiply by).  Channels >= C are padding: y = d = 0 there (softplus(0) != 0).
// HBM-bound: 2 B read + 2 (+2) B written per element, 16-byte accesses (Cp is a multiple of 32, so a vector never
// straddles a pixel).
__global__ void act_map_kernel(__nv_bfloat16* __restrict__ zy, __nv_bfloat16* __restrict__ d, size_t n_vec, int C,
                               int Cp, int act) {
    const int vec_per_px = Cp / 8;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n_vec; v += (size_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(v % vec_per_px) * 8;
        uint4 raw = reinterpret_cast<const uint4*>(zy)[v];
        uint32_t in[4] = {raw.x, raw.y, raw.z, raw.w}, oy[4], od[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float y0, d0, y1, d1;
            act_value_grad(bf16_lo(in[e]), act, &y0, &d0);
            act_value_grad(bf16_hi(in[e]), act, &y1, &d1);
            if (c0 + 2 * e >= C) y0 = d0 = 0.0f;
            if (c0 + 2 * e + 1 >= C) y1 = d1 = 0.0f;
            oy[e] = pack_bf16x2(y0, y1);
            od[e] = pack_bf16x2(d0, d1);
        }
        reinterpret_cast<uint4*>(zy)[v] = make_uint4(oy[0], oy[1], oy[2], oy[3]);
        if (d != nullptr) reinterpret_cast<uint4*>(d)[v] = make_uint4(od[0], od[1], od[2], od[3]);
    }
}

}  // namespace onr

extern "C" {

int onr_act_map(void* zy_bf16, void* d_bf16, size_t pixels, int C, int Cp, int act, void* stream) {
    using namespace onr;
    ONR_REQUIRE(zy_bf16 != nullptr && Cp % 32 == 0 && C >= 1 && C <= Cp, "act_map: bad shape (C %d Cp %d)", C, Cp);
    ONR_REQUIRE(act >= 0 && act < kActCount, "act_map: unknown activation code %d", act);
    if (pixels == 0) return 0;
    const size_t n_vec = pixels * (size_t)(Cp / 8);
    act_map_kernel<<<layout_grid(n_vec, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<__nv_bfloat16*>(zy_bf16), reinterpret_cast<__nv_bfloat16*>(d_bf16), n_vec, C, Cp, act);
    ONR_LAUNCH_CHECK();
    return 0;
}

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
