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

// dst += src over bf16 maps (fp32 add, one rounding): the gradient a multi-resolution head sends into a block output
// joins the gradient the next block's dgrad has already written there (reference model.py:615-623 with sin_res=False).
__global__ void add_bf16_kernel(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ src, size_t n_vec) {
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n_vec; v += (size_t)gridDim.x * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4*>(dst)[v], b = __ldg(reinterpret_cast<const uint4*>(src) + v);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(bf16_lo(aw[e]) + bf16_lo(bw[e]), bf16_hi(aw[e]) + bf16_hi(bw[e]));
        reinterpret_cast<uint4*>(dst)[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// F.adaptive_avg_pool2d on NCHW fp32 planes (reference main_train.py:239: the frame pooled to each head's resolution):
// output (oh, ow) averages rows [floor(oh*H/Ho), ceil((oh+1)*H/Ho)) x the same in w, summed in row-major order.
__global__ void adaptive_avg_pool_kernel(const float* __restrict__ src, int planes, int H, int W, int Ho, int Wo,
                                         float* __restrict__ dst) {
    const size_t total = (size_t)planes * Ho * Wo;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ow = (int)(idx % Wo);
        const int oh = (int)((idx / Wo) % Ho);
        const size_t pl = idx / ((size_t)Wo * Ho);
        const int h0 = (int)(((long long)oh * H) / Ho), h1 = (int)((((long long)oh + 1) * H + Ho - 1) / Ho);
        const int w0 = (int)(((long long)ow * W) / Wo), w1 = (int)((((long long)ow + 1) * W + Wo - 1) / Wo);
        const float* p = src + pl * (size_t)H * W;
        float acc = 0.0f;
        for (int h = h0; h < h1; ++h)
            for (int w = w0; w < w1; ++w) acc += p[(size_t)h * W + w];
        dst[idx] = acc / (float)((h1 - h0) * (w1 - w0));
    }
}

}  // namespace onr

extern "C" {

int onr_add_bf16(void* dst_bf16, const void* src_bf16, size_t n, void* stream) {
    using namespace onr;
    ONR_REQUIRE(dst_bf16 != nullptr && src_bf16 != nullptr && n % 8 == 0, "add_bf16: element count must be a multiple of 8");
    if (n == 0) return 0;
    add_bf16_kernel<<<layout_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<__nv_bfloat16*>(dst_bf16), reinterpret_cast<const __nv_bfloat16*>(src_bf16), n / 8);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_adaptive_avg_pool(const float* src, int planes, int H, int W, int Ho, int Wo, float* dst, void* stream) {
    using namespace onr;
    ONR_REQUIRE(src != nullptr && dst != nullptr && planes >= 1 && H >= 1 && W >= 1 && Ho >= 1 && Wo >= 1,
                "adaptive_avg_pool: bad shape");
    adaptive_avg_pool_kernel<<<layout_grid((size_t)planes * Ho * Wo, 256), 256, 0, (cudaStream_t)stream>>>(
        src, planes, H, W, Ho, Wo, dst);
    ONR_LAUNCH_CHECK();
    return 0;
}

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
