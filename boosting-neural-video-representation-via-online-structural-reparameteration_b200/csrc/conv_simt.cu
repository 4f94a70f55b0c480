// conv_simt.cu — straightforward CUDA-core versions of the three convolution passes, on exactly the
// layouts the tcgen05 kernels use.  They exist so that the tensor-core kernels can be cross-checked on
// the device at full problem sizes (tests/ and csrc/selftest.cu); the product path never calls them.
// Also holds the NCHW fp32 <-> NHWC bf16 converters used at the NeRVBlock module boundary.
#include "onr_common.cuh"
#include "selftest_kernels.h"

namespace onr {

// A-operand element of the un-shuffled view: tensor [B][H*s][W*s][Cp], k = i*(s*Cp) + jc.
__device__ __forceinline__ float a_view(const __nv_bfloat16* a, int b, int h, int w, int k, int H, int W,
                                        int Cp, int s) {
    if (h < 0 || h >= H || w < 0 || w >= W) return 0.0f;
    const int jcn = s * Cp;
    const int i = k / jcn, jc = k - i * jcn;
    const size_t idx = ((size_t)(b * H * s + h * s + i) * (W * s) + (size_t)w * s) * Cp + jc;
    return __bfloat162float(a[idx]);
}

__global__ void simt_conv_kernel(onr_conv_desc d) {
    const int k_tap = d.a_s * d.a_s * d.a_cp;
    const size_t total = (size_t)d.B * d.H * d.W * d.n_total;
    const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(d.a);
    const __nv_bfloat16* Wt = reinterpret_cast<const __nv_bfloat16*>(d.w);
    const int sign = d.kind == ONR_CONV_DGRAD ? -1 : 1;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(idx % d.n_total);
        size_t pix = idx / d.n_total;
        const int w = (int)(pix % d.W);
        pix /= d.W;
        const int h = (int)(pix % d.H);
        const int b = (int)(pix / d.H);
        float acc = 0.0f;
        for (int tap = 0; tap < 9; ++tap) {
            const int hh = h + (tap / 3 - 1) * sign, ww = w + (tap % 3 - 1) * sign;
            if (hh < 0 || hh >= d.H || ww < 0 || ww >= d.W) continue;
            const __nv_bfloat16* wrow = Wt + ((size_t)tap * d.n_rows + n) * k_tap;
            for (int k = 0; k < k_tap; ++k)
                acc = fmaf(a_view(A, b, hh, ww, k, d.H, d.W, d.a_cp, d.a_s), __bfloat162float(wrow[k]), acc);
        }
        const int out_jc = d.out_s * d.out_cp;
        const int oi = n / out_jc, ojc = n - oi * out_jc;
        const size_t o = ((size_t)(b * d.H * d.out_s + h * d.out_s + oi) * (d.W * d.out_s) +
                          (size_t)w * d.out_s) * d.out_cp + ojc;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
        if (d.kind == ONR_CONV_DGRAD) {
            const __nv_bfloat16* dm = reinterpret_cast<const __nv_bfloat16*>(d.dmul);
            const float m = __bfloat162float(dm[((size_t)(b * d.H + h) * d.W + w) * d.n_total + n]);
            out[o] = __float2bfloat16(acc * m);
        } else {
            const float z = acc + d.bias_p[n];
            const float sg = 1.0f / (1.0f + expf(-z));
            const float y = z * sg;
            out[o] = __float2bfloat16(y);
            if (d.kind == ONR_CONV_FPROP_TRAIN)
                reinterpret_cast<__nv_bfloat16*>(d.out_d)[o] = __float2bfloat16(sg + y * (1.0f - sg));
        }
    }
}

// one thread per (n, tap, ci); serial loop over all pixels.
__global__ void simt_wgrad_kernel(onr_wgrad_desc d) {
    const int n_pre = d.s * d.s * d.dz_cp;
    const size_t total = (size_t)n_pre * 9 * d.x_cp;
    const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(d.x);
    const __nv_bfloat16* DZ = reinterpret_cast<const __nv_bfloat16*>(d.dz);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int ci = (int)(idx % d.x_cp);
        const int tap = (int)((idx / d.x_cp) % 9);
        const int n = (int)(idx / ((size_t)d.x_cp * 9));
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        float acc = 0.0f;
        for (int b = 0; b < d.B; ++b)
            for (int h = 0; h < d.H; ++h) {
                const int hh = h + dh;
                if (hh < 0 || hh >= d.H) continue;
                for (int w = 0; w < d.W; ++w) {
                    const int ww = w + dw;
                    if (ww < 0 || ww >= d.W) continue;
                    const float g = a_view(DZ, b, h, w, n, d.H, d.W, d.dz_cp, d.s);
                    const float x = __bfloat162float(X[((size_t)(b * d.H + hh) * d.W + ww) * d.x_cp + ci]);
                    acc = fmaf(g, x, acc);
                }
            }
        d.dKp[idx] += acc;
        if (tap == 4 && ci == 0) {
            float sb = 0.0f;
            for (int b = 0; b < d.B; ++b)
                for (int h = 0; h < d.H; ++h)
                    for (int w = 0; w < d.W; ++w) sb += a_view(DZ, b, h, w, n, d.H, d.W, d.dz_cp, d.s);
            d.dbias_p[n] += sb;
        }
    }
}

static inline int grid_for(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = (size_t)num_sms() * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace onr

extern "C" {

int onr_simt_conv(const onr_conv_desc* d, void* stream) {
    using namespace onr;
    ONR_REQUIRE(d != nullptr, "null desc");
    const size_t total = (size_t)d->B * d->H * d->W * d->n_total;
    simt_conv_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*d);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_simt_wgrad(const onr_wgrad_desc* d, void* stream) {
    using namespace onr;
    ONR_REQUIRE(d != nullptr, "null desc");
    const size_t total = (size_t)d->s * d->s * d->dz_cp * 9 * d->x_cp;
    simt_wgrad_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(*d);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
