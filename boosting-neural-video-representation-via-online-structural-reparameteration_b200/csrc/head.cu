// head.cu — RGB head: 1x1 conv C->3 + bias + (tanh+1)/2 (or sigmoid), forward and backward.
// Reference: model.py:601 (head_layer), model.py:620-623 (forward).  Bandwidth-bound: the forward reads
// the last block's NHWC bf16 activation once; the backward reads y and SiLU'(z) once and writes dz once.
// Thread layout: 8-channel (16-byte) chunks, (Cp/8) consecutive threads per pixel, so a warp touches one
// contiguous span of the activation.
#include "onr_common.cuh"

namespace onr {

constexpr int kHeadMaxC = 128;
constexpr int kHeadMaxChunks = kHeadMaxC / 8;

// Every kernel here gives a thread one fixed 8-channel piece (ch = tid % chunks) of a pixel that advances by a
// whole number of block-sized pixel groups, so the head weights of that piece live in registers, consecutive
// threads read consecutive 16-byte pieces, and the inner loops carry no division (all indices are 32-bit: the host
// wrappers reject npix * chunks >= 2^31).
__device__ __forceinline__ void head_load_w(const float* __restrict__ Wh, int C, int ch, float w[3][8]) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = ch * 8 + e;
            w[k][e] = c < C ? __ldg(Wh + k * C + c) : 0.0f;
        }
}

// kOne: a single image (batch 1, the training case) needs no division, which also keeps the loops free of branches
// so that every load of an unrolled pass is issued before the first use.
template <bool kOne>
__device__ __forceinline__ uint32_t head_img_offset(uint32_t pix, uint32_t HW) {
    if (kOne) return pix;
    const uint32_t b = pix / HW;
    return b * 3u * HW + (pix - b * HW);
}

__device__ __forceinline__ float head_act(float a, int use_sigmoid) {
    return use_sigmoid ? 1.0f / (1.0f + expf(-a)) : (tanhf(a) + 1.0f) * 0.5f;
}

// blockDim.x = ppb * chunks.  Each pass covers kFwdU groups of ppb pixels: every thread reduces its 8 channels of
// one pixel per group against the three weight rows, the `chunks` partials of a pixel meet in shared memory
// (double-buffered: one barrier per pass), and the first 3*ppb threads finish the pixels and write the planar fp32
// image with unit stride.  The next pass's activations are already in flight while the current one is reduced.
constexpr int kFwdU = 2;
template <bool kOne>
__global__ void __launch_bounds__(512)
head_fwd_kernel(const __nv_bfloat16* __restrict__ y, uint32_t npix, uint32_t HW, int C, int Cp,
                const float* __restrict__ Wh, const float* __restrict__ bh, int use_sigmoid,
                float* __restrict__ img) {
    __shared__ float part[2][kFwdU][3][512 + 512 / 4];   // [buf][u][k][pl * (chunks + 1) + ch]
    const int chunks = Cp / 8;
    const int ppb = blockDim.x / chunks;
    const int pl = threadIdx.x / chunks, ch = threadIdx.x - pl * chunks;
    float w[3][8];
    head_load_w(Wh, C, ch, w);
    const int fk = threadIdx.x / ppb, fp = threadIdx.x - fk * ppb;   // finishing role: channel fk of pixel fp
    const float fb = fk < 3 ? __ldg(bh + fk) : 0.0f;
    const uint32_t span = kFwdU * ppb, step = gridDim.x * span;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    auto load = [&](uint32_t base, uint4 v[kFwdU]) {
#pragma unroll
        for (int u = 0; u < kFwdU; ++u) {
            const uint32_t pix = base + u * ppb + pl;
            v[u] = (base < npix && pix < npix) ? __ldg(reinterpret_cast<const uint4*>(y) + pix * chunks + ch) : zero;
        }
    };
    uint4 nxt[kFwdU];
    load(blockIdx.x * span, nxt);
    int buf = 0;
    for (uint32_t base = blockIdx.x * span; base < npix; base += step, buf ^= 1) {
        uint4 cur[kFwdU];
#pragma unroll
        for (int u = 0; u < kFwdU; ++u) cur[u] = nxt[u];
        load(base + step, nxt);
#pragma unroll
        for (int u = 0; u < kFwdU; ++u) {
            const uint32_t q[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
            float a[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = bf16_lo(q[e]), hi = bf16_hi(q[e]);
#pragma unroll
                for (int k = 0; k < 3; ++k) a[k] = fmaf(hi, w[k][2 * e + 1], fmaf(lo, w[k][2 * e], a[k]));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) part[buf][u][k][pl * (chunks + 1) + ch] = a[k];
        }
        __syncthreads();
        if (fk < 3) {
#pragma unroll
            for (int u = 0; u < kFwdU; ++u) {
                const uint32_t pix = base + u * ppb + fp;
                if (pix >= npix) continue;
                float s = fb;
                const float* pp = &part[buf][u][fk][fp * (chunks + 1)];
                for (int c = 0; c < chunks; ++c) s += pp[c];
                img[head_img_offset<kOne>(pix, HW) + fk * HW] = head_act(s, use_sigmoid);
            }
        }
    }
}

// Backward, split in two independent streaming kernels so that only the part the rest of the backward depends on
// (dz) sits on the critical path; the weight/bias gradient reduction can overlap the block kernels.
//   g_pre[k] = gimg[k] * d(act)/d(pre)  with  (tanh+1)/2 -> 2 o (1-o),  sigmoid -> o (1-o)
//   dz[px, c]  = (sum_k g_pre[k] Wh[k, c]) * SiLU'(z)[px, c]
//   gWh[k, c] += sum_px g_pre[k] y[px, c] ;  gbh[k] += sum_px g_pre[k]
template <bool kOne>
__device__ __forceinline__ void head_gpre(const float* __restrict__ gimg, const float* __restrict__ img,
                                          uint32_t pix, uint32_t HW, int use_sigmoid, float gp[3]) {
    const uint32_t io = head_img_offset<kOne>(pix, HW);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float o = __ldg(img + io + k * HW);
        const float g = __ldg(gimg + io + k * HW);
        gp[k] = use_sigmoid ? g * o * (1.0f - o) : g * 2.0f * o * (1.0f - o);
    }
}

constexpr int kDzUnroll = 4;   // independent 16-byte loads in flight per thread
constexpr int kGwUnroll = 4;

template <bool kOne>
__global__ void __launch_bounds__(256, 4)
head_bwd_dz_kernel(const float* __restrict__ gimg, const float* __restrict__ img,
                   const __nv_bfloat16* __restrict__ dsilu, uint32_t npix, uint32_t HW, int C, int Cp,
                   const float* __restrict__ Wh, int use_sigmoid, __nv_bfloat16* __restrict__ dz) {
    const int chunks = Cp / 8;
    const int ppb = blockDim.x / chunks;
    const int pl = threadIdx.x / chunks, ch = threadIdx.x - pl * chunks;
    float w[3][8];
    head_load_w(Wh, C, ch, w);
    const uint32_t step = gridDim.x * ppb;
    for (uint32_t pix0 = blockIdx.x * ppb + pl; pix0 < npix; pix0 += kDzUnroll * step) {
        uint4 dv[kDzUnroll];
        float gp[kDzUnroll][3];
        // out-of-range members of the pass re-read pixel pix0 (loads stay unconditional), their stores are skipped
#pragma unroll
        for (int u = 0; u < kDzUnroll; ++u) {
            const uint32_t pix = pix0 + u * step < npix ? pix0 + u * step : pix0;
            dv[u] = __ldg(reinterpret_cast<const uint4*>(dsilu) + pix * chunks + ch);
            head_gpre<kOne>(gimg, img, pix, HW, use_sigmoid, gp[u]);
        }
#pragma unroll
        for (int u = 0; u < kDzUnroll; ++u) {
            const uint32_t pix = pix0 + u * step;
            const uint32_t du[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
            uint32_t out[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    d0 = fmaf(gp[u][k], w[k][2 * e], d0);
                    d1 = fmaf(gp[u][k], w[k][2 * e + 1], d1);
                }
                out[e] = pack_bf16x2(d0 * bf16_lo(du[e]), d1 * bf16_hi(du[e]));
            }
            if (pix < npix)
                reinterpret_cast<uint4*>(dz)[pix * chunks + ch] = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
}

// blockDim.x = chunks * lanes; thread (lane, ch) walks pixels lane, lane + lanes*gridDim, ... for its 8 channels.
template <bool kOne>
__global__ void __launch_bounds__(384)
head_bwd_gw_kernel(const float* __restrict__ gimg, const float* __restrict__ img,
                                   const __nv_bfloat16* __restrict__ y, uint32_t npix, uint32_t HW, int C, int Cp,
                                   int use_sigmoid, float* __restrict__ gWh, float* __restrict__ gbh) {
    __shared__ float sg[3 * kHeadMaxC + 3];
    for (int i = threadIdx.x; i < 3 * Cp + 3; i += blockDim.x) sg[i] = 0.0f;
    __syncthreads();
    const int chunks = Cp / 8;
    const int lanes = blockDim.x / chunks;
    const int lane = threadIdx.x / chunks, ch = threadIdx.x - lane * chunks;
    float gw[3][8];
    float gb[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) gw[k][e] = 0.0f;
    {
        const uint32_t stride = gridDim.x * lanes;
        for (uint32_t pix0 = blockIdx.x * lanes + lane; pix0 < npix; pix0 += kGwUnroll * stride) {
            uint4 yv[kGwUnroll];
            float gp[kGwUnroll][3];
#pragma unroll
            for (int u = 0; u < kGwUnroll; ++u) {
                const bool ok = pix0 + u * stride < npix;
                const uint32_t pix = ok ? pix0 + u * stride : pix0;
                yv[u] = __ldg(reinterpret_cast<const uint4*>(y) + pix * chunks + ch);
                head_gpre<kOne>(gimg, img, pix, HW, use_sigmoid, gp[u]);
                if (!ok) gp[u][0] = gp[u][1] = gp[u][2] = 0.0f;   // contributes nothing
            }
#pragma unroll
            for (int u = 0; u < kGwUnroll; ++u) {
                const uint32_t yu[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float y0 = bf16_lo(yu[e]), y1 = bf16_hi(yu[e]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        gw[k][e * 2] = fmaf(gp[u][k], y0, gw[k][e * 2]);
                        gw[k][e * 2 + 1] = fmaf(gp[u][k], y1, gw[k][e * 2 + 1]);
                    }
                }
                gb[0] += gp[u][0];
                gb[1] += gp[u][1];
                gb[2] += gp[u][2];
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&sg[k * Cp + ch * 8 + e], gw[k][e]);
            if (ch == 0) atomicAdd(&sg[3 * Cp + k], gb[k]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * Cp; i += blockDim.x) {
        const int k = i / Cp, c = i % Cp;
        if (c < C) atomicAdd(&gWh[k * C + c], sg[i]);
    }
    if (threadIdx.x < 3) atomicAdd(&gbh[threadIdx.x], sg[3 * Cp + threadIdx.x]);
}

// Both halves in one pass over the pixels (same thread layout): reads y and SiLU'(z), writes dz, reduces gWh / gbh.
// Less total work than the two separate kernels (g_pre and the loop overhead are shared), but the reduction then
// sits on the critical path; onr_head_bwd uses it, the split entry points remain for callers that overlap them.
constexpr int kFusedUnroll = 2;
template <bool kOne>
__global__ void __launch_bounds__(384)
head_bwd_fused_kernel(const float* __restrict__ gimg, const float* __restrict__ img,
                      const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dsilu, uint32_t npix,
                      uint32_t HW, int C, int Cp, const float* __restrict__ Wh, int use_sigmoid,
                      float* __restrict__ gWh, float* __restrict__ gbh, __nv_bfloat16* __restrict__ dz) {
    __shared__ float sg[3 * kHeadMaxC + 3];
    for (int i = threadIdx.x; i < 3 * Cp + 3; i += blockDim.x) sg[i] = 0.0f;
    __syncthreads();
    const int chunks = Cp / 8;
    const int lanes = blockDim.x / chunks;
    const int lane = threadIdx.x / chunks, ch = threadIdx.x - lane * chunks;
    float w[3][8], gw[3][8];
    head_load_w(Wh, C, ch, w);
    float gb[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) gw[k][e] = 0.0f;
    const uint32_t stride = gridDim.x * lanes;
    for (uint32_t pix0 = blockIdx.x * lanes + lane; pix0 < npix; pix0 += kFusedUnroll * stride) {
        uint4 yv[kFusedUnroll], dv[kFusedUnroll];
        float gp[kFusedUnroll][3];
#pragma unroll
        for (int u = 0; u < kFusedUnroll; ++u) {
            const bool ok = pix0 + u * stride < npix;
            const uint32_t pix = ok ? pix0 + u * stride : pix0;
            yv[u] = __ldg(reinterpret_cast<const uint4*>(y) + pix * chunks + ch);
            dv[u] = __ldg(reinterpret_cast<const uint4*>(dsilu) + pix * chunks + ch);
            head_gpre<kOne>(gimg, img, pix, HW, use_sigmoid, gp[u]);
            if (!ok) gp[u][0] = gp[u][1] = gp[u][2] = 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kFusedUnroll; ++u) {
            const uint32_t pix = pix0 + u * stride;
            const uint32_t yu[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
            const uint32_t du[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
            uint32_t out[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float y0 = bf16_lo(yu[e]), y1 = bf16_hi(yu[e]);
                float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    gw[k][e * 2] = fmaf(gp[u][k], y0, gw[k][e * 2]);
                    gw[k][e * 2 + 1] = fmaf(gp[u][k], y1, gw[k][e * 2 + 1]);
                    d0 = fmaf(gp[u][k], w[k][2 * e], d0);
                    d1 = fmaf(gp[u][k], w[k][2 * e + 1], d1);
                }
                out[e] = pack_bf16x2(d0 * bf16_lo(du[e]), d1 * bf16_hi(du[e]));
            }
            gb[0] += gp[u][0];
            gb[1] += gp[u][1];
            gb[2] += gp[u][2];
            if (pix < npix)
                reinterpret_cast<uint4*>(dz)[pix * chunks + ch] = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&sg[k * Cp + ch * 8 + e], gw[k][e]);
        if (ch == 0) atomicAdd(&sg[3 * Cp + k], gb[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * Cp; i += blockDim.x) {
        const int k = i / Cp, c = i % Cp;
        if (c < C) atomicAdd(&gWh[k * C + c], sg[i]);
    }
    if (threadIdx.x < 3) atomicAdd(&gbh[threadIdx.x], sg[3 * Cp + threadIdx.x]);
}

// threads per block: a multiple of `chunks` close to `target`, holding whole pixels
static inline int head_threads(int chunks, int target) { return (target / chunks) * chunks; }
static inline bool head_fits_u32(size_t npix, int chunks) { return npix * (size_t)chunks * 3 < (1ull << 31); }

}  // namespace onr

extern "C" {

int onr_head_fwd(const void* y, int B, int H, int W, int C, int Cp, const float* Wh, const float* bh,
                 int use_sigmoid, float* img, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    const int chunks = Cp / 8;
    ONR_REQUIRE(head_fits_u32(npix, chunks), "head: too many pixels for 32-bit indexing");
    const int threads = head_threads(chunks, 384);
    const int ppb = threads / chunks;
    ONR_REQUIRE(3 * ppb <= threads, "head: block too small to finish its pixels");
    int grid = (int)((npix + (size_t)ppb * kFwdU - 1) / ((size_t)ppb * kFwdU));
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    auto kern = B == 1 ? head_fwd_kernel<true> : head_fwd_kernel<false>;
    kern<<<grid, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(y), (uint32_t)npix,
                                                    (uint32_t)(H * W), C, Cp, Wh, bh, use_sigmoid, img);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_head_bwd_dz(const float* gimg, const float* img, const void* dsilu, int B, int H, int W, int C, int Cp,
                    const float* Wh, int use_sigmoid, void* dz, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    const int chunks = Cp / 8;
    ONR_REQUIRE(head_fits_u32(npix, chunks), "head: too many pixels for 32-bit indexing");
    const int threads = head_threads(chunks, 256);
    const int ppb = threads / chunks;
    size_t grid = (npix + (size_t)ppb * kDzUnroll - 1) / ((size_t)ppb * kDzUnroll);
    if (grid > (size_t)num_sms() * 8) grid = (size_t)num_sms() * 8;
    auto kern = B == 1 ? head_bwd_dz_kernel<true> : head_bwd_dz_kernel<false>;
    kern<<<(int)grid, threads, 0, (cudaStream_t)stream>>>(
        gimg, img, reinterpret_cast<const __nv_bfloat16*>(dsilu), (uint32_t)npix, (uint32_t)(H * W), C, Cp, Wh,
        use_sigmoid, reinterpret_cast<__nv_bfloat16*>(dz));
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_head_bwd_gw(const float* gimg, const float* img, const void* y, int B, int H, int W, int C, int Cp,
                    int use_sigmoid, float* gWh, float* gbh, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    const int chunks = Cp / 8;
    ONR_REQUIRE(head_fits_u32(npix, chunks), "head: too many pixels for 32-bit indexing");
    const int lanes = 384 / chunks;
    const int threads = lanes * chunks;
    int grid = (int)((npix + lanes - 1) / lanes);
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    auto kern = B == 1 ? head_bwd_gw_kernel<true> : head_bwd_gw_kernel<false>;
    kern<<<grid, threads, 0, (cudaStream_t)stream>>>(
        gimg, img, reinterpret_cast<const __nv_bfloat16*>(y), (uint32_t)npix, (uint32_t)(H * W), C, Cp,
        use_sigmoid, gWh, gbh);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_head_bwd(const float* gimg, const float* img, const void* y, const void* dsilu, int B, int H, int W,
                 int C, int Cp, const float* Wh, int use_sigmoid, float* gWh, float* gbh, void* dz,
                 void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    const int chunks = Cp / 8;
    ONR_REQUIRE(head_fits_u32(npix, chunks), "head: too many pixels for 32-bit indexing");
    const int lanes = 384 / chunks;
    const int threads = lanes * chunks;
    int grid = (int)((npix + lanes - 1) / lanes);
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    auto kern = B == 1 ? head_bwd_fused_kernel<true> : head_bwd_fused_kernel<false>;
    kern<<<grid, threads, 0, (cudaStream_t)stream>>>(
        gimg, img, reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(dsilu),
        (uint32_t)npix, (uint32_t)(H * W), C, Cp, Wh, use_sigmoid, gWh, gbh, reinterpret_cast<__nv_bfloat16*>(dz));
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
