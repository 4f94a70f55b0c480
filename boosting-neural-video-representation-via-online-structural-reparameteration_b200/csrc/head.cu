// head.cu — RGB head: 1x1 conv C->3 + bias + (tanh+1)/2 (or sigmoid), forward and backward.
// Reference: model.py:601 (head_layer), model.py:620-623 (forward).  Bandwidth-bound: the forward reads
// the last block's NHWC bf16 activation once; the backward reads y and SiLU'(z) once and writes dz once.
// Thread layout: 8-channel (16-byte) chunks, (Cp/8) consecutive threads per pixel, so a warp touches one
// contiguous span of the activation.
#include "onr_common.cuh"

namespace onr {

constexpr int kHeadMaxC = 128;

__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ y, size_t npix, int HW, int C, int Cp,
                                const float* __restrict__ Wh, const float* __restrict__ bh, int use_sigmoid,
                                float* __restrict__ img) {
    __shared__ float sw[3 * kHeadMaxC];
    for (int i = threadIdx.x; i < 3 * Cp; i += blockDim.x) {
        const int k = i / Cp, c = i % Cp;
        sw[i] = c < C ? Wh[k * C + c] : 0.0f;
    }
    __syncthreads();
    const float b0 = bh[0], b1 = bh[1], b2 = bh[2];
    const int chunks = Cp / 8;
    for (size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pix < npix;
         pix += (size_t)gridDim.x * blockDim.x) {
        const uint4* yp = reinterpret_cast<const uint4*>(y + pix * Cp);
        float a0 = b0, a1 = b1, a2 = b2;
        for (int ch = 0; ch < chunks; ++ch) {
            const uint4 v = __ldg(yp + ch);
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = bf16_lo(u[e]), hi = bf16_hi(u[e]);
                const int c = ch * 8 + e * 2;
                a0 = fmaf(lo, sw[c], a0);           a0 = fmaf(hi, sw[c + 1], a0);
                a1 = fmaf(lo, sw[Cp + c], a1);      a1 = fmaf(hi, sw[Cp + c + 1], a1);
                a2 = fmaf(lo, sw[2 * Cp + c], a2);  a2 = fmaf(hi, sw[2 * Cp + c + 1], a2);
            }
        }
        const size_t b = pix / HW, hw = pix % HW;
        float* o = img + b * 3 * (size_t)HW + hw;
        if (use_sigmoid) {
            o[0] = 1.0f / (1.0f + expf(-a0));
            o[HW] = 1.0f / (1.0f + expf(-a1));
            o[2 * (size_t)HW] = 1.0f / (1.0f + expf(-a2));
        } else {
            o[0] = (tanhf(a0) + 1.0f) * 0.5f;
            o[HW] = (tanhf(a1) + 1.0f) * 0.5f;
            o[2 * (size_t)HW] = (tanhf(a2) + 1.0f) * 0.5f;
        }
    }
}

// Backward, split in two independent streaming kernels so that only the part the rest of the backward depends on
// (dz) sits on the critical path; the weight/bias gradient reduction can overlap the block kernels.
//   g_pre[k] = gimg[k] * d(act)/d(pre)  with  (tanh+1)/2 -> 2 o (1-o),  sigmoid -> o (1-o)
//   dz[px, c]  = (sum_k g_pre[k] Wh[k, c]) * SiLU'(z)[px, c]
//   gWh[k, c] += sum_px g_pre[k] y[px, c] ;  gbh[k] += sum_px g_pre[k]
// Thread layout for both: 8-channel (16-byte) pieces, (Cp/8) consecutive threads per pixel.
__device__ __forceinline__ void head_gpre(const float* __restrict__ gimg, const float* __restrict__ img, size_t pix,
                                          int HW, int use_sigmoid, float gp[3]) {
    const size_t b = pix / HW, hw = pix % HW;
    const size_t io = b * 3 * (size_t)HW + hw;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float o = __ldg(img + io + k * (size_t)HW);
        const float g = __ldg(gimg + io + k * (size_t)HW);
        gp[k] = use_sigmoid ? g * o * (1.0f - o) : g * 2.0f * o * (1.0f - o);
    }
}

__global__ void __launch_bounds__(256)
head_bwd_dz_kernel(const float* __restrict__ gimg, const float* __restrict__ img,
                   const __nv_bfloat16* __restrict__ dsilu, size_t npix, int HW, int C, int Cp,
                   const float* __restrict__ Wh, int use_sigmoid, __nv_bfloat16* __restrict__ dz) {
    __shared__ float sw[3 * kHeadMaxC];
    for (int i = threadIdx.x; i < 3 * Cp; i += blockDim.x) {
        const int k = i / Cp, c = i % Cp;
        sw[i] = c < C ? Wh[k * C + c] : 0.0f;
    }
    __syncthreads();
    const int chunks = Cp / 8;
    const size_t total = npix * chunks;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // 4 independent 16-byte loads in flight per thread
    for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
        uint4 dv[4];
        float gp[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t i = i0 + u * stride;
            if (i < total) {
                dv[u] = __ldg(reinterpret_cast<const uint4*>(dsilu) + i);
                head_gpre(gimg, img, i / chunks, HW, use_sigmoid, gp[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t i = i0 + u * stride;
            if (i >= total) continue;
            const int ch = (int)(i % chunks);
            const uint32_t du[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
            uint32_t out[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = ch * 8 + e * 2;
                float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    d0 = fmaf(gp[u][k], sw[k * Cp + c], d0);
                    d1 = fmaf(gp[u][k], sw[k * Cp + c + 1], d1);
                }
                out[e] = pack_bf16x2(d0 * bf16_lo(du[e]), d1 * bf16_hi(du[e]));
            }
            reinterpret_cast<uint4*>(dz)[i] = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
}

// blockDim.x = chunks * lanes; thread (lane, ch) walks pixels lane, lane + lanes*gridDim, ... for its 8 channels.
__global__ void head_bwd_gw_kernel(const float* __restrict__ gimg, const float* __restrict__ img,
                                   const __nv_bfloat16* __restrict__ y, size_t npix, int HW, int C, int Cp,
                                   int use_sigmoid, float* __restrict__ gWh, float* __restrict__ gbh) {
    __shared__ float sg[3 * kHeadMaxC + 3];
    for (int i = threadIdx.x; i < 3 * Cp + 3; i += blockDim.x) sg[i] = 0.0f;
    __syncthreads();
    const int chunks = Cp / 8;
    const int lanes = blockDim.x / chunks;
    const int ch = threadIdx.x % chunks, lane = threadIdx.x / chunks;
    float gw[3][8];
    float gb[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) gw[k][e] = 0.0f;
    if (lane < lanes) {
        const size_t stride = (size_t)gridDim.x * lanes;
        for (size_t pix0 = (size_t)blockIdx.x * lanes + lane; pix0 < npix; pix0 += 2 * stride) {
            uint4 yv[2];
            float gp[2][3];
            bool ok[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const size_t pix = pix0 + u * stride;
                ok[u] = pix < npix;
                if (ok[u]) {
                    yv[u] = __ldg(reinterpret_cast<const uint4*>(y + pix * Cp) + ch);
                    head_gpre(gimg, img, pix, HW, use_sigmoid, gp[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!ok[u]) continue;
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
                if (ch == 0) {
                    gb[0] += gp[u][0];
                    gb[1] += gp[u][1];
                    gb[2] += gp[u][2];
                }
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

}  // namespace onr

extern "C" {

int onr_head_fwd(const void* y, int B, int H, int W, int C, int Cp, const float* Wh, const float* bh,
                 int use_sigmoid, float* img, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    int grid = (int)((npix + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    head_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(y), npix,
                                                           H * W, C, Cp, Wh, bh, use_sigmoid, img);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_head_bwd_dz(const float* gimg, const float* img, const void* dsilu, int B, int H, int W, int C, int Cp,
                    const float* Wh, int use_sigmoid, void* dz, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    const size_t total = npix * (Cp / 8);
    size_t grid = (total + 256 * 4 - 1) / (256 * 4);
    if (grid > (size_t)num_sms() * 8) grid = (size_t)num_sms() * 8;
    head_bwd_dz_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(
        gimg, img, reinterpret_cast<const __nv_bfloat16*>(dsilu), npix, H * W, C, Cp, Wh, use_sigmoid,
        reinterpret_cast<__nv_bfloat16*>(dz));
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_head_bwd_gw(const float* gimg, const float* img, const void* y, int B, int H, int W, int C, int Cp,
                    int use_sigmoid, float* gWh, float* gbh, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    const int chunks = Cp / 8;
    const int lanes = 384 / chunks;
    const int threads = lanes * chunks;
    int grid = (int)((npix + lanes - 1) / lanes);
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    head_bwd_gw_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
        gimg, img, reinterpret_cast<const __nv_bfloat16*>(y), npix, H * W, C, Cp, use_sigmoid, gWh, gbh);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_head_bwd(const float* gimg, const float* img, const void* y, const void* dsilu, int B, int H, int W,
                 int C, int Cp, const float* Wh, int use_sigmoid, float* gWh, float* gbh, void* dz,
                 void* stream) {
    int rc = onr_head_bwd_dz(gimg, img, dsilu, B, H, W, C, Cp, Wh, use_sigmoid, dz, stream);
    if (rc) return rc;
    return onr_head_bwd_gw(gimg, img, y, B, H, W, C, Cp, use_sigmoid, gWh, gbh, stream);
}

}  // extern "C"
