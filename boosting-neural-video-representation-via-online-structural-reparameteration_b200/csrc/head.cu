// head.cu — RGB head: 1x1 conv C->3 + bias + (tanh+1)/2 (or sigmoid), forward and backward.
// Reference: model.py:601 (head_layer), model.py:620-623 (forward).  Bandwidth-bound: the forward reads
// the last block's NHWC bf16 activation once; the backward reads y and SiLU'(z) once and writes dz once.
// Thread layout: 8-channel (16-byte) chunks, (Cp/8) consecutive threads per pixel, so a warp touches one
// contiguous span of the activation.
#include "onr_common.cuh"
#include "onr_ptx.cuh"

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

// The same fused backward fed by the bulk-copy engine (single image, H*W % 4 == 0): one persistent CTA per SM, a
// producer warp streams 128-pixel tiles of y, SiLU'(z) and the six gimg / img plane segments into a ring of shared
// memory stages (cp.async.bulk + mbarrier transaction counts), 32 * chunks consumer threads reduce them.  The whole
// ring is in flight while a tile is being consumed, which a register-fed loop at this register count cannot do.
// SiLU and its derivative from the pre-activation, with the arithmetic of the convolution epilogue (conv_igemm.cu)
__device__ __forceinline__ void silu_from_z(float z, float& y, float& d) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    const float sg = fmaf(0.5f, t, 0.5f);
    y = z * sg;
    d = fmaf(y, 1.0f - sg, sg);
}

constexpr int kHbPix = 128;          // pixels per stage
constexpr int kHbMaxStages = 4;
struct HeadStream {
    const float* gimg; const float* img; const __nv_bfloat16* y; const __nv_bfloat16* dsilu;
    uint32_t npix; int C, Cp; const float* Wh; int use_sigmoid;
    float* gWh; float* gbh; __nv_bfloat16* dz; int stages;
};
// kZ: `a.y` holds the pre-activation z and `a.dsilu` is unused — one activation stream instead of two.
template <bool kZ>
__global__ void __launch_bounds__(544, 1) head_bwd_stream_kernel(const HeadStream a) {
    extern __shared__ __align__(128) uint8_t hs_smem[];
    __shared__ float sg[3 * kHeadMaxC + 3];
    __shared__ __align__(8) uint64_t bars[2 * kHbMaxStages];
    const int chunks = a.Cp / 8;
    constexpr uint32_t kActs = kZ ? 1u : 2u;
    const uint32_t act_bytes = kHbPix * a.Cp * 2, stage_bytes = kActs * act_bytes + 6 * kHbPix * 4;
    const int S = a.stages;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kHbMaxStages]);
    const int n_cons_warps = chunks;           // 32 * chunks consumer threads
    for (int i = threadIdx.x; i < 3 * a.Cp + 3; i += blockDim.x) sg[i] = 0.0f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, n_cons_warps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t ntiles = (a.npix + kHbPix - 1) / kHbPix;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        // ---------------------------------------------------------------- producer
        if (elect_one()) {
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int s = it % S;
                if (it >= (uint32_t)S) mbar_wait(empty0 + 8 * s, ((it / S) - 1) & 1);
                const uint32_t pix0 = t * kHbPix;
                const uint32_t np = min((uint32_t)kHbPix, a.npix - pix0);
                const uint32_t dst = smem_u32(hs_smem) + s * stage_bytes;
                const uint32_t bar = full0 + 8 * s;
                mbar_expect_tx(bar, kActs * np * a.Cp * 2 + 6 * np * 4);
                bulk_load_1d(dst, a.y + (size_t)pix0 * a.Cp, np * a.Cp * 2, bar);
                if (!kZ) bulk_load_1d(dst + act_bytes, a.dsilu + (size_t)pix0 * a.Cp, np * a.Cp * 2, bar);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    bulk_load_1d(dst + kActs * act_bytes + k * kHbPix * 4, a.gimg + (size_t)k * a.npix + pix0, np * 4,
                                 bar);
                    bulk_load_1d(dst + kActs * act_bytes + (3 + k) * kHbPix * 4, a.img + (size_t)k * a.npix + pix0,
                                 np * 4, bar);
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- consumers
        const int ct = threadIdx.x - 32;
        const int lane_px = ct / chunks, ch = ct - lane_px * chunks;   // 32 pixel lanes x chunks
        float w[3][8], gw[3][8];
        head_load_w(a.Wh, a.C, ch, w);
        float gb[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) gw[k][e] = 0.0f;
        uint32_t it = 0;
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % S;
            mbar_wait(full0 + 8 * s, (it / S) & 1);
            const uint8_t* st = hs_smem + (size_t)s * stage_bytes;
            const uint4* sy = reinterpret_cast<const uint4*>(st);
            const uint4* sd = reinterpret_cast<const uint4*>(st + (kZ ? 0u : act_bytes));
            const float* sgi = reinterpret_cast<const float*>(st + kActs * act_bytes);
            const uint32_t pix0 = t * kHbPix;
            const uint32_t np = min((uint32_t)kHbPix, a.npix - pix0);
#pragma unroll
            for (int j = 0; j < kHbPix / 32; ++j) {
                const uint32_t p = j * 32 + lane_px;
                if (p >= np) continue;
                const uint4 yv = sy[p * chunks + ch], dv = sd[p * chunks + ch];
                float gp[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float g = sgi[k * kHbPix + p], o = sgi[(3 + k) * kHbPix + p];
                    gp[k] = a.use_sigmoid ? g * o * (1.0f - o) : g * 2.0f * o * (1.0f - o);
                }
                const uint32_t yu[4] = {yv.x, yv.y, yv.z, yv.w};
                const uint32_t du[4] = {dv.x, dv.y, dv.z, dv.w};
                uint32_t out[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float y0 = bf16_lo(yu[e]), y1 = bf16_hi(yu[e]);
                    float s0 = bf16_lo(du[e]), s1 = bf16_hi(du[e]);
                    if (kZ) {
                        silu_from_z(y0, y0, s0);
                        silu_from_z(y1, y1, s1);
                    }
                    float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        gw[k][e * 2] = fmaf(gp[k], y0, gw[k][e * 2]);
                        gw[k][e * 2 + 1] = fmaf(gp[k], y1, gw[k][e * 2 + 1]);
                        d0 = fmaf(gp[k], w[k][2 * e], d0);
                        d1 = fmaf(gp[k], w[k][2 * e + 1], d1);
                    }
                    out[e] = pack_bf16x2(d0 * s0, d1 * s1);
                }
                gb[0] += gp[0]; gb[1] += gp[1]; gb[2] += gp[2];
                reinterpret_cast<uint4*>(a.dz)[(size_t)(pix0 + p) * chunks + ch] =
                    make_uint4(out[0], out[1], out[2], out[3]);
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(empty0 + 8 * s);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&sg[k * a.Cp + ch * 8 + e], gw[k][e]);
            if (ch == 0) atomicAdd(&sg[3 * a.Cp + k], gb[k]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * a.Cp; i += blockDim.x) {
        const int k = i / a.Cp, c = i % a.Cp;
        if (c < a.C) atomicAdd(&a.gWh[k * a.C + c], sg[i]);
    }
    if (threadIdx.x < 3) atomicAdd(&a.gbh[threadIdx.x], sg[3 * a.Cp + threadIdx.x]);
}

// Forward in the same streaming form: the ring holds 128-pixel tiles of y.  Four lanes share a pixel, lane q
// owning the 16-byte chunks q, q+4, ... (CPT of them) with their weights in registers; two xor-shuffles finish the
// three dot products inside the warp, so consumer warps never meet at a barrier and run decoupled from each other.
struct HeadFwdStream {
    const __nv_bfloat16* y; uint32_t npix; int C, Cp; const float* Wh; const float* bh; int use_sigmoid;
    float* img; int stages;
};
constexpr int kHfConsumers = 512;   // 16 warps x 8 pixels: one 128-pixel tile per pass
template <int CPT, bool kZ>
__global__ void __launch_bounds__(32 + kHfConsumers, 1) head_fwd_stream_kernel(const HeadFwdStream a) {
    extern __shared__ __align__(128) uint8_t hs_smem[];
    __shared__ __align__(8) uint64_t bars[2 * 8];
    constexpr int chunks = 4 * CPT;
    const uint32_t stage_bytes = kHbPix * chunks * 16;
    const int S = a.stages;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, kHfConsumers / 32);   // one arrival per consumer warp
        }
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t ntiles = (a.npix + kHbPix - 1) / kHbPix;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int s = it % S;
                if (it >= (uint32_t)S) mbar_wait(empty0 + 8 * s, ((it / S) - 1) & 1);
                const uint32_t pix0 = t * kHbPix;
                const uint32_t np = min((uint32_t)kHbPix, a.npix - pix0);
                const uint32_t bar = full0 + 8 * s;
                mbar_expect_tx(bar, np * chunks * 16);
                bulk_load_1d(smem_u32(hs_smem) + s * stage_bytes, a.y + (size_t)pix0 * a.Cp, np * chunks * 16, bar);
            }
        }
        return;
    }
    const int ct = threadIdx.x - 32;
    const int q = ct & 3, pl = ct >> 2;           // lane within the pixel, pixel within the pass (0..63)
    float w[CPT][3][8];
#pragma unroll
    for (int i = 0; i < CPT; ++i) head_load_w(a.Wh, a.C, q + 4 * i, w[i]);
    const float bias = q < 3 ? __ldg(a.bh + q) : 0.0f;
    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % S;
        mbar_wait(full0 + 8 * s, (it / S) & 1);
        const uint4* sy = reinterpret_cast<const uint4*>(hs_smem + (size_t)s * stage_bytes);
        const uint32_t pix0 = t * kHbPix;
        const uint32_t np = min((uint32_t)kHbPix, a.npix - pix0);
        float res[kHbPix * 4 / kHfConsumers];
#pragma unroll
        for (int j = 0; j < kHbPix * 4 / kHfConsumers; ++j) {
            const uint32_t p = j * (kHfConsumers / 4) + pl;
            const uint32_t pc = p < np ? p : 0;       // lanes beyond a partial tile re-read pixel 0 (not stored)
            float acc[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const uint4 v = sy[pc * chunks + q + 4 * i];
                const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float lo = bf16_lo(u[e]), hi = bf16_hi(u[e]);
                    if (kZ) {
                        float unused;
                        silu_from_z(lo, lo, unused);
                        silu_from_z(hi, hi, unused);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        acc[k] = fmaf(hi, w[i][k][2 * e + 1], fmaf(lo, w[i][k][2 * e], acc[k]));
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
                acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
            }
            res[j] = q == 0 ? acc[0] : (q == 1 ? acc[1] : acc[2]);
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty0 + 8 * s);   // the stage has been read
#pragma unroll
        for (int j = 0; j < kHbPix * 4 / kHfConsumers; ++j) {
            const uint32_t p = j * (kHfConsumers / 4) + pl;
            if (q < 3 && p < np) a.img[(size_t)q * a.npix + pix0 + p] = head_act(res[j] + bias, a.use_sigmoid);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Wide heads (more than kHeadMaxC channels): the multi-resolution heads of wide early stages (sin_res=False with the
// reference's default widths: 1024 / 512 / 256 channels on 45x80 .. 270x480 maps, model.py:598-608).  Small maps off
// the north-star path: plain kernels, one warp per pixel with the lanes striding over the channels.
__global__ void head_wide_fwd_kernel(const __nv_bfloat16* __restrict__ y, uint32_t npix, uint32_t HW, int C, int Cp,
                                     const float* __restrict__ Wh, const float* __restrict__ bh, int use_sigmoid,
                                     float* __restrict__ img) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t pix = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pix < npix; pix += warps) {
        const __nv_bfloat16* row = y + (size_t)pix * Cp;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        for (int c = lane; c < C; c += 32) {
            const float v = __bfloat162float(row[c]);
            a0 = fmaf(v, __ldg(Wh + c), a0);
            a1 = fmaf(v, __ldg(Wh + C + c), a1);
            a2 = fmaf(v, __ldg(Wh + 2 * C + c), a2);
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
        if (lane < 3) {
            const float a = (lane == 0 ? a0 : (lane == 1 ? a1 : a2)) + __ldg(bh + lane);
            const uint32_t b = pix / HW;
            img[(size_t)b * 3u * HW + (size_t)lane * HW + (pix - b * HW)] = head_act(a, use_sigmoid);
        }
    }
}

// dz[p,c] = (sum_k Wh[k,c] gpre[p,k]) * d[p,c];  gWh[k,c] += sum_p gpre[p,k] y[p,c];  gbh[k] += sum_p gpre[p,k].
// A block owns a run of pixels; thread t owns channels t, t + blockDim, ...: its weight-gradient partials stay in
// registers over the run (3 per owned channel, at most kWideOwn channels) and are combined with atomics at the end.
constexpr int kWideOwn = 4;              // channels per thread: C <= kWideOwn * 256
__global__ void __launch_bounds__(256)
head_wide_bwd_kernel(const float* __restrict__ gimg, const float* __restrict__ img, const __nv_bfloat16* __restrict__ y,
                     const __nv_bfloat16* __restrict__ dact, uint32_t npix, uint32_t HW, int C, int Cp,
                     const float* __restrict__ Wh, int use_sigmoid, float* __restrict__ gWh, float* __restrict__ gbh,
                     __nv_bfloat16* __restrict__ dz, uint32_t px_per_block) {
    __shared__ float sgp[3];
    const uint32_t p0 = blockIdx.x * px_per_block;
    const uint32_t p1 = min(npix, p0 + px_per_block);
    float w[kWideOwn][3], acc[kWideOwn][3], gb[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < kWideOwn; ++j) {
        const int c = threadIdx.x + j * 256;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            w[j][k] = c < C ? __ldg(Wh + k * C + c) : 0.0f;
            acc[j][k] = 0.0f;
        }
    }
    for (uint32_t pix = p0; pix < p1; ++pix) {
        __syncthreads();
        if (threadIdx.x < 3) {
            const uint32_t b = pix / HW;
            const size_t o = (size_t)b * 3u * HW + (size_t)threadIdx.x * HW + (pix - b * HW);
            const float g = gimg[o], v = img[o];
            sgp[threadIdx.x] = use_sigmoid ? g * v * (1.0f - v) : g * 2.0f * v * (1.0f - v);
        }
        __syncthreads();
        const float g0 = sgp[0], g1 = sgp[1], g2 = sgp[2];
        if (threadIdx.x == 0) { gb[0] += g0; gb[1] += g1; gb[2] += g2; }
#pragma unroll
        for (int j = 0; j < kWideOwn; ++j) {
            const int c = threadIdx.x + j * 256;
            if (c < Cp) {
                const size_t o = (size_t)pix * Cp + c;
                float out = 0.0f;
                if (c < C) {
                    const float yv = __bfloat162float(y[o]);
                    acc[j][0] = fmaf(g0, yv, acc[j][0]);
                    acc[j][1] = fmaf(g1, yv, acc[j][1]);
                    acc[j][2] = fmaf(g2, yv, acc[j][2]);
                    out = (w[j][0] * g0 + w[j][1] * g1 + w[j][2] * g2) * __bfloat162float(dact[o]);
                }
                dz[o] = __float2bfloat16(out);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kWideOwn; ++j) {
        const int c = threadIdx.x + j * 256;
        if (c < C)
#pragma unroll
            for (int k = 0; k < 3; ++k) atomicAdd(gWh + k * C + c, acc[j][k]);
    }
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(gbh + k, gb[k]);
}

// threads per block: a multiple of `chunks` close to `target`, holding whole pixels
static inline int head_threads(int chunks, int target) { return (target / chunks) * chunks; }
static inline bool head_fits_u32(size_t npix, int chunks) { return npix * (size_t)chunks * 3 < (1ull << 31); }

}  // namespace onr

template <bool kZ>
static int head_fwd_stream_launch(const void* y, size_t npix, int C, int Cp, const float* Wh, const float* bh,
                                  int use_sigmoid, float* img, void* stream) {
    using namespace onr;
    const int chunks = Cp / 8;
    {
        const size_t stage_bytes = (size_t)kHbPix * Cp * 2;
        int stages = (int)((160 * 1024) / stage_bytes);
        if (stages > 8) stages = 8;
        auto kern = chunks == 4 ? head_fwd_stream_kernel<1, kZ> : chunks == 8 ? head_fwd_stream_kernel<2, kZ>
                  : chunks == 12 ? head_fwd_stream_kernel<3, kZ> : head_fwd_stream_kernel<4, kZ>;
        static bool attr_set[4] = {false, false, false, false};
        if (!attr_set[chunks / 4 - 1]) {
            ONR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[chunks / 4 - 1] = true;
        }
        HeadFwdStream hs{reinterpret_cast<const __nv_bfloat16*>(y), (uint32_t)npix, C, Cp, Wh, bh, use_sigmoid, img,
                         stages};
        const int ntiles = (int)((npix + kHbPix - 1) / kHbPix);
        const int grid = ntiles < num_sms() ? ntiles : num_sms();
        kern<<<grid, 32 + kHfConsumers, stages * stage_bytes, (cudaStream_t)stream>>>(hs);
        ONR_LAUNCH_CHECK();
        return 0;
    }
}

template <bool kZ>
static int head_bwd_stream_launch(const float* gimg, const float* img, const void* y, const void* dsilu, size_t npix,
                                  int C, int Cp, const float* Wh, int use_sigmoid, float* gWh, float* gbh, void* dz,
                                  void* stream) {
    using namespace onr;
    const int chunks = Cp / 8;
    const size_t stage_bytes = (kZ ? 1 : 2) * (size_t)kHbPix * Cp * 2 + 6 * kHbPix * 4;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > kHbMaxStages) stages = kHbMaxStages;
    ONR_REQUIRE(stages >= 2, "head: stage does not fit shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        ONR_CUDA(cudaFuncSetAttribute(head_bwd_stream_kernel<kZ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      200 * 1024));
        attr_set = true;
    }
    HeadStream hs{gimg, img, reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(dsilu),
                  (uint32_t)npix, C, Cp, Wh, use_sigmoid, gWh, gbh, reinterpret_cast<__nv_bfloat16*>(dz), stages};
    const int ntiles = (int)((npix + kHbPix - 1) / kHbPix);
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    head_bwd_stream_kernel<kZ><<<grid, 32 + 32 * chunks, stages * stage_bytes, (cudaStream_t)stream>>>(hs);
    ONR_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int onr_head_fwd_z(const void* z, int B, int H, int W, int C, int Cp, const float* Wh, const float* bh,
                   int use_sigmoid, float* img, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    ONR_REQUIRE(B == 1 && head_fits_u32(npix, Cp / 8), "head_fwd_z: single image only");
    return head_fwd_stream_launch<true>(z, npix, C, Cp, Wh, bh, use_sigmoid, img, stream);
}

int onr_head_fwd(const void* y, int B, int H, int W, int C, int Cp, const float* Wh, const float* bh,
                 int use_sigmoid, float* img, void* stream) {
    using namespace onr;
    const size_t npix = (size_t)B * H * W;
    if (Cp > kHeadMaxC) {                       // wide head of an early multi-resolution stage
        ONR_REQUIRE(Cp % 32 == 0 && C <= Cp && npix < (1ull << 31), "head: unsupported shape");
        size_t grid = (npix * 32 + 255) / 256;
        if (grid > (size_t)num_sms() * 8) grid = (size_t)num_sms() * 8;
        head_wide_fwd_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(y), (uint32_t)npix, (uint32_t)(H * W), C, Cp, Wh, bh, use_sigmoid, img);
        ONR_LAUNCH_CHECK();
        return 0;
    }
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const int chunks = Cp / 8;
    ONR_REQUIRE(head_fits_u32(npix, chunks), "head: too many pixels for 32-bit indexing");
    static const bool no_stream = getenv("ONR_HEAD_STREAM") && atoi(getenv("ONR_HEAD_STREAM")) == 0;
    if (B == 1 && !no_stream) return head_fwd_stream_launch<false>(y, npix, C, Cp, Wh, bh, use_sigmoid, img, stream);
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

int onr_head_bwd_z(const float* gimg, const float* img, const void* z, int B, int H, int W, int C, int Cp,
                   const float* Wh, int use_sigmoid, float* gWh, float* gbh, void* dz, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const size_t npix = (size_t)B * H * W;
    ONR_REQUIRE(B == 1 && npix % 4 == 0 && head_fits_u32(npix, Cp / 8), "head_bwd_z: single image with H*W % 4 == 0 only");
    return head_bwd_stream_launch<true>(gimg, img, z, nullptr, npix, C, Cp, Wh, use_sigmoid, gWh, gbh, dz, stream);
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
    const size_t npix = (size_t)B * H * W;
    if (Cp > kHeadMaxC) {                       // wide head of an early multi-resolution stage
        ONR_REQUIRE(Cp % 32 == 0 && C <= Cp && Cp <= kWideOwn * 256 && npix < (1ull << 31),
                    "head: at most %d channels", kWideOwn * 256);
        uint32_t ppb = (uint32_t)((npix + (size_t)num_sms() * 4 - 1) / ((size_t)num_sms() * 4));
        if (ppb < 8) ppb = 8;
        const int grid = (int)((npix + ppb - 1) / ppb);
        head_wide_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
            gimg, img, reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(dsilu),
            (uint32_t)npix, (uint32_t)(H * W), C, Cp, Wh, use_sigmoid, gWh, gbh, reinterpret_cast<__nv_bfloat16*>(dz), ppb);
        ONR_LAUNCH_CHECK();
        return 0;
    }
    ONR_REQUIRE(Cp % 32 == 0 && Cp <= kHeadMaxC && C <= Cp, "head: unsupported channels");
    const int chunks = Cp / 8;
    ONR_REQUIRE(head_fits_u32(npix, chunks), "head: too many pixels for 32-bit indexing");
    static const bool no_stream = getenv("ONR_HEAD_STREAM") && atoi(getenv("ONR_HEAD_STREAM")) == 0;
    if (B == 1 && npix % 4 == 0 && !no_stream)
        return head_bwd_stream_launch<false>(gimg, img, y, dsilu, npix, C, Cp, Wh, use_sigmoid, gWh, gbh, dz, stream);
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
