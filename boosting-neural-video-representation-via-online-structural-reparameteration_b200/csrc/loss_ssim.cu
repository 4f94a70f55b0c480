// loss_ssim.cu — Fusion6 loss (0.7 * L1 + 0.3 * (1 - SSIM)) with its analytic backward, PSNR, MS-SSIM.
//
// Reference: utils.py:159-160 (Fusion6), utils.py:191-199 (psnr_fn), utils.py:201-211 (msssim_fn); the
// SSIM itself is the third-party pytorch_msssim==0.2.1 (requirements.txt:3), restated here from its
// published algorithm: 11-tap Gaussian (sigma 1.5, normalised), separable VALID filtering (vertical pass
// then horizontal) of X, Y, X*X, Y*Y, X*Y per channel; C1 = 1e-4, C2 = 9e-4;
//   cs = (2 s12 + C2) / (s1 + s2 + C2),  ssim = (2 m1 m2 + C1) / (m1^2 + m2^2 + C1) * cs,
// mean over the valid map per (batch, channel), then over channels; ms_ssim: 5 scales, ReLU on cs/ssim,
// avg_pool2d(kernel 2, padding = dim % 2) between scales, weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333).
//
// Backward (the reference relies on autograd): with m = F(X), q = F(X^2), r = F(XY) (F = the filter),
//   dS/dm = [2 m2 (A2 - A1) - 2 m S (B2 - B1)] / (B1 B2),  dS/dq = -S / B2,  dS/dr = 2 A1 / (B1 B2)
//   dS/dX = Ft(dS/dm) + 2 X Ft(dS/dq) + Y Ft(dS/dr)          (Ft = transposed, "full" filter)
// Two tiled kernels: pass 1 produces the SSIM statistics and the three coefficient maps, pass 2 applies
// Ft, adds the L1 sign term and accumulates |d| and d^2 for L1 / PSNR.
#include "onr_common.cuh"

namespace onr {

constexpr int kWin = 11;
constexpr int kTS = 32;
constexpr int kHalo = kWin - 1;
constexpr int kIn = kTS + kHalo;   // 42
constexpr int kPitch = kIn + 1;    // 43
constexpr int kVR = 8;             // output rows per thread in the vertical passes
constexpr int kHC = 4;             // output columns per thread in the horizontal passes

// The taps travel as a by-value kernel argument (constant bank), so nothing is uploaded at run time and
// every launch is CUDA-graph capturable.
struct Gauss { float g[kWin]; };

static Gauss make_gauss() {
    Gauss w;
    float sum = 0.0f;
    for (int i = 0; i < kWin; ++i) {
        const float c = (float)(i - kWin / 2);
        w.g[i] = expf(-(c * c) / (2.0f * 1.5f * 1.5f));
        sum += w.g[i];
    }
    for (int i = 0; i < kWin; ++i) w.g[i] /= sum;
    return w;
}

__device__ __forceinline__ float block_sum(float v, float* sred) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    float t = 0.0f;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? sred[lane] : 0.0f;
        t = warp_sum(t);
    }
    return t;  // valid in thread 0
}

// grid (tiles_x, tiles_y, planes).  plane_acc[plane*2 + {0,1}] += {sum ssim_map, sum cs_map}.
template <bool kGrad>
__global__ void __launch_bounds__(256)
ssim_stats_kernel(const Gauss gw, const float* __restrict__ X, const float* __restrict__ Y, int H, int W,
                  double* __restrict__ plane_acc, float* __restrict__ coef /* [3][planes][Hv][Wv] */,
                  int planes) {
    __shared__ float sx[kIn][kPitch];
    __shared__ float sy[kIn][kPitch];
    __shared__ float sv[5][kTS][kPitch];
    __shared__ float sred[8];
    const int Hv = H - kHalo, Wv = W - kHalo;
    const int plane = blockIdx.z;
    const int oy0 = blockIdx.y * kTS, ox0 = blockIdx.x * kTS;
    const float* xp = X + (size_t)plane * H * W;
    const float* yp = Y + (size_t)plane * H * W;
    for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
        const int r = i / kIn, c = i % kIn;
        const int gy = oy0 + r, gx = ox0 + c;
        const bool in = gy < H && gx < W;
        sx[r][c] = in ? xp[(size_t)gy * W + gx] : 0.0f;
        sy[r][c] = in ? yp[(size_t)gy * W + gx] : 0.0f;
    }
    __syncthreads();
    // vertical pass with a register sliding window: a thread owns one column and kVR consecutive output rows, so
    // each staged value is read once per kVR outputs (tap order per output is unchanged: k ascending)
    for (int i = threadIdx.x; i < (kTS / kVR) * kIn; i += 256) {
        const int c = i % kIn, r0 = (i / kIn) * kVR;
        // the five filtered quantities travel as two fp32 pairs (x,y), (xx,yy) + xy: 2 FFMA2 + 1 FFMA per tap
        uint64_t a01[kVR], a23[kVR];
        float a4[kVR];
#pragma unroll
        for (int o = 0; o < kVR; ++o) { a01[o] = 0ull; a23[o] = 0ull; a4[o] = 0.0f; }
#pragma unroll
        for (int j = 0; j < kVR + kHalo; ++j) {
            const float x = sx[r0 + j][c], y = sy[r0 + j][c];
            const uint64_t v01 = pack2(x, y), v23 = pack2(x * x, y * y);
            const float v4 = x * y;
#pragma unroll
            for (int o = 0; o < kVR; ++o) {
                const int k = j - o;
                if (k >= 0 && k < kWin) {
                    fma2_s(a01[o], gw.g[k], v01);
                    fma2_s(a23[o], gw.g[k], v23);
                    a4[o] = fmaf(gw.g[k], v4, a4[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kVR; ++o) {
            sv[0][r0 + o][c] = lo2(a01[o]); sv[1][r0 + o][c] = hi2(a01[o]);
            sv[2][r0 + o][c] = lo2(a23[o]); sv[3][r0 + o][c] = hi2(a23[o]);
            sv[4][r0 + o][c] = a4[o];
        }
    }
    __syncthreads();
    const float C1 = 1e-4f, C2 = 9e-4f;
    float ssim_sum = 0.0f, cs_sum = 0.0f;
    // horizontal pass: a thread owns one row and kHC consecutive output columns; a warp covers 4 rows x 8 column
    // groups (with the 43-float pitch its 32 shared reads fall in 32 distinct banks, and its global stores stay
    // within four 128-byte row segments)
    for (int i = threadIdx.x; i < kTS * (kTS / kHC); i += 256) {
        const int c0 = (i % (kTS / kHC)) * kHC, r = i / (kTS / kHC);
        uint64_t a01[kHC], a23[kHC];
        float a4[kHC];
#pragma unroll
        for (int o = 0; o < kHC; ++o) { a01[o] = 0ull; a23[o] = 0ull; a4[o] = 0.0f; }
#pragma unroll
        for (int j = 0; j < kHC + kHalo; ++j) {
            const uint64_t v01 = pack2(sv[0][r][c0 + j], sv[1][r][c0 + j]);
            const uint64_t v23 = pack2(sv[2][r][c0 + j], sv[3][r][c0 + j]);
            const float v4 = sv[4][r][c0 + j];
#pragma unroll
            for (int o = 0; o < kHC; ++o) {
                const int k = j - o;
                if (k >= 0 && k < kWin) {
                    fma2_s(a01[o], gw.g[k], v01);
                    fma2_s(a23[o], gw.g[k], v23);
                    a4[o] = fmaf(gw.g[k], v4, a4[o]);
                }
            }
        }
        float acc[kHC][5];
#pragma unroll
        for (int o = 0; o < kHC; ++o) {
            acc[o][0] = lo2(a01[o]); acc[o][1] = hi2(a01[o]);
            acc[o][2] = lo2(a23[o]); acc[o][3] = hi2(a23[o]);
            acc[o][4] = a4[o];
        }
        const int oy = oy0 + r;
#pragma unroll
        for (int o = 0; o < kHC; ++o) {
            const int ox = ox0 + c0 + o;
            if (oy >= Hv || ox >= Wv) continue;
            const float m1 = acc[o][0], m2 = acc[o][1], q = acc[o][2], p = acc[o][3], rr = acc[o][4];
            const float m1s = m1 * m1, m2s = m2 * m2, m12 = m1 * m2;
            const float s1 = q - m1s, s2 = p - m2s, s12 = rr - m12;
            const float A1 = 2.0f * m12 + C1, A2 = 2.0f * s12 + C2;
            const float B1 = m1s + m2s + C1, B2 = s1 + s2 + C2;
            const float cs = A2 / B2;
            const float S = (A1 / B1) * cs;
            ssim_sum += S;
            cs_sum += cs;
            if (kGrad) {
                const float inv = 1.0f / (B1 * B2);
                const size_t oo = ((size_t)plane * Hv + oy) * Wv + ox;
                const size_t mapsz = (size_t)planes * Hv * Wv;
                coef[oo] = (2.0f * m2 * (A2 - A1) - 2.0f * m1 * S * (B2 - B1)) * inv;
                coef[mapsz + oo] = -S / B2;
                coef[2 * mapsz + oo] = 2.0f * A1 * inv;
            }
        }
    }
    const float t0 = block_sum(ssim_sum, sred);
    const float t1 = block_sum(cs_sum, sred);
    if (threadIdx.x == 0) {
        atomicAdd(&plane_acc[plane * 2], (double)t0);
        atomicAdd(&plane_acc[plane * 2 + 1], (double)t1);
    }
}

// grid (tiles_x, tiles_y, planes) over INPUT pixels. gacc[0] += sum |d|, gacc[1] += sum d^2.
__global__ void __launch_bounds__(256)
fusion_bwd_kernel(const Gauss gw, const float* __restrict__ X, const float* __restrict__ Y, int H, int W,
                  const float* __restrict__ coef, int planes, float k_l1, float k_ssim, float k_mse,
                  float* __restrict__ grad, double* __restrict__ gacc) {
    __shared__ float sc[3][kIn][kPitch];
    __shared__ float sv[3][kTS][kPitch];
    __shared__ float sred[8];
    const int Hv = H - kHalo, Wv = W - kHalo;
    const int plane = blockIdx.z;
    const int y0 = blockIdx.y * kTS, x0 = blockIdx.x * kTS;
    const size_t mapsz = (size_t)planes * Hv * Wv;
    for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
        const int r = i / kIn, c = i % kIn;
        const int oy = y0 - kHalo + r, ox = x0 - kHalo + c;
        const bool in = oy >= 0 && oy < Hv && ox >= 0 && ox < Wv;
        const size_t o = ((size_t)plane * Hv + oy) * Wv + ox;
#pragma unroll
        for (int qn = 0; qn < 3; ++qn) sc[qn][r][c] = in ? coef[qn * mapsz + o] : 0.0f;
    }
    __syncthreads();
    // transposed ("full") filter, same sliding-window layout as the statistics kernel; the window is walked
    // downwards so that each output still accumulates its taps in ascending k
    for (int i = threadIdx.x; i < (kTS / kVR) * kIn; i += 256) {
        const int c = i % kIn, r0 = (i / kIn) * kVR;
        uint64_t a01[kVR];
        float a2[kVR];
#pragma unroll
        for (int o = 0; o < kVR; ++o) { a01[o] = 0ull; a2[o] = 0.0f; }
#pragma unroll
        for (int j = kVR + kHalo - 1; j >= 0; --j) {
            const uint64_t v01 = pack2(sc[0][r0 + j][c], sc[1][r0 + j][c]);
            const float v2 = sc[2][r0 + j][c];
#pragma unroll
            for (int o = 0; o < kVR; ++o) {
                const int k = o + kHalo - j;
                if (k >= 0 && k < kWin) {
                    fma2_s(a01[o], gw.g[k], v01);
                    a2[o] = fmaf(gw.g[k], v2, a2[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kVR; ++o) {
            sv[0][r0 + o][c] = lo2(a01[o]); sv[1][r0 + o][c] = hi2(a01[o]); sv[2][r0 + o][c] = a2[o];
        }
    }
    __syncthreads();
    float l1 = 0.0f, l2 = 0.0f;
    for (int i = threadIdx.x; i < kTS * (kTS / kHC); i += 256) {
        const int c0 = (i % (kTS / kHC)) * kHC, r = i / (kTS / kHC);
        uint64_t a01[kHC];
        float a2[kHC];
#pragma unroll
        for (int o = 0; o < kHC; ++o) { a01[o] = 0ull; a2[o] = 0.0f; }
#pragma unroll
        for (int j = kHC + kHalo - 1; j >= 0; --j) {
            const uint64_t v01 = pack2(sv[0][r][c0 + j], sv[1][r][c0 + j]);
            const float v2 = sv[2][r][c0 + j];
#pragma unroll
            for (int o = 0; o < kHC; ++o) {
                const int k = o + kHalo - j;
                if (k >= 0 && k < kWin) {
                    fma2_s(a01[o], gw.g[k], v01);
                    a2[o] = fmaf(gw.g[k], v2, a2[o]);
                }
            }
        }
        float acc[kHC][3];
#pragma unroll
        for (int o = 0; o < kHC; ++o) { acc[o][0] = lo2(a01[o]); acc[o][1] = hi2(a01[o]); acc[o][2] = a2[o]; }
        const int y = y0 + r;
#pragma unroll
        for (int o = 0; o < kHC; ++o) {
            const int x = x0 + c0 + o;
            if (y >= H || x >= W) continue;
            const size_t idx = ((size_t)plane * H + y) * W + x;
            const float xv = X[idx], yv = Y[idx];
            const float d = xv - yv;
            l1 += fabsf(d);
            l2 = fmaf(d, d, l2);
            if (grad) {
                const float sgn = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
                const float g = k_l1 * sgn + k_ssim * (acc[o][0] + 2.0f * xv * acc[o][1] + yv * acc[o][2]);
                grad[idx] = k_mse != 0.0f ? fmaf(k_mse, d, g) : g;   // k_mse == 0 keeps the Fusion6 bits unchanged
            }
        }
    }
    const float t0 = block_sum(l1, sred);
    const float t1 = block_sum(l2, sred);
    if (threadIdx.x == 0) {
        atomicAdd(&gacc[0], (double)t0);
        atomicAdd(&gacc[1], (double)t1);
    }
}

__global__ void fusion_finalize_kernel(const double* __restrict__ gacc, const double* __restrict__ plane_acc,
                                       int planes, double n_all, double n_valid, float w_l1, float w_ssim,
                                       float w_mse, float* __restrict__ out5) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int p = 0; p < planes; ++p) s += plane_acc[p * 2] / n_valid;
    const double ssim = s / planes;
    const double l1 = gacc[0] / n_all;
    const double mse = gacc[1] / n_all;
    out5[0] = (float)(w_l1 * l1 + w_ssim * (1.0 - ssim) + w_mse * mse);
    out5[1] = (float)l1;
    out5[2] = (float)ssim;
    out5[3] = (float)mse;
    out5[4] = (float)(-10.0 * log10(mse));
}

__global__ void scale_by_device_scalar_kernel(float4* __restrict__ x, size_t n4, float* __restrict__ tail, int ntail,
                                              const float* __restrict__ s) {
    const float k = __ldg(s);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = x[i];
        v.x *= k; v.y *= k; v.z *= k; v.w *= k;
        x[i] = v;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] *= k;
}

// avg_pool2d(kernel 2, stride 2, padding (ph, pw), count_include_pad) of both images of a scale in one launch
// (blockIdx.y selects prediction / target)
__global__ void avgpool2_kernel(const float* __restrict__ in0, const float* __restrict__ in1, int planes, int H,
                                int W, int ph, int pw, int Ho, int Wo, float* __restrict__ out0,
                                float* __restrict__ out1) {
    const float* __restrict__ in = blockIdx.y ? in1 : in0;
    float* __restrict__ out = blockIdx.y ? out1 : out0;
    const size_t total = (size_t)planes * Ho * Wo;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % Wo);
        const int oy = (int)((idx / Wo) % Ho);
        const size_t pl = idx / ((size_t)Wo * Ho);
        const float* p = in + pl * H * W;
        float s = 0.0f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int y = 2 * oy - ph + dy, x = 2 * ox - pw + dx;
                if (y >= 0 && y < H && x >= 0 && x < W) s += p[(size_t)y * W + x];
            }
        out[idx] = s * 0.25f;
    }
}

struct MsDims { int H[5], W[5]; };

__global__ void msssim_finalize_kernel(const double* __restrict__ acc /* [5][planes][2] */,
                                       const double* __restrict__ acc0 /* scale 0: [planes][2] */, int planes,
                                       MsDims dims, float* __restrict__ out1) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double wts[5] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
    double total = 0.0;
    for (int p = 0; p < planes; ++p) {
        double prod = 1.0;
        for (int l = 0; l < 5; ++l) {
            const double nv = (double)(dims.H[l] - kHalo) * (double)(dims.W[l] - kHalo);
            const double* a = l == 0 ? acc0 + (size_t)p * 2 : acc + ((size_t)l * planes + p) * 2;
            double v = (l < 4 ? a[1] : a[0]) / nv;
            // the reference does this in fp32: relu then pow
            float vf = (float)v;
            vf = vf > 0.0f ? vf : 0.0f;
            prod *= (double)powf(vf, (float)wts[l]);
        }
        total += prod;
    }
    out1[0] = (float)(total / planes);
}

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace onr

extern "C" {

size_t onr_loss_workspace_bytes(int B, int H, int W) {
    using namespace onr;
    const size_t planes = (size_t)B * 3;
    const size_t Hv = H > kHalo ? H - kHalo : 0, Wv = W > kHalo ? W - kHalo : 0;
    return align256(16 * sizeof(double)) + align256(planes * 2 * sizeof(double)) +
           3 * planes * Hv * Wv * sizeof(float);
}

int onr_scale_by_device_scalar(float* x, size_t n, const float* scalar_dev, void* stream) {
    using namespace onr;
    ONR_REQUIRE(((uintptr_t)x & 15) == 0, "scale: x must be 16-byte aligned");
    const size_t n4 = n / 4;
    int grid = (int)((n4 + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    if (grid < 1) grid = 1;
    scale_by_device_scalar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(x), n4, x + n4 * 4,
                                                                          (int)(n - n4 * 4), scalar_dev);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_fusion6_fwd_bwd(const float* pred, const float* target, int B, int H, int W, float w_l1,
                        float w_ssim, float grad_scale, float* out5, float* grad_pred, void* work,
                        void* stream) {
    return onr_fusion_loss(pred, target, B, H, W, w_l1, 0.0f, w_ssim, grad_scale, out5, grad_pred, work, stream);
}

int onr_fusion_loss(const float* pred, const float* target, int B, int H, int W, float w_l1, float w_mse,
                    float w_ssim, float grad_scale, float* out5, float* grad_pred, void* work, void* stream) {
    using namespace onr;
    ONR_REQUIRE(H > kHalo && W > kHalo, "fusion6: image smaller than the 11x11 SSIM window");
    const Gauss gw = make_gauss();
    cudaStream_t st = (cudaStream_t)stream;
    const int planes = B * 3;
    const int Hv = H - kHalo, Wv = W - kHalo;
    uint8_t* wp = reinterpret_cast<uint8_t*>(work);
    double* gacc = reinterpret_cast<double*>(wp);
    double* pacc = reinterpret_cast<double*>(wp + align256(16 * sizeof(double)));
    float* coef = reinterpret_cast<float*>(wp + align256(16 * sizeof(double)) +
                                           align256((size_t)planes * 2 * sizeof(double)));
    ONR_CUDA(cudaMemsetAsync(wp, 0, align256(16 * sizeof(double)) + align256((size_t)planes * 2 * sizeof(double)),
                             st));
    dim3 g1(ceil_div(Wv, kTS), ceil_div(Hv, kTS), planes);
    if (grad_pred) ssim_stats_kernel<true><<<g1, 256, 0, st>>>(gw, pred, target, H, W, pacc, coef, planes);
    else ssim_stats_kernel<false><<<g1, 256, 0, st>>>(gw, pred, target, H, W, pacc, coef, planes);
    ONR_LAUNCH_CHECK();
    const double n_all = (double)planes * H * W, n_valid = (double)Hv * Wv;
    const float k_l1 = (float)(w_l1 / n_all) * grad_scale;
    const float k_ssim = (float)(-(double)w_ssim / (n_valid * planes)) * grad_scale;
    const float k_mse = (float)(2.0 * (double)w_mse / n_all) * grad_scale;
    dim3 g2(ceil_div(W, kTS), ceil_div(H, kTS), planes);
    fusion_bwd_kernel<<<g2, 256, 0, st>>>(gw, pred, target, H, W, coef, planes, k_l1, k_ssim, k_mse, grad_pred, gacc);
    ONR_LAUNCH_CHECK();
    fusion_finalize_kernel<<<1, 32, 0, st>>>(gacc, pacc, planes, n_all, n_valid, w_l1, w_ssim, w_mse, out5);
    ONR_LAUNCH_CHECK();
    return 0;
}

static void ms_dims(int H, int W, onr::MsDims* d) {
    d->H[0] = H;
    d->W[0] = W;
    for (int l = 1; l < 5; ++l) {
        const int ph = d->H[l - 1] % 2, pw = d->W[l - 1] % 2;
        d->H[l] = (d->H[l - 1] + 2 * ph - 2) / 2 + 1;
        d->W[l] = (d->W[l - 1] + 2 * pw - 2) / 2 + 1;
    }
}

size_t onr_msssim_workspace_bytes(int B, int H, int W) {
    using namespace onr;
    MsDims d;
    ms_dims(H, W, &d);
    const size_t planes = (size_t)B * 3;
    size_t bytes = align256(5 * planes * 2 * sizeof(double));
    for (int l = 1; l < 5; ++l) bytes += 2 * align256(planes * d.H[l] * d.W[l] * sizeof(float));
    return bytes;
}

int onr_msssim(const float* pred, const float* target, int B, int H, int W, float* out1, void* work,
               const void* loss_work, void* stream) {
    using namespace onr;
    MsDims d;
    ms_dims(H, W, &d);
    ONR_REQUIRE(d.H[4] > kHalo && d.W[4] > kHalo, "ms-ssim: image side must exceed 160 pixels");
    const Gauss gw = make_gauss();
    cudaStream_t st = (cudaStream_t)stream;
    const int planes = B * 3;
    uint8_t* wp = reinterpret_cast<uint8_t*>(work);
    double* acc = reinterpret_cast<double*>(wp);
    const size_t acc_bytes = align256((size_t)5 * planes * 2 * sizeof(double));
    ONR_CUDA(cudaMemsetAsync(acc, 0, acc_bytes, st));
    uint8_t* cur = wp + acc_bytes;
    const float* x = pred;
    const float* y = target;
    // scale 0 is the plain SSIM of the loss: when the caller hands over the workspace of an onr_fusion6_fwd_bwd
    // call on the same images (already ordered before this one), its per-plane sums are reused
    const double* acc0 = loss_work ? reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(loss_work) +
                                                                     align256(16 * sizeof(double)))
                                   : acc;
    for (int l = 0; l < 5; ++l) {
        const int Hl = d.H[l], Wl = d.W[l];
        if (l > 0 || !loss_work) {
            dim3 g(ceil_div(Wl - kHalo, kTS), ceil_div(Hl - kHalo, kTS), planes);
            ssim_stats_kernel<false><<<g, 256, 0, st>>>(gw, x, y, Hl, Wl, acc + (size_t)l * planes * 2, nullptr,
                                                        planes);
            ONR_LAUNCH_CHECK();
        }
        if (l < 4) {
            const int Ho = d.H[l + 1], Wo = d.W[l + 1];
            const size_t nb = align256((size_t)planes * Ho * Wo * sizeof(float));
            float* nx = reinterpret_cast<float*>(cur);
            float* ny = reinterpret_cast<float*>(cur + nb);
            cur += 2 * nb;
            const size_t total = (size_t)planes * Ho * Wo;
            int grid = (int)((total + 255) / 256);
            if (grid > num_sms() * 16) grid = num_sms() * 16;
            avgpool2_kernel<<<dim3(grid, 2), 256, 0, st>>>(x, y, planes, Hl, Wl, Hl % 2, Wl % 2, Ho, Wo, nx, ny);
            ONR_LAUNCH_CHECK();
            x = nx;
            y = ny;
        }
    }
    msssim_finalize_kernel<<<1, 32, 0, st>>>(acc, acc0, planes, d, out1);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
