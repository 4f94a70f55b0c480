// fold.cu — ERB online structural re-parameterisation on the device, in fp32.
//
// Forward  (reference model.py:450-516):
//   T[o,i,h,w]    = sum_m W2[o,m,h,w] * W1[m,i]                  (_fuse_1x1_3x3_1x1_branch, :510)
//   Kseq[p,i,h,w] = sum_o W3[p,o] * T[o,i,h,w]                   (:513-515; the 9x `repeat` of W3 is
//                                                                 never materialised here)
//   K = W3x3 + padH(W1x3) + padW(W3x1) + Kseq ;  b = b3x3 + b1x3 + b3x1   (:475-476, :495-496)
// Backward (what autograd derives for the lines above):
//   g3x3 += dK ; g1x3 += dK[:,:,1,:] ; g3x1 += dK[:,:,:,1] ; gb* += db
//   gW3[p,o] += sum dK[p,.] T[o,.] ; dT = W3^T dK ; gW2[o,m,.] += sum_i dT[o,i,.] W1[m,i] ;
//   gW1[m,i] += sum_{o,hw} W2[o,m,hw] dT[o,i,hw]
// All contractions go through one strided fp32 GEMM kernel (32x32 tiles, four-way in-CTA split of each K slab).
// Also: packing of an OIHW fp32 kernel into the bf16 implicit-GEMM operand layouts and back.
#include "onr_common.cuh"

namespace onr {

// offset(idx) = (idx / div) * hi + (idx % div) * lo   — lets a GEMM index run over (channel, tap) pairs.
// A plain strided axis has div = 2^30 and takes the multiply-only path.
struct Axis {
    int div;
    long long hi, lo;
};
__device__ __forceinline__ int axis_off(const Axis& a, int idx) {
    if (idx < a.div) return idx * (int)a.lo;
    return (idx / a.div) * (int)a.hi + (idx % a.div) * (int)a.lo;
}
static inline long long axis_extent(const Axis& a, int n) {   // largest offset the axis produces for idx < n
    if (n <= 0) return 0;
    const int idx = n - 1;
    if (idx < a.div) return (long long)idx * a.lo;
    return (long long)(idx / a.div) * a.hi + (long long)(a.div - 1) * a.lo;
}

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    int M, N, K;
    Axis am, ak, bk, bn, cm, cn;
    int accumulate;  // C += A*B when non-zero
    int k_chunk;     // split-K: blockIdx.z handles k in [z*k_chunk, (z+1)*k_chunk); partials are combined
                     // with atomicAdd (only legal with accumulate != 0, i.e. C already holds its addend)
    int contiguous_c;  // C is a plain row-major [M][N] array (used to zero it before a split-K overwrite)
};

// fp32 GEMM over strided/composite axes: 32x32 output tile per 256-thread CTA.  The folds are many small
// contractions (tens to a few hundred tiles, K from 52 to 5850), so the tile is small to fill the machine without
// a cross-CTA split (the forward fold must be deterministic) and the CTA's 8 warps split every K slab four ways:
// k-group kg = tid/64 multiplies k in [kg*GK/4, (kg+1)*GK/4) of the slab with a 4x4 register block per thread
// (two 16-byte shared loads per 8 packed FFMA2 = 16 FMAs); the four partial tiles are summed in a fixed order at the end.
// AKF / BKF: the operand is staged with k running fastest across a warp (8 k x 4 rows per warp: whole 32-byte
// sectors of a k-contiguous operand, conflict-free shared stores) or with the m / n index fastest.
constexpr int GT = 32;        // tile edge
constexpr int GP = GT + 4;    // shared pitch (floats), keeps rows 16-byte aligned
// SK: both k axes are plain strides (all contractions but gW1), so the slab loop carries no axis decomposition.
template <int GK, bool AKF, bool BKF, bool SK>
__global__ void __launch_bounds__(256) fold_gemm_kernel(const GemmArgs g) {
    constexpr int L = GT * GK / 256;   // staged elements per thread per operand per slab
    constexpr int KB = GK / 8;         // 8-wide k blocks per slab
    constexpr int KPG = GK / 4;        // k per k-group per slab
    __shared__ __align__(16) float As[GK][GP];
    __shared__ __align__(16) float Bs[GK][GP];
    __shared__ float red[3][16][64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int k_begin = blockIdx.z * g.k_chunk;
    const int k_end = min(g.K, k_begin + g.k_chunk);
    const bool split = gridDim.z > 1;

    // staging coordinates: element l of this thread is (a_k(l), a_m(l)) of the slab
    auto st_k = [&](bool kf, int l) { return kf ? (warp % KB) * 8 + (lane & 7) : warp + 8 * l; };
    auto st_r = [&](bool kf, int l) { return kf ? ((warp + 8 * l) / KB) * 4 + (lane >> 3) : lane; };
    int a_off[L], b_off[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const int m = m0 + st_r(AKF, l), n = n0 + st_r(BKF, l);
        a_off[l] = m < g.M ? axis_off(g.am, m) : -1;
        b_off[l] = n < g.N ? axis_off(g.bn, n) : -1;
    }
    float ra[L], rb[L];
    const int ak_lo = (int)g.ak.lo, bk_lo = (int)g.bk.lo;
    auto k_off_a = [&](int k) { return SK ? k * ak_lo : axis_off(g.ak, k); };
    auto k_off_b = [&](int k) { return SK ? k * bk_lo : axis_off(g.bk, k); };
    auto fetch = [&](int k0) {
        if (AKF) {
            const int k = k0 + st_k(true, 0);
            const int ko = k < k_end ? k_off_a(k) : -1;
#pragma unroll
            for (int l = 0; l < L; ++l) ra[l] = (ko >= 0 && a_off[l] >= 0) ? __ldg(g.A + a_off[l] + ko) : 0.0f;
        } else {
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const int k = k0 + st_k(false, l);
                ra[l] = (k < k_end && a_off[l] >= 0) ? __ldg(g.A + a_off[l] + k_off_a(k)) : 0.0f;
            }
        }
        if (BKF) {
            const int k = k0 + st_k(true, 0);
            const int ko = k < k_end ? k_off_b(k) : -1;
#pragma unroll
            for (int l = 0; l < L; ++l) rb[l] = (ko >= 0 && b_off[l] >= 0) ? __ldg(g.B + b_off[l] + ko) : 0.0f;
        } else {
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const int k = k0 + st_k(false, l);
                rb[l] = (k < k_end && b_off[l] >= 0) ? __ldg(g.B + b_off[l] + k_off_b(k)) : 0.0f;
            }
        }
    };

    const int kg = tid >> 6, t = tid & 63, tx = t & 7, ty = t >> 3;
    uint64_t acc2[4][2] = {};   // 4x4 register block as fp32 pairs (columns 0-1, 2-3): 8 FFMA2 per k
    fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += GK) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            As[st_k(AKF, l)][st_r(AKF, l)] = ra[l];
            Bs[st_k(BKF, l)][st_r(BKF, l)] = rb[l];
        }
        __syncthreads();
        if (k0 + GK < k_end) fetch(k0 + GK);
#pragma unroll
        for (int kk = 0; kk < KPG; ++kk) {
            const int k = kg * KPG + kk;
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const uint64_t b01 = pack2(b.x, b.y), b23 = pack2(b.z, b.w);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fma2_s(acc2[i][0], av[i], b01);
                fma2_s(acc2[i][1], av[i], b23);
            }
        }
        __syncthreads();
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[i][0] = lo2(acc2[i][0]); acc[i][1] = hi2(acc2[i][0]);
        acc[i][2] = lo2(acc2[i][1]); acc[i][3] = hi2(acc2[i][1]);
    }
    // fixed-order sum of the four k-group partials
    if (kg > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[kg - 1][i * 4 + j][t] = acc[i][j];
    }
    __syncthreads();
    if (kg > 0) return;
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += red[q][i * 4 + j][t];
    int c_n[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        c_n[j] = n < g.N ? axis_off(g.cn, n) : -1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        float* crow = g.C + axis_off(g.cm, m);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c_n[j] < 0) continue;
            float* c = crow + c_n[j];
            if (split) atomicAdd(c, acc[i][j]);
            else *c = g.accumulate ? *c + acc[i][j] : acc[i][j];
        }
    }
}

template <int GK, bool SK>
static void launch_gemm_gk(const GemmArgs& g, dim3 grid, bool akf, bool bkf, cudaStream_t st) {
    if (akf && bkf) fold_gemm_kernel<GK, true, true, SK><<<grid, 256, 0, st>>>(g);
    else if (akf) fold_gemm_kernel<GK, true, false, SK><<<grid, 256, 0, st>>>(g);
    else if (bkf) fold_gemm_kernel<GK, false, true, SK><<<grid, 256, 0, st>>>(g);
    else fold_gemm_kernel<GK, false, false, SK><<<grid, 256, 0, st>>>(g);
}

// allow_split: combine split-K partials with atomics (summation order, hence the last fp32 bits, then varies
// from run to run).  Used for gradients only; the forward fold stays deterministic so that a deploy
// checkpoint reproduces the train-state decode bit for bit (reference main_train.py:332-349).
static int launch_gemm(GemmArgs g, cudaStream_t st, bool allow_split = false) {
    ONR_REQUIRE(axis_extent(g.am, g.M) + axis_extent(g.ak, g.K) < (1ll << 31) &&
                axis_extent(g.bn, g.N) + axis_extent(g.bk, g.K) < (1ll << 31) &&
                axis_extent(g.cm, g.M) + axis_extent(g.cn, g.N) < (1ll << 31),
                "fold gemm: operand too large for 32-bit offsets");
    // 64-deep slabs unless K would waste more than a tenth of them
    const int GK = (long long)ceil_div(g.K, 64) * 64 * 10 <= (long long)g.K * 11 ? 64 : 32;
    const int tiles = ceil_div(g.N, GT) * ceil_div(g.M, GT);
    int splits = 1;
    if (!g.accumulate && allow_split && g.contiguous_c && tiles < num_sms()) {
        // overwrite semantics with split-K: zero C, then accumulate partials atomically
        ONR_CUDA(cudaMemsetAsync(g.C, 0, (size_t)g.M * g.N * sizeof(float), st));
        g.accumulate = 1;
    }
    if (g.accumulate && allow_split) {
        // enough CTAs to cover the machine twice, but at least two slabs of work per CTA
        splits = ceil_div(2 * num_sms(), tiles);
        const int max_splits = ceil_div(g.K, 2 * GK);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    g.k_chunk = ceil_div(ceil_div(g.K, splits), GK) * GK;
    splits = ceil_div(g.K, g.k_chunk);
    dim3 grid(ceil_div(g.N, GT), ceil_div(g.M, GT), splits);
    // stage an operand k-fastest when its k stride is 1 (composite axes: when the inner run is contiguous)
    const bool akf = g.ak.lo == 1, bkf = g.bk.lo == 1;
    const bool sk = g.ak.div >= g.K && g.bk.div >= g.K;
    if (GK == 64 && sk) launch_gemm_gk<64, true>(g, grid, akf, bkf, st);
    else if (GK == 64) launch_gemm_gk<64, false>(g, grid, akf, bkf, st);
    else if (sk) launch_gemm_gk<32, true>(g, grid, akf, bkf, st);
    else launch_gemm_gk<32, false>(g, grid, akf, bkf, st);
    ONR_LAUNCH_CHECK();
    return 0;
}
static inline Axis ax(long long stride) { return Axis{1 << 30, 0, stride}; }   // idx * stride
static inline Axis ax2(int div, long long hi, long long lo) { return Axis{div, hi, lo}; }

// K = W3x3 + padH(W1x3) + padW(W3x1) (+ nothing else yet), bias = b3x3 + b1x3 + b3x1
__global__ void fold_base_kernel(const float* __restrict__ w3x3, const float* __restrict__ b3x3,
                                 const float* __restrict__ w1x3, const float* __restrict__ b1x3,
                                 const float* __restrict__ w3x1, const float* __restrict__ b3x1, int Cin,
                                 int Cout, float* __restrict__ K, float* __restrict__ bias) {
    const size_t total = (size_t)Cout * Cin * 9;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int hw = (int)(idx % 9);
        const size_t oi = idx / 9;
        const int h = hw / 3, w = hw % 3;
        float v13 = (h == 1) ? w1x3[oi * 3 + w] : 0.0f;   // 1x3 fills the middle row  (F.pad (0,0,1,1))
        float v31 = (w == 1) ? w3x1[oi * 3 + h] : 0.0f;   // 3x1 fills the middle column (F.pad (1,1,0,0))
        K[idx] = w3x3[idx] + (v13 + v31);
        if (idx < (size_t)Cout) bias[idx] = b3x3[idx] + (b1x3[idx] + b3x1[idx]);
    }
}

__global__ void fold_bwd_direct_kernel(const float* __restrict__ dK, const float* __restrict__ db, int Cin,
                                       int Cout, float* __restrict__ g3x3, float* __restrict__ gb3x3,
                                       float* __restrict__ g1x3, float* __restrict__ gb1x3,
                                       float* __restrict__ g3x1, float* __restrict__ gb3x1) {
    const size_t total = (size_t)Cout * Cin * 9;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int hw = (int)(idx % 9);
        const size_t oi = idx / 9;
        const int h = hw / 3, w = hw % 3;
        const float g = dK[idx];
        g3x3[idx] += g;
        if (h == 1) g1x3[oi * 3 + w] += g;
        if (w == 1) g3x1[oi * 3 + h] += g;
        if (idx < (size_t)Cout) {
            const float b = db[idx];
            gb3x3[idx] += b;
            gb1x3[idx] += b;
            gb3x1[idx] += b;
        }
    }
}

// OIHW fp32 -> wf[9][Npad][Cpi], wd[9][Cpi_rows][Nk] (taps flipped), bias_p[Npad]
__global__ void pack_weights_kernel(const float* __restrict__ K, const float* __restrict__ bias, int Cin,
                                    int Cnew, int s, int Npad, int Cpi_rows, int Cpi, int Cpo,
                                    __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                    float* __restrict__ bias_p) {
    const int Nk = s * s * Cpo;
    const size_t nwf = (size_t)9 * Npad * Cpi;
    const size_t nwd = wd ? (size_t)9 * Cpi_rows * Nk : 0;   // decode-only callers pass wd = NULL
    const size_t total = nwf + nwd + Npad;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        if (idx < nwf) {
            const int ci = (int)(idx % Cpi);
            const int n = (int)((idx / Cpi) % Npad);
            const int tap = (int)(idx / ((size_t)Cpi * Npad));
            float v = 0.0f;
            if (n < Nk && ci < Cin) {
                const int ij = n / Cpo, c = n % Cpo;
                if (c < Cnew) v = K[((size_t)(c * s * s + ij) * Cin + ci) * 9 + tap];
            }
            wf[idx] = __float2bfloat16(v);
        } else if (idx < nwf + nwd) {
            const size_t j = idx - nwf;
            const int n = (int)(j % Nk);
            const int ci = (int)((j / Nk) % Cpi_rows);
            const int tap = (int)(j / ((size_t)Nk * Cpi_rows));
            float v = 0.0f;
            if (ci < Cin) {
                const int ij = n / Cpo, c = n % Cpo;
                // dgrad walks taps with the opposite sign, so it can use the same tap index as fprop
                if (c < Cnew) v = K[((size_t)(c * s * s + ij) * Cin + ci) * 9 + tap];
            }
            wd[j] = __float2bfloat16(v);
        } else {
            const int n = (int)(idx - nwf - nwd);
            float v = 0.0f;
            if (n < Nk) {
                const int ij = n / Cpo, c = n % Cpo;
                if (c < Cnew) v = bias[c * s * s + ij];
            }
            bias_p[n] = v;
        }
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dKp, const float* __restrict__ dbias_p,
                                    int Cin, int Cnew, int s, int Cpi, int Cpo, float* __restrict__ dK,
                                    float* __restrict__ dbias) {
    const int Cout = Cnew * s * s;
    const size_t total = (size_t)Cout * Cin * 9;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % 9);
        const int ci = (int)((idx / 9) % Cin);
        const int o = (int)(idx / ((size_t)9 * Cin));
        const int c = o / (s * s), ij = o % (s * s);
        const int n = ij * Cpo + c;
        dK[idx] = dKp[((size_t)n * 9 + tap) * Cpi + ci];
        if (idx < (size_t)Cout) {
            const int o2 = (int)idx;
            const int c2 = o2 / (s * s), ij2 = o2 % (s * s);
            dbias[o2] = dbias_p[ij2 * Cpo + c2];
        }
    }
}

static inline int ew_grid(size_t total) {
    size_t g = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace onr

extern "C" {

int onr_erb_fold_fwd(const float* w3x3, const float* b3x3, const float* w1x3, const float* b1x3,
                     const float* w3x1, const float* b3x1, const float* w1, const float* w2, const float* w3,
                     int Cin, int Cout, float* K, float* bias, float* T, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cin > 0 && Cout > 0, "fold: bad channels");
    cudaStream_t st = (cudaStream_t)stream;
    const int C2 = 2 * Cin;
    fold_base_kernel<<<ew_grid((size_t)Cout * Cin * 9), 256, 0, st>>>(w3x3, b3x3, w1x3, b1x3, w3x1, b3x1, Cin,
                                                                     Cout, K, bias);
    ONR_LAUNCH_CHECK();
    // T[(o,hw), i] = sum_m W2[o,m,hw] * W1[m,i]
    GemmArgs g1{w2, w1, T, Cout * 9, Cin, C2,
                ax2(9, (long long)C2 * 9, 1), ax(9), ax(Cin), ax(1),
                ax2(9, (long long)Cin * 9, 1), ax(9), 0, 0, 0};
    int rc = launch_gemm(g1, st);
    if (rc) return rc;
    // K[p, (i,hw)] += sum_o W3[p,o] * T[o,(i,hw)]
    GemmArgs g2{w3, T, K, Cout, Cin * 9, Cout,
                ax(Cout), ax(1), ax((long long)Cin * 9), ax(1),
                ax((long long)Cin * 9), ax(1), 1, 0, 0};
    return launch_gemm(g2, st);
}

int onr_erb_fold_bwd(const float* dK, const float* dbias, const float* w1, const float* w2, const float* w3,
                     const float* T, int Cin, int Cout, float* g3x3, float* gb3x3, float* g1x3, float* gb1x3,
                     float* g3x1, float* gb3x1, float* gw1, float* gw2, float* gw3, float* dT, void* stream) {
    using namespace onr;
    ONR_REQUIRE(Cin > 0 && Cout > 0, "fold: bad channels");
    cudaStream_t st = (cudaStream_t)stream;
    const int C2 = 2 * Cin;
    const long long CK = (long long)Cin * 9;
    fold_bwd_direct_kernel<<<ew_grid((size_t)Cout * Cin * 9), 256, 0, st>>>(dK, dbias, Cin, Cout, g3x3, gb3x3,
                                                                           g1x3, gb1x3, g3x1, gb3x1);
    ONR_LAUNCH_CHECK();
    // gW3[p,o] += sum_x dK[p,x] T[o,x]
    GemmArgs a{dK, T, gw3, Cout, Cout, (int)CK, ax(CK), ax(1), ax(1), ax(CK), ax(Cout), ax(1), 1, 0, 0};
    int rc = launch_gemm(a, st, true);
    if (rc) return rc;
    // dT[o,x] = sum_p W3[p,o] dK[p,x]
    GemmArgs b{w3, dK, dT, Cout, (int)CK, Cout, ax(1), ax(Cout), ax(CK), ax(1), ax(CK), ax(1), 0, 0, 1};
    rc = launch_gemm(b, st, true);
    if (rc) return rc;
    // gW2[(o,hw), m] += sum_i dT[o,i,hw] W1[m,i]
    GemmArgs c{dT, w1, gw2, Cout * 9, C2, Cin,
               ax2(9, CK, 1), ax(9), ax(1), ax(Cin),
               ax2(9, (long long)C2 * 9, 1), ax(9), 1, 0, 0};
    rc = launch_gemm(c, st, true);
    if (rc) return rc;
    // gW1[m,i] += sum_{(o,hw)} W2[o,m,hw] dT[o,i,hw]
    GemmArgs d{w2, dT, gw1, C2, Cin, Cout * 9,
               ax(9), ax2(9, (long long)C2 * 9, 1), ax2(9, CK, 1), ax(9),
               ax(Cin), ax(1), 1, 0, 0};
    return launch_gemm(d, st, true);
}

int onr_pack_weights(const float* K, const float* bias, int Cin, int Cnew, int s, int Npad, int Cpi_rows,
                     void* wf, void* wd, float* bias_p, void* stream) {
    using namespace onr;
    const int Cpi = pad32(Cin), Cpo = pad32(Cnew);
    ONR_REQUIRE(Npad >= s * s * Cpo && Cpi_rows >= Cpi, "pack_weights: padded sizes too small");
    const size_t total = (size_t)9 * Npad * Cpi + (size_t)9 * Cpi_rows * s * s * Cpo + Npad;
    pack_weights_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(
        K, bias, Cin, Cnew, s, Npad, Cpi_rows, Cpi, Cpo, reinterpret_cast<__nv_bfloat16*>(wf),
        reinterpret_cast<__nv_bfloat16*>(wd), bias_p);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_unpack_wgrad(const float* dKp, const float* dbias_p, int Cin, int Cnew, int s, float* dK,
                     float* dbias, void* stream) {
    using namespace onr;
    const int Cpi = pad32(Cin), Cpo = pad32(Cnew);
    unpack_wgrad_kernel<<<ew_grid((size_t)Cnew * s * s * Cin * 9), 256, 0, (cudaStream_t)stream>>>(
        dKp, dbias_p, Cin, Cnew, s, Cpi, Cpo, dK, dbias);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
