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
// All contractions go through one strided fp32 GEMM kernel (64x64x16 tiles, 4x4 register blocks).
// Also: packing of an OIHW fp32 kernel into the bf16 implicit-GEMM operand layouts and back.
#include "onr_common.cuh"

namespace onr {

// offset(idx) = (idx / div) * hi + (idx % div) * lo   — lets a GEMM index run over (channel, tap) pairs
struct Axis {
    int div;
    long long hi, lo;
};
__device__ __forceinline__ long long axis_off(const Axis& a, int idx) {
    return (long long)(idx / a.div) * a.hi + (long long)(idx % a.div) * a.lo;
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

constexpr int GK = 16;

// TM x TN output tile per 256-thread block, GK-deep K slabs.  The next slab is fetched into registers while the
// current one is multiplied (the contractions are small and latency-bound: K <= 5850, a few hundred CTAs at most).
template <int TM, int TN>
__global__ void __launch_bounds__(256) strided_gemm_kernel(const GemmArgs g) {
    constexpr int RM = TM / 16, RN = TN / 16;              // per-thread register block
    constexpr int LA = TM * GK / 256, LB = TN * GK / 256;  // elements each thread stages per slab
    __shared__ float As[GK][TM + 1];
    __shared__ float Bs[GK][TN + 1];
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[RM][RN] = {};
    const int k_begin = blockIdx.z * g.k_chunk;
    const int k_end = min(g.K, k_begin + g.k_chunk);
    const bool split = gridDim.z > 1;
    // loop-invariant parts of the staging addresses
    long long a_off[LA], b_off[LB];
    int a_m[LA], a_k[LA], b_k[LB], b_n[LB];
#pragma unroll
    for (int l = 0; l < LA; ++l) {
        const int e = threadIdx.x + l * 256;
        a_m[l] = e / GK; a_k[l] = e % GK;
        a_off[l] = (m0 + a_m[l] < g.M) ? axis_off(g.am, m0 + a_m[l]) : -1;
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
        const int e = threadIdx.x + l * 256;
        b_k[l] = e / TN; b_n[l] = e % TN;
        b_off[l] = (n0 + b_n[l] < g.N) ? axis_off(g.bn, n0 + b_n[l]) : -1;
    }
    float ra[LA], rb[LB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int l = 0; l < LA; ++l) {
            const int k = k0 + a_k[l];
            ra[l] = (a_off[l] >= 0 && k < k_end) ? g.A[a_off[l] + axis_off(g.ak, k)] : 0.0f;
        }
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const int k = k0 + b_k[l];
            rb[l] = (b_off[l] >= 0 && k < k_end) ? g.B[axis_off(g.bk, k) + b_off[l]] : 0.0f;
        }
    };
    fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += GK) {
#pragma unroll
        for (int l = 0; l < LA; ++l) As[a_k[l]][a_m[l]] = ra[l];
#pragma unroll
        for (int l = 0; l < LB; ++l) Bs[b_k[l]][b_n[l]] = rb[l];
        __syncthreads();
        if (k0 + GK < k_end) fetch(k0 + GK);
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float a[RM], b[RN];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[k][ty * RM + i];
#pragma unroll
            for (int j = 0; j < RN; ++j) b[j] = Bs[k][tx * RN + j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty * RM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int n = n0 + tx * RN + j;
            if (n >= g.N) continue;
            float* c = g.C + axis_off(g.cm, m) + axis_off(g.cn, n);
            if (split) atomicAdd(c, acc[i][j]);
            else *c = g.accumulate ? *c + acc[i][j] : acc[i][j];
        }
    }
}

// allow_split: combine split-K partials with atomics (summation order, hence the last fp32 bits, then varies
// from run to run).  Used for gradients only; the forward fold stays deterministic so that a deploy
// checkpoint reproduces the train-state decode bit for bit (reference main_train.py:332-349).
static int launch_gemm(GemmArgs g, cudaStream_t st, bool allow_split = false) {
    // 64x64 tiles when they already cover the machine, 32x32 tiles (4x the CTAs) otherwise
    const int tiles64 = ceil_div(g.N, 64) * ceil_div(g.M, 64);
    const int T = tiles64 >= num_sms() ? 64 : 32;
    const int tiles = ceil_div(g.N, T) * ceil_div(g.M, T);
    int splits = 1;
    if (!g.accumulate && allow_split && g.contiguous_c && tiles < num_sms()) {
        // overwrite semantics with split-K: zero C, then accumulate partials atomically
        ONR_CUDA(cudaMemsetAsync(g.C, 0, (size_t)g.M * g.N * sizeof(float), st));
        g.accumulate = 1;
    }
    if (g.accumulate && allow_split) {
        // enough CTAs to cover the machine twice, but at least 64 k-elements of work per CTA
        splits = ceil_div(2 * num_sms(), tiles);
        const int max_splits = ceil_div(g.K, 64);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    g.k_chunk = ceil_div(ceil_div(g.K, splits), GK) * GK;
    splits = ceil_div(g.K, g.k_chunk);
    dim3 grid(ceil_div(g.N, T), ceil_div(g.M, T), splits);
    if (T == 64) strided_gemm_kernel<64, 64><<<grid, 256, 0, st>>>(g);
    else strided_gemm_kernel<32, 32><<<grid, 256, 0, st>>>(g);
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
