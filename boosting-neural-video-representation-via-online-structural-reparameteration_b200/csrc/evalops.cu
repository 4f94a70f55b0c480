// evalops.cu — weight transforms of the eval / decode path.
//  * global magnitude pruning (reference main_eval.py:572-587, torch.nn.utils.prune.global_unstructured with
//    L1Unstructured): the k-th smallest |w| over all prunable tensors is found with an exact radix select
//    on the fp32 bit patterns (|w| as uint32 is order-preserving), then applied as a mask;
//  * per-row affine quantisation (reference utils.py:11-67 quantize_per_tensor): min/max over NON-ZERO
//    entries, scale = (max-min)/2^bit, q = round((t-min)/(scale+1e-19)) (round-half-even like torch.round),
//    new = min + scale*q.  The reference loops over rows in Python (3744 iterations for stem.2).
#include "onr_common.cuh"
#include <float.h>

namespace onr {

__global__ void __launch_bounds__(256)
abs_radix_hist_kernel(const float* __restrict__ w, size_t n, uint32_t prefix, uint32_t prefix_mask, int shift,
                      unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t bits = __float_as_uint(w[i]) & 0x7fffffffu;
        if ((bits & prefix_mask) == prefix) atomicAdd(&sh[(bits >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void magnitude_mask_kernel(const float* __restrict__ w, size_t n, float thr, float* __restrict__ mask,
                                      float* __restrict__ w_out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = w[i];
        const float mk = fabsf(v) > thr ? 1.0f : 0.0f;
        if (mask) mask[i] = mk;
        if (w_out) w_out[i] = v * mk;
    }
}

// order-preserving float <-> uint mapping for atomicMin/Max
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void quant_init_kernel(uint32_t* __restrict__ mm, int rows) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        mm[2 * i] = 0xffffffffu;      // running min (ordered encoding)
        mm[2 * i + 1] = 0u;           // running max
    }
}

// grid (chunks, rows)
__global__ void __launch_bounds__(256)
quant_minmax_kernel(const float* __restrict__ t, size_t cols, uint32_t* __restrict__ mm) {
    const int row = blockIdx.y;
    const float* p = t + (size_t)row * cols;
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cols; i += (size_t)gridDim.x * blockDim.x) {
        const float v = p[i];
        if (v != 0.0f) {
            const uint32_t o = f2ord(v);
            lo = min(lo, o);
            hi = max(hi, o);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo != 0xffffffffu) atomicMin(&mm[2 * row], lo);
        if (hi != 0u) atomicMax(&mm[2 * row + 1], hi);
    }
}

__global__ void __launch_bounds__(256)
quant_apply_kernel(const float* __restrict__ t, size_t cols, int bit, const uint32_t* __restrict__ mm,
                   float* __restrict__ q_out, float* __restrict__ new_out) {
    const int row = blockIdx.y;
    const uint32_t lo = mm[2 * row], hi = mm[2 * row + 1];
    const bool empty = lo == 0xffffffffu;           // no non-zero entry: reference uses [0, 0]
    const float t_min = empty ? 0.0f : ord2f(lo);
    const float t_max = empty ? 0.0f : ord2f(hi);
    const float scale = __fdiv_rn(__fsub_rn(t_max, t_min), exp2f((float)bit));
    const float denom = __fadd_rn(scale, 1e-19f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cols; i += (size_t)gridDim.x * blockDim.x) {
        const size_t idx = (size_t)row * cols + i;
        const float q = rintf(__fdiv_rn(__fsub_rn(t[idx], t_min), denom));
        if (q_out) q_out[idx] = q;
        if (new_out) new_out[idx] = __fadd_rn(t_min, __fmul_rn(scale, q));
    }
}

// uint8 frame -> fp32 in [0,1]: exactly torchvision ToTensor's `.float().div(255)` (reference model.py:64-65)
__global__ void frame_u8_to_f32_kernel(const uint8_t* __restrict__ src, size_t n, float* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __fdiv_rn((float)src[i], 255.0f);
}

// g[i] *= m[i]: the gradient side of torch.nn.utils.prune's `weight = weight_orig * weight_mask` re-parameterisation
// (reference main_eval.py:213-545, prune-then-finetune), applied to the whole flat gradient buffer in one pass.
__global__ void mul_inplace_kernel(float* __restrict__ g, const float* __restrict__ m, size_t n) {
    const size_t n4 = n / 4;
    float4* g4 = reinterpret_cast<float4*>(g);
    const float4* m4 = reinterpret_cast<const float4*>(m);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = g4[i];
        const float4 b = __ldg(m4 + i);
        a.x *= b.x; a.y *= b.y; a.z *= b.z; a.w *= b.w;
        g4[i] = a;
    }
    for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        g[i] *= m[i];
}

static inline int grid1d(size_t n) {
    size_t g = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace onr

extern "C" {

int onr_abs_radix_hist(const float* w, size_t n, uint32_t prefix, uint32_t prefix_mask, int shift,
                       unsigned long long* hist256, void* stream) {
    using namespace onr;
    ONR_REQUIRE(shift >= 0 && shift <= 24, "radix shift out of range");
    abs_radix_hist_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(w, n, prefix, prefix_mask, shift, hist256);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_apply_magnitude_mask(const float* w, size_t n, float thr, float* mask, float* w_out, void* stream) {
    using namespace onr;
    magnitude_mask_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(w, n, thr, mask, w_out);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_mul_inplace_f32(float* g, const float* m, size_t n, void* stream) {
    using namespace onr;
    ONR_REQUIRE(g != nullptr && m != nullptr, "mul_inplace: null argument");
    ONR_REQUIRE(((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0, "mul_inplace: buffers must be 16-byte aligned");
    if (n == 0) return 0;
    mul_inplace_kernel<<<grid1d((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(g, m, n);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_frame_u8_to_f32(const void* src_u8, size_t n, float* dst, void* stream) {
    using namespace onr;
    frame_u8_to_f32_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(src_u8),
                                                                       n, dst);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_quant_rows(const float* t, int rows, size_t cols, int bit, float* q_out, float* new_out, void* scratch,
                   void* stream) {
    using namespace onr;
    ONR_REQUIRE(rows >= 1 && rows <= 65535 && cols >= 1 && bit >= 1 && bit <= 16, "quant_rows: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* mm = reinterpret_cast<uint32_t*>(scratch);   // [rows][2]
    quant_init_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(mm, rows);
    ONR_LAUNCH_CHECK();
    size_t chunks = (cols + 256 * 8 - 1) / (256 * 8);
    if (chunks > 256) chunks = 256;
    dim3 grid((unsigned)chunks, (unsigned)rows);
    quant_minmax_kernel<<<grid, 256, 0, st>>>(t, cols, mm);
    ONR_LAUNCH_CHECK();
    quant_apply_kernel<<<grid, 256, 0, st>>>(t, cols, bit, mm, q_out, new_out);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
