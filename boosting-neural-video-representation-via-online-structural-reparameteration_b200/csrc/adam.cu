// adam.cu — fused multi-tensor Adam step (+ gradient averaging and zero-grad) in one launch.
// Reference: torch.optim.Adam as constructed at main_train.py:196 (betas=(args.beta, 0.999), eps 1e-8,
// no weight decay / amsgrad) and stepped at main_train.py:248-250; the learning rate comes from
// adjust_lr (utils.py:240-259) and is read from device memory so the step is CUDA-graph replayable.
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// 28 B of HBM traffic per parameter (p,g,m,v read; p,m,v written; g re-zeroed): bandwidth-bound,
// 128-bit accesses, grid = (blocks per tensor, tensors).
#include "onr_common.cuh"

namespace onr {

constexpr size_t kAdamPerBlock = 256 * 4 * 4;      // elements per block per sweep (4 float4 per thread)

struct AdamEntry {
    uint64_t p, g, m, v, n;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamEntry* __restrict__ table, const float* __restrict__ lr_dev,
                  const int* __restrict__ step_dev, float beta1, float beta2, float eps, float grad_scale,
                  int zero_grad) {
    const AdamEntry e = table[blockIdx.y];
    const size_t n = (size_t)e.n;
    // blocks beyond this tensor's share exit immediately; small tensors use one block
    const size_t my_blocks = (n + kAdamPerBlock - 1) / kAdamPerBlock;
    if (blockIdx.x >= my_blocks) return;
    const size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (start >= n) return;
    float* __restrict__ p = reinterpret_cast<float*>(e.p);
    float* __restrict__ g = reinterpret_cast<float*>(e.g);
    float* __restrict__ m = reinterpret_cast<float*>(e.m);
    float* __restrict__ v = reinterpret_cast<float*>(e.v);
    const float lr = *lr_dev;
    const int t = *step_dev;
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)t));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)t));
    const float step_size = lr / bc1;
    const size_t stride = (size_t)(my_blocks < gridDim.x ? my_blocks : gridDim.x) * blockDim.x * 4;
    const bool vec_ok = ((e.p | e.g | e.m | e.v) & 15ull) == 0;
    for (size_t i = start; i < n; i += stride) {
        if (vec_ok && i + 4 <= n) {
            float4 pv = *reinterpret_cast<float4*>(p + i);
            float4 gv = *reinterpret_cast<float4*>(g + i);
            float4 mv = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            float* pa = &pv.x; float* ga = &gv.x; float* ma = &mv.x; float* va = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gr = ga[k] * grad_scale;
                ma[k] = ma[k] + (gr - ma[k]) * (1.0f - beta1);
                va[k] = va[k] * beta2 + (1.0f - beta2) * gr * gr;
                const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
                pa[k] = pa[k] - step_size * (ma[k] / denom);
            }
            *reinterpret_cast<float4*>(p + i) = pv;
            *reinterpret_cast<float4*>(m + i) = mv;
            *reinterpret_cast<float4*>(v + i) = vv;
            if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (size_t j = i; j < n && j < i + 4; ++j) {
                const float gr = g[j] * grad_scale;
                const float mm = m[j] + (gr - m[j]) * (1.0f - beta1);
                const float vv = v[j] * beta2 + (1.0f - beta2) * gr * gr;
                const float denom = sqrtf(vv) / bc2_sqrt + eps;
                p[j] = p[j] - step_size * (mm / denom);
                m[j] = mm;
                v[j] = vv;
                if (zero_grad) g[j] = 0.0f;
            }
        }
    }
}

// One thread: t = ++step; lr(t) per adjust_lr (reference utils.py:240-259) evaluated in double like Python.
// The t-th optimizer step is iteration i = (t-1) % steps_per_epoch of epoch (t-1) / steps_per_epoch and uses
// cur_epoch = epoch + i / data_size (data_size = len(dataset), main_train.py:216, :247).
__global__ void sched_tick_kernel(int* __restrict__ step_dev, float* __restrict__ lr_dev, double lr0,
                                  int steps_per_epoch, int data_size, int warmup, int epochs, int lr_type) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int t = *step_dev + 1;
    *step_dev = t;
    const int epoch = ((t - 1) / steps_per_epoch) % epochs;
    const int it = (t - 1) % steps_per_epoch;
    const double e = (double)epoch + (double)it / (double)data_size;
    double mult = 1.0;
    if (lr_type == 0) mult = 0.5 * (cos(3.141592653589793 * (e - warmup) / (double)(epochs - warmup)) + 1.0);
    if (e < (double)warmup) mult = 0.1 + 0.9 * e / (double)warmup;
    *lr_dev = (float)(lr0 * mult);
}

}  // namespace onr

extern "C" int onr_sched_tick(int* step_dev, float* lr_dev, double lr0, int steps_per_epoch, int data_size,
                              int warmup, int epochs, int lr_type, void* stream) {
    using namespace onr;
    ONR_REQUIRE(steps_per_epoch >= 1 && data_size >= 1 && epochs >= 1 && (lr_type == 0 || lr_type == 1),
                "sched_tick: bad schedule (lr_type 0 = cosine, 1 = const)");
    sched_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_dev, lr_dev, lr0, steps_per_epoch, data_size, warmup,
                                                         epochs, lr_type);
    ONR_LAUNCH_CHECK();
    return 0;
}

extern "C" int onr_adam_multi(const uint64_t* table, int n_tensors, size_t max_numel, const float* lr_dev,
                              const int* step_dev, float beta1, float beta2, float eps, float grad_scale,
                              int zero_grad, void* stream) {
    using namespace onr;
    ONR_REQUIRE(n_tensors >= 1 && n_tensors <= 65535, "adam: bad tensor count %d", n_tensors);
    size_t bx = (max_numel + kAdamPerBlock - 1) / kAdamPerBlock;   // ~4 float4 per thread on the largest tensor
    if (bx < 1) bx = 1;
    if (bx > 296) bx = 296;                                        // two blocks per SM on the big tensors
    dim3 grid((unsigned)bx, (unsigned)n_tensors);
    adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AdamEntry*>(table), lr_dev,
                                                             step_dev, beta1, beta2, eps, grad_scale, zero_grad);
    ONR_LAUNCH_CHECK();
    return 0;
}
