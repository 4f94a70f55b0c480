// adam.cu — fused multi-tensor Adam step (+ gradient averaging and zero-grad) in one launch.
// Reference: torch.optim.Adam as constructed at main_train.py:196 (betas=(args.beta, 0.999), eps 1e-8,
// no weight decay / amsgrad) and stepped at main_train.py:248-250; the learning rate comes from
// adjust_lr (utils.py:240-259) and is read from device memory so the step is CUDA-graph replayable.
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// 28 B of HBM traffic per parameter (p,g,m,v read; p,m,v written; g re-zeroed): bandwidth-bound,
// 128-bit accesses.  The grid is flat: block i works on one kAdamPerBlock-element piece of one tensor, found by
// walking the (shared-memory copy of the) table, so no block is launched only to exit; the bias corrections
// are evaluated in double by one thread per block while the others already have their loads in flight.
#include "onr_common.cuh"

namespace onr {

constexpr int kAdamUnroll = 4;                               // float4 per array per thread
constexpr size_t kAdamPerBlock = 256 * 4 * kAdamUnroll;      // elements per block
constexpr int kAdamMaxTensors = 512;                         // table entries staged in shared memory per launch

struct AdamEntry {
    uint64_t p, g, m, v, n;
};

// Bias corrections in double (as torch.optim.Adam evaluates them in Python), once per step:
// hyp[0] = lr / (1 - b1^t), hyp[1] = sqrt(1 - b2^t).
__global__ void adam_hyper_kernel(const float* __restrict__ lr_dev, const int* __restrict__ step_dev, float beta1,
                                  float beta2, float* __restrict__ hyp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int t = *step_dev;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    hyp[0] = *lr_dev / (float)bc1;
    hyp[1] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamEntry* __restrict__ table, int n_tensors, const float* __restrict__ hyp,
                  float beta1, float beta2, float eps, float grad_scale, int zero_grad) {
    __shared__ AdamEntry s_tab[kAdamMaxTensors];
    __shared__ float s_hyp[2];
    __shared__ int s_sel[2];
    for (int i = threadIdx.x; i < n_tensors; i += blockDim.x) s_tab[i] = table[i];
    if (threadIdx.x >= 254) s_hyp[threadIdx.x - 254] = hyp[threadIdx.x - 254];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned rest = blockIdx.x;
        int sel = -1;
        for (int i = 0; i < n_tensors; ++i) {
            const unsigned nb = (unsigned)((s_tab[i].n + kAdamPerBlock - 1) / kAdamPerBlock);
            if (rest < nb) { sel = i; break; }
            rest -= nb;
        }
        s_sel[0] = sel;
        s_sel[1] = (int)rest;
    }
    __syncthreads();
    if (s_sel[0] < 0) return;
    const AdamEntry e = s_tab[s_sel[0]];
    const size_t n = (size_t)e.n;
    const size_t base = (size_t)s_sel[1] * kAdamPerBlock;
    float* __restrict__ p = reinterpret_cast<float*>(e.p);
    float* __restrict__ g = reinterpret_cast<float*>(e.g);
    float* __restrict__ m = reinterpret_cast<float*>(e.m);
    float* __restrict__ v = reinterpret_cast<float*>(e.v);
    const float step_size = s_hyp[0], bc2_sqrt = s_hyp[1];
    const bool vec_ok = ((e.p | e.g | e.m | e.v) & 15ull) == 0;
    // Each thread owns kAdamUnroll groups of 4 consecutive elements; a group is moved as one 128-bit access when
    // the tensor is 16-byte aligned and the group lies inside it, element by element otherwise (tensor tails,
    // unaligned storage).  Every load is issued before the first use.
    float4 pv[kAdamUnroll], gv[kAdamUnroll], mv[kAdamUnroll], vv[kAdamUnroll];
    int mode[kAdamUnroll];   // 2: vector, 1: element-wise, 0: nothing
#pragma unroll
    for (int u = 0; u < kAdamUnroll; ++u) {
        const size_t i = base + ((size_t)u * 256 + threadIdx.x) * 4;
        mode[u] = (vec_ok && i + 4 <= n) ? 2 : (i < n ? 1 : 0);
        pv[u] = gv[u] = mv[u] = vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mode[u] == 2) {
            pv[u] = *reinterpret_cast<const float4*>(p + i);
            gv[u] = *reinterpret_cast<const float4*>(g + i);
            mv[u] = *reinterpret_cast<const float4*>(m + i);
            vv[u] = *reinterpret_cast<const float4*>(v + i);
        } else if (mode[u] == 1) {
            float* pa = &pv[u].x; float* ga = &gv[u].x; float* ma = &mv[u].x; float* va = &vv[u].x;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k < n) { pa[k] = p[i + k]; ga[k] = g[i + k]; ma[k] = m[i + k]; va[k] = v[i + k]; }
        }
    }
#pragma unroll
    for (int u = 0; u < kAdamUnroll; ++u) {
        const size_t i = base + ((size_t)u * 256 + threadIdx.x) * 4;
        float* pa = &pv[u].x; float* ga = &gv[u].x; float* ma = &mv[u].x; float* va = &vv[u].x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = ga[k] * grad_scale;
            ma[k] = ma[k] + (gr - ma[k]) * (1.0f - beta1);
            va[k] = va[k] * beta2 + (1.0f - beta2) * gr * gr;
            const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
            pa[k] = pa[k] - step_size * (ma[k] / denom);
        }
        if (mode[u] == 2) {
            *reinterpret_cast<float4*>(p + i) = pv[u];
            *reinterpret_cast<float4*>(m + i) = mv[u];
            *reinterpret_cast<float4*>(v + i) = vv[u];
            if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (mode[u] == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k < n) {
                    p[i + k] = pa[k]; m[i + k] = ma[k]; v[i + k] = va[k];
                    if (zero_grad) g[i + k] = 0.0f;
                }
        }
    }
}

// One thread: t = ++step; lr(t) per adjust_lr (reference utils.py:240-259) evaluated in double like Python.
// The t-th optimizer step is iteration i = (t-1) % steps_per_epoch of epoch (t-1) / steps_per_epoch and uses
// cur_epoch = epoch + i / data_size (data_size = len(dataset), main_train.py:216, :247).
// epoch_offset / epoch_mod: the prune-then-finetune loop (reference main_eval.py:446-466) restarts the optimizer (Adam
// step count from 1) but continues the epoch numbering at the checkpoint's epoch and evaluates
// adjust_lr(epoch % (start_epoch + finetune_epochs)) against the ORIGINAL --epochs / warm-up.
__global__ void sched_tick_kernel(int* __restrict__ step_dev, float* __restrict__ lr_dev, double lr0,
                                  int steps_per_epoch, int data_size, int warmup, int epochs, int lr_type,
                                  int epoch_offset, int epoch_mod) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int t = *step_dev + 1;
    *step_dev = t;
    const int epoch = (epoch_offset + (t - 1) / steps_per_epoch) % epoch_mod;
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
                                                         epochs, lr_type, 0, epochs);
    ONR_LAUNCH_CHECK();
    return 0;
}

extern "C" int onr_sched_tick_ex(int* step_dev, float* lr_dev, double lr0, int steps_per_epoch, int data_size,
                                 int warmup, int epochs, int lr_type, int epoch_offset, int epoch_mod, void* stream) {
    using namespace onr;
    ONR_REQUIRE(steps_per_epoch >= 1 && data_size >= 1 && epochs >= 1 && (lr_type == 0 || lr_type == 1),
                "sched_tick: bad schedule (lr_type 0 = cosine, 1 = const)");
    ONR_REQUIRE(epoch_offset >= 0 && epoch_mod >= 1, "sched_tick: bad epoch offset / modulus");
    sched_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_dev, lr_dev, lr0, steps_per_epoch, data_size, warmup,
                                                         epochs, lr_type, epoch_offset, epoch_mod);
    ONR_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t onr_adam_block_elems(void) { return onr::kAdamPerBlock; }
extern "C" int onr_adam_max_tensors(void) { return onr::kAdamMaxTensors; }

extern "C" int onr_adam_multi(const uint64_t* table, int n_tensors, size_t total_blocks, const float* lr_dev,
                              const int* step_dev, float* hyp_scratch, float beta1, float beta2, float eps,
                              float grad_scale, int zero_grad, void* stream) {
    using namespace onr;
    ONR_REQUIRE(n_tensors >= 1 && n_tensors <= kAdamMaxTensors, "adam: %d tensors in one call (limit %d)",
                n_tensors, kAdamMaxTensors);
    ONR_REQUIRE(total_blocks >= 1 && total_blocks < (1ull << 31), "adam: bad block count");
    ONR_REQUIRE(hyp_scratch != nullptr, "adam: hyp_scratch (2 floats of device memory) is required");
    adam_hyper_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(lr_dev, step_dev, beta1, beta2, hyp_scratch);
    ONR_LAUNCH_CHECK();
    adam_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const AdamEntry*>(table), n_tensors, hyp_scratch, beta1, beta2, eps, grad_scale, zero_grad);
    ONR_LAUNCH_CHECK();
    return 0;
}
