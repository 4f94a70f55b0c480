// fold_branches.cu — online structural re-parameterisation of the ACB / RepVGG / DBB / ECB branch sets
// (reference model.py:345-393, explicit multi-branch forward :541-565, SeqConv3x3 :191-300) into the single 3x3
// kernel + bias the tcgen05 convolution consumes, and the backward that scatters dK / dbias to the branch parameters.
// The arithmetic is in fold_branches.cuh (shared with the CPU check in tests/); here are the launches.
//
// fp32 SIMT on purpose: these are the reference's ablation baselines, a few MFLOP..GFLOP of contraction per step with
// channel counts of 26..112; the ERB fold, the north-star path, has its own tensor-core implementation (fold_tc.cu).
// Every gradient element is produced by exactly one thread (or one warp, reduced in a fixed order): bit-reproducible,
// which the data-parallel replicas rely on.
#include "onr_common.cuh"
#include "fold_branches.cuh"

namespace onr {

__global__ void branch_fold_k_kernel(const onr_branch_set s, float* __restrict__ K) {
    const size_t total = (size_t)s.cout * s.cin * 9;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx % 9);
        const int i = (int)((idx / 9) % s.cin);
        const int o = (int)(idx / ((size_t)9 * s.cin));
        K[idx] = branch_k_elem(s, o, i, t);
    }
}

__global__ void branch_fold_b_kernel(const onr_branch_set s, float* __restrict__ bias) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o < s.cout) bias[o] = branch_b_elem(s, o);
}

__global__ void branch_bwd_oi_kernel(const onr_branch_set s, const onr_branch_set g, const float* __restrict__ dK) {
    const size_t total = (size_t)s.cout * s.cin;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
        branch_bwd_oi(s, g, dK, (int)(idx / s.cin), (int)(idx % s.cin));
}

__global__ void branch_bwd_w2_kernel(const onr_branch_set s, const float* __restrict__ dK, float* __restrict__ gw2) {
    const int cm = 2 * s.cin;
    const size_t total = (size_t)s.cout * cm * 9;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx % 9);
        const int m = (int)((idx / 9) % cm);
        const int o = (int)(idx / ((size_t)9 * cm));
        gw2[idx] = branch_bwd_w2_elem(s, dK, o, m, t);
    }
}

// one warp per (m, i): the lanes stride over (o, t), fixed-order shuffle reduction
__global__ void branch_bwd_w1_kernel(const onr_branch_set s, const float* __restrict__ dK, float* __restrict__ gw1) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int total = 2 * s.cin * s.cin;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < total; p += warps) {
        const int m = p / s.cin, i = p - m * s.cin;
        const float v = warp_sum(branch_bwd_w1_partial(s, dK, m, i, lane, 32));
        if (lane == 0) gw1[p] = v;
    }
}

// one warp per output channel: the lanes stride over the input channels for the SeqConv3x3 scale gradients (fixed-order
// shuffle reduction), lane 0 writes
__global__ void branch_bwd_o_kernel(const onr_branch_set s, const onr_branch_set g, const float* __restrict__ dK,
                                    const float* __restrict__ db) {
    const int lane = threadIdx.x & 31;
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= s.cout) return;                                     // warp-uniform
    float dot[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int e = 0; e < 3; ++e)
        if (s.edge_k0[e]) dot[e] = warp_sum(branch_bwd_scale_partial(s, dK, e, o, lane, 32));
    if (lane == 0) branch_bwd_o(s, g, db, o, dot);
}

static inline int grid_of(size_t n, int block) {
    size_t gsz = (n + block - 1) / block;
    const size_t cap = (size_t)num_sms() * 8;
    return (int)(gsz < cap ? (gsz ? gsz : 1) : cap);
}

static int check_set(const onr_branch_set* s) {
    ONR_REQUIRE(s != nullptr && s->cin >= 1 && s->cout >= 1, "branch fold: bad channel counts");
    ONR_REQUIRE((s->seq_w1 == nullptr) == (s->seq_w2 == nullptr), "branch fold: the 1x1 -> 3x3 branch needs both kernels");
    for (int e = 0; e < 3; ++e)
        if (s->edge_k0[e])
            ONR_REQUIRE(s->edge_b0[e] && s->edge_scale[e] && s->edge_bias[e] && s->edge_mask[e],
                        "branch fold: SeqConv3x3 branch %d needs k0, b0, scale, bias and mask", e);
    return 0;
}

}  // namespace onr

extern "C" {

int onr_branch_fold_fwd(const onr_branch_set* w, float* K, float* bias, void* stream) {
    using namespace onr;
    if (int rc = check_set(w)) return rc;
    ONR_REQUIRE(K != nullptr && bias != nullptr, "branch fold: null output");
    cudaStream_t st = (cudaStream_t)stream;
    branch_fold_k_kernel<<<grid_of((size_t)w->cout * w->cin * 9, 256), 256, 0, st>>>(*w, K);
    ONR_LAUNCH_CHECK();
    branch_fold_b_kernel<<<ceil_div(w->cout, 128), 128, 0, st>>>(*w, bias);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_branch_fold_bwd(const onr_branch_set* w, const float* dK, const float* dbias, const onr_branch_set* g,
                        void* stream) {
    using namespace onr;
    if (int rc = check_set(w)) return rc;
    ONR_REQUIRE(dK != nullptr && dbias != nullptr && g != nullptr, "branch fold backward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    branch_bwd_oi_kernel<<<grid_of((size_t)w->cout * w->cin, 128), 128, 0, st>>>(*w, *g, dK);
    ONR_LAUNCH_CHECK();
    if (w->seq_w1) {
        if (g->seq_w2) {
            branch_bwd_w2_kernel<<<grid_of((size_t)w->cout * 2 * w->cin * 9, 256), 256, 0, st>>>(*w, dK, g->seq_w2);
            ONR_LAUNCH_CHECK();
        }
        if (g->seq_w1) {
            branch_bwd_w1_kernel<<<grid_of((size_t)2 * w->cin * w->cin * 32, 256), 256, 0, st>>>(*w, dK, g->seq_w1);
            ONR_LAUNCH_CHECK();
        }
    }
    branch_bwd_o_kernel<<<ceil_div(w->cout, 4), 128, 0, st>>>(*w, *g, dK, dbias);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
