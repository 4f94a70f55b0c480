// fold_branches.cuh — element arithmetic of the structural re-parameterisation of the reference's OTHER linear branch
// sets (model.py:345-393: ACB, RepVGG, DBB, ECB + SeqConv3x3 :191-300) into one 3x3 kernel + bias, and of its
// backward.  The reference trains these blocks with an explicit multi-branch forward (model.py:541-565) and cannot
// deploy them; every branch is linear in x, so the sum of the branches IS one 3x3 convolution with zero padding:
//
//   3x3                      K += W                                       b += bias
//   1x3 (pad (0,1))          K[:, :, 1, :] += W                           b += bias          (ACB)
//   3x1 (pad (1,0))          K[:, :, :, 1] += W                           b += bias          (ACB)
//   1x1                      K[:, :, 1, 1] += W                           b += bias          (RepVGG, DBB)
//   1x1 -> 3x3 (no biases)   K[o,i,t] += sum_m W2[o,m,t] W1[m,i]                             (DBB, ECB)
//   1x1 -> AvgPool3x3(pad 1, count_include_pad)   K[o,i,t] += W[o,i] / 9                     (DBB)
//   SeqConv3x3 (1x1 with bias, border padded WITH that bias, depthwise scale*mask 3x3 + bias; model.py:277-300):
//                            K[o,i,t] += scale[o] mask[o,t] k0[o,i]       b[o] += b0[o] scale[o] sum_t mask[o,t] + bias[o]
//
// (the bias-valued border padding of SeqConv3x3 is what makes the last line exact at the image border as well.)
// The functions are __host__ __device__ so that tests/ can compile this very arithmetic with g++ and check it against
// the oracle on the CPU before any GPU time is spent; the product only ever runs them inside the kernels of
// fold_branches.cu.
#pragma once

#ifndef ONR_HD
#if defined(__CUDACC__)
#define ONR_HD __host__ __device__ __forceinline__
#else
#define ONR_HD inline
#endif
#endif

#include "../../include/orepnerv.h"

namespace onr {

ONR_HD float branch_k_elem(const onr_branch_set& s, int o, int i, int t) {
    const int cin = s.cin, h = t / 3, w = t - h * 3;
    const size_t oi = (size_t)o * cin + i;
    float k = s.w3x3 ? s.w3x3[oi * 9 + t] : 0.0f;
    if (s.w1x3 && h == 1) k += s.w1x3[oi * 3 + w];
    if (s.w3x1 && w == 1) k += s.w3x1[oi * 3 + h];
    if (s.w1x1 && t == 4) k += s.w1x1[oi];
    if (s.avg_w) k += s.avg_w[oi] / 9.0f;
    for (int e = 0; e < 3; ++e)
        if (s.edge_k0[e]) k += s.edge_scale[e][o] * s.edge_mask[e][(size_t)o * 9 + t] * s.edge_k0[e][oi];
    if (s.seq_w1) {
        const int cm = 2 * cin;
        float acc = 0.0f;
        for (int m = 0; m < cm; ++m) acc += s.seq_w2[((size_t)o * cm + m) * 9 + t] * s.seq_w1[(size_t)m * cin + i];
        k += acc;
    }
    return k;
}

ONR_HD float edge_mask_sum(const float* mask, int o) {
    float ms = 0.0f;
    for (int t = 0; t < 9; ++t) ms += mask[(size_t)o * 9 + t];
    return ms;
}

ONR_HD float branch_b_elem(const onr_branch_set& s, int o) {
    float b = s.b3x3 ? s.b3x3[o] : 0.0f;
    if (s.b1x3) b += s.b1x3[o];
    if (s.b3x1) b += s.b3x1[o];
    if (s.b1x1) b += s.b1x1[o];
    for (int e = 0; e < 3; ++e)
        if (s.edge_k0[e])
            b += s.edge_b0[e][o] * s.edge_scale[e][o] * edge_mask_sum(s.edge_mask[e], o) + s.edge_bias[e][o];
    return b;
}

// gradients that depend on one (o, i) column of dK: the direct branches and the SeqConv3x3 1x1 kernels.  OVERWRITES.
ONR_HD void branch_bwd_oi(const onr_branch_set& s, const onr_branch_set& g, const float* dK, int o, int i) {
    const int cin = s.cin;
    const size_t oi = (size_t)o * cin + i;
    float dk[9], sum = 0.0f;
    for (int t = 0; t < 9; ++t) {
        dk[t] = dK[oi * 9 + t];
        sum += dk[t];
    }
    if (g.w3x3)
        for (int t = 0; t < 9; ++t) g.w3x3[oi * 9 + t] = dk[t];
    if (g.w1x3)
        for (int w = 0; w < 3; ++w) g.w1x3[oi * 3 + w] = dk[3 + w];
    if (g.w3x1)
        for (int h = 0; h < 3; ++h) g.w3x1[oi * 3 + h] = dk[h * 3 + 1];
    if (g.w1x1) g.w1x1[oi] = dk[4];
    if (g.avg_w) g.avg_w[oi] = sum / 9.0f;
    for (int e = 0; e < 3; ++e)
        if (g.edge_k0[e]) {
            float acc = 0.0f;
            for (int t = 0; t < 9; ++t) acc += s.edge_mask[e][(size_t)o * 9 + t] * dk[t];
            g.edge_k0[e][oi] = s.edge_scale[e][o] * acc;
        }
}

// d seq_w2[o,m,t] = sum_i dK[o,i,t] W1[m,i]
ONR_HD float branch_bwd_w2_elem(const onr_branch_set& s, const float* dK, int o, int m, int t) {
    const int cin = s.cin;
    float acc = 0.0f;
    for (int i = 0; i < cin; ++i) acc += dK[((size_t)o * cin + i) * 9 + t] * s.seq_w1[(size_t)m * cin + i];
    return acc;
}

// partial of d seq_w1[m,i] = sum_{o,t} W2[o,m,t] dK[o,i,t] over idx = o*9+t in {start, start+stride, ...}
ONR_HD float branch_bwd_w1_partial(const onr_branch_set& s, const float* dK, int m, int i, int start, int stride) {
    const int cin = s.cin, cm = 2 * cin, n = s.cout * 9;
    float acc = 0.0f;
    for (int idx = start; idx < n; idx += stride) {
        const int o = idx / 9, t = idx - o * 9;
        acc += s.seq_w2[((size_t)o * cm + m) * 9 + t] * dK[((size_t)o * cin + i) * 9 + t];
    }
    return acc;
}

// partial of sum_i k0[o,i] sum_t mask[o,t] dK[o,i,t] over i in {start, start+stride, ...}: the dK-dependent part of
// d scale[o] of SeqConv3x3 branch e
ONR_HD float branch_bwd_scale_partial(const onr_branch_set& s, const float* dK, int e, int o, int start, int stride) {
    const int cin = s.cin;
    float acc = 0.0f;
    for (int i = start; i < cin; i += stride) {
        const size_t oi = (size_t)o * cin + i;
        float a = 0.0f;
        for (int t = 0; t < 9; ++t) a += s.edge_mask[e][(size_t)o * 9 + t] * dK[oi * 9 + t];
        acc += a * s.edge_k0[e][oi];
    }
    return acc;
}

// gradients indexed by the output channel alone: every bias, and scale / b0 / bias of the SeqConv3x3 branches.
// `scale_dot[e]` = the complete sum of branch_bwd_scale_partial over i.  OVERWRITES.
ONR_HD void branch_bwd_o(const onr_branch_set& s, const onr_branch_set& g, const float* db, int o,
                         const float* scale_dot) {
    const float d = db[o];
    if (g.b3x3) g.b3x3[o] = d;
    if (g.b1x3) g.b1x3[o] = d;
    if (g.b3x1) g.b3x1[o] = d;
    if (g.b1x1) g.b1x1[o] = d;
    for (int e = 0; e < 3; ++e)
        if (s.edge_k0[e]) {
            const float ms = edge_mask_sum(s.edge_mask[e], o);
            if (g.edge_bias[e]) g.edge_bias[e][o] = d;
            if (g.edge_b0[e]) g.edge_b0[e][o] = d * s.edge_scale[e][o] * ms;
            if (g.edge_scale[e]) g.edge_scale[e][o] = scale_dot[e] + d * s.edge_b0[e][o] * ms;
        }
}

}  // namespace onr
