// stem.cu — frame-index positional encoding + the two-layer stem MLP, forward and backward.
//
// Forward (reference utils.py:121-129 PositionalEncoding.forward; model.py:174-188 MLP;
// model.py:612-613 stem + view):
//   e[b,2i] = sin((t_b * f_i) * pi), e[b,2i+1] = cos(...)   with f_i = fp32(lbase**i) from the host
//   h1 = SiLU(W1 e + b1);  o = SiLU(W2 h1 + b2);  x0[b,h,w,c] = o[b, c*fh*fw + h*fw + w]  (NHWC bf16)
// Backward: rank-1 gW2, W2^T g for dh1, then the 80x512 layer.
// W2 (7.7 MB at fc 9_16_26, 33 MB at 9_16_112) is read exactly once per pass with 128-bit loads.
#include "onr_common.cuh"
#include "act.cuh"

namespace onr {

// `act`: activation code of act.cuh (reference MLP(act=...), model.py:174-188).  Code 0 (swish, the north-star
// configuration) keeps its original expressions bit for bit; the other codes go through act_value_grad.

__global__ void pe_layer1_kernel(const float* __restrict__ t_norm, const float* __restrict__ freqs, int levels,
                                 const float* __restrict__ W1, const float* __restrict__ b1, int hid,
                                 float* __restrict__ embed, float* __restrict__ pre1, float* __restrict__ h1,
                                 int act) {
    extern __shared__ float se[];  // [2*levels]
    const int b = blockIdx.x;
    const int E = 2 * levels;
    if (t_norm != nullptr) {
        const float t = t_norm[b];
        const float pi = 3.14159265358979323846f;
        for (int i = threadIdx.x; i < levels; i += blockDim.x) {
            const float v = __fmul_rn(__fmul_rn(t, freqs[i]), pi);   // same rounding order as the reference
            const float s = sinf(v), c = cosf(v);
            se[2 * i] = s;
            se[2 * i + 1] = c;
            if (blockIdx.y == 0) {
                embed[(size_t)b * E + 2 * i] = s;
                embed[(size_t)b * E + 2 * i + 1] = c;
            }
        }
    } else {
        // the caller already holds the embedding (Generator.forward(embed) of the reference API)
        for (int i = threadIdx.x; i < E; i += blockDim.x) se[i] = embed[(size_t)b * E + i];
    }
    __syncthreads();
    // the hidden units are spread over gridDim.y blocks (every block recomputes the 2*levels encoding values, which is
    // cheaper than one SM pulling all of W1 through its own load path); the per-unit arithmetic order is unchanged
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < hid; j += gridDim.y * blockDim.x) {
        const float* w = W1 + (size_t)j * E;
        float acc = 0.0f;
        for (int e = 0; e < E; ++e) acc = fmaf(w[e], se[e], acc);
        acc += b1[j];
        pre1[(size_t)b * hid + j] = acc;
        float hv = acc / (1.0f + expf(-acc));
        if (act != 0) {
            float dd;
            act_value_grad(acc, act, &hv, &dd);
        }
        h1[(size_t)b * hid + j] = hv;
    }
}

__global__ void pos_encoding_kernel(const float* __restrict__ t_norm, int B, const float* __restrict__ freqs,
                                    int levels, float* __restrict__ embed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * levels) return;
    const int b = idx / levels, i = idx % levels;
    const float v = __fmul_rn(__fmul_rn(t_norm[b], freqs[i]), 3.14159265358979323846f);
    embed[(size_t)b * 2 * levels + 2 * i] = sinf(v);
    embed[(size_t)b * 2 * levels + 2 * i + 1] = cosf(v);
}

// one warp per output position p = (h*fw + w)*Cp + c of the NHWC stem output
__global__ void stem_layer2_kernel(const float* __restrict__ h1, int B, int hid, const float* __restrict__ W2,
                                   const float* __restrict__ b2, int fc_dim, int fh, int fw, int Cp,
                                   __nv_bfloat16* __restrict__ x0, __nv_bfloat16* __restrict__ dstem, int act) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int total = fh * fw * Cp;
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < total; p += gridDim.x * warps_per_block) {
        const int c = p % Cp;
        const int hw = p / Cp;
        if (c >= fc_dim) {
            if (lane == 0)
                for (int b = 0; b < B; ++b) {
                    x0[(size_t)b * total + p] = __float2bfloat16(0.0f);
                    dstem[(size_t)b * total + p] = __float2bfloat16(0.0f);
                }
            continue;
        }
        const int o = c * fh * fw + hw;
        const float4* wrow = reinterpret_cast<const float4*>(W2 + (size_t)o * hid);
        for (int b = 0; b < B; ++b) {
            const float4* hv = reinterpret_cast<const float4*>(h1 + (size_t)b * hid);
            float acc = 0.0f;
            for (int j = lane; j < hid / 4; j += 32) {
                const float4 w = __ldg(wrow + j), h = hv[j];
                acc = fmaf(w.x, h.x, acc);
                acc = fmaf(w.y, h.y, acc);
                acc = fmaf(w.z, h.z, acc);
                acc = fmaf(w.w, h.w, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                const float z = acc + b2[o];
                const float sg = 1.0f / (1.0f + expf(-z));
                float y = z * sg;
                float dy = sg + y * (1.0f - sg);
                if (act != 0) act_value_grad(z, act, &y, &dy);
                x0[(size_t)b * total + p] = __float2bfloat16(y);
                dstem[(size_t)b * total + p] = __float2bfloat16(dy);
            }
        }
    }
}

// Each block owns a chunk of output rows o; thread t owns hidden units j = t, t+blockDim, ...
// gW2[o,j] += g*h1[j];  gb2[o] += g;  dh1[b,j] += W2[o,j]*g  (block-partial, then atomics).
constexpr int kStemRows = 16;
constexpr int kStemMaxJ = 8;
__global__ void stem_bwd_layer2_kernel(const __nv_bfloat16* __restrict__ g0, int B, const float* __restrict__ h1,
                                       int hid, const float* __restrict__ W2, int fc_dim, int fh, int fw, int Cp,
                                       float* __restrict__ gW2, float* __restrict__ gb2,
                                       float* __restrict__ dh1) {
    const int b = blockIdx.y;
    const int n_out = fc_dim * fh * fw;
    const int o_begin = blockIdx.x * kStemRows;
    const int o_end = min(n_out, o_begin + kStemRows);
    float hreg[kStemMaxJ], acc[kStemMaxJ];
#pragma unroll
    for (int u = 0; u < kStemMaxJ; ++u) {
        const int j = threadIdx.x + u * blockDim.x;
        hreg[u] = j < hid ? h1[(size_t)b * hid + j] : 0.0f;
        acc[u] = 0.0f;
    }
    for (int o = o_begin; o < o_end; ++o) {
        const int c = o / (fh * fw), hw = o % (fh * fw);
        const float g = __bfloat162float(g0[((size_t)b * fh * fw + hw) * Cp + c]);
        if (threadIdx.x == 0) atomicAdd(&gb2[o], g);
#pragma unroll
        for (int u = 0; u < kStemMaxJ; ++u) {
            const int j = threadIdx.x + u * blockDim.x;
            if (j < hid) {
                const size_t idx = (size_t)o * hid + j;
                acc[u] = fmaf(__ldg(W2 + idx), g, acc[u]);
                if (B == 1) gW2[idx] += g * hreg[u];
                else atomicAdd(&gW2[idx], g * hreg[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kStemMaxJ; ++u) {
        const int j = threadIdx.x + u * blockDim.x;
        if (j < hid) atomicAdd(&dh1[(size_t)b * hid + j], acc[u]);
    }
}

// Batch-1 fast path (the training configuration): hid % 4 == 0, gradients are plain read-modify-writes.  Thread t owns
// the four hidden units 4t..4t+3 (+512 per extra pass); a block owns kStemFastRows output rows, all of whose W2 /
// gW2 rows are in flight at once (the kernel moves 3 x n_out x hid floats and nothing else of size).
constexpr int kStemFastRows = 8;
constexpr int kStemFastU = 2;   // hid <= 128 * 4 * kStemFastU
__global__ void __launch_bounds__(128)
stem_bwd_layer2_b1_kernel(const __nv_bfloat16* __restrict__ g0, const float* __restrict__ h1, int hid,
                          const float* __restrict__ W2, int fc_dim, int fh, int fw, int Cp,
                          float* __restrict__ gW2, float* __restrict__ gb2, float* __restrict__ dh1) {
    const int n_out = fc_dim * fh * fw, HW = fh * fw;
    const int o_begin = blockIdx.x * kStemFastRows;
    float g[kStemFastRows];
#pragma unroll
    for (int r = 0; r < kStemFastRows; ++r) {
        const int o = o_begin + r;
        g[r] = 0.0f;
        if (o < n_out) {
            const int c = o / HW, hw = o - c * HW;
            g[r] = __bfloat162float(g0[(size_t)hw * Cp + c]);
        }
    }
    if (threadIdx.x < kStemFastRows && o_begin + threadIdx.x < n_out) gb2[o_begin + threadIdx.x] += g[threadIdx.x];
#pragma unroll
    for (int u = 0; u < kStemFastU; ++u) {
        const int j = (threadIdx.x + u * 128) * 4;
        if (j >= hid) break;
        const float4 h = *reinterpret_cast<const float4*>(h1 + j);
        float4 w[kStemFastRows], gwv[kStemFastRows];
#pragma unroll
        for (int r = 0; r < kStemFastRows; ++r) {
            const size_t idx = (size_t)min(o_begin + r, n_out - 1) * hid + j;
            w[r] = __ldg(reinterpret_cast<const float4*>(W2 + idx));
            gwv[r] = *reinterpret_cast<const float4*>(gW2 + idx);
        }
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int r = 0; r < kStemFastRows; ++r) {
            if (o_begin + r >= n_out) break;
            acc.x = fmaf(w[r].x, g[r], acc.x); acc.y = fmaf(w[r].y, g[r], acc.y);
            acc.z = fmaf(w[r].z, g[r], acc.z); acc.w = fmaf(w[r].w, g[r], acc.w);
            gwv[r].x = fmaf(g[r], h.x, gwv[r].x); gwv[r].y = fmaf(g[r], h.y, gwv[r].y);
            gwv[r].z = fmaf(g[r], h.z, gwv[r].z); gwv[r].w = fmaf(g[r], h.w, gwv[r].w);
            *reinterpret_cast<float4*>(gW2 + (size_t)(o_begin + r) * hid + j) = gwv[r];
        }
        atomicAdd(&dh1[j], acc.x); atomicAdd(&dh1[j + 1], acc.y);
        atomicAdd(&dh1[j + 2], acc.z); atomicAdd(&dh1[j + 3], acc.w);
    }
}

__global__ void stem_bwd_layer1_kernel(const float* __restrict__ dh1, const float* __restrict__ pre1,
                                       const float* __restrict__ embed, int B, int hid, int E,
                                       float* __restrict__ gW1, float* __restrict__ gb1, int act) {
    // one thread per (j, e) pair plus bias; sums over the batch
    const int total = hid * E;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / E, e = idx % E;
        float gw = 0.0f, gb = 0.0f;
        for (int b = 0; b < B; ++b) {
            const float z = pre1[(size_t)b * hid + j];
            const float sg = 1.0f / (1.0f + expf(-z));
            float da = sg + z * sg * (1.0f - sg);
            if (act != 0) {
                float yy;
                act_value_grad(z, act, &yy, &da);
            }
            const float dpre = dh1[(size_t)b * hid + j] * da;
            gw = fmaf(dpre, embed[(size_t)b * E + e], gw);
            gb += dpre;
        }
        gW1[idx] += gw;
        if (e == 0) gb1[j] += gb;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Data-parallel exchange of the stem gradients in FACTORED form.  For one frame gW2 = g (x) h1 is a rank-1 matrix
// (7.7 MB at S720, 33 MB at L720): all-reducing it is by far the largest and — coming last in the backward — the only
// fully exposed piece of the gradient exchange.  Instead every rank publishes its factors
//     slot = [ g[n_out] | h1[hid] | dpre1[hid] | embed[E] ]          (fp32, ~21 KB at S720)
// the slots are all-gathered (latency-bound), and every rank forms the SUMS over ranks itself, in rank order, so all
// replicas obtain bit-identical gradients:  gW2 = sum_r g_r (x) h1_r,  gb2 = sum_r g_r,
// gW1 = sum_r dpre1_r (x) embed_r,  gb1 = sum_r dpre1_r.
__global__ void __launch_bounds__(128)
stem_factors_local_kernel(const __nv_bfloat16* __restrict__ g0, const float* __restrict__ W2, int hid, int fc_dim,
                          int fh, int fw, int Cp, float* __restrict__ slot_g, float* __restrict__ dh1) {
    // dh1[j] += sum_o W2[o][j] g[o] over this block's kStemFastRows rows; also converts g to fp32 into the slot
    const int n_out = fc_dim * fh * fw, HW = fh * fw;
    const int o_begin = blockIdx.x * kStemFastRows;
    float g[kStemFastRows];
#pragma unroll
    for (int r = 0; r < kStemFastRows; ++r) {
        const int o = o_begin + r;
        g[r] = 0.0f;
        if (o < n_out) {
            const int c = o / HW, hw = o - c * HW;
            g[r] = __bfloat162float(g0[(size_t)hw * Cp + c]);
        }
    }
    if (threadIdx.x < kStemFastRows && o_begin + threadIdx.x < n_out) slot_g[o_begin + threadIdx.x] = g[threadIdx.x];
#pragma unroll
    for (int u = 0; u < kStemFastU; ++u) {
        const int j = (threadIdx.x + u * 128) * 4;
        if (j >= hid) break;
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int r = 0; r < kStemFastRows; ++r) {
            if (o_begin + r >= n_out) break;
            const float4 w = __ldg(reinterpret_cast<const float4*>(W2 + (size_t)(o_begin + r) * hid + j));
            acc.x = fmaf(w.x, g[r], acc.x); acc.y = fmaf(w.y, g[r], acc.y);
            acc.z = fmaf(w.z, g[r], acc.z); acc.w = fmaf(w.w, g[r], acc.w);
        }
        atomicAdd(&dh1[j], acc.x); atomicAdd(&dh1[j + 1], acc.y);
        atomicAdd(&dh1[j + 2], acc.z); atomicAdd(&dh1[j + 3], acc.w);
    }
}

// slot tail: h1, dpre1 = dh1 * SiLU'(pre1), embed
__global__ void stem_factors_tail_kernel(const float* __restrict__ h1, const float* __restrict__ dh1,
                                         const float* __restrict__ pre1, const float* __restrict__ embed, int hid, int E,
                                         float* __restrict__ slot_h1, float* __restrict__ slot_dpre1,
                                         float* __restrict__ slot_embed, int act) {
    const int n = hid > E ? hid : E;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        if (j < hid) {
            const float z = pre1[j];
            const float sg = 1.0f / (1.0f + expf(-z));
            float da = sg + z * sg * (1.0f - sg);
            if (act != 0) {
                float yy;
                act_value_grad(z, act, &yy, &da);
            }
            slot_h1[j] = h1[j];
            slot_dpre1[j] = dh1[j] * da;
        }
        if (j < E) slot_embed[j] = embed[j];
    }
}

// gW2[o][j] = sum_r g_r[o] h1_r[j]  (j in float4 pieces), gb2[o] = sum_r g_r[o];  OVERWRITES
__global__ void __launch_bounds__(256)
stem_from_factors_w2_kernel(const float* __restrict__ slots, int K, size_t stride, int n_out, int hid,
                            float* __restrict__ gW2, float* __restrict__ gb2) {
    const int q4 = hid / 4;
    const size_t total = (size_t)n_out * q4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx / q4), j = (int)(idx % q4) * 4;
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float gsum = 0.0f;
        for (int r = 0; r < K; ++r) {
            const float* sl = slots + (size_t)r * stride;
            const float g = sl[o];
            const float4 h = *reinterpret_cast<const float4*>(sl + n_out + j);
            acc.x = fmaf(g, h.x, acc.x); acc.y = fmaf(g, h.y, acc.y);
            acc.z = fmaf(g, h.z, acc.z); acc.w = fmaf(g, h.w, acc.w);
            gsum += g;
        }
        *reinterpret_cast<float4*>(gW2 + (size_t)o * hid + j) = acc;
        if (j == 0) gb2[o] = gsum;
    }
}

// gW1[j][e] = sum_r dpre1_r[j] embed_r[e], gb1[j] = sum_r dpre1_r[j];  OVERWRITES
__global__ void stem_from_factors_w1_kernel(const float* __restrict__ slots, int K, size_t stride, int n_out, int hid,
                                            int E, float* __restrict__ gW1, float* __restrict__ gb1) {
    const int total = hid * E;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / E, e = idx % E;
        float gw = 0.0f, gb = 0.0f;
        for (int r = 0; r < K; ++r) {
            const float* sl = slots + (size_t)r * stride;
            const float dpre = sl[n_out + hid + j];
            gw = fmaf(dpre, sl[n_out + 2 * hid + e], gw);
            gb += dpre;
        }
        gW1[idx] = gw;
        if (e == 0) gb1[j] = gb;
    }
}

}  // namespace onr

extern "C" {

size_t onr_stem_factor_floats(int fc_dim, int fh, int fw, int hid, int emb_len) {
    const size_t n = (size_t)fc_dim * fh * fw + 2 * (size_t)hid + emb_len;
    return (n + 63) / 64 * 64;      // slots stay 256-byte aligned
}

int onr_stem_bwd_factors(const void* g0, const float* embed, int emb_len, const float* pre1, const float* h1, int hid,
                         const float* W2, int fc_dim, int fh, int fw, int Cp, float* slot, float* scratch_dh1,
                         void* stream) {
    return onr_stem_bwd_factors_act(g0, embed, emb_len, pre1, h1, hid, W2, fc_dim, fh, fw, Cp, slot, scratch_dh1, 0,
                                    stream);
}

int onr_stem_bwd_factors_act(const void* g0, const float* embed, int emb_len, const float* pre1, const float* h1,
                             int hid, const float* W2, int fc_dim, int fh, int fw, int Cp, float* slot,
                             float* scratch_dh1, int act, void* stream) {
    using namespace onr;
    ONR_REQUIRE(act >= 0 && act < kActCount, "stem: unknown activation code %d", act);
    ONR_REQUIRE(hid % 4 == 0 && hid <= 128 * 4 * kStemFastU, "stem factors: unsupported widths");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_out = fc_dim * fh * fw;
    ONR_CUDA(cudaMemsetAsync(scratch_dh1, 0, (size_t)hid * sizeof(float), st));
    stem_factors_local_kernel<<<ceil_div(n_out, kStemFastRows), 128, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(g0), W2, hid, fc_dim, fh, fw, Cp, slot, scratch_dh1);
    ONR_LAUNCH_CHECK();
    stem_factors_tail_kernel<<<ceil_div(hid > emb_len ? hid : emb_len, 256), 256, 0, st>>>(h1, scratch_dh1, pre1, embed, hid, emb_len,
                                                                 slot + n_out, slot + n_out + hid, slot + n_out + 2 * hid, act);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_stem_grads_from_factors(const float* slots, int K, int emb_len, int hid, int fc_dim, int fh, int fw,
                                float* gW1, float* gb1, float* gW2, float* gb2, void* stream) {
    using namespace onr;
    ONR_REQUIRE(K >= 1 && hid % 4 == 0, "stem factors: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_out = fc_dim * fh * fw;
    const size_t stride = onr_stem_factor_floats(fc_dim, fh, fw, hid, emb_len);
    size_t blocks = ((size_t)n_out * (hid / 4) + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    stem_from_factors_w2_kernel<<<(int)blocks, 256, 0, st>>>(slots, K, stride, n_out, hid, gW2, gb2);
    ONR_LAUNCH_CHECK();
    stem_from_factors_w1_kernel<<<ceil_div(hid * emb_len, 256), 256, 0, st>>>(slots, K, stride, n_out, hid, emb_len, gW1,
                                                                             gb1);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_pe_stem_fwd(const float* t_norm, int B, const float* freqs, int levels, const float* W1,
                    const float* b1, int hid, const float* W2, const float* b2, int fc_dim, int fh, int fw,
                    int Cp, float* embed, float* pre1, float* h1, void* x0, void* dstem, void* stream) {
    return onr_pe_stem_fwd_act(t_norm, B, freqs, levels, W1, b1, hid, W2, b2, fc_dim, fh, fw, Cp, embed, pre1, h1, x0,
                               dstem, 0, stream);
}

int onr_pe_stem_fwd_act(const float* t_norm, int B, const float* freqs, int levels, const float* W1,
                        const float* b1, int hid, const float* W2, const float* b2, int fc_dim, int fh, int fw,
                        int Cp, float* embed, float* pre1, float* h1, void* x0, void* dstem, int act, void* stream) {
    using namespace onr;
    ONR_REQUIRE(act >= 0 && act < kActCount, "stem: unknown activation code %d", act);
    ONR_REQUIRE(B >= 1 && levels >= 1 && hid % 4 == 0 && Cp % 32 == 0 && Cp >= fc_dim, "stem: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    pe_layer1_kernel<<<dim3(B, ceil_div(hid, 64)), 64, 2 * levels * sizeof(float), st>>>(t_norm, freqs, levels, W1, b1,
                                                                                         hid, embed, pre1, h1, act);
    ONR_LAUNCH_CHECK();
    const int total = fh * fw * Cp;
    const int wpb = 8;
    int grid = ceil_div(total, wpb);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    stem_layer2_kernel<<<grid, wpb * 32, 0, st>>>(h1, B, hid, W2, b2, fc_dim, fh, fw, Cp,
                                                  reinterpret_cast<__nv_bfloat16*>(x0),
                                                  reinterpret_cast<__nv_bfloat16*>(dstem), act);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_pos_encoding(const float* t_norm, int B, const float* freqs, int levels, float* embed, void* stream) {
    using namespace onr;
    ONR_REQUIRE(B >= 1 && levels >= 1, "pos_encoding: bad shape");
    pos_encoding_kernel<<<ceil_div(B * levels, 128), 128, 0, (cudaStream_t)stream>>>(t_norm, B, freqs, levels,
                                                                                    embed);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_stem_bwd(const void* g0, int B, const float* embed, int emb_len, const float* pre1, const float* h1,
                 int hid, const float* W2, int fc_dim, int fh, int fw, int Cp, float* gW1, float* gb1,
                 float* gW2, float* gb2, float* scratch_dh1, void* stream) {
    return onr_stem_bwd_act(g0, B, embed, emb_len, pre1, h1, hid, W2, fc_dim, fh, fw, Cp, gW1, gb1, gW2, gb2,
                            scratch_dh1, 0, stream);
}

int onr_stem_bwd_act(const void* g0, int B, const float* embed, int emb_len, const float* pre1, const float* h1,
                     int hid, const float* W2, int fc_dim, int fh, int fw, int Cp, float* gW1, float* gb1,
                     float* gW2, float* gb2, float* scratch_dh1, int act, void* stream) {
    using namespace onr;
    ONR_REQUIRE(act >= 0 && act < kActCount, "stem: unknown activation code %d", act);
    ONR_REQUIRE(hid <= 128 * kStemMaxJ, "stem: hidden width %d too large", hid);
    cudaStream_t st = (cudaStream_t)stream;
    ONR_CUDA(cudaMemsetAsync(scratch_dh1, 0, (size_t)B * hid * sizeof(float), st));
    const int n_out = fc_dim * fh * fw;
    if (B == 1 && hid % 4 == 0 && hid <= 128 * 4 * kStemFastU) {
        stem_bwd_layer2_b1_kernel<<<ceil_div(n_out, kStemFastRows), 128, 0, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(g0), h1, hid, W2, fc_dim, fh, fw, Cp, gW2, gb2, scratch_dh1);
    } else {
        dim3 grid(ceil_div(n_out, kStemRows), B);
        stem_bwd_layer2_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g0), B, h1, hid, W2,
                                                     fc_dim, fh, fw, Cp, gW2, gb2, scratch_dh1);
    }
    ONR_LAUNCH_CHECK();
    stem_bwd_layer1_kernel<<<ceil_div(hid * emb_len, 256), 256, 0, st>>>(scratch_dh1, pre1, embed, B, hid,
                                                                        emb_len, gW1, gb1, act);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
