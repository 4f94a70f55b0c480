// act.cuh — the reference's activation table (model.py:86-117, ActivationLayer) and the derivatives PyTorch's autograd
// uses for it, as value/derivative pairs of the PRE-activation z.
//
//   code  --act        y(z)                                    dy/dz (as torch's backward formulas evaluate it)
//   0     swish        z * sigmoid(z)                          s + y (1 - s)
//   1     relu         max(z, 0)                               z > 0
//   2     leaky        z > 0 ? z : 0.01 z                      z > 0 ? 1 : 0.01
//   3     leaky01      z > 0 ? z : 0.1 z                       z > 0 ? 1 : 0.1
//   4     relu6        min(max(z, 0), 6)                       0 < z < 6
//   5     gelu         z/2 (1 + erf(z / sqrt 2))               (1 + erf(z / sqrt 2)) / 2 + z exp(-z^2/2) / sqrt(2 pi)
//   6     softplus     z > 20 ? z : log1p(exp(z))              z > 20 ? 1 : sigmoid(z)
//   7     hardswish    z relu6(z + 3) / 6                      z <= -3 ? 0 : (z < 3 ? z / 3 + 1/2 : 1)
//   8     sin          sin z                                   cos z
//
// swish is the north-star activation and stays fused in the tcgen05 epilogue (conv_igemm.cu); the others run through
// the pre-activation mode of that kernel (ONR_CONV_FPROP_Z) followed by onr_act_map (layout.cu).
#pragma once

#ifndef ONR_HD
#if defined(__CUDACC__)
#define ONR_HD __host__ __device__ __forceinline__
#else
#define ONR_HD inline
#endif
#endif

#include <math.h>

namespace onr {

constexpr int kActSwish = 0, kActCount = 9;

ONR_HD void act_value_grad(float z, int act, float* y, float* d) {
    switch (act) {
        default:
        case 0: {
            const float s = 1.0f / (1.0f + expf(-z));
            const float v = z * s;
            *y = v;
            *d = s + v * (1.0f - s);
            break;
        }
        case 1: *y = z > 0.0f ? z : 0.0f; *d = z > 0.0f ? 1.0f : 0.0f; break;
        case 2: *y = z > 0.0f ? z : 0.01f * z; *d = z > 0.0f ? 1.0f : 0.01f; break;
        case 3: *y = z > 0.0f ? z : 0.1f * z; *d = z > 0.0f ? 1.0f : 0.1f; break;
        case 4: *y = fminf(fmaxf(z, 0.0f), 6.0f); *d = (z > 0.0f && z < 6.0f) ? 1.0f : 0.0f; break;
        case 5: {
            const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
            *y = z * cdf;
            *d = cdf + z * expf(-0.5f * z * z) * 0.39894228040143267794f;
            break;
        }
        case 6: {
            const bool lin = z > 20.0f;
            *y = lin ? z : log1pf(expf(z));
            *d = lin ? 1.0f : 1.0f / (1.0f + expf(-z));
            break;
        }
        case 7: {
            *y = z * fminf(fmaxf(z + 3.0f, 0.0f), 6.0f) / 6.0f;
            *d = z <= -3.0f ? 0.0f : (z < 3.0f ? z / 3.0f + 0.5f : 1.0f);
            break;
        }
        case 8: *y = sinf(z); *d = cosf(z); break;
    }
}

}  // namespace onr
