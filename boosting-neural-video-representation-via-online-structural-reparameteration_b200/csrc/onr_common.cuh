// onr_common.cuh — error plumbing, launch helpers and TMA descriptor construction shared by all
// translation units of liborepnerv.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/orepnerv.h"

namespace onr {

void set_error(const char* fmt, ...);
void count_launch();   // every kernel launch of the library bumps a process-wide counter (onr_launch_count)

#define ONR_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            ::onr::set_error(__VA_ARGS__);    \
            return -1;                        \
        }                                     \
    } while (0)

#define ONR_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::onr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                             __FILE__, __LINE__);                                        \
            return static_cast<int>(_e);                                                 \
        }                                                                                \
    } while (0)

#define ONR_LAUNCH_CHECK()                                                               \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            ::onr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                        \
            return static_cast<int>(_e);                                                 \
        }                                                                                \
        ::onr::count_launch();                                                           \
    } while (0)

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int pad32(int c) { return (c + 31) / 32 * 32; }

// 5-D TMA view of an NHWC bf16 activation [B][H*s][W*s][Cp] as the pre-PixelShuffle tensor
// (jc = j*Cp + c  |  w  |  i  |  h  |  b); s = 1 gives the plain NHWC tensor.
// box = {32 channels (64 B, SWIZZLE_64B), box_w, 1, box_h, 1}.
// box = {inner channels (32 -> SWIZZLE_64B, 64 -> SWIZZLE_128B), box_w, 1, box_h, 1}.
int make_act_tmap(CUtensorMap* map, const void* ptr, int B, int H, int W, int Cp, int s, int box_w,
                  int box_h, int inner = 32);
// 3-D TMA view of packed weights [taps][rows][k] bf16; box = {inner, box_rows, 1}.
int make_weight_tmap(CUtensorMap* map, const void* ptr, int taps, int rows, int k, int box_rows, int inner = 32);

// 2-D TMA view of a K-major fp32 matrix [rows][k] with row pitch `pitch` floats (multiple of 4);
// box = {32 floats = 128 B (SWIZZLE_128B), box_rows}.  Reads past `k` / `rows` are zero-filled.
int make_f32_2d_tmap(CUtensorMap* map, const void* ptr, int rows, int k, int pitch, int box_rows);

// ----------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float silu_f(float z) { return z / (1.0f + __expf(-z)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// Packed fp32 FMA (sm_100 FFMA2): two IEEE fma.rn per instruction, so results are bit-identical to two scalar
// FFMAs at half the issue slots.  fma2_s broadcasts a scalar multiplicand (ptxas folds the {a,a} pair into the
// instruction's scalar-operand form, no extra move).
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi2(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ void fma2_s(uint64_t& acc, float a, uint64_t b) {
    const uint64_t aa = pack2(a, a);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(aa), "l"(b));
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

}  // namespace onr
