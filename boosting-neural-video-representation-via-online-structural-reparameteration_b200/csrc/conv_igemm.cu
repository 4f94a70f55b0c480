// conv_igemm.cu — 3x3 / stride 1 / pad 1 convolution as a tcgen05 implicit GEMM for sm_100a.
//
// Replaces F.conv2d + nn.PixelShuffle + nn.SiLU of NeRVBlock.forward (reference model.py:539, :523,
// :520 and :567) and, with mirrored taps and the un-shuffled dZ view as the A operand, the data
// gradient that autograd derives for it (main_train.py:249).
//
//   D[pixel, n] = sum_{tap, k} A[pixel + off(tap), k] * Wt[tap][n][k]
//
// The first version of this kernel (profiles/r01_ncu_full_*_v1.txt) was bound by TMA load traffic out of
// L2 (~5.5 TB/s), not by the tensor pipe: every tap re-loaded its own shifted A tile and every 128-pixel
// tile re-streamed the whole weight matrix.  This version spends far fewer operand bytes per MAC:
//
// * A sub-tile is 16 rows x 8 columns of pixels.  For a horizontal tap dw the producer loads ONE TMA box of
//   (16*MS + 2) rows x 8 columns x 32 channels (halo rows included, out-of-image pixels zero-filled by the
//   TMA unit).  Because 8 pixels x 64 B = 512 B is exactly the SWIZZLE_64B repeat, the three vertical taps
//   dh = -1,0,+1 are the SAME smem box read at start offsets (1+dh)*512 B — 3x fewer A bytes, no im2col.
// * MS sub-tiles (stacked vertically, MS*block_n <= 512 TMEM columns) share every weight tile: the weights
//   are streamed once per MS*128 pixels instead of once per 128.
// * A boxes and weight tiles travel through two independent mbarrier rings fed by two producer warps.
// * warp 0 = A producer, warp 1 = UMMA issuer (one thread), warp 2 = weight producer, warps 3..6 = epilogue
//   (tcgen05.ld -> bias/SiLU/SiLU' or dgrad scaling -> bf16 -> swizzled smem -> TMA store through the
//   PixelShuffle view, so the shuffle is pure addressing).
// * persistent: grid = min(tiles, #SM); TMEM double-buffered when 2*MS*block_n <= 512.
#include "onr_common.cuh"
#include "onr_ptx.cuh"

#include <stdlib.h>

namespace onr {

constexpr int kSubH = 16;                          // sub-tile: 16 x 8 pixels = 128 accumulator rows
constexpr int kSubW = 8;
constexpr int kChunkK = 32;                        // bf16 elements per K step (64 bytes)
constexpr int kRowBytes = kSubW * 64;              // one image row of a box = 512 B = SWIZZLE_64B repeat
constexpr int kStageOutBytes = 128 * 64;           // one 128 x 32 bf16 staging tile
constexpr int kThreads = 224;
constexpr int kMaxBlockN = 256;
constexpr int kMaxMS = 4;
constexpr int kMaxDynSmem = 232448;                // 227 KB: per-block opt-in maximum on sm_100
constexpr int kMaxRing = 8;

struct ConvParams {
    int H, W, B;
    int tiles_w, tiles_h, n_tiles, total_tiles;
    int block_n, ms;
    int chunks, jc_chunks, sign;
    int n_total, acc_bufs;
    int na, nb;              // ring depths
    int a_bytes, b_bytes;    // slot sizes
    int mode;
    int out_jc;              // channels per shuffle row i of the output view (out_s * out_cp)
    const float* bias;
    const __nv_bfloat16* dmul;
    // optional per-CTA cycle counters (selftest "prof" mode): [cta][8] =
    // {total, mma wait A, mma wait B, mma wait tmem_empty, epi wait tmem_full, epi wait store, epi busy, tiles}
    long long* prof;
};

__device__ __forceinline__ long long clk() { return clock64(); }

struct __align__(8) SmemBarriers {
    uint64_t a_full[kMaxRing], a_empty[kMaxRing];
    uint64_t b_full[kMaxRing], b_empty[kMaxRing];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct TileCoord {
    int b, h0, w0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int tile) {
    TileCoord t;
    const int nt = tile % p.n_tiles;
    int mt = tile / p.n_tiles;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    t.b = mt / p.tiles_h;
    t.h0 = th * kSubH * p.ms;
    t.w0 = tw * kSubW;
    t.n0 = nt * p.block_n;
    return t;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmD,
                  const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [na x A slot] [nb x B slot] [staging 2 x 2 x 8 KB] [barriers]
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_ring = smem_base;
    const uint32_t b_ring = a_ring + p.na * p.a_bytes;
    const uint32_t staging = b_ring + p.nb * p.b_bytes;
    SmemBarriers* bars =
        reinterpret_cast<SmemBarriers*>(smem_raw + (staging + 4 * kStageOutBytes - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(smem_u32(&bars->a_full[s]), 1);
            mbar_init(smem_u32(&bars->a_empty[s]), 1);
            mbar_init(smem_u32(&bars->b_full[s]), 1);
            mbar_init(smem_u32(&bars->b_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->tmem_full[b]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[b]), 128);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmY);
        if (p.mode == ONR_CONV_FPROP_TRAIN) tma_prefetch_desc(&tmD);
    }
    if (warp == 2 && lane == 0) tma_prefetch_desc(&tmB);
    if (warp == 1) {
        tmem_alloc(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================================================================== A producer
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord t = tile_coord(p, tile);
                for (int ch = 0; ch < p.chunks; ++ch) {
                    const int ii = ch / p.jc_chunks;
                    const int jc0 = (ch - ii * p.jc_chunks) * kChunkK;
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        mbar_wait(smem_u32(&bars->a_empty[slot]), phase ^ 1);
                        const uint32_t full = smem_u32(&bars->a_full[slot]);
                        mbar_expect_tx(full, p.a_bytes);
                        tma_load_5d(a_ring + slot * p.a_bytes, &tmA, full, jc0, t.w0 + dwi - 1, ii, t.h0 - 1, t.b);
                        if (++slot == (uint32_t)p.na) { slot = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================================== weight producer
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord t = tile_coord(p, tile);
                for (int ch = 0; ch < p.chunks; ++ch) {
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        const int kw = 1 + (dwi - 1) * p.sign;
                        for (int dhi = 0; dhi < 3; ++dhi) {
                            const int kh = 1 + (dhi - 1) * p.sign;
                            mbar_wait(smem_u32(&bars->b_empty[slot]), phase ^ 1);
                            const uint32_t full = smem_u32(&bars->b_full[slot]);
                            mbar_expect_tx(full, p.b_bytes);
                            tma_load_3d(b_ring + slot * p.b_bytes, &tmB, full, ch * kChunkK, t.n0, kh * 3 + kw);
                            if (++slot == (uint32_t)p.nb) { slot = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== UMMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.block_n, 0, 0);
            uint32_t aslot = 0, aphase = 0, bslot = 0, bphase = 0;
            int it = 0;
            const bool prof = p.prof != nullptr;
            long long t_start = prof ? clk() : 0, w_a = 0, w_b = 0, w_t = 0, t0 = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int buf = it % p.acc_bufs;
                const uint32_t acc_phase = (uint32_t)(it / p.acc_bufs) & 1u;
                if (prof) t0 = clk();
                mbar_wait(smem_u32(&bars->tmem_empty[buf]), acc_phase ^ 1);
                if (prof) w_t += clk() - t0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * p.ms * p.block_n;
                uint32_t first = 1;
                for (int ch = 0; ch < p.chunks; ++ch) {
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        if (prof) t0 = clk();
                        mbar_wait(smem_u32(&bars->a_full[aslot]), aphase);
                        if (prof) w_a += clk() - t0;
                        const uint32_t a_s = a_ring + aslot * p.a_bytes;
                        for (int dhi = 0; dhi < 3; ++dhi) {
                            if (prof) t0 = clk();
                            mbar_wait(smem_u32(&bars->b_full[bslot]), bphase);
                            if (prof) w_b += clk() - t0;
                            tc_fence_after();
                            const uint32_t b_s = b_ring + bslot * p.b_bytes;
#pragma unroll
                            for (int k = 0; k < kChunkK / 16; ++k) {
                                const uint64_t bdesc = make_smem_desc(b_s + k * 32, 16, 512, SWZ_64B);
                                for (int ms = 0; ms < p.ms; ++ms) {
                                    // rows of sub-tile ms for vertical tap dh: box rows (dhi + 16*ms) ...
                                    const uint64_t adesc =
                                        make_smem_desc(a_s + (dhi + kSubH * ms) * kRowBytes + k * 32, 16, 512, SWZ_64B);
                                    umma_bf16(d_tmem + ms * p.block_n, adesc, bdesc, idesc, (first && k == 0) ? 0u : 1u);
                                }
                            }
                            first = 0;
                            umma_commit(smem_u32(&bars->b_empty[bslot]));
                            if (++bslot == (uint32_t)p.nb) { bslot = 0; bphase ^= 1; }
                        }
                        umma_commit(smem_u32(&bars->a_empty[aslot]));
                        if (++aslot == (uint32_t)p.na) { aslot = 0; aphase ^= 1; }
                    }
                }
                umma_commit(smem_u32(&bars->tmem_full[buf]));
            }
            if (prof) {
                long long* o = p.prof + (size_t)blockIdx.x * 8;
                o[0] = clk() - t_start; o[1] = w_a; o[2] = w_b; o[3] = w_t; o[7] = it;
            }
        }
    } else {
        // ===================================================================== epilogue (warps 3..6)
        const int q = warp & 3;             // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;      // accumulator row == pixel inside the 16 x 8 sub-tile
        const int hl = row >> 3, wl = row & 7;
        const bool store_thread = (warp == 3 && lane == 0);
        const uint32_t swz = (uint32_t)((row >> 1) & 3);
        uint32_t chunk_ctr = 0;
        int it = 0;
        const bool prof = p.prof != nullptr && store_thread;
        long long e_full = 0, e_store = 0, e_busy = 0, t0 = 0, t1 = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const TileCoord t = tile_coord(p, tile);
            const int buf = it % p.acc_bufs;
            const uint32_t acc_phase = (uint32_t)(it / p.acc_bufs) & 1u;
            if (prof) t0 = clk();
            mbar_wait(smem_u32(&bars->tmem_full[buf]), acc_phase);
            if (prof) { t1 = clk(); e_full += t1 - t0; }
            tc_fence_after();
            const int nchunks = p.block_n / 32;
            for (int ms = 0; ms < p.ms; ++ms) {
                const int hs0 = t.h0 + ms * kSubH;
                if (hs0 >= p.H) break;      // sub-tile entirely below the image (uniform across the CTA)
                for (int c = 0; c < nchunks; ++c) {
                    const int n = t.n0 + c * 32;
                    if (n >= p.n_total) break;
                    const uint32_t sbuf = chunk_ctr & 1u;
                    ++chunk_ctr;
                    if (prof) t0 = clk();
                    if (store_thread) tma_store_wait_read<1>();
                    if (prof) e_store += clk() - t0;
                    named_bar_sync(1, 128);
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * p.ms + ms) * p.block_n + c * 32, r);
                    tmem_ld_wait();
                    const uint32_t ybuf = staging + sbuf * 2 * kStageOutBytes;
                    const uint32_t dbuf = ybuf + kStageOutBytes;
                    if (p.mode == ONR_CONV_DGRAD) {
                        const int h = hs0 + hl, w = t.w0 + wl;
                        uint4 dv[4];
                        if (h < p.H && w < p.W) {
                            const uint4* dp = reinterpret_cast<const uint4*>(
                                p.dmul + ((size_t)(t.b * p.H + h) * p.W + w) * p.n_total + n);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dv[j] = __ldg(dp + j);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) dv[j] = make_uint4(0, 0, 0, 0);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t dw4[4] = {dv[j].x, dv[j].y, dv[j].z, dv[j].w};
                            uint32_t o[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float a0 = __uint_as_float(r[j * 8 + e * 2]) * bf16_lo(dw4[e]);
                                const float a1 = __uint_as_float(r[j * 8 + e * 2 + 1]) * bf16_hi(dw4[e]);
                                o[e] = pack_bf16x2(a0, a1);
                            }
                            const uint32_t addr = ybuf + row * 64 + ((j ^ swz) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(o[0]),
                                         "r"(o[1]), "r"(o[2]), "r"(o[3])
                                         : "memory");
                        }
                    } else {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b0 = __ldg(bp + j * 2), b1 = __ldg(bp + j * 2 + 1);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            float yv[8], dv[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float z = __uint_as_float(r[j * 8 + e]) + bb[e];
                                const float sg = fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);
                                const float y = z * sg;
                                yv[e] = y;
                                dv[e] = fmaf(y, 1.0f - sg, sg);
                            }
                            const uint32_t off = row * 64 + ((j ^ swz) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(ybuf + off),
                                         "r"(pack_bf16x2(yv[0], yv[1])), "r"(pack_bf16x2(yv[2], yv[3])),
                                         "r"(pack_bf16x2(yv[4], yv[5])), "r"(pack_bf16x2(yv[6], yv[7]))
                                         : "memory");
                            if (p.mode == ONR_CONV_FPROP_TRAIN)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dbuf + off),
                                             "r"(pack_bf16x2(dv[0], dv[1])), "r"(pack_bf16x2(dv[2], dv[3])),
                                             "r"(pack_bf16x2(dv[4], dv[5])), "r"(pack_bf16x2(dv[6], dv[7]))
                                             : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (store_thread) {
                        const int oi = n / p.out_jc;
                        const int ojc = n - oi * p.out_jc;
                        tma_store_5d(&tmY, ybuf, ojc, t.w0, oi, hs0, t.b);
                        if (p.mode == ONR_CONV_FPROP_TRAIN) tma_store_5d(&tmD, dbuf, ojc, t.w0, oi, hs0, t.b);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
            if (prof) e_busy += clk() - t1;
        }
        if (store_thread) tma_store_wait_all<0>();
        if (prof) {
            long long* o = p.prof + (size_t)blockIdx.x * 8;
            o[4] = e_full; o[5] = e_store; o[6] = e_busy;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace onr

// ------------------------------------------------------------------------------------------- host
struct onr_conv_plan {
    CUtensorMap tmA, tmB, tmY, tmD;
    onr::ConvParams p;
    int grid;
    size_t smem;
};

extern "C" {

// N tiling rule shared with the weight packer: block_n <= 256 (one UMMA per K slice), multiple of 32,
// chosen to minimise zero padding.
int onr_conv_tile_n(int n_total, int* block_n, int* n_tiles) {
    ONR_REQUIRE(n_total > 0 && n_total % 32 == 0, "n_total must be a positive multiple of 32");
    const int nt0 = (n_total + onr::kMaxBlockN - 1) / onr::kMaxBlockN;
    int best_nt = nt0, best_bn = ((n_total + nt0 - 1) / nt0 + 31) / 32 * 32;
    for (int nt = nt0; nt <= nt0 + 3; ++nt) {
        const int bn = ((n_total + nt - 1) / nt + 31) / 32 * 32;
        if (bn * nt < best_bn * best_nt) { best_nt = nt; best_bn = bn; }
    }
    if (block_n) *block_n = best_bn;
    if (n_tiles) *n_tiles = best_nt;
    return 0;
}

int onr_conv_plan_create(onr_conv_plan** out, const onr_conv_desc* d) {
    using namespace onr;
    ONR_REQUIRE(out && d, "null argument");
    ONR_REQUIRE(d->kind >= 0 && d->kind <= 2, "bad conv kind %d", d->kind);
    ONR_REQUIRE(d->a_cp % 32 == 0 && d->out_cp % 32 == 0 && d->n_total % 32 == 0,
                "channel counts must be multiples of 32 (a_cp %d out_cp %d n %d)", d->a_cp, d->out_cp,
                d->n_total);
    ONR_REQUIRE(d->n_total == d->out_s * d->out_s * d->out_cp, "n_total must equal out_s^2*out_cp");
    ONR_REQUIRE(d->B >= 1 && d->H >= 1 && d->W >= 1, "bad grid");
    int block_n = 0, n_tiles = 0;
    onr_conv_tile_n(d->n_total, &block_n, &n_tiles);
    ONR_REQUIRE(d->n_rows >= block_n * n_tiles, "packed weights need %d rows, got %d", block_n * n_tiles,
                d->n_rows);
    if (d->kind == ONR_CONV_DGRAD) ONR_REQUIRE(d->dmul != nullptr, "dgrad needs dmul");
    else ONR_REQUIRE(d->bias_p != nullptr, "fprop needs bias_p");
    onr_conv_plan* pl = new onr_conv_plan();
    ConvParams& p = pl->p;
    p.H = d->H; p.W = d->W; p.B = d->B;
    p.block_n = block_n;
    p.n_tiles = n_tiles;
    // sub-tiles per CTA tile: as many as TMEM allows, but keep >= 2 tiles per SM when the layer is large enough
    int ms_max = 512 / block_n;
    if (ms_max > kMaxMS) ms_max = kMaxMS;
    const int k_tap0 = d->a_s * d->a_s * d->a_cp;
    int ms = 1;
    {
        // pick the sub-tile count that minimises (waves over the SMs) x (operand bytes per CTA tile)
        double best = 1e300;
        for (int m = 1; m <= ms_max; ++m) {
            const long long tiles = (long long)d->B * ceil_div(d->H, kSubH * m) * ceil_div(d->W, kSubW) * n_tiles;
            const double waves = (double)((tiles + num_sms() - 1) / num_sms());
            const double bytes = (double)(k_tap0 / kChunkK) * (3.0 * (kSubH * m + 2) * kRowBytes + 9.0 * block_n * 64);
            const double cost = waves * bytes;
            if (cost < best * 0.999) { best = cost; ms = m; }
        }
        const char* env_ms = getenv("ONR_CONV_MS");     // experiments only
        if (env_ms && atoi(env_ms) >= 1 && atoi(env_ms) <= ms_max) ms = atoi(env_ms);
    }
    p.ms = ms;
    p.tiles_w = ceil_div(d->W, kSubW);
    p.tiles_h = ceil_div(d->H, kSubH * ms);
    p.total_tiles = d->B * p.tiles_h * p.tiles_w * n_tiles;
    const int k_tap = d->a_s * d->a_s * d->a_cp;
    p.chunks = k_tap / kChunkK;
    p.jc_chunks = d->a_s * d->a_cp / kChunkK;
    p.sign = d->kind == ONR_CONV_DGRAD ? -1 : 1;
    p.n_total = d->n_total;
    p.acc_bufs = (2 * ms * block_n <= 512) ? 2 : 1;
    p.mode = d->kind;
    p.out_jc = d->out_s * d->out_cp;
    p.bias = d->bias_p;
    p.dmul = reinterpret_cast<const __nv_bfloat16*>(d->dmul);
    p.prof = nullptr;
    const int box_h = kSubH * ms + 2;
    p.a_bytes = box_h * kRowBytes;                      // multiple of 512
    p.a_bytes = (p.a_bytes + 1023) / 1024 * 1024;       // keep every slot 1024-aligned
    p.b_bytes = block_n * 64;                           // multiple of 2048
    // ring depths: ~96 KB of A boxes at most, the rest for weight tiles
    const int budget = kMaxDynSmem - 1024 - 4 * kStageOutBytes - (int)sizeof(SmemBarriers) - 1024;
    int na = 3;
    while (na > 2 && na * p.a_bytes > budget / 2) --na;
    int nb = (budget - na * p.a_bytes) / p.b_bytes;
    if (nb > kMaxRing) nb = kMaxRing;
    ONR_REQUIRE(nb >= 2, "conv plan: shared memory too small for block_n %d ms %d", block_n, ms);
    // spend what is left on deeper A ring
    while (na < kMaxRing && (na + 1) * p.a_bytes + nb * p.b_bytes <= budget && na < 4) ++na;
    p.na = na;
    p.nb = nb;
    pl->smem = 1024 + (size_t)na * p.a_bytes + (size_t)nb * p.b_bytes + 4 * kStageOutBytes + sizeof(SmemBarriers);
    pl->grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    int rc = make_act_tmap(&pl->tmA, d->a, d->B, d->H, d->W, d->a_cp, d->a_s, kSubW, box_h);
    if (!rc) rc = make_weight_tmap(&pl->tmB, d->w, 9, d->n_rows, k_tap, block_n);
    if (!rc) rc = make_act_tmap(&pl->tmY, d->out, d->B, d->H, d->W, d->out_cp, d->out_s, kSubW, kSubH);
    if (!rc)
        rc = make_act_tmap(&pl->tmD, d->kind == ONR_CONV_FPROP_TRAIN ? d->out_d : d->out, d->B, d->H, d->W,
                           d->out_cp, d->out_s, kSubW, kSubH);
    if (rc) { delete pl; return rc; }
    // The attribute is per function, not per plan: raise it once to the opt-in maximum (227 KB on sm_100).
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kMaxDynSmem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%d) failed: %s", kMaxDynSmem, cudaGetErrorString(e));
            delete pl;
            return (int)e;
        }
        attr_set = true;
    }
    if (pl->smem > (size_t)kMaxDynSmem) {
        set_error("conv plan needs %zu bytes of shared memory", pl->smem);
        delete pl;
        return -1;
    }
    *out = pl;
    return 0;
}

int onr_conv_plan_run(const onr_conv_plan* pl, void* stream) {
    using namespace onr;
    ONR_REQUIRE(pl != nullptr, "null plan");
    conv_igemm_kernel<<<pl->grid, kThreads, pl->smem, (cudaStream_t)stream>>>(pl->tmA, pl->tmB, pl->tmY,
                                                                             pl->tmD, pl->p);
    ONR_LAUNCH_CHECK();
    return 0;
}

// Profiling hook (selftest only): per-CTA cycle counters, 8 x int64 per CTA, or NULL to switch off.
int onr_conv_plan_set_prof(onr_conv_plan* pl, long long* prof_dev, int* grid) {
    ONR_REQUIRE(pl != nullptr, "null plan");
    pl->p.prof = prof_dev;
    if (grid) *grid = pl->grid;
    return 0;
}

// Tiling the plan chose (for logs / tests): block_n, n_tiles, sub-tiles per CTA tile, ring depths.
int onr_conv_plan_info(const onr_conv_plan* pl, int* block_n, int* n_tiles, int* ms, int* na, int* nb) {
    ONR_REQUIRE(pl != nullptr, "null plan");
    if (block_n) *block_n = pl->p.block_n;
    if (n_tiles) *n_tiles = pl->p.n_tiles;
    if (ms) *ms = pl->p.ms;
    if (na) *na = pl->p.na;
    if (nb) *nb = pl->p.nb;
    return 0;
}

void onr_conv_plan_destroy(onr_conv_plan* pl) { delete pl; }

}  // extern "C"
