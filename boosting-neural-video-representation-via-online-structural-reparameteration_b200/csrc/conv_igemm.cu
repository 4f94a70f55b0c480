// conv_igemm.cu — 3x3 / stride 1 / pad 1 convolution as a tcgen05 implicit GEMM for sm_100a.
//
// Replaces F.conv2d + nn.PixelShuffle + nn.SiLU of NeRVBlock.forward (reference model.py:539, :523,
// :520 and :567) and, with flipped taps and the un-shuffled dZ view as the A operand, the data
// gradient that autograd derives for it (main_train.py:249).
//
//   D[pixel, n] = sum_{tap, k} A[pixel + off(tap), k] * Wt[tap][n][k]
//
// * M tile  = 128 pixels = an 8 x 16 spatial patch.  For every tap the A tile is ONE TMA box of the
//   NHWC activation shifted by (dh, dw); the TMA unit zero-fills the halo, so padding costs nothing
//   and no im2col buffer exists.  K advances in 32-channel (64 B, SWIZZLE_64B) chunks.
// * N tile  = block_n <= 384 fp32 accumulator columns in TMEM (2 UMMAs of block_n/2 when > 256).
// * warp 0 = TMA producer, warp 1 = UMMA issuer (one thread), warps 2..5 = epilogue
//   (tcgen05.ld -> bias/SiLU/SiLU' or dgrad scaling -> bf16 -> swizzled smem -> TMA store through the
//   PixelShuffle view, so the shuffle is pure addressing).
// * persistent: grid = min(tiles, #SM), static round-robin tile order, TMEM double-buffered when
//   2*block_n <= 512 so the epilogue of tile t overlaps the MMAs of tile t+1.
#include "onr_common.cuh"
#include "onr_ptx.cuh"

namespace onr {

constexpr int kTileH = 8;
constexpr int kTileW = 16;
constexpr int kBlockM = 128;
constexpr int kChunkK = 32;                       // bf16 elements per K step (64 bytes)
constexpr int kABytes = kBlockM * kChunkK * 2;    // 8192
constexpr int kStageOutBytes = kBlockM * 64;      // one 128 x 32 bf16 staging tile
constexpr int kThreads = 192;
constexpr int kMaxBlockN = 384;
constexpr int kSmemBudget = 220 * 1024;
constexpr int kMaxDynSmem = 232448;   // 227 KB: per-block opt-in maximum on sm_100

struct ConvParams {
    int H, W, B;
    int tiles_w, tiles_h, m_tiles, n_tiles, total_tiles;
    int block_n, n_sub, sub_n;
    int chunks, jc_chunks, sign;
    int n_total, acc_bufs, stages;
    int mode;
    int out_jc;  // channels per shuffle row i of the output view (out_s * out_cp)
    const float* bias;
    const __nv_bfloat16* dmul;
};

struct __align__(8) SmemBarriers {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmD,
                  const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x (A | B)] [staging 2 x 2 x 8 KB] [barriers]
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = kABytes + p.block_n * 64;
    const uint32_t staging = smem_base + p.stages * stage_bytes;
    SmemBarriers* bars = reinterpret_cast<SmemBarriers*>(smem_raw + (staging + 4 * kStageOutBytes - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->tmem_full[b]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[b]), 128);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmY);
        if (p.mode == ONR_CONV_FPROP_TRAIN) tma_prefetch_desc(&tmD);
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const int ksteps = 9 * p.chunks;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                int mt = tile / p.n_tiles;
                const int tw = mt % p.tiles_w;
                mt /= p.tiles_w;
                const int th = mt % p.tiles_h;
                const int b = mt / p.tiles_h;
                const int h0 = th * kTileH, w0 = tw * kTileW, n0 = nt * p.block_n;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int tap = ks / p.chunks;
                    const int ch = ks - tap * p.chunks;
                    const int ii = ch / p.jc_chunks;
                    const int jc0 = (ch - ii * p.jc_chunks) * kChunkK;
                    const int dh = (tap / 3 - 1) * p.sign;
                    const int dw = (tap % 3 - 1) * p.sign;
                    mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
                    const uint32_t full = smem_u32(&bars->full[stage]);
                    const uint32_t a_s = smem_base + stage * stage_bytes;
                    const uint32_t b_s = a_s + kABytes;
                    mbar_expect_tx(full, stage_bytes);
                    tma_load_5d(a_s, &tmA, full, jc0, w0 + dw, ii, h0 + dh, b);
                    for (int sub = 0; sub < p.n_sub; ++sub)
                        tma_load_3d(b_s + sub * p.sub_n * 64, &tmB, full, ch * kChunkK, n0 + sub * p.sub_n,
                                    tap);
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== UMMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(kBlockM, p.sub_n, 0, 0);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int buf = it % p.acc_bufs;
                const uint32_t acc_phase = (uint32_t)(it / p.acc_bufs) & 1u;
                mbar_wait(smem_u32(&bars->tmem_empty[buf]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * p.block_n;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(smem_u32(&bars->full[stage]), phase);
                    tc_fence_after();
                    const uint32_t a_s = smem_base + stage * stage_bytes;
                    const uint32_t b_s = a_s + kABytes;
#pragma unroll
                    for (int k = 0; k < kChunkK / 16; ++k) {
                        const uint64_t adesc = make_smem_desc(a_s + k * 32, 16, 512, SWZ_64B);
                        for (int sub = 0; sub < p.n_sub; ++sub) {
                            const uint64_t bdesc =
                                make_smem_desc(b_s + sub * p.sub_n * 64 + k * 32, 16, 512, SWZ_64B);
                            umma_bf16(d_tmem + sub * p.sub_n, adesc, bdesc, idesc, (ks | k) != 0);
                        }
                    }
                    umma_commit(smem_u32(&bars->empty[stage]));
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(smem_u32(&bars->tmem_full[buf]));
            }
        }
    } else {
        // ===================================================================== epilogue (4 warps)
        const int q = warp & 3;             // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;      // accumulator row == pixel inside the 8x16 patch
        const int hl = row >> 4, wl = row & 15;
        const bool store_thread = (warp == 2 && lane == 0);
        const uint32_t swz = (uint32_t)((row >> 1) & 3);
        uint32_t chunk_ctr = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int nt = tile % p.n_tiles;
            int mt = tile / p.n_tiles;
            const int tw = mt % p.tiles_w;
            mt /= p.tiles_w;
            const int th = mt % p.tiles_h;
            const int b = mt / p.tiles_h;
            const int h0 = th * kTileH, w0 = tw * kTileW, n0 = nt * p.block_n;
            const int buf = it % p.acc_bufs;
            const uint32_t acc_phase = (uint32_t)(it / p.acc_bufs) & 1u;
            mbar_wait(smem_u32(&bars->tmem_full[buf]), acc_phase);
            tc_fence_after();
            const int nchunks = p.block_n / 32;
            for (int c = 0; c < nchunks; ++c) {
                const int n = n0 + c * 32;
                if (n >= p.n_total) break;
                const uint32_t sbuf = chunk_ctr & 1u;
                ++chunk_ctr;
                if (store_thread) tma_store_wait_read<1>();
                named_bar_sync(1, 128);
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * p.block_n + c * 32, r);
                tmem_ld_wait();
                const uint32_t ybuf = staging + sbuf * 2 * kStageOutBytes;
                const uint32_t dbuf = ybuf + kStageOutBytes;
                if (p.mode == ONR_CONV_DGRAD) {
                    const int h = h0 + hl, w = w0 + wl;
                    uint4 dv[4];
                    if (h < p.H && w < p.W) {
                        const uint4* dp = reinterpret_cast<const uint4*>(
                            p.dmul + ((size_t)(b * p.H + h) * p.W + w) * p.n_total + n);
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
                    tma_store_5d(&tmY, ybuf, ojc, w0, oi, h0, b);
                    if (p.mode == ONR_CONV_FPROP_TRAIN) tma_store_5d(&tmD, dbuf, ojc, w0, oi, h0, b);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
        }
        if (store_thread) tma_store_wait_all<0>();
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

int onr_conv_tile_n(int n_total, int* block_n, int* n_tiles) {
    ONR_REQUIRE(n_total > 0 && n_total % 32 == 0, "n_total must be a positive multiple of 32");
    const int nt = (n_total + onr::kMaxBlockN - 1) / onr::kMaxBlockN;
    int bn = ((n_total + nt - 1) / nt + 31) / 32 * 32;
    if (block_n) *block_n = bn;
    if (n_tiles) *n_tiles = nt;
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
    onr_conv_plan* pl = new onr_conv_plan();
    ConvParams& p = pl->p;
    p.H = d->H; p.W = d->W; p.B = d->B;
    p.tiles_w = ceil_div(d->W, kTileW);
    p.tiles_h = ceil_div(d->H, kTileH);
    p.m_tiles = d->B * p.tiles_h * p.tiles_w;
    p.n_tiles = n_tiles;
    p.total_tiles = p.m_tiles * n_tiles;
    p.block_n = block_n;
    p.n_sub = block_n > 256 ? 2 : 1;
    p.sub_n = block_n / p.n_sub;
    const int k_tap = d->a_s * d->a_s * d->a_cp;
    p.chunks = k_tap / kChunkK;
    p.jc_chunks = d->a_s * d->a_cp / kChunkK;
    p.sign = d->kind == ONR_CONV_DGRAD ? -1 : 1;
    p.n_total = d->n_total;
    p.acc_bufs = (2 * block_n <= 512) ? 2 : 1;
    p.mode = d->kind;
    p.out_jc = d->out_s * d->out_cp;
    p.bias = d->bias_p;
    p.dmul = reinterpret_cast<const __nv_bfloat16*>(d->dmul);
    const int stage_bytes = kABytes + block_n * 64;
    int stages = (kSmemBudget - 4 * kStageOutBytes - 1024 - (int)sizeof(SmemBarriers)) / stage_bytes;
    if (stages > 8) stages = 8;
    p.stages = stages;
    pl->smem = 1024 + (size_t)stages * stage_bytes + 4 * kStageOutBytes + sizeof(SmemBarriers);
    pl->grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    if (d->kind == ONR_CONV_DGRAD) ONR_REQUIRE(d->dmul != nullptr, "dgrad needs dmul");
    else ONR_REQUIRE(d->bias_p != nullptr, "fprop needs bias_p");
    int rc = make_act_tmap(&pl->tmA, d->a, d->B, d->H, d->W, d->a_cp, d->a_s, kTileW, kTileH);
    if (!rc) rc = make_weight_tmap(&pl->tmB, d->w, 9, d->n_rows, k_tap, p.sub_n);
    if (!rc) rc = make_act_tmap(&pl->tmY, d->out, d->B, d->H, d->W, d->out_cp, d->out_s, kTileW, kTileH);
    if (!rc)
        rc = make_act_tmap(&pl->tmD, d->kind == ONR_CONV_FPROP_TRAIN ? d->out_d : d->out, d->B, d->H, d->W,
                           d->out_cp, d->out_s, kTileW, kTileH);
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
    ONR_REQUIRE(pl->smem <= (size_t)kMaxDynSmem, "conv plan needs %zu bytes of shared memory", pl->smem);
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

void onr_conv_plan_destroy(onr_conv_plan* pl) { delete pl; }

}  // extern "C"
