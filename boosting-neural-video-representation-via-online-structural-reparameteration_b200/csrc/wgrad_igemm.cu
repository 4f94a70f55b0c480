// wgrad_igemm.cu — weight gradient of the 3x3 block convolution as a tcgen05 GEMM whose reduction
// dimension is the pixel index (autograd's convolution_backward weight branch for model.py:539 / :523,
// run from main_train.py:249).
//
//   dKp[n][kh*3+kw][ci] = sum_{b,h,w} dZ[b,h,w,n] * X[b, h+kh-1, w+kw-1, ci]
//
// Both operands are NHWC, i.e. the reduction index (pixel) is the slow axis: they are consumed as
// MN-major UMMA operands straight from the TMA boxes (rows = pixels, 64-byte swizzled channel
// chunks), so no transposed copy of the activations is ever written.
//
// Work decomposition: a CTA owns (n_tile of 128 dZ channels, kw) and a slice of the image.  It walks
// image rows with a rolling window: the X row tile r (shifted by kw-1 columns) is loaded once and
// multiplied against dZ rows r+1, r, r-1 (kh = 0,1,2), each of which is also loaded exactly once.
// Accumulators: 3 x Cin_p fp32 columns of TMEM.  Partial sums of the image slices are combined with
// red.global.add.v4.f32 into dKp (zeroed by the caller).
//
// Input channels beyond 128 (fc dim 128 with expansion 8, fc_hw_dim 9_16_156, ...): the X channels are cut into
// chunks of <= 128 and the chunk index becomes one more job coordinate — every chunk is an independent GEMM over the
// same dZ rows that owns the columns [c0, c0 + cc) of dKp; chunk 0 also carries the bias column.
//
// Bias gradient for free: in the kw = 1 jobs the centre-tap (kh = 1) MMAs see a B operand that is 32 channels
// wider — a constant tile whose channel 0 is 1.0 — so one extra accumulator column collects
// dbias_p[n] = sum_pixels dZ[pixel, n] on the tensor pipe (+11 % MMA work in one of three jobs) instead of a
// separate pass over dZ (it was 94 us per step).
#include "onr_common.cuh"
#include "onr_ptx.cuh"

namespace onr {

constexpr int kWgPx = 64;          // pixels (reduction elements) per pipeline stage
constexpr int kWgBox = kWgPx * 64; // bytes of one 32-channel box
constexpr int kWgStages = 6;
// warps 0..3 = epilogue, warp 4 = TMA producer, warp 5 = UMMA issuer (highest id: never starved by the others)
constexpr int kWgThreads = 192;
constexpr int kWgWarpProd = 4, kWgWarpMma = 5;
constexpr int kWgRowsPerUnit = 24;

struct WgradParams {
    int B, H, W;
    int s, jc_chunks, n_pre, n_tiles;
    int xb, x_cp;          // xb: X boxes per stage of a full chunk; x_cp: ALL padded input channels (pitch of dKp)
    int nchunks;           // ceil(x_cp / 128)
    int wchunks, hunits, units_total, splits, rows_per_unit;
    float* dKp;
    float* dbias_p;
};

struct __align__(8) WgBarriers {
    uint64_t full[kWgStages];
    uint64_t empty[kWgStages];
    uint64_t acc_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX,
                   const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = (4 + p.xb + 1) * kWgBox;    // dZ boxes | X boxes | constant ones box
    WgBarriers* bars =
        reinterpret_cast<WgBarriers*>(smem_raw + (smem_base - smem_u32(smem_raw)) + kWgStages * stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int chunk = job / (p.n_tiles * 3), jj = job % (p.n_tiles * 3);
    const int n_tile = jj / 3, kw = jj % 3;
    const int c0 = chunk * 128;                              // first X channel of this CTA's chunk
    const int cc = min(128, p.x_cp - c0);                    // its width (multiple of 32)
    const int xbc = cc / 32;
    const int u0 = (int)((long long)split * p.units_total / p.splits);
    const int u1 = (int)((long long)(split + 1) * p.units_total / p.splits);

    // constant "ones" box of every stage: [64 pixel rows][32 channels] bf16, SWIZZLE_64B layout, channel 0 = 1.0
    for (int i = threadIdx.x; i < kWgStages * kWgPx; i += blockDim.x) {
        const int st = i / kWgPx, r = i % kWgPx;
        uint4* row = reinterpret_cast<uint4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + st * stage_bytes +
                                              (4 + p.xb) * kWgBox + r * 64);
        const int phys0 = (r >> 1) & 3;                     // physical 16-byte piece holding logical piece 0
#pragma unroll
        for (int j = 0; j < 4; ++j) row[j] = make_uint4(j == phys0 ? 0x00003F80u : 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();                               // generic-proxy writes -> visible to the tensor pipe
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        mbar_init(smem_u32(&bars->acc_full), 1);
        fence_mbar_init();
    }
    if (warp == kWgWarpProd && lane == 0) {
        tma_prefetch_desc(&tmDz);
        tma_prefetch_desc(&tmX);
    }
    if (warp == kWgWarpMma) {
        tmem_alloc(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == kWgWarpProd) {
        {   // warp-uniform loop, issue predicated by elect_one (see conv_igemm.cu)
            uint32_t g = 0;
            for (int u = u0; u < u1; ++u) {
                const int hu = u % p.hunits;
                const int wc = (u / p.hunits) % p.wchunks;
                const int b = u / (p.hunits * p.wchunks);
                const int r0 = hu * p.rows_per_unit;
                const int r1 = min(p.H, r0 + p.rows_per_unit);
                const int wbase = wc * kWgPx;
                for (int r = r0 - 2; r < r1; ++r, ++g) {
                    const uint32_t stage = g % kWgStages;
                    const uint32_t phase = (g / kWgStages) & 1u;
                    mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
                    if (elect_one()) {
                        const uint32_t full = smem_u32(&bars->full[stage]);
                        const uint32_t dz_s = smem_base + stage * stage_bytes;
                        const uint32_t x_s = dz_s + 4 * kWgBox;
                        const bool with_x = r >= r0;
                        mbar_expect_tx(full, (with_x ? (4 + xbc) : 4) * kWgBox);
                        // dZ row r+1, four 32-channel boxes of this CTA's 128-channel n tile
                        for (int qb = 0; qb < 4; ++qb) {
                            const int qn = n_tile * 4 + qb;  // 32-channel chunk of n'
                            const int ii = qn / p.jc_chunks;
                            const int jc0 = (qn - ii * p.jc_chunks) * 32;
                            tma_load_5d(dz_s + qb * kWgBox, &tmDz, full, jc0, wbase, ii, r + 1, b);
                        }
                        if (with_x)
                            for (int xb = 0; xb < xbc; ++xb)
                                tma_load_5d(x_s + xb * kWgBox, &tmX, full, c0 + xb * 32, wbase + kw - 1, 0, r, b);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kWgWarpMma) {
        {
            const uint32_t idesc = make_idesc_bf16(128, cc, 1, 1);
            // centre tap + ones column: chunk 0 only (a full chunk whenever there are several, so the constant box
            // at (4 + p.xb) * kWgBox directly follows its X boxes)
            const uint32_t idesc_b = chunk == 0 ? make_idesc_bf16(128, cc + 32, 1, 1) : idesc;
            uint32_t g = 0;
            uint32_t started = 0;  // bit kh set once acc[kh] holds data
            for (int u = u0; u < u1; ++u) {
                const int hu = u % p.hunits;
                const int r0 = hu * p.rows_per_unit;
                const int r1 = min(p.H, r0 + p.rows_per_unit);
                for (int r = r0 - 2; r < r1; ++r, ++g) {
                    const uint32_t stage = g % kWgStages;
                    const uint32_t phase = (g / kWgStages) & 1u;
                    mbar_wait(smem_u32(&bars->full[stage]), phase);
                    tc_fence_after();
                    if (elect_one()) {
                      if (r >= r0) {
                        const uint32_t x_s = smem_base + stage * stage_bytes + 4 * kWgBox;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            // dZ row (r - kh + 1) lives in the stage loaded at iteration g - kh
                            const uint32_t dz_s = smem_base + ((g - kh) % kWgStages) * stage_bytes;
#pragma unroll
                            for (int kk = 0; kk < kWgPx / 16; ++kk) {
                                const uint64_t adesc = make_smem_desc(dz_s + kk * 1024, kWgBox, 512, SWZ_64B);
                                const uint64_t bdesc = make_smem_desc(x_s + kk * 1024, kWgBox, 512, SWZ_64B);
                                // accumulator columns: kh=0 -> [0,Cp), kh=2 -> [Cp,2Cp), kh=1 -> [2Cp, 3Cp(+32))
                                const uint32_t col = kh == 0 ? 0u : (kh == 2 ? (uint32_t)cc : 2u * cc);
                                umma_bf16(tmem_base + col, adesc, bdesc, (kh == 1 && kw == 1) ? idesc_b : idesc,
                                          ((started >> kh) & 1u) | (kk != 0));
                            }
                        }
                      }
                      if (g >= 2) umma_commit(smem_u32(&bars->empty[(g - 2) % kWgStages]));
                    }
                    __syncwarp();
                    if (r >= r0) started = 7u;
                }
            }
            if (elect_one()) umma_commit(smem_u32(&bars->acc_full));
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int n = n_tile * 128 + row;
        mbar_wait(smem_u32(&bars->acc_full), 0);
        tc_fence_after();
        const int cchunks = cc / 32;
        for (int kh = 0; kh < 3; ++kh) {
            const uint32_t colb = kh == 0 ? 0u : (kh == 2 ? (uint32_t)cc : 2u * cc);
            for (int c = 0; c < cchunks; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + colb + c * 32, r);
                tmem_ld_wait();
                if (n < p.n_pre) {
                    float* dst = p.dKp + ((size_t)n * 9 + kh * 3 + kw) * p.x_cp + c0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst + j * 4),
                                     "f"(__uint_as_float(r[j * 4])), "f"(__uint_as_float(r[j * 4 + 1])),
                                     "f"(__uint_as_float(r[j * 4 + 2])), "f"(__uint_as_float(r[j * 4 + 3]))
                                     : "memory");
                }
            }
        }
        if (kw == 1 && chunk == 0 && p.dbias_p != nullptr) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + 3u * cc, r);
            tmem_ld_wait();
            if (n < p.n_pre) atomicAdd(p.dbias_p + n, __uint_as_float(r[0]));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWgWarpMma) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace onr

struct onr_wgrad_plan {
    CUtensorMap tmDz, tmX;
    onr::WgradParams p;
    int grid;
    size_t smem;
    const void* dz;
    float* dbias_p;
    int dz_cp;
};

extern "C" {

int onr_wgrad_plan_create(onr_wgrad_plan** out, const onr_wgrad_desc* d) {
    using namespace onr;
    ONR_REQUIRE(out && d, "null argument");
    ONR_REQUIRE(d->x_cp % 32 == 0 && d->dz_cp % 32 == 0 && d->x_cp >= 32 && d->dz_cp >= 32,
                "wgrad: channel counts must be positive multiples of 32 (x_cp %d dz_cp %d)", d->x_cp, d->dz_cp);
    ONR_REQUIRE(d->s >= 1 && d->B >= 1 && d->H >= 1 && d->W >= 1, "wgrad: bad grid");
    onr_wgrad_plan* pl = new onr_wgrad_plan();
    WgradParams& p = pl->p;
    p.B = d->B; p.H = d->H; p.W = d->W;
    p.s = d->s;
    p.jc_chunks = d->s * d->dz_cp / 32;
    p.n_pre = d->s * d->s * d->dz_cp;
    p.n_tiles = ceil_div(p.n_pre, 128);
    p.x_cp = d->x_cp;
    p.nchunks = ceil_div(d->x_cp, 128);
    p.xb = (d->x_cp < 128 ? d->x_cp : 128) / 32;
    p.wchunks = ceil_div(d->W, kWgPx);
    p.rows_per_unit = kWgRowsPerUnit;
    p.hunits = ceil_div(d->H, p.rows_per_unit);
    p.units_total = d->B * p.wchunks * p.hunits;
    const int jobs = p.nchunks * p.n_tiles * 3;
    int splits = num_sms() / jobs;
    if (splits < 1) splits = 1;
    if (splits > p.units_total) splits = p.units_total;
    p.splits = splits;
    p.dKp = d->dKp;
    p.dbias_p = d->dbias_p;
    pl->grid = jobs * splits;
    pl->smem = 1024 + (size_t)kWgStages * (4 + p.xb + 1) * kWgBox + sizeof(WgBarriers);
    pl->dz = d->dz;
    pl->dbias_p = d->dbias_p;
    pl->dz_cp = d->dz_cp;
    int rc = make_act_tmap(&pl->tmDz, d->dz, d->B, d->H, d->W, d->dz_cp, d->s, kWgPx, 1);
    if (!rc) rc = make_act_tmap(&pl->tmX, d->x, d->B, d->H, d->W, d->x_cp, 1, kWgPx, 1);
    if (rc) { delete pl; return rc; }
    static bool attr_set = false;   // per-function attribute: raise once to the sm_100 opt-in maximum
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(wgrad smem) failed: %s", cudaGetErrorString(e));
            delete pl;
            return (int)e;
        }
        attr_set = true;
    }
    ONR_REQUIRE(pl->smem <= 232448, "wgrad plan needs %zu bytes of shared memory", pl->smem);
    *out = pl;
    return 0;
}

int onr_wgrad_plan_run(const onr_wgrad_plan* pl, void* stream) {
    using namespace onr;
    ONR_REQUIRE(pl != nullptr, "null plan");
    const WgradParams& p = pl->p;
    wgrad_igemm_kernel<<<pl->grid, kWgThreads, pl->smem, (cudaStream_t)stream>>>(pl->tmDz, pl->tmX, p);
    ONR_LAUNCH_CHECK();
    return 0;
}

void onr_wgrad_plan_destroy(onr_wgrad_plan* pl) { delete pl; }

}  // extern "C"
