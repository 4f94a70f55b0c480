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
// * K advances in 64-channel chunks with 128-byte rows / SWIZZLE_128B wherever the channel count allows and
//   a 32-channel / SWIZZLE_64B tail otherwise (96 = 64 + 32).  Cycle counters on v2
//   (profiles/r01_conv_v2_role_cycle_counters.txt) showed the UMMA operand fetch from 64-byte-swizzled rows
//   running at ~53 B/clk, i.e. the tensor pipe idling on shared memory; 128-byte rows double that.
// * warp 0 = A producer, warp 1 = UMMA issuer (one thread), warp 2 = weight producer, warps 3..10 = epilogue
//   (two warps per TMEM lane quarter: tcgen05.ld -> bias/SiLU/SiLU' or dgrad scaling -> bf16 -> swizzled smem
//   -> TMA store through the PixelShuffle view, so the shuffle is pure addressing).
// * persistent: grid = min(tiles, #SM); the accumulator is ALWAYS double-buffered in TMEM
//   (2*MS*block_n <= 512) so the epilogue of tile t overlaps the MMAs of tile t+1.
#include "onr_common.cuh"
#include "onr_ptx.cuh"

#include <stdlib.h>

namespace onr {

constexpr int kSubH = 16;                          // sub-tile: 16 x 8 pixels = 128 accumulator rows
constexpr int kSubW = 8;
constexpr int kStageOutBytes = 128 * 128;          // one 128 x 64 bf16 staging tile (or two 128 x 32 tiles)
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
// Warp roles: warps 0..7 = epilogue, 8 = A producer, 9 = weight producer, 10.. = UMMA issuers (warp 10 also owns
// the TMEM allocation).
constexpr int kWarpProdA = 8, kWarpProdB = 9, kWarpMma0 = 10;
constexpr int kMaxIssuers = 2;                     // UMMA issuer warps (alternate K-groups of a tile)
constexpr int kWarpMma = kWarpMma0;                // TMEM owner
constexpr int kThreads = (kWarpMma0 + kMaxIssuers) * 32;
constexpr int kMaxBlockN = 256;
constexpr int kMaxMS = 4;
constexpr int kMaxDynSmem = 232448;                // 227 KB: per-block opt-in maximum on sm_100
constexpr int kMaxRing = 8;
// Work items (tiles) are handed out DYNAMICALLY: a persistent CTA takes work item blockIdx.x first and every further one
// from a global atomic counter.  With the static round-robin assignment a CTA that starts late — because a side-stream
// kernel (wgrad, the ERB fold GEMMs, MS-SSIM) still holds its SM — kept its full share of tiles and the whole kernel
// waited for it; now it simply processes fewer.
constexpr int kSchedRing = 4;

struct ConvParams {
    int H, W, B;
    int tiles_w, tiles_h, n_tiles, total_tiles;
    int cluster, px_tiles, pair_tiles;   // CTAs per cluster (1 or 2), pixel tiles, work items per cluster
    int ksplit;                          // DGRAD of a small layer: K-groups of a tile split over ksplit work items
    float* kacc;                         // [B*H*W][n_total] fp32 partial sums of the split (zeroed per launch)
    int dynamic;                         // work items after a CTA's first one come from a global atomic counter
    int* sched;                          // [0] next work item - gridDim, [1] CTAs finished (both zero between launches)
    int block_n, ms;
    int chunks, cpi, n64, cj, sign;   // chunks per tap, chunks per shuffle row i, 64-wide chunks per i, channels per i
    int n_total, acc_bufs, issuers;
    int na, nb;              // ring depths
    int a_bytes, b_bytes;    // slot sizes
    int mode;
    int out_jc;              // channels per shuffle row i of the output view (out_s * out_cp)
    const float* bias;
    const __nv_bfloat16* dmul;
    // FPROP_HEAD: RGB head fused into the epilogue
    const float* head_w;     // [3][head_c]
    const float* head_b;     // [3]
    float* img;              // fp32 NCHW [B][3][H*out_s][W*out_s]
    int head_c, use_sigmoid, out_s;
    // optional per-CTA cycle counters (selftest "prof" mode): [cta][8] =
    // {total, mma wait A, mma wait B, mma wait tmem_empty, epi wait tmem_full, epi wait store, epi busy, tiles}
    long long* prof;
};

__device__ __forceinline__ long long clk() { return clock64(); }

struct __align__(8) SmemBarriers {
    uint64_t a_full[kMaxRing], a_empty[kMaxRing];
    uint64_t b_full[kMaxRing], b_empty[kMaxRing];
    uint64_t tmem_full[2], tmem_empty[2], turn[2];
    // work-item feed: the A-producer warp fetches work items and publishes them to the other roles through this ring
    uint64_t sched_full[kSchedRing], sched_empty[kSchedRing];
    int sched_tile[kSchedRing];
    uint32_t tmem_base;
    // bias of the current tile's block_n channels (fprop), double-buffered over tiles: read from shared memory in the
    // epilogue instead of 32 global loads per thread and iteration, whose latency eight epilogue warps cannot hide
    // (role cycle counters: the fprop epilogue was busy 74 % of the kernel and set its pace)
    alignas(16) float bias_s[2][kMaxBlockN];
    // FPROP_HEAD: head weights of the tile's channels (zero in the channel padding), bias, and the per-row partial
    // dot products of the two column halves
    alignas(16) float head_w_s[3][kMaxBlockN];
    float head_b_s[4];
    float head_red[2][128][3];
};

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Chunk {
    int ii, jc0, width;
};
__device__ __forceinline__ Chunk chunk_of(const ConvParams& p, int ch) {
    Chunk c;
    c.ii = ch / p.cpi;
    const int k = ch - c.ii * p.cpi;
    c.width = k < p.n64 ? 64 : 32;
    c.jc0 = k < p.n64 ? k * 64 : p.n64 * 64;
    return c;
}

struct TileCoord {
    int b, h0, w0, n0;
    int g0, g1;    // K-groups (chunk, dw) of the tile this work item covers: all of them unless the plan splits K
    bool skip;     // padding work item of an odd-sized CTA pair: runs the loads / MMAs of its neighbour, stores nothing
};
// Work item q of a cluster = (n tile, `cluster` horizontally adjacent pixel tiles); CTA `rank` takes pixel tile
// (q / n_tiles) * cluster + rank.  Both CTAs of a pair therefore walk the SAME weight tiles in the same order, which is
// what lets them share every weight tile through TMA multicast.
__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int q, int rank) {
    TileCoord t;
    const int groups = p.chunks * 3;
    const int part = q % p.ksplit;               // (ksplit > 1 only with cluster == 1)
    q /= p.ksplit;
    t.g0 = (int)((long long)part * groups / p.ksplit);
    t.g1 = (int)((long long)(part + 1) * groups / p.ksplit);
    const int nt = q % p.n_tiles;
    int mt = (q / p.n_tiles) * p.cluster + rank;
    t.skip = mt >= p.px_tiles;
    if (t.skip) mt = p.px_tiles - 1;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    t.b = mt / p.tiles_h;
    t.h0 = th * kSubH * p.ms;
    t.w0 = tw * kSubW;
    t.n0 = nt * p.block_n;
    return t;
}

// Consumer side of the work-item feed (every role but the A producer): all lanes wait, lane 0 releases the slot.
struct TileFeed {
    uint32_t slot = 0, phase = 0;
    __device__ __forceinline__ int next(SmemBarriers* bars) {
        mbar_wait(smem_u32(&bars->sched_full[slot]), phase);
        const int tile = *reinterpret_cast<volatile int*>(&bars->sched_tile[slot]);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(smem_u32(&bars->sched_empty[slot]));
        if (++slot == kSchedRing) { slot = 0; phase ^= 1; }
        return tile;
    }
};

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA64, const __grid_constant__ CUtensorMap tmA32,
                  const __grid_constant__ CUtensorMap tmB64, const __grid_constant__ CUtensorMap tmB32,
                  const __grid_constant__ CUtensorMap tmY64, const __grid_constant__ CUtensorMap tmY32,
                  const __grid_constant__ CUtensorMap tmD64, const __grid_constant__ CUtensorMap tmD32,
                  const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [na x A slot] [nb x B slot] [staging: y 16 KB | d 16 KB] [barriers]
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_ring = smem_base;
    const uint32_t b_ring = a_ring + p.na * p.a_bytes;
    const uint32_t staging = b_ring + p.nb * p.b_bytes;
    SmemBarriers* bars =
        reinterpret_cast<SmemBarriers*>(smem_raw + (staging + 2 * kStageOutBytes - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = p.cluster > 1 ? (int)cluster_ctarank() : 0;
    const int q0 = blockIdx.x / p.cluster, qstep = gridDim.x / p.cluster;
    const uint16_t mc_mask = (uint16_t)((1u << p.cluster) - 1u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(smem_u32(&bars->a_full[s]), 1);
            mbar_init(smem_u32(&bars->a_empty[s]), 1);
            mbar_init(smem_u32(&bars->b_full[s]), 1);
            // a weight slot is refilled (by multicast, in every CTA of the cluster) once ALL CTAs have consumed it
            mbar_init(smem_u32(&bars->b_empty[s]), p.cluster);
        }
        for (int s = 0; s < kSchedRing; ++s) {
            mbar_init(smem_u32(&bars->sched_full[s]), 1);
            // released by the weight producer, every issuer warp and every epilogue warp
            mbar_init(smem_u32(&bars->sched_empty[s]), 1 + p.issuers + kEpiWarps);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->tmem_full[b]), p.issuers);
            mbar_init(smem_u32(&bars->turn[b]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[b]), kEpiThreads);
        }
        fence_mbar_init();
    }
    if (warp == kWarpProdA && lane == 0) {
        tma_prefetch_desc(&tmA64);
        tma_prefetch_desc(&tmA32);
        tma_prefetch_desc(&tmY64);
        tma_prefetch_desc(&tmY32);
    }
    if (warp == kWarpProdB && lane == 0) {
        tma_prefetch_desc(&tmB64);
        tma_prefetch_desc(&tmB32);
    }
    if (warp == kWarpMma) {
        tmem_alloc(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();     // the peer's barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == kWarpProdA) {
        // ===================================================================== A producer
        // (service warps run their loops warp-uniformly and predicate only the issue with elect_one: under a
        //  `lane == 0` branch the compiler wraps every UTMALDG / UTCHMMA in an ELECT..BRA.U.ANY serialisation loop
        //  with R2UR moves — measured 285 vs 195 cycles per MMA in csrc/mma_bench.cu)
        {
            uint32_t slot = 0, phase = 0;
            uint32_t fslot = 0, fphase = 0;
            for (int n = 0;; ++n) {
                // fetch the next work item and publish it (the end marker too, so that every role terminates)
                mbar_wait(smem_u32(&bars->sched_empty[fslot]), fphase ^ 1);
                if (elect_one()) {
                    int next = q0 + n * qstep;
                    if (p.dynamic && n > 0) next = qstep + atomicAdd(p.sched, 1);
                    bars->sched_tile[fslot] = next;
                    mbar_arrive(smem_u32(&bars->sched_full[fslot]));
                }
                __syncwarp();
                const int tile = *reinterpret_cast<volatile int*>(&bars->sched_tile[fslot]);
                if (++fslot == kSchedRing) { fslot = 0; fphase ^= 1; }
                if (tile >= p.pair_tiles) break;
                const TileCoord t = tile_coord(p, tile, rank);
                for (int ch = 0; ch < p.chunks; ++ch) {
                    const Chunk c = chunk_of(p, ch);
                    const CUtensorMap* map = c.width == 64 ? &tmA64 : &tmA32;
                    const uint32_t bytes = (uint32_t)(kSubH * p.ms + 2) * kSubW * c.width * 2;
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        const int g = ch * 3 + dwi;
                        if (g < t.g0 || g >= t.g1) continue;
                        mbar_wait(smem_u32(&bars->a_empty[slot]), phase ^ 1);
                        if (elect_one()) {
                            const uint32_t full = smem_u32(&bars->a_full[slot]);
                            mbar_expect_tx(full, bytes);
                            tma_load_5d(a_ring + slot * p.a_bytes, map, full, c.jc0, t.w0 + dwi - 1, c.ii, t.h0 - 1, t.b);
                        }
                        __syncwarp();
                        if (++slot == (uint32_t)p.na) { slot = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kWarpProdB) {
        // ===================================================================== weight producer
        {
            // cluster of 2: each CTA fetches HALF of every weight tile (block_n / 2 rows) and multicasts it into both
            // CTAs, so a weight byte leaves L2 once per pair of pixel tiles (the kernel is bound by L2 -> SM operand
            // traffic, profiles/r01_ncu_full_*_block4_v4.txt); the full barrier of each CTA counts both halves.
            uint32_t slot = 0, phase = 0;
            const int half_rows = p.block_n / p.cluster;
            TileFeed feed;
            for (int tile = feed.next(bars); tile < p.pair_tiles; tile = feed.next(bars)) {
                const TileCoord t = tile_coord(p, tile, rank);
                for (int ch = 0; ch < p.chunks; ++ch) {
                    const Chunk c = chunk_of(p, ch);
                    const CUtensorMap* map = c.width == 64 ? &tmB64 : &tmB32;
                    const uint32_t bytes = (uint32_t)p.block_n * c.width * 2;
                    const uint32_t half_bytes = (uint32_t)half_rows * c.width * 2;
                    const int k0 = c.ii * p.cj + c.jc0;
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        const int g = ch * 3 + dwi;
                        if (g < t.g0 || g >= t.g1) continue;
                        const int kw = 1 + (dwi - 1) * p.sign;
                        for (int dhi = 0; dhi < 3; ++dhi) {
                            const int kh = 1 + (dhi - 1) * p.sign;
                            mbar_wait(smem_u32(&bars->b_empty[slot]), phase ^ 1);
                            if (elect_one()) {
                                const uint32_t full = smem_u32(&bars->b_full[slot]);
                                mbar_expect_tx(full, bytes);
                                if (p.cluster > 1)
                                    tma_load_3d_mc(b_ring + slot * p.b_bytes + rank * half_bytes, map, full, k0,
                                                   t.n0 + rank * half_rows, kh * 3 + kw, mc_mask);
                                else
                                    tma_load_3d(b_ring + slot * p.b_bytes, map, full, k0, t.n0, kh * 3 + kw);
                            }
                            __syncwarp();
                            if (++slot == (uint32_t)p.nb) { slot = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp >= kWarpMma0 && warp < kWarpMma0 + p.issuers) {
        // ===================================================================== UMMA issuers
        // A K-group = one A box (chunk, dw) and its three vertical taps: 3 * (width/16) * MS MMAs, then one burst
        // of commits (3 weight slots + the A slot).  csrc/mma_bench.cu: a commit drains the issuing thread's MMAs
        // (~600 cycles) before that thread can issue again, but a second issuer warp overlaps the drain
        // (N=192: 1949 -> 3848 MAC/clk/SM).  So the issuer warps take alternate K-groups of the SAME tile.
        //
        // Ordering protocol (alias-free by construction): groups WAIT FOR THEIR DATA in strict global order.  The
        // owner of group g waits for `turn[me]` (arrived by the owner of group g-1 once ITS data had landed), then
        // for its own group's full barriers, passes the turn, issues and commits (= drains).  By induction every
        // earlier phase of every ring slot is complete when an owner waits, so a parity wait can never see a stale
        // phase — even if an issuer warp is starved for microseconds by co-resident side-stream kernels (this was
        // an intermittent deadlock with a looser scheme).  Nobody waits on a barrier phase it does not consume.
        const int me = warp - kWarpMma0;
        const uint32_t idesc = make_idesc_bf16(128, p.block_n, 0, 0);
        uint32_t aslot = 0, aphase = 0, bslot = 0, bphase = 0, gidx = 0, turn_waits = 0;
        int it = 0;
        const bool prof = p.prof != nullptr && me == 0;
        long long t_start = prof ? clk() : 0, w_a = 0, w_b = 0, w_t = 0, t0 = 0;
        TileFeed feed;
        for (int tile = feed.next(bars); tile < p.pair_tiles; tile = feed.next(bars), ++it) {
            const int buf = it % p.acc_bufs;
            const uint32_t acc_phase = (uint32_t)(it / p.acc_bufs) & 1u;
            if (prof) t0 = clk();
            mbar_wait(smem_u32(&bars->tmem_empty[buf]), acc_phase ^ 1);
            if (prof) w_t += clk() - t0;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * p.ms * p.block_n;
            const TileCoord tc = tile_coord(p, tile, rank);
            bool first_group = true;
            for (int ch = 0; ch < p.chunks; ++ch) {
                const Chunk c = chunk_of(p, ch);
                const uint32_t row_b = c.width == 64 ? 1024u : 512u;   // bytes per image row of the box == SBO
                const uint32_t lay = c.width == 64 ? SWZ_128B : SWZ_64B;
                const int ksl = c.width / 16;
                for (int dwi = 0; dwi < 3; ++dwi) {
                    if (ch * 3 + dwi < tc.g0 || ch * 3 + dwi >= tc.g1) continue;     // split K: not this work item's group
                    const bool mine = (gidx % (uint32_t)p.issuers) == (uint32_t)me;
                    uint32_t b_s[3], b_bar[3], b_full[3], b_par[3];
                    uint32_t bs = bslot, bp = bphase;
#pragma unroll
                    for (int dhi = 0; dhi < 3; ++dhi) {
                        b_s[dhi] = b_ring + bs * p.b_bytes;
                        b_bar[dhi] = smem_u32(&bars->b_empty[bs]);
                        b_full[dhi] = smem_u32(&bars->b_full[bs]);
                        b_par[dhi] = bp;
                        if (++bs == (uint32_t)p.nb) { bs = 0; bp ^= 1; }
                    }
                    if (mine) {
                        if (p.issuers > 1 && gidx > 0) {
                            mbar_wait(smem_u32(&bars->turn[me]), turn_waits & 1u);   // group gidx-1 has been issued
                            ++turn_waits;
                        }
                        if (prof) t0 = clk();
                        mbar_wait(smem_u32(&bars->a_full[aslot]), aphase);
                        if (prof) { const long long t1 = clk(); w_a += t1 - t0; t0 = t1; }
#pragma unroll
                        for (int dhi = 0; dhi < 3; ++dhi) mbar_wait(b_full[dhi], b_par[dhi]);
                        if (prof) w_b += clk() - t0;
                        tc_fence_after();
                        const uint32_t a_s = a_ring + aslot * p.a_bytes;
                        if (elect_one()) {
                            // The turn only has to guarantee that every earlier group has finished WAITING for its
                            // data, so it is passed on before issuing (both issuers then issue concurrently; issue
                            // rate, not the tensor pipe, limits N = 96 MMAs).  Exception: a tile's first group
                            // overwrites the accumulator and must reach the tensor pipe first.
                            // Decode plans (FPROP_INFER) keep strict issue order instead: the accumulation order, hence
                            // every output bit, is then reproducible run to run (a decoder must be deterministic).
                            const bool early = !first_group && p.mode != ONR_CONV_FPROP_INFER && p.mode != ONR_CONV_FPROP_HEAD;
                            if (p.issuers > 1 && early) mbar_arrive(smem_u32(&bars->turn[me ^ 1]));
#pragma unroll
                            for (int dhi = 0; dhi < 3; ++dhi) {
                                for (int k = 0; k < ksl; ++k) {
                                    const uint64_t bdesc = make_smem_desc(b_s[dhi] + k * 32, 16, row_b, lay);
                                    for (int ms = 0; ms < p.ms; ++ms) {
                                        // rows of sub-tile ms for vertical tap dh: box rows (dhi + 16*ms) ...
                                        const uint64_t adesc =
                                            make_smem_desc(a_s + (dhi + kSubH * ms) * row_b + k * 32, 16, row_b, lay);
                                        umma_bf16(d_tmem + ms * p.block_n, adesc, bdesc, idesc,
                                                  (first_group && dhi == 0 && k == 0) ? 0u : 1u);
                                    }
                                }
                            }
                            if (p.issuers > 1 && !early) mbar_arrive(smem_u32(&bars->turn[me ^ 1]));
#pragma unroll
                            for (int dhi = 0; dhi < 3; ++dhi) {
                                if (p.cluster > 1) umma_commit_mc(b_bar[dhi], mc_mask);   // frees the slot in both CTAs
                                else umma_commit(b_bar[dhi]);
                            }
                            umma_commit(smem_u32(&bars->a_empty[aslot]));
                        }
                        __syncwarp();
                    }
                    bslot = bs;
                    bphase = bp;
                    if (++aslot == (uint32_t)p.na) { aslot = 0; aphase ^= 1; }
                    ++gidx;
                    first_group = false;
                }
            }
            if (elect_one()) umma_commit(smem_u32(&bars->tmem_full[buf]));   // one arrival per issuer
            __syncwarp();
        }
        if (prof && lane == 0) {
            long long* o = p.prof + (size_t)blockIdx.x * 8;
            o[0] = clk() - t_start; o[1] = w_a; o[2] = w_b; o[3] = w_t; o[7] = it;
        }
    } else if (warp < kEpiWarps) {
        // ===================================================================== epilogue (warps 0..7)
        // Two warps per TMEM lane quarter: warp half 0 handles columns [0,32) of every 64-column group,
        // half 1 columns [32,64).  One iteration = 64 accumulator columns -> one 128-byte-row store tile
        // (SWIZZLE_128B) when the output channel range allows it, otherwise two 64-byte-row tiles.
        const int ew = warp;
        const int q = warp & 3;             // TMEM lane quarter this warp may touch
        const int half = ew >> 2;           // warps 0..3 -> 0, warps 4..7 -> 1
        const int row = q * 32 + lane;      // accumulator row == pixel inside the 16 x 8 sub-tile
        const int hl = row >> 3, wl = row & 7;
        const bool store_thread = (ew == 0 && lane == 0);
        if (p.mode == ONR_CONV_FPROP_HEAD) {
            // block_n == out_cp: every tile holds all channels of one PixelShuffle sub-pixel, in channel order
            for (int i = threadIdx.x; i < 3 * p.block_n; i += kEpiThreads) {
                const int k = i / p.block_n, c = i - k * p.block_n;
                bars->head_w_s[k][c] = c < p.head_c ? __ldg(p.head_w + k * p.head_c + c) : 0.0f;
            }
            if (threadIdx.x < 3) bars->head_b_s[threadIdx.x] = __ldg(p.head_b + threadIdx.x);
            named_bar_sync(1, kEpiThreads);
        }
        uint32_t iter_ctr = 0;
        int it = 0;
        const bool prof = p.prof != nullptr && store_thread;
        long long e_full = 0, e_store = 0, e_busy = 0, t0 = 0, t1 = 0;
        TileFeed feed;
        for (int tile = feed.next(bars); tile < p.pair_tiles; tile = feed.next(bars), ++it) {
            const TileCoord t = tile_coord(p, tile, rank);
            const int buf = it % p.acc_bufs;
            const uint32_t acc_phase = (uint32_t)(it / p.acc_bufs) & 1u;
            if (prof) t0 = clk();
            if (p.mode != ONR_CONV_DGRAD) {
                // (a thread is never two tiles ahead of another — the per-iteration barriers below — so writing buffer
                //  it & 1 cannot race with readers of the previous tile; the first barrier of the loop publishes it)
                for (int i = threadIdx.x; i < p.block_n; i += kEpiThreads)
                    bars->bias_s[it & 1][i] = t.n0 + i < p.n_total ? __ldg(p.bias + t.n0 + i) : 0.0f;
            }
            mbar_wait(smem_u32(&bars->tmem_full[buf]), acc_phase);
            if (prof) { t1 = clk(); e_full += t1 - t0; }
            tc_fence_after();
            const int ngroups = (p.block_n + 63) / 64;
            for (int ms = 0; ms < p.ms; ++ms) {
                const int hs0 = t.h0 + ms * kSubH;
                if (hs0 >= p.H || t.skip) break;      // sub-tile entirely below the image / padding work item
                if (p.mode == ONR_CONV_FPROP_HEAD) {
                    // ---- fused RGB head: SiLU -> 1x1 conv C->3 -> bias -> (tanh+1)/2, straight from the accumulator
                    float pk[3] = {0.0f, 0.0f, 0.0f};
                    // the two warps of a lane quarter split the tile's columns evenly, in 16-column pieces:
                    // half 0 takes [0, block_n/2), half 1 the rest (block_n is a multiple of 32)
                    const int c_begin = half * (p.block_n / 2), c_end = c_begin + p.block_n / 2;
                    for (int col = c_begin; col < c_end; col += 16) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * p.ms + ms) * p.block_n + col, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float z = __uint_as_float(r[e]) + bars->bias_s[it & 1][col + e];
                            const float y = z * fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);
                            pk[0] = fmaf(y, bars->head_w_s[0][col + e], pk[0]);
                            pk[1] = fmaf(y, bars->head_w_s[1][col + e], pk[1]);
                            pk[2] = fmaf(y, bars->head_w_s[2][col + e], pk[2]);
                        }
                    }
                    bars->head_red[half][row][0] = pk[0];
                    bars->head_red[half][row][1] = pk[1];
                    bars->head_red[half][row][2] = pk[2];
                    named_bar_sync(1, kEpiThreads);
                    if (half == 0) {
                        const int h = hs0 + hl, w = t.w0 + wl;
                        if (h < p.H && w < p.W) {
                            // n tile index == sub-pixel (i, j) of the PixelShuffle
                            const int nt = t.n0 / p.block_n, si = nt / p.out_s, sj = nt - si * p.out_s;
                            const size_t Ho = (size_t)p.H * p.out_s, Wo = (size_t)p.W * p.out_s;
                            const size_t px = ((size_t)h * p.out_s + si) * Wo + (size_t)w * p.out_s + sj;
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const float v = bars->head_red[0][row][k] + bars->head_red[1][row][k] + bars->head_b_s[k];
                                const float o = p.use_sigmoid ? 1.0f / (1.0f + __expf(-v)) : (tanhf(v) + 1.0f) * 0.5f;
                                p.img[((size_t)t.b * 3 + k) * Ho * Wo + px] = o;
                            }
                        }
                    }
                    named_bar_sync(2, kEpiThreads);
                    continue;
                }
                for (int g = 0; g < ngroups; ++g) {
                    const int n_g = t.n0 + g * 64;              // first output channel of the group
                    if (n_g >= p.n_total) break;
                    const int col = g * 64 + half * 32;         // this warp's columns inside the tile
                    const int n = t.n0 + col;
                    const bool valid = col < p.block_n && n < p.n_total;
                    // wide store: all 64 channels valid, same shuffle row i, 64-aligned inside it
                    const int oi = n_g / p.out_jc;
                    const int ojc = n_g - oi * p.out_jc;
                    const bool wide = (g * 64 + 64 <= p.block_n) && (n_g + 64 <= p.n_total) && (ojc % 64 == 0) &&
                                      (ojc + 64 <= p.out_jc);
                    const uint32_t sbuf = 0;   // single staging buffer: the weight ring needs the shared memory more
                    ++iter_ctr;
                    if (prof) t0 = clk();
                    if (store_thread) tma_store_wait_read<0>();
                    if (prof) e_store += clk() - t0;
                    named_bar_sync(1, kEpiThreads);
                    const uint32_t ybuf = staging + sbuf * 2 * kStageOutBytes;
                    const uint32_t dbuf = ybuf + kStageOutBytes;
                    if (valid && p.ksplit > 1) {
                        // split K: add this work item's partial sums into the fp32 scratch; conv_split_finish_kernel
                        // applies the epilogue once every part has landed
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * p.ms + ms) * p.block_n + col, r);
                        tmem_ld_wait();
                        const int h = hs0 + hl, w = t.w0 + wl;
                        if (h < p.H && w < p.W) {
                            float* dst = p.kacc + ((size_t)(t.b * p.H + h) * p.W + w) * p.n_total + n;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst + j * 4),
                                             "f"(__uint_as_float(r[j * 4])), "f"(__uint_as_float(r[j * 4 + 1])),
                                             "f"(__uint_as_float(r[j * 4 + 2])), "f"(__uint_as_float(r[j * 4 + 3]))
                                             : "memory");
                        }
                    } else if (valid) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * p.ms + ms) * p.block_n + col, r);
                        tmem_ld_wait();
                        // staging address of this thread's 16-byte piece j (0..3) of its 32 channels
                        //   wide  : [128 rows][128 B], SWIZZLE_128B: piece (4*half + j) ^ (row & 7)
                        //   narrow: tile `half` of [128 rows][64 B], SWIZZLE_64B: piece j ^ ((row >> 1) & 3)
                        const uint32_t rbase = wide ? row * 128u : half * (kStageOutBytes / 2) + row * 64u;
                        const uint32_t sx = wide ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                        const uint32_t jofs = wide ? half * 4u : 0u;
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
                                const uint32_t addr = ybuf + rbase + (((jofs + j) ^ sx) << 4);
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(o[0]),
                                             "r"(o[1]), "r"(o[2]), "r"(o[3])
                                             : "memory");
                            }
                        } else {
                            const float4* bp = reinterpret_cast<const float4*>(&bars->bias_s[it & 1][col]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 b0 = bp[j * 2], b1 = bp[j * 2 + 1];
                                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                                float yv[8], dv[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float z = __uint_as_float(r[j * 8 + e]) + bb[e];
                                    if (p.mode == ONR_CONV_FPROP_Z) {      // pre-activation only (uniform branch)
                                        yv[e] = z;
                                        dv[e] = 0.0f;
                                    } else {
                                        const float sg = fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);
                                        const float y = z * sg;
                                        yv[e] = y;
                                        dv[e] = fmaf(y, 1.0f - sg, sg);
                                    }
                                }
                                const uint32_t off = rbase + (((jofs + j) ^ sx) << 4);
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
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, kEpiThreads);
                    if (store_thread && p.ksplit == 1) {
                        const bool train = p.mode == ONR_CONV_FPROP_TRAIN;
                        if (wide) {
                            tma_store_5d(&tmY64, ybuf, ojc, t.w0, oi, hs0, t.b);
                            if (train) tma_store_5d(&tmD64, dbuf, ojc, t.w0, oi, hs0, t.b);
                        } else {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int colh = g * 64 + hh * 32;
                                const int nh = t.n0 + colh;
                                if (colh < p.block_n && nh < p.n_total) {
                                    const int oih = nh / p.out_jc;
                                    const int ojh = nh - oih * p.out_jc;
                                    tma_store_5d(&tmY32, ybuf + hh * (kStageOutBytes / 2), ojh, t.w0, oih, hs0, t.b);
                                    if (train)
                                        tma_store_5d(&tmD32, dbuf + hh * (kStageOutBytes / 2), ojh, t.w0, oih, hs0, t.b);
                                }
                            }
                        }
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

    // the peer may still multicast weight tiles into this CTA's ring / arrive on its barriers until it is done as well
    if (p.cluster > 1) cluster_sync_all();
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (p.dynamic && threadIdx.x == 0) {
        // the last CTA to finish re-arms the counters for the next launch (no memset node per launch)
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
        }
    }
}

// Split-K dgrad, second half: out = bf16(partial sums x SiLU'(z) of the producer), NHWC with out_s == 1.
__global__ void conv_split_finish_kernel(const float* __restrict__ acc, const __nv_bfloat16* __restrict__ dmul,
                                         __nv_bfloat16* __restrict__ out, size_t n) {
    for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2; i < n;
         i += (size_t)gridDim.x * blockDim.x * 2) {
        const float2 a = *reinterpret_cast<const float2*>(acc + i);
        const uint32_t d = *reinterpret_cast<const uint32_t*>(dmul + i);
        *reinterpret_cast<uint32_t*>(out + i) = pack_bf16x2(a.x * bf16_lo(d), a.y * bf16_hi(d));
    }
}

}  // namespace onr

// ------------------------------------------------------------------------------------------- host
struct onr_conv_plan {
    CUtensorMap tmA64, tmA32, tmB64, tmB32, tmY64, tmY32, tmD64, tmD32;
    onr::ConvParams p;
    int grid;
    size_t smem;
    int* sched = nullptr;      // two device counters of the dynamic work-item feed (owned by the plan)
    float* kacc = nullptr;     // split-K scratch (owned by the plan)
    size_t kacc_elems = 0;
    void* out = nullptr;       // split-K: final destination of the finishing kernel
    ~onr_conv_plan() {
        if (sched) cudaFree(sched);
        if (kacc) cudaFree(kacc);
    }
};

extern "C" {

// N tiling rule shared with the weight packer: block_n is a multiple of 32, <= 256 (one UMMA per K slice and room
// for two accumulator buffers), chosen to minimise operand bytes per useful MAC:
//   bytes/tile ~ 3 * (16*ms + 2) * 8  (A box rows per 32 channels)  +  9 * block_n  (weight rows)
//   MACs/tile  ~ ms * 128 * block_n * 9,   ms = min(4, 256 / block_n) sub-tiles,   x padding waste block_n*n_tiles/N.
// (384 -> 3 x 128 with 2 sub-tiles; 96 -> 96 with 2 sub-tiles; 800 -> 7 x 128.)
int onr_conv_tile_n(int n_total, int* block_n, int* n_tiles) {
    ONR_REQUIRE(n_total > 0 && n_total % 32 == 0, "n_total must be a positive multiple of 32");
    double best = 1e300;
    int best_bn = 32, best_nt = n_total / 32;
    for (int bn = 32; bn <= onr::kMaxBlockN; bn += 32) {
        const int nt = (n_total + bn - 1) / bn;
        if (nt > 1 && bn < 64) continue;                      // tiny N tiles only for tiny N
        int ms = 256 / bn;
        if (ms > onr::kMaxMS) ms = onr::kMaxMS;
        const double bytes = 3.0 * (16 * ms + 2) * 8 + 9.0 * bn;
        const double macs = (double)ms * 128 * bn * 9;
        const double cost = bytes / macs * ((double)bn * nt / n_total);
        if (cost < best * 0.9999) { best = cost; best_bn = bn; best_nt = nt; }
    }
    if (block_n) *block_n = best_bn;
    if (n_tiles) *n_tiles = best_nt;
    return 0;
}

int onr_conv_plan_create(onr_conv_plan** out, const onr_conv_desc* d) {
    using namespace onr;
    ONR_REQUIRE(out && d, "null argument");
    ONR_REQUIRE(d->kind >= 0 && d->kind <= 4, "bad conv kind %d", d->kind);
    ONR_REQUIRE(d->a_cp % 32 == 0 && d->out_cp % 32 == 0 && d->n_total % 32 == 0,
                "channel counts must be multiples of 32 (a_cp %d out_cp %d n %d)", d->a_cp, d->out_cp,
                d->n_total);
    ONR_REQUIRE(d->n_total == d->out_s * d->out_s * d->out_cp, "n_total must equal out_s^2*out_cp");
    ONR_REQUIRE(d->B >= 1 && d->H >= 1 && d->W >= 1, "bad grid");
    int block_n = 0, n_tiles = 0;
    onr_conv_tile_n(d->n_total, &block_n, &n_tiles);
    if (d->kind == ONR_CONV_FPROP_HEAD) {
        // one N tile per PixelShuffle sub-pixel, so that an accumulator row holds every channel of one output pixel
        ONR_REQUIRE(d->out_cp <= kMaxBlockN && d->head_w && d->head_b && d->img && d->head_c > 0 && d->head_c <= d->out_cp,
                    "fused head needs out_cp <= %d and the head tensors", kMaxBlockN);
        block_n = d->out_cp;
        n_tiles = d->out_s * d->out_s;
    }
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
    int ms_max = 256 / block_n;          // always leave room for two accumulator buffers
    if (ms_max < 1) ms_max = 1;
    if (ms_max > kMaxMS) ms_max = kMaxMS;
    const int k_tap0 = d->a_s * d->a_s * d->a_cp;
    int ms = 1;
    {
        // pick the sub-tile count that minimises (waves over the SMs) x (operand bytes per CTA tile)
        double best = 1e300;
        for (int m = 1; m <= ms_max; ++m) {
            const long long tiles = (long long)d->B * ceil_div(d->H, kSubH * m) * ceil_div(d->W, kSubW) * n_tiles;
            const double waves = (double)((tiles + num_sms() - 1) / num_sms());
            const double bytes = (double)(k_tap0 / 32) * (3.0 * (kSubH * m + 2) * kSubW * 64 + 9.0 * block_n * 64);
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
    p.px_tiles = d->B * p.tiles_h * p.tiles_w;
    // CTA pairs that share every weight tile by TMA multicast: ONR_CONV_CLUSTER=2
    // (measured on B200: 0.169 vs 0.167 ms for block 4 — the kernel is not bound by L2 reads, each SM still ingests
    //  every weight byte — so the pairing is off unless asked for)
    p.cluster = 1;
    if (const char* e = getenv("ONR_CONV_CLUSTER")) {
        const int v = atoi(e);
        if (v == 1 || (v == 2 && block_n % 32 == 0)) p.cluster = v;
    }
    p.pair_tiles = ceil_div(p.px_tiles, p.cluster) * n_tiles;
    // Split K for the data gradient of a SMALL layer: block 0's dgrad is two tiles with K = 9 x 800 channels — two
    // CTAs issuing 540 N=32 MMAs each while 146 SMs idle (46 us, on the critical path).  Its (chunk, dw) K-groups are
    // dealt out to ksplit work items per tile (>= 3 groups each) whose partial sums meet in an fp32 scratch
    // (red.global.add); conv_split_finish_kernel applies the x SiLU' epilogue.  Measured on B200: the kernel itself goes
    // from 43 to 20 us (block 0) and 23 to 20 us (block 1), but the training step does not move (1.2045 vs 1.1992 ms):
    // by then the step waits for the side-stream chains (wgrad -> exchange -> fold backward of blocks 1 and 0), not for
    // this kernel.  Hence opt-in: ONR_CONV_KSPLIT=1.
    p.ksplit = 1;
    p.kacc = nullptr;
    if (d->kind == ONR_CONV_DGRAD && d->out_s == 1 && p.cluster == 1 && p.total_tiles * 2 <= num_sms() &&
        getenv("ONR_CONV_KSPLIT") && atoi(getenv("ONR_CONV_KSPLIT")) == 1) {
        const int groups = d->a_s * (d->a_s * d->a_cp / 64 + ((d->a_s * d->a_cp) % 64 ? 1 : 0)) * 3;
        int ks = num_sms() / p.total_tiles;
        if (ks > groups / 3) ks = groups / 3;
        if (ks > 1) p.ksplit = ks;
    }
    p.pair_tiles *= p.ksplit;
    p.dynamic = 0;      // measured on B200: 1.193 (dynamic) vs 1.172 ms/step (static) - the static order keeps the L2 locality; knob only
    if (const char* e = getenv("ONR_CONV_DYNAMIC")) p.dynamic = (atoi(e) != 0 && p.cluster == 1) ? 1 : 0;
    p.sched = nullptr;
    const int k_tap = d->a_s * d->a_s * d->a_cp;
    p.cj = d->a_s * d->a_cp;                      // channels per shuffle row i of the A view
    p.n64 = p.cj / 64;
    p.cpi = p.n64 + ((p.cj % 64) ? 1 : 0);
    p.chunks = d->a_s * p.cpi;
    p.sign = d->kind == ONR_CONV_DGRAD ? -1 : 1;
    p.n_total = d->n_total;
    p.acc_bufs = 2;
    p.issuers = kMaxIssuers;
    if (const char* e = getenv("ONR_CONV_ISSUERS")) {     // experiments only
        const int v = atoi(e);
        if (v >= 1 && v <= kMaxIssuers) p.issuers = v;
    }
    p.mode = d->kind;
    p.out_jc = d->out_s * d->out_cp;
    p.bias = d->bias_p;
    p.dmul = reinterpret_cast<const __nv_bfloat16*>(d->dmul);
    p.head_w = d->head_w; p.head_b = d->head_b; p.img = d->img;
    p.head_c = d->head_c; p.use_sigmoid = d->use_sigmoid; p.out_s = d->out_s;
    p.prof = nullptr;
    const int box_h = kSubH * ms + 2;
    const int wmax = p.n64 > 0 ? 64 : 32;               // widest chunk this plan uses
    p.a_bytes = box_h * kSubW * wmax * 2;               // multiple of 1024 (64-wide) / 512 (32-wide)
    p.a_bytes = (p.a_bytes + 1023) / 1024 * 1024;       // keep every slot 1024-aligned
    p.b_bytes = block_n * wmax * 2;                     // multiple of 2048
    // ring depths: ~96 KB of A boxes at most, the rest for weight tiles
    const int budget = kMaxDynSmem - 1024 - 2 * kStageOutBytes - (int)sizeof(SmemBarriers) - 1024;
    // one A box feeds three weight tiles, so two or three A slots are enough; weight tiles get the rest
    int na = 3;
    while (na > 2 && (na * p.a_bytes > budget / 2 || (budget - na * p.a_bytes) / p.b_bytes < 6)) --na;
    int nb = (budget - na * p.a_bytes) / p.b_bytes;
    if (nb > kMaxRing) nb = kMaxRing;
    ONR_REQUIRE(nb >= 3, "conv plan: shared memory too small for block_n %d ms %d", block_n, ms);
    // spend what is left on a deeper A ring
    while (na < 4 && (na + 1) * p.a_bytes + nb * p.b_bytes <= budget) ++na;
    p.na = na;
    p.nb = nb;
    pl->smem = 1024 + (size_t)na * p.a_bytes + (size_t)nb * p.b_bytes + 2 * kStageOutBytes + sizeof(SmemBarriers);
    {
        const int max_clusters = num_sms() / p.cluster;
        pl->grid = (p.pair_tiles < max_clusters ? p.pair_tiles : max_clusters) * p.cluster;
    }
    // 64-channel maps only exist when the channel extent allows them; otherwise they alias the 32-wide ones
    const int a64 = p.cj >= 64 ? 64 : 32, o64 = p.out_jc >= 64 ? 64 : 32, b64 = k_tap >= 64 ? 64 : 32;
    const void* out_ptr = d->kind == ONR_CONV_FPROP_HEAD ? d->a : d->out;     // (fused head: no bf16 output; maps unused)
    const void* outd = d->kind == ONR_CONV_FPROP_TRAIN ? d->out_d : out_ptr;
    int rc = make_act_tmap(&pl->tmA64, d->a, d->B, d->H, d->W, d->a_cp, d->a_s, kSubW, box_h, a64);
    if (!rc) rc = make_act_tmap(&pl->tmA32, d->a, d->B, d->H, d->W, d->a_cp, d->a_s, kSubW, box_h, 32);
    if (!rc) rc = make_weight_tmap(&pl->tmB64, d->w, 9, d->n_rows, k_tap, block_n / p.cluster, b64);
    if (!rc) rc = make_weight_tmap(&pl->tmB32, d->w, 9, d->n_rows, k_tap, block_n / p.cluster, 32);
    if (!rc) rc = make_act_tmap(&pl->tmY64, out_ptr, d->B, d->H, d->W, d->out_cp, d->out_s, kSubW, kSubH, o64);
    if (!rc) rc = make_act_tmap(&pl->tmY32, out_ptr, d->B, d->H, d->W, d->out_cp, d->out_s, kSubW, kSubH, 32);
    if (!rc) rc = make_act_tmap(&pl->tmD64, outd, d->B, d->H, d->W, d->out_cp, d->out_s, kSubW, kSubH, o64);
    if (!rc) rc = make_act_tmap(&pl->tmD32, outd, d->B, d->H, d->W, d->out_cp, d->out_s, kSubW, kSubH, 32);
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
    if (p.ksplit > 1) {
        pl->kacc_elems = (size_t)d->B * d->H * d->W * d->n_total;
        cudaError_t e = cudaMalloc(&pl->kacc, pl->kacc_elems * sizeof(float));
        if (e != cudaSuccess) {
            set_error("conv plan: split-K scratch allocation failed: %s", cudaGetErrorString(e));
            delete pl;
            return (int)e;
        }
        p.kacc = pl->kacc;
        pl->out = d->out;
        pl->grid = p.pair_tiles < num_sms() ? p.pair_tiles : num_sms();
    }
    if (p.dynamic) {
        cudaError_t e = cudaMalloc(&pl->sched, 2 * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(pl->sched, 0, 2 * sizeof(int));
        if (e != cudaSuccess) {
            set_error("conv plan: counter allocation failed: %s", cudaGetErrorString(e));
            delete pl;
            return (int)e;
        }
        p.sched = pl->sched;
    }
    *out = pl;
    return 0;
}

int onr_conv_plan_run(const onr_conv_plan* pl, void* stream) {
    using namespace onr;
    ONR_REQUIRE(pl != nullptr, "null plan");
    if (pl->p.ksplit > 1) ONR_CUDA(cudaMemsetAsync(pl->kacc, 0, pl->kacc_elems * sizeof(float), (cudaStream_t)stream));
    if (pl->p.cluster > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pl->grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = pl->smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pl->p.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ONR_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel, pl->tmA64, pl->tmA32, pl->tmB64, pl->tmB32, pl->tmY64,
                                    pl->tmY32, pl->tmD64, pl->tmD32, pl->p));
    } else {
        conv_igemm_kernel<<<pl->grid, kThreads, pl->smem, (cudaStream_t)stream>>>(
            pl->tmA64, pl->tmA32, pl->tmB64, pl->tmB32, pl->tmY64, pl->tmY32, pl->tmD64, pl->tmD32, pl->p);
    }
    ONR_LAUNCH_CHECK();
    if (pl->p.ksplit > 1) {
        size_t blocks = (pl->kacc_elems / 2 + 255) / 256;
        if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
        conv_split_finish_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
            pl->kacc, pl->p.dmul, reinterpret_cast<__nv_bfloat16*>(pl->out), pl->kacc_elems);
        ONR_LAUNCH_CHECK();
    }
    return 0;
}

int onr_conv_plan_set_head(onr_conv_plan* pl, float* img, const float* head_w, const float* head_b) {
    ONR_REQUIRE(pl != nullptr && pl->p.mode == ONR_CONV_FPROP_HEAD && img && head_w && head_b, "set_head: not a fused-head plan");
    pl->p.img = img;
    pl->p.head_w = head_w;
    pl->p.head_b = head_b;
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
