// fold_tc.cu — ERB online structural re-parameterisation on the tensor cores, fp32-accurate.
//
// Same arithmetic as fold.cu (reference model.py:450-516 and the gradients autograd derives for it, SURVEY.md
// 8a-A3), but the six contractions of a block run as tcgen05 GEMMs in 3xTF32 split precision:
//     x ~= hi + lo,  hi = rn_tf32(x),  lo = rn_tf32(x - hi)      (both exactly TF32 numbers, |x - hi - lo| <= 2^-24 |x|)
//     a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b          (fp32 accumulation in TMEM; error ~2^-22 per product, unbiased)
// which keeps the 1e-5 relative gate of north_star with two decimal orders of margin, where a single TF32 or bf16
// pass would not.  L720's block 0 (Cin 112 -> Cout 2800) makes this 8.5 GMAC per fold: far out of reach of SIMT FMAs.
//
// Internal layout: every 4-D tensor X[o][i][kh][kw] is handled "tap-major", X'[o][x], x = (kh*3+kw)*Cin + i, which is
// the K-major operand order of the convolution kernels (pack / wgrad un-pack become plain copies) and makes every
// GEMM operand a dense K-major matrix that TMA can fetch:
//   forward   F1: Tt[x][o]   = sum_m W2p[(hw,o)][m] * W1t[i][m]           (M = 9 Cout, N = Cin,   K = 2 Cin)
//             F2: Kt[p][x]  += sum_o W3[p][o]       * Tt[x][o]            (M = Cout,   N = 9 Cin, K = Cout)
//   backward  B1: gW3[p][o]  = sum_x dKq[p][x]      * Tq[o][x]            (M = Cout,   N = Cout,  K = 9 Cin)
//             B2: dT[o][x]   = sum_p W3t[o][p]      * dKt[x][p]           (M = Cout,   N = 9 Cin, K = Cout)
//             B3: gW2[o][m][hw]  = sum_i dTa[(o,hw)][i] * W1[m][i]        (M = 9 Cout, N = 2 Cin, K = Cin)
//             B4: gW1[m][i]  = sum_(hw,o) W2c[m][(hw,o)] * dTb[i][(hw,o)] (M = 2 Cin,  N = Cin,   K = 9 Cout)
// A GEMM's epilogue writes its result through up to two two-level affine index maps, optionally already split into
// hi/lo and/or transposed, i.e. directly in the operand layout of the GEMM that consumes it — no intermediate
// reshuffling kernels.  Split-K: the CTAs that share an output tile form a thread-block cluster (2, 4 or 8 CTAs); each
// parks its partial tile in its own shared memory and then reduces one row slice of the tile over distributed shared
// memory in rank order — no global workspace, no atomics, no extra launch, and every result is bit-reproducible
// (replicas of a data-parallel run stay bit-identical; a deploy checkpoint reproduces the train-state decode).
#include "onr_common.cuh"
#include "onr_ptx.cuh"

#include <stdlib.h>

namespace onr {

constexpr int kFtStages = 3;
// warp 0 = TMA producer, warp 1 = UMMA issuer (+ TMEM owner), warps 2..17 = epilogue.  Sixteen epilogue warps: the
// write-out is a long chain of dependent address arithmetic per element, and with one warp per scheduler every
// instruction latency of it was exposed (ONR_FOLD_PROF=1: the write-out took 4x the GEMM); four warps per scheduler
// hide it.
constexpr int kFtEpiWarps = 16;
constexpr int kFtEpiThreads = kFtEpiWarps * 32;
constexpr int kFtThreads = 64 + kFtEpiThreads;
constexpr int kFtKBox = 32;       // fp32 elements per pipeline stage and operand row: 128-byte rows, SWIZZLE_128B
constexpr int kFtMaxBn = 128;
constexpr int kFtNoDiv = 1 << 30;   // "this index is not split" in an output map

struct FtOut {
    float* hi;          // destination (fp32 result, or the hi part when lo != NULL)
    float* lo;          // NULL, or destination of the lo part (same offsets)
    int rdiv, cdiv;     // offset(r, c) = (r / rdiv) * rs_hi + (r % rdiv) * rs_lo + (c / cdiv) * cs_hi + (c % cdiv) * cs_lo
    long long rs_hi, rs_lo, cs_hi, cs_lo;
    int accumulate;     // dst += value (fp32 destinations only)
};

struct FtParams {
    int M, N, K;
    int bn, n_tiles, m_tiles, splits, k_stages, tmem_cols, n_out;
    FtOut out[2];
    long long* prof;      // debugging aid (ONR_FOLD_PROF=1): per-CTA clock64 stamps [grid][8], else NULL
};

// The tensor core adds into its fp32 accumulator with truncation: over a long reduction that is a bias which grows
// linearly with the number of accumulating MMAs (measured: 4e-5 of the result after ~1200 of them, L720 block 0).  So the
// accumulator is "promoted" every kFtChunk pipeline stages (48 MMAs): the epilogue warps add the TMEM chunk into fp32
// registers (round-to-nearest adds) while the issuer continues in the other TMEM buffer.
constexpr int kFtChunk = 4;

struct __align__(8) FtBarriers {
    uint64_t full[kFtStages], empty[kFtStages], chunk_full[2], chunk_empty[2];
    uint32_t tmem_base;
};

// x = hi + lo with BOTH parts exactly representable in TF32 and both roundings to nearest: truncating instead (or
// letting the tensor core drop the low 13 bits of lo) biases every product by ~1e-7 of its magnitude in the same
// direction, which adds up to a few 1e-5 of the result over the K = 25 200 reduction of L720's block 0 (measured).
__device__ __forceinline__ float rna_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = rna_tf32(v);
    lo = rna_tf32(v - hi);
}

// Column walker of one output map: offset of column c without a division per element.
struct FtCol {
    long long off;
    int lo;     // c % cdiv
};
__device__ __forceinline__ FtCol ft_col_begin(const FtOut& q, long long roff, int c) {
    FtCol w;
    const int hi = c / q.cdiv;
    w.lo = c - hi * q.cdiv;
    w.off = roff + (long long)hi * q.cs_hi + (long long)w.lo * q.cs_lo;
    return w;
}
__device__ __forceinline__ void ft_col_step(const FtOut& q, FtCol& w, int step) {
    w.lo += step;
    w.off += (long long)step * q.cs_lo;
    while (w.lo >= q.cdiv) {
        w.lo -= q.cdiv;
        w.off += q.cs_hi - (long long)q.cdiv * q.cs_lo;
    }
}
__device__ __forceinline__ void ft_store(const FtOut& q, long long off, float v) {
    if (q.lo != nullptr) {
        float h, l;
        split_tf32(v, h, l);
        q.hi[off] = h;
        q.lo[off] = l;
    } else if (q.accumulate) {
        q.hi[off] += v;
    } else {
        q.hi[off] = v;
    }
}

__global__ void __launch_bounds__(kFtThreads, 1)
fold_tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                    const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                    const FtParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_bytes = 128u * 128u, b_bytes = (uint32_t)p.bn * 128u;
    const uint32_t stage_bytes = 2u * a_bytes + 2u * b_bytes;          // A hi | A lo | B hi | B lo
    FtBarriers* bars =
        reinterpret_cast<FtBarriers*>(smem_raw + (smem_base - smem_u32(smem_raw)) + kFtStages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
    const int ks0 = (int)((long long)split * p.k_stages / p.splits);
    const int ks1 = (int)((long long)(split + 1) * p.k_stages / p.splits);
    const long long t_begin = p.prof ? clock64() : 0;
#define FT_STAMP(i) do { if (p.prof) p.prof[(size_t)blockIdx.x * 8 + (i)] = clock64() - t_begin; } while (0)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kFtStages; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->chunk_full[b]), 1);
            mbar_init(smem_u32(&bars->chunk_empty[b]), kFtEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmAh);
        tma_prefetch_desc(&tmAl);
        tma_prefetch_desc(&tmBh);
        tma_prefetch_desc(&tmBl);
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) FT_STAMP(0);                      // set-up done
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t part_pitch = (uint32_t)p.bn + 4u;      // floats; rows stay 16-byte aligned for the 128-bit remote loads

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        for (int ks = ks0, it = 0; ks < ks1; ++ks, ++it) {
            const uint32_t stage = it % kFtStages, phase = (it / kFtStages) & 1u;
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
            if (elect_one()) {
                const uint32_t full = smem_u32(&bars->full[stage]);
                const uint32_t s0 = smem_base + stage * stage_bytes;
                mbar_expect_tx(full, stage_bytes);
                tma_load_2d(s0, &tmAh, full, ks * kFtKBox, mt * 128);
                tma_load_2d(s0 + a_bytes, &tmAl, full, ks * kFtKBox, mt * 128);
                tma_load_2d(s0 + 2 * a_bytes, &tmBh, full, ks * kFtKBox, nt * p.bn);
                tma_load_2d(s0 + 2 * a_bytes + b_bytes, &tmBl, full, ks * kFtKBox, nt * p.bn);
            }
            __syncwarp();
        }
        if (lane == 0) FT_STAMP(1);                         // every TMA issued
    } else if (warp == 1) {
        // ------------------------------------------------------------------ UMMA issuer: 3xTF32
        const uint32_t idesc = make_idesc_tf32(128, (uint32_t)p.bn);
        for (int ks = ks0, it = 0; ks < ks1; ++ks, ++it) {
            const uint32_t stage = it % kFtStages, phase = (it / kFtStages) & 1u;
            const int ch = it / kFtChunk, first = (it % kFtChunk) == 0;
            const uint32_t d_tmem = tmem_base + (uint32_t)(ch & 1) * (uint32_t)(p.tmem_cols / 2);
            if (first) {
                mbar_wait(smem_u32(&bars->chunk_empty[ch & 1]), ((uint32_t)(ch >> 1) & 1u) ^ 1u);
                tc_fence_after();
            }
            mbar_wait(smem_u32(&bars->full[stage]), phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ah = smem_base + stage * stage_bytes, al = ah + a_bytes;
                const uint32_t bh = ah + 2 * a_bytes, bl = bh + b_bytes;
#pragma unroll
                for (int k = 0; k < kFtKBox / 8; ++k) {
                    const uint64_t dah = make_smem_desc(ah + k * 32, 16, 1024, SWZ_128B);
                    const uint64_t dal = make_smem_desc(al + k * 32, 16, 1024, SWZ_128B);
                    const uint64_t dbh = make_smem_desc(bh + k * 32, 16, 1024, SWZ_128B);
                    const uint64_t dbl = make_smem_desc(bl + k * 32, 16, 1024, SWZ_128B);
                    // small terms first, then the leading product
                    umma_tf32(d_tmem, dal, dbh, idesc, (!first || k > 0) ? 1u : 0u);
                    umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                    umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                }
                umma_commit(smem_u32(&bars->empty[stage]));
                if ((it % kFtChunk) == kFtChunk - 1 || ks == ks1 - 1) umma_commit(smem_u32(&bars->chunk_full[ch & 1]));
            }
            __syncwarp();
        }
        if (lane == 0) FT_STAMP(2);                         // every MMA issued
    } else {
        // ------------------------------------------------------------------ epilogue, part 1 (warps 2..5)
        // The accumulator tile is parked in shared memory first (thread = row; the operand ring is idle by now: every
        // MMA has completed and no TMA is in flight); part 2 then writes it out with the lanes of a warp running along
        // whichever index is contiguous in each destination, so that every global store is coalesced.  (Writing
        // straight from the TMEM row-per-thread layout made 32 sectors per store instruction for the column-contiguous
        // destinations and took 8x longer than the GEMM itself; ONR_FOLD_PROF=1 stamps.)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        // warp w may only touch TMEM lane quarter w % 4; the four warps of a quarter take one 32-column chunk each
        // (block_n <= 128)
        const int c0 = 32 * ((warp - 2) >> 2);
        const bool has_cols = c0 < p.bn;
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
        const int n_chunks = (ks1 - ks0 + kFtChunk - 1) / kFtChunk;
        for (int ch = 0; ch < n_chunks; ++ch) {
            mbar_wait(smem_u32(&bars->chunk_full[ch & 1]), (uint32_t)(ch >> 1) & 1u);
            tc_fence_after();
            if (has_cols) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch & 1) * (uint32_t)(p.tmem_cols / 2) +
                                       (uint32_t)c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->chunk_empty[ch & 1]));
        }
        if (threadIdx.x == 64) FT_STAMP(3);                 // accumulator complete
        float* mine = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
        if (has_cols) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < p.bn) mine[row * part_pitch + c0 + j] = acc[j];
        }
    }
    if (threadIdx.x == 64) FT_STAMP(4);                     // tile parked

    // rows [r0, r0 + nrows) of the tile are this CTA's to write
    int r0 = 0, nrows = 128;
    if (p.splits > 1) {
        // split-K: every CTA of the cluster reduces rows [split * 128/S, (split+1) * 128/S) of the tile, reading the S
        // partials over distributed shared memory in rank order, and leaves the sums in its own copy of those rows
        // (which no peer reads)
        cluster_sync_all();
        if (threadIdx.x == 64) FT_STAMP(5);                 // cluster barrier passed
        nrows = 128 / p.splits;
        r0 = split * nrows;
        if (warp >= 2) {
            float* mine = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
            const int t = (int)threadIdx.x - 64;
            const int q4 = p.bn / 4, total = nrows * q4;      // 128-bit pieces: a remote load costs the same for 4 or 16 B
            constexpr int kRb = 2;      // pieces per thread whose remote loads are all in flight before the first sum
            for (int e0 = t; e0 < total; e0 += kFtEpiThreads * kRb) {
                float4 v[kRb][8];
                int idx[kRb];
#pragma unroll
                for (int e = 0; e < kRb; ++e) {
                    const int el = e0 + e * kFtEpiThreads;
                    const bool ok = el < total;
                    const int rr = r0 + (ok ? el / q4 : 0), cc = ok ? (el % q4) * 4 : 0;
                    idx[e] = ok ? (int)(rr * part_pitch + cc) : -1;
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        v[e][s] = (ok && s < p.splits) ? ld_dsmem_f32x4(smem_base + (uint32_t)idx[e] * 4u, (uint32_t)s)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int e = 0; e < kRb; ++e) {
                    if (idx[e] < 0) continue;
                    float4 acc = v[e][0];      // summed in rank order (x + 0 is exact)
#pragma unroll
                    for (int s = 1; s < 8; ++s) {
                        acc.x += v[e][s].x; acc.y += v[e][s].y; acc.z += v[e][s].z; acc.w += v[e][s].w;
                    }
                    *reinterpret_cast<float4*>(mine + idx[e]) = acc;
                }
            }
        }
        // (the epilogue warps' own writes become visible to each other at the named barrier below)
    }
    if (warp >= 2) {
        // ------------------------------------------------------------------ epilogue, part 2: coalesced write-out
        named_bar_sync(1, kFtEpiThreads);
        if (threadIdx.x == 64) FT_STAMP(6);                 // slice reduced
        const float* tile = reinterpret_cast<const float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
        const int t = (int)threadIdx.x - 64, ew = t >> 5;
        const int n0 = nt * p.bn;
        for (int o = 0; o < p.n_out; ++o) {
            // by VALUE: the map lives in the kernel-parameter constant bank, and a reference indexed by the run-time `o`
            // turns every field access of the inner loops into an indexed LDC (measured: ~270 cycles per element)
            const FtOut q = o == 0 ? p.out[0] : p.out[1];
            if (q.cs_lo == 1) {
                // columns are contiguous in this destination: a warp takes a row, its lanes run along the columns
                // (block_n <= 128: at most four columns per lane).  Two rows are in flight per warp, and the old values
                // of an accumulating destination are all requested before the first one is used, so the L2 round trip
                // is paid per pair of rows, not per element.
                constexpr int kCols = kFtMaxBn / 32, kRows = 2;
                for (int rb = r0 + ew * kRows; rb < r0 + nrows; rb += kFtEpiWarps * kRows) {
                    long long offs[kRows][kCols];
                    float val[kRows][kCols], old[kRows][kCols];
#pragma unroll
                    for (int i = 0; i < kRows; ++i) {
                        const int rr = rb + i, gm = mt * 128 + rr;
                        const bool rok = rr < r0 + nrows && gm < p.M;
                        const long long roff = rok ? (long long)(gm / q.rdiv) * q.rs_hi + (long long)(gm % q.rdiv) * q.rs_lo : 0;
                        FtCol w = ft_col_begin(q, roff, n0 + lane);
#pragma unroll
                        for (int k = 0; k < kCols; ++k) {
                            const int cc = lane + 32 * k;
                            const bool ok = rok && cc < p.bn && n0 + cc < p.N;
                            offs[i][k] = ok ? w.off : -1;
                            val[i][k] = ok ? tile[rr * part_pitch + cc] : 0.0f;
                            ft_col_step(q, w, 32);
                        }
                    }
                    if (q.accumulate && q.lo == nullptr) {
#pragma unroll
                        for (int i = 0; i < kRows; ++i)
#pragma unroll
                            for (int k = 0; k < kCols; ++k) old[i][k] = offs[i][k] >= 0 ? q.hi[offs[i][k]] : 0.0f;
#pragma unroll
                        for (int i = 0; i < kRows; ++i)
#pragma unroll
                            for (int k = 0; k < kCols; ++k)
                                if (offs[i][k] >= 0) q.hi[offs[i][k]] = old[i][k] + val[i][k];
                    } else {
#pragma unroll
                        for (int i = 0; i < kRows; ++i)
#pragma unroll
                            for (int k = 0; k < kCols; ++k)
                                if (offs[i][k] >= 0) ft_store(q, offs[i][k], val[i][k]);
                    }
                }
            } else if (q.rs_lo == 1 && q.rdiv <= 32 && q.cs_lo == q.rdiv && q.cdiv == kFtNoDiv) {
                // interleaved destination (gW2[o][m][hw]: row = (o, hw), column = m): for one o the (column, hw) pairs
                // of the tile form ONE contiguous run of memory, index v = column * rdiv + hw.  A warp takes an o, its
                // lanes run along v: full, coalesced sectors instead of 36-byte fragments per store.
                const int rd = q.rdiv;
                const int gm_lo = mt * 128 + r0, gm_hi = min(mt * 128 + r0 + nrows, p.M);     // rows [gm_lo, gm_hi)
                const int ncols = min(p.bn, p.N - n0);
                for (int g = gm_lo / rd + ew; g * rd < gm_hi; g += kFtEpiWarps) {
                    const long long base = (long long)g * q.rs_hi + (long long)n0 * rd;
                    for (int v = lane; v < ncols * rd; v += 32) {
                        const int c = v / rd, rl = v - c * rd;
                        const int gm = g * rd + rl;
                        if (gm >= gm_lo && gm < gm_hi) ft_store(q, base + v, tile[(gm - mt * 128) * part_pitch + c]);
                    }
                }
            } else {
                // rows are (piecewise) contiguous: consecutive threads take consecutive rows of one column; eight columns
                // per thread are read from shared memory before the first store
                const int rl = t % nrows, cgroups = kFtEpiThreads / nrows;
                const int gm = mt * 128 + r0 + rl;
                if (gm < p.M) {
                    const long long roff = (long long)(gm / q.rdiv) * q.rs_hi + (long long)(gm % q.rdiv) * q.rs_lo;
                    FtCol w = ft_col_begin(q, roff, n0 + t / nrows);
                    constexpr int kB = 8;
                    for (int cb = t / nrows; cb < p.bn; cb += kB * cgroups) {
                        long long offs[kB];
                        float val[kB];
#pragma unroll
                        for (int k = 0; k < kB; ++k) {
                            const int cc = cb + k * cgroups;
                            const bool ok = cc < p.bn && n0 + cc < p.N;
                            offs[k] = ok ? w.off : -1;
                            val[k] = ok ? tile[(r0 + rl) * part_pitch + cc] : 0.0f;
                            ft_col_step(q, w, cgroups);
                        }
#pragma unroll
                        for (int k = 0; k < kB; ++k)
                            if (offs[k] >= 0) ft_store(q, offs[k], val[k]);
                    }
                }
            }
        }
    }
    if (p.splits > 1) cluster_sync_relaxed();  // nobody leaves while a peer may still read its partial tile
    if (threadIdx.x == 64) FT_STAMP(7);

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------- operand preparation
struct FoldDims {
    int ci, co, c2, X, K9;          // Cin, Cout, 2 Cin, 9 Cin, 9 Cout
    int cip, cop, c2p, Xp, K9p;     // row pitches (floats), multiples of 32
};

struct PrepFwdArgs {
    const float *w3x3, *b3x3, *w1x3, *b1x3, *w3x1, *b3x1, *w1, *w2, *w3;
    float *W2p_h, *W2p_l, *W1t_h, *W1t_l, *W3_h, *W3_l, *Kt, *bias;
    float *W3t_h, *W3t_l, *W1_h, *W1_l, *W2c_h, *W2c_l;   // backward operands (NULL when not training)
    FoldDims d;
};

__device__ __forceinline__ void store_split(float* h, float* l, size_t off, float v) {
    float a, b;
    split_tf32(v, a, b);
    h[off] = a;
    l[off] = b;
}

__global__ void fold_prep_fwd_kernel(const PrepFwdArgs a) {
    const FoldDims d = a.d;
    // 32-bit index arithmetic throughout (the host checks every segment fits): divisions by run-time values are the
    // cost of these kernels
    const uint32_t n0 = (uint32_t)d.K9 * d.c2;          // W2p
    const uint32_t n1 = n0 + (uint32_t)d.ci * d.c2;     // W1t
    const uint32_t n2 = n1 + (uint32_t)d.co * d.co;     // W3
    const uint32_t n3 = n2 + (uint32_t)d.co * d.X;      // base -> Kt (+ bias)
    const bool train = a.W3t_h != nullptr;
    const uint32_t n4 = n3 + (train ? (uint32_t)d.co * d.co : 0u);      // W3t
    const uint32_t n5 = n4 + (train ? (uint32_t)d.c2 * d.ci : 0u);      // W1
    const uint32_t n6 = n5 + (train ? (uint32_t)d.c2 * d.K9 : 0u);      // W2c
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n6;
         idx += gridDim.x * blockDim.x) {
        if (idx < n0) {
            const int m = (int)(idx % d.c2), r = (int)(idx / d.c2);
            const int hw = r / d.co, o = r % d.co;
            store_split(a.W2p_h, a.W2p_l, (size_t)r * d.c2p + m, a.w2[((size_t)o * d.c2 + m) * 9 + hw]);
        } else if (idx < n1) {
            const uint32_t j = idx - n0;
            const int m = (int)(j % d.c2), i = (int)(j / d.c2);
            store_split(a.W1t_h, a.W1t_l, (size_t)i * d.c2p + m, a.w1[(size_t)m * d.ci + i]);
        } else if (idx < n2) {
            const uint32_t j = idx - n1;
            const int o = (int)(j % d.co), pp = (int)(j / d.co);
            store_split(a.W3_h, a.W3_l, (size_t)pp * d.cop + o, a.w3[j]);
        } else if (idx < n3) {
            const uint32_t j = idx - n2;
            const int x = (int)(j % d.X), pp = (int)(j / d.X);
            const int hw = x / d.ci, i = x % d.ci;
            const int h = hw / 3, w = hw % 3;
            const size_t oi = (size_t)pp * d.ci + i;
            const float v13 = (h == 1) ? a.w1x3[oi * 3 + w] : 0.0f;   // 1x3 fills the middle row    (model.py:495)
            const float v31 = (w == 1) ? a.w3x1[oi * 3 + h] : 0.0f;   // 3x1 fills the middle column (model.py:496)
            a.Kt[j] = a.w3x3[oi * 9 + hw] + (v13 + v31);
            if (j < (uint32_t)d.co) a.bias[j] = a.b3x3[j] + (a.b1x3[j] + a.b3x1[j]);
        } else if (idx < n4) {
            const uint32_t j = idx - n3;
            const int pp = (int)(j % d.co), o = (int)(j / d.co);
            store_split(a.W3t_h, a.W3t_l, (size_t)o * d.cop + pp, a.w3[(size_t)pp * d.co + o]);
        } else if (idx < n5) {
            const uint32_t j = idx - n4;
            const int i = (int)(j % d.ci), m = (int)(j / d.ci);
            store_split(a.W1_h, a.W1_l, (size_t)m * d.cip + i, a.w1[j]);
        } else {
            const uint32_t j = idx - n5;
            const int k = (int)(j % d.K9), m = (int)(j / d.K9);
            const int hw = k / d.co, o = k % d.co;
            store_split(a.W2c_h, a.W2c_l, (size_t)m * d.K9p + k, a.w2[((size_t)o * d.c2 + m) * 9 + hw]);
        }
    }
}

struct PrepBwdArgs {
    const float *dKt, *dbias;       // tap-major folded-kernel gradient [co][X], [co]
    float *dKq_h, *dKq_l, *dKtt_h, *dKtt_l;
    float *g3x3, *gb3x3, *g1x3, *gb1x3, *g3x1, *gb3x1;
    FoldDims d;
};

__global__ void fold_prep_bwd_kernel(const PrepBwdArgs a) {
    const FoldDims d = a.d;
    const uint32_t n0 = (uint32_t)d.co * d.X;     // dKq + direct gradients
    const uint32_t n1 = n0 + n0;                  // dKt transposed
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n1;
         idx += gridDim.x * blockDim.x) {
        if (idx < n0) {
            const int x = (int)(idx % d.X), pp = (int)(idx / d.X);
            const float g = a.dKt[idx];
            store_split(a.dKq_h, a.dKq_l, (size_t)pp * d.Xp + x, g);
            const int hw = x / d.ci, i = x % d.ci;
            const int h = hw / 3, w = hw % 3;
            const size_t oi = (size_t)pp * d.ci + i;
            a.g3x3[oi * 9 + hw] = g;
            if (h == 1) a.g1x3[oi * 3 + w] = g;
            if (w == 1) a.g3x1[oi * 3 + h] = g;
            if (idx < (uint32_t)d.co) {
                const float b = a.dbias[idx];
                a.gb3x3[idx] = b;
                a.gb1x3[idx] = b;
                a.gb3x1[idx] = b;
            }
        } else {
            const uint32_t j = idx - n0;
            const int pp = (int)(j % d.co), x = (int)(j / d.co);
            store_split(a.dKtt_h, a.dKtt_l, (size_t)x * d.cop + pp, a.dKt[(size_t)pp * d.X + x]);
        }
    }
}

// tap-major [Cout][9][Cin] <-> reference OIHW [Cout][Cin][3][3]
__global__ void tapmajor_permute_kernel(const float* __restrict__ src, float* __restrict__ dst, int ci, int co,
                                        int to_oihw) {
    const uint32_t total = (uint32_t)co * ci * 9;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += gridDim.x * blockDim.x) {
        const int hw = (int)(idx % 9);
        const int i = (int)((idx / 9) % ci);
        const uint32_t o = idx / (9u * ci);
        const uint32_t t = (o * 9 + hw) * ci + i;
        if (to_oihw) dst[idx] = src[t];
        else dst[t] = src[idx];
    }
}

// Tap-major fp32 kernel -> wf[9][Npad][Cpi], wd[9][Cpi_rows][Nk], bias_p[Npad]  (see pack_weights_kernel in fold.cu)
__global__ void pack_weights_t_kernel(const float* __restrict__ Kt, const float* __restrict__ bias, int Cin, int Cnew,
                                      int s, int Npad, int Cpi_rows, int Cpi, int Cpo, __nv_bfloat16* __restrict__ wf,
                                      __nv_bfloat16* __restrict__ wd, float* __restrict__ bias_p) {
    const int Nk = s * s * Cpo;
    const uint32_t X = 9u * Cin;
    const uint32_t nwf = 9u * Npad * Cpi;
    const uint32_t nwd = wd ? 9u * Cpi_rows * Nk : 0u;
    const uint32_t total = nwf + nwd + Npad;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += gridDim.x * blockDim.x) {
        if (idx < nwf) {
            const int ci = (int)(idx % Cpi);
            const int n = (int)((idx / Cpi) % Npad);
            const int tap = (int)(idx / ((uint32_t)Cpi * Npad));
            float v = 0.0f;
            if (n < Nk && ci < Cin) {
                const int ij = n / Cpo, c = n % Cpo;
                if (c < Cnew) v = Kt[(size_t)(c * s * s + ij) * X + (size_t)tap * Cin + ci];
            }
            wf[idx] = __float2bfloat16(v);
        } else if (idx < nwf + nwd) {
            const uint32_t j = idx - nwf;
            const int n = (int)(j % Nk);
            const int ci = (int)((j / Nk) % Cpi_rows);
            const int tap = (int)(j / ((uint32_t)Nk * Cpi_rows));
            float v = 0.0f;
            if (ci < Cin) {
                const int ij = n / Cpo, c = n % Cpo;
                if (c < Cnew) v = Kt[(size_t)(c * s * s + ij) * X + (size_t)tap * Cin + ci];
            }
            wd[j] = __float2bfloat16(v);
        } else {
            const int n = (int)(idx - nwf - nwd);
            float v = 0.0f;
            if (n < Nk) {
                const int ij = n / Cpo, c = n % Cpo;
                if (c < Cnew) v = bias[c * s * s + ij];
            }
            bias_p[n] = v;
        }
    }
}

// wgrad accumulators dKp[Nk][9][Cpi], dbias_p[Nk] (n' order) -> tap-major dKt[Cout][9][Cin], dbias[Cout]
__global__ void unpack_wgrad_t_kernel(const float* __restrict__ dKp, const float* __restrict__ dbias_p, int Cin,
                                      int Cnew, int s, int Cpi, int Cpo, float* __restrict__ dKt,
                                      float* __restrict__ dbias) {
    const int Cout = Cnew * s * s;
    const uint32_t total = (uint32_t)Cout * Cin * 9;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += gridDim.x * blockDim.x) {
        const int ci = (int)(idx % Cin);
        const int tap = (int)((idx / Cin) % 9);
        const int o = (int)(idx / (9u * Cin));
        const int c = o / (s * s), ij = o % (s * s);
        dKt[idx] = dKp[((size_t)(ij * Cpo + c) * 9 + tap) * Cpi + ci];
        if (idx < (uint32_t)Cout) {
            const int o2 = (int)idx;
            dbias[o2] = dbias_p[(o2 % (s * s)) * Cpo + o2 / (s * s)];
        }
    }
}

static inline int ft_grid(size_t total) {
    size_t g = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

struct FtGemm {
    CUtensorMap tmAh, tmAl, tmBh, tmBl;
    FtParams p;
    int grid;
    size_t smem;
};

}  // namespace onr

struct onr_fold_plan {
    onr::FoldDims d;
    int train;
    // workspace carve-up
    float *W2p_h, *W2p_l, *W1t_h, *W1t_l, *W3_h, *W3_l, *Tt_h, *Tt_l, *Tq_h, *Tq_l;
    float *W3t_h, *W3t_l, *W1_h, *W1_l, *W2c_h, *W2c_l, *dKq_h, *dKq_l, *dKtt_h, *dKtt_l, *dTa_h, *dTa_l, *dTb_h, *dTb_l;
    onr::FtGemm F1, F2, B1, B2, B3, B4;
};

namespace onr {

static size_t r256(size_t floats) { return (floats + 63) / 64 * 64; }

struct FoldLayout {
    size_t off[32];
    size_t total_floats;
};

static void choose_tiles(int M, int N, int K, int* bn, int* n_tiles, int* m_tiles, int* splits, int* k_stages) {
    const int nt = ceil_div(N, kFtMaxBn);
    int b = ceil_div(ceil_div(N, nt), 16) * 16;
    if (b < 16) b = 16;
    *bn = b;
    *n_tiles = nt;
    *m_tiles = ceil_div(M, 128);
    *k_stages = ceil_div(K, kFtKBox);
    const int tiles = nt * *m_tiles;
    // Split K only to bound the LENGTH of a CTA's K loop (~kFtSplitStages pipeline stages), never just to fill idle
    // SMs: these GEMMs run on side streams beside the persistent convolution kernels, every CTA of theirs takes a
    // whole SM away from those for its (latency-bound) lifetime, and the contention costs more than the fold gains
    // (ONR_FOLD_SPLIT_STAGES overrides; the splits of a tile form a thread-block cluster: a power of two <= 8).
    static const int split_stages = getenv("ONR_FOLD_SPLIT_STAGES") ? atoi(getenv("ONR_FOLD_SPLIT_STAGES")) : 16;
    int cap = num_sms() / tiles;
    const int want = ceil_div(*k_stages, split_stages > 0 ? split_stages : 16);
    if (cap > want) cap = want;
    if (cap > *k_stages / 2) cap = *k_stages / 2;
    int s = 1;
    while (s * 2 <= cap && s * 2 <= 8) s *= 2;
    *splits = s;
}

static FoldDims make_dims(int ci, int co) {
    FoldDims d;
    d.ci = ci; d.co = co; d.c2 = 2 * ci; d.X = 9 * ci; d.K9 = 9 * co;
    d.cip = pad32(d.ci); d.cop = pad32(d.co); d.c2p = pad32(d.c2); d.Xp = pad32(d.X); d.K9p = pad32(d.K9);
    return d;
}

// order of the arrays in the workspace (each hi/lo pair adjacent)
enum { A_W2p, A_W1t, A_W3, A_Tt, A_Tq, A_W3t, A_W1, A_W2c, A_dKq, A_dKtt, A_dTa, A_dTb, A_COUNT };

static size_t arr_floats(const FoldDims& d, int a) {
    switch (a) {
        case A_W2p: return (size_t)d.K9 * d.c2p;
        case A_W1t: return (size_t)d.ci * d.c2p;
        case A_W3: return (size_t)d.co * d.cop;
        case A_Tt: return (size_t)d.X * d.cop;
        case A_Tq: return (size_t)d.co * d.Xp;
        case A_W3t: return (size_t)d.co * d.cop;
        case A_W1: return (size_t)d.c2 * d.cip;
        case A_W2c: return (size_t)d.c2 * d.K9p;
        case A_dKq: return (size_t)d.co * d.Xp;
        case A_dKtt: return (size_t)d.X * d.cop;
        case A_dTa: return (size_t)d.K9 * d.cip;
        case A_dTb: return (size_t)d.ci * d.K9p;
    }
    return 0;
}


static void fold_layout(const FoldDims& d, int train, FoldLayout* L) {
    size_t off = 0;
    for (int a = 0; a < A_COUNT; ++a) {
        const bool used = train || a <= A_Tt;
        L->off[a] = off;
        if (used) off += 2 * r256(arr_floats(d, a));
    }
    L->total_floats = off;
}

static int make_gemm(FtGemm* g, const float* Ah, const float* Al, int M, int K, int lda, const float* Bh,
                     const float* Bl, int N, int ldb) {
    FtParams& p = g->p;
    p.M = M; p.N = N; p.K = K;
    choose_tiles(M, N, K, &p.bn, &p.n_tiles, &p.m_tiles, &p.splits, &p.k_stages);
    int cols = 32;
    while (cols < p.bn) cols *= 2;
    p.tmem_cols = 2 * cols;      // two accumulator buffers (chunked promotion)
    p.n_out = 0;
    p.prof = nullptr;
    g->grid = p.n_tiles * p.m_tiles * p.splits;
    g->smem = 1024 + (size_t)kFtStages * (2 * 128 * 128 + 2 * (size_t)p.bn * 128) + sizeof(FtBarriers);
    int rc = make_f32_2d_tmap(&g->tmAh, Ah, M, K, lda, 128);
    if (!rc) rc = make_f32_2d_tmap(&g->tmAl, Al, M, K, lda, 128);
    if (!rc) rc = make_f32_2d_tmap(&g->tmBh, Bh, N, K, ldb, p.bn);
    if (!rc) rc = make_f32_2d_tmap(&g->tmBl, Bl, N, K, ldb, p.bn);
    return rc;
}

static void add_out(FtGemm* g, float* hi, float* lo, int rdiv, long long rs_hi, long long rs_lo, int cdiv,
                    long long cs_hi, long long cs_lo, int accumulate) {
    FtOut& o = g->p.out[g->p.n_out++];
    o.hi = hi; o.lo = lo; o.rdiv = rdiv; o.cdiv = cdiv;
    o.rs_hi = rs_hi; o.rs_lo = rs_lo; o.cs_hi = cs_hi; o.cs_lo = cs_lo;
    o.accumulate = accumulate;
}

static int run_gemm_prof(const FtGemm& g0, cudaStream_t st);

static int run_gemm(const FtGemm& g, cudaStream_t st) {
    static const bool prof = getenv("ONR_FOLD_PROF") != nullptr;
    if (prof && g.p.prof == nullptr) return run_gemm_prof(g, st);
    if (g.p.splits > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(g.grid);
        cfg.blockDim = dim3(kFtThreads);
        cfg.dynamicSmemBytes = g.smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = g.p.splits;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ONR_CUDA(cudaLaunchKernelEx(&cfg, fold_tc_gemm_kernel, g.tmAh, g.tmAl, g.tmBh, g.tmBl, g.p));
    } else {
        fold_tc_gemm_kernel<<<g.grid, kFtThreads, g.smem, st>>>(g.tmAh, g.tmAl, g.tmBh, g.tmBl, g.p);
    }
    ONR_LAUNCH_CHECK();
    return 0;
}

// ONR_FOLD_PROF=1: run the GEMM with per-CTA cycle stamps and print their averages (debugging only; synchronises)
static int run_gemm_prof(const FtGemm& g0, cudaStream_t st) {
    FtGemm g = g0;
    long long* buf = nullptr;
    ONR_CUDA(cudaMalloc(&buf, (size_t)g.grid * 8 * sizeof(long long)));
    ONR_CUDA(cudaMemset(buf, 0, (size_t)g.grid * 8 * sizeof(long long)));
    g.p.prof = buf;
    int rc = run_gemm(g, st);
    if (rc) return rc;
    ONR_CUDA(cudaStreamSynchronize(st));
    long long* h = new long long[(size_t)g.grid * 8];
    ONR_CUDA(cudaMemcpy(h, buf, (size_t)g.grid * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg[8] = {0}, mx[8] = {0};
    for (int c = 0; c < g.grid; ++c)
        for (int i = 0; i < 8; ++i) {
            avg[i] += (double)h[c * 8 + i] / g.grid;
            if ((double)h[c * 8 + i] > mx[i]) mx[i] = (double)h[c * 8 + i];
        }
    fprintf(stderr, "[fold gemm M%d N%d K%d bn%d splits%d grid%d] avg cycles: setup %.0f tma-issued %.0f mma-issued %.0f "
            "acc %.0f epi %.0f cbar %.0f reduced %.0f end %.0f (max end %.0f)\n", g.p.M, g.p.N, g.p.K, g.p.bn, g.p.splits,
            g.grid, avg[0], avg[1], avg[2], avg[3], avg[4], avg[5], avg[6], avg[7], mx[7]);
    delete[] h;
    cudaFree(buf);
    return 0;
}

constexpr int kNoDiv = kFtNoDiv;

}  // namespace onr

extern "C" {

size_t onr_fold_workspace_bytes(int Cin, int Cout, int train) {
    using namespace onr;
    FoldLayout L;
    fold_layout(make_dims(Cin, Cout), train, &L);
    return L.total_floats * sizeof(float) + 1024;
}

int onr_fold_plan_create(onr_fold_plan** out, int Cin, int Cout, void* workspace, int train) {
    using namespace onr;
    ONR_REQUIRE(out && workspace && Cin > 0 && Cout > 0, "fold plan: bad arguments");
    ONR_REQUIRE(((uintptr_t)workspace & 1023) == 0, "fold plan: workspace must be 1024-byte aligned");
    {
        const long long c2 = 2ll * Cin, k9 = 9ll * Cout, x = 9ll * Cin;
        const long long most = k9 * c2 * 2 + (long long)Cout * Cout * 2 + (long long)Cout * x + (long long)Cin * c2 * 2;
        ONR_REQUIRE(most < (1ll << 31) && 9ll * pad32(Cout) * pad32(Cin) * 2 < (1ll << 31),
                    "fold plan: block too large for the 32-bit index arithmetic of the operand kernels");
    }
    onr_fold_plan* pl = new onr_fold_plan();
    pl->train = train;
    const FoldDims d = pl->d = make_dims(Cin, Cout);
    FoldLayout L;
    fold_layout(d, train, &L);
    float* base = reinterpret_cast<float*>(workspace);
    float** slots[A_COUNT][2] = {
        {&pl->W2p_h, &pl->W2p_l}, {&pl->W1t_h, &pl->W1t_l}, {&pl->W3_h, &pl->W3_l}, {&pl->Tt_h, &pl->Tt_l},
        {&pl->Tq_h, &pl->Tq_l}, {&pl->W3t_h, &pl->W3t_l}, {&pl->W1_h, &pl->W1_l}, {&pl->W2c_h, &pl->W2c_l},
        {&pl->dKq_h, &pl->dKq_l}, {&pl->dKtt_h, &pl->dKtt_l}, {&pl->dTa_h, &pl->dTa_l}, {&pl->dTb_h, &pl->dTb_l}};
    for (int a = 0; a < A_COUNT; ++a) {
        const bool used = train || a <= A_Tt;
        *slots[a][0] = used ? base + L.off[a] : nullptr;
        *slots[a][1] = used ? base + L.off[a] + r256(arr_floats(d, a)) : nullptr;
    }
    int rc = 0;
    // F1: Tt[(hw,i)][o] (and Tq[o][(hw,i)] when training) = W2p[(hw,o)][m] . W1t[i][m]
    rc = make_gemm(&pl->F1, pl->W2p_h, pl->W2p_l, d.K9, d.c2, d.c2p, pl->W1t_h, pl->W1t_l, d.ci, d.c2p);
    if (rc) { delete pl; return rc; }
    add_out(&pl->F1, pl->Tt_h, pl->Tt_l, d.co, (long long)d.ci * d.cop, 1, kNoDiv, 0, d.cop, 0);
    if (train) add_out(&pl->F1, pl->Tq_h, pl->Tq_l, d.co, d.ci, d.Xp, kNoDiv, 0, 1, 0);
    // F2: Kt[p][x] += W3[p][o] . Tt[x][o]      (Kt is bound per call)
    rc = make_gemm(&pl->F2, pl->W3_h, pl->W3_l, d.co, d.co, d.cop, pl->Tt_h, pl->Tt_l, d.X, d.cop);
    if (rc) { delete pl; return rc; }
    add_out(&pl->F2, nullptr, nullptr, kNoDiv, 0, d.X, kNoDiv, 0, 1, 1);
    if (train) {
        // B1: gW3[p][o] += dKq[p][x] . Tq[o][x]
        rc = make_gemm(&pl->B1, pl->dKq_h, pl->dKq_l, d.co, d.X, d.Xp, pl->Tq_h, pl->Tq_l, d.co, d.Xp);
        if (rc) { delete pl; return rc; }
        add_out(&pl->B1, nullptr, nullptr, kNoDiv, 0, d.co, kNoDiv, 0, 1, 0);
        // B2: dT[o][x] = W3t[o][p] . dKtt[x][p]  ->  dTa[(o,hw)][i] and dTb[i][(hw,o)], both split
        rc = make_gemm(&pl->B2, pl->W3t_h, pl->W3t_l, d.co, d.co, d.cop, pl->dKtt_h, pl->dKtt_l, d.X, d.cop);
        if (rc) { delete pl; return rc; }
        add_out(&pl->B2, pl->dTa_h, pl->dTa_l, kNoDiv, 0, 9ll * d.cip, d.ci, d.cip, 1, 0);
        add_out(&pl->B2, pl->dTb_h, pl->dTb_l, kNoDiv, 0, 1, d.ci, d.co, d.K9p, 0);
        // B3: gW2[o][m][hw] += dTa[(o,hw)][i] . W1[m][i]
        rc = make_gemm(&pl->B3, pl->dTa_h, pl->dTa_l, d.K9, d.ci, d.cip, pl->W1_h, pl->W1_l, d.c2, d.cip);
        if (rc) { delete pl; return rc; }
        add_out(&pl->B3, nullptr, nullptr, 9, 9ll * d.c2, 1, kNoDiv, 0, 9, 0);
        // B4: gW1[m][i] += W2c[m][(hw,o)] . dTb[i][(hw,o)]
        rc = make_gemm(&pl->B4, pl->W2c_h, pl->W2c_l, d.c2, d.K9, d.K9p, pl->dTb_h, pl->dTb_l, d.ci, d.K9p);
        if (rc) { delete pl; return rc; }
        add_out(&pl->B4, nullptr, nullptr, kNoDiv, 0, d.ci, kNoDiv, 0, 1, 0);
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fold_tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(fold gemm smem) failed: %s", cudaGetErrorString(e));
            delete pl;
            return (int)e;
        }
        attr_set = true;
    }
    *out = pl;
    return 0;
}

void onr_fold_plan_destroy(onr_fold_plan* pl) { delete pl; }

int onr_fold_plan_fwd(onr_fold_plan* pl, const float* w3x3, const float* b3x3, const float* w1x3, const float* b1x3,
                      const float* w3x1, const float* b3x1, const float* w1, const float* w2, const float* w3,
                      float* Kt, float* bias, void* stream) {
    using namespace onr;
    ONR_REQUIRE(pl && Kt && bias, "fold fwd: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const FoldDims& d = pl->d;
    PrepFwdArgs a{w3x3, b3x3, w1x3, b1x3, w3x1, b3x1, w1, w2, w3,
                  pl->W2p_h, pl->W2p_l, pl->W1t_h, pl->W1t_l, pl->W3_h, pl->W3_l, Kt, bias,
                  pl->W3t_h, pl->W3t_l, pl->W1_h, pl->W1_l, pl->W2c_h, pl->W2c_l, d};
    size_t total = (size_t)d.K9 * d.c2 + (size_t)d.ci * d.c2 + (size_t)d.co * d.co + (size_t)d.co * d.X;
    if (pl->train) total += (size_t)d.co * d.co + (size_t)d.c2 * d.ci + (size_t)d.c2 * d.K9;
    fold_prep_fwd_kernel<<<ft_grid(total), 256, 0, st>>>(a);
    ONR_LAUNCH_CHECK();
    int rc = run_gemm(pl->F1, st);
    if (rc) return rc;
    FtGemm f2 = pl->F2;
    f2.p.out[0].hi = Kt;
    return run_gemm(f2, st);
}

int onr_fold_plan_bwd(onr_fold_plan* pl, const float* dKt, const float* dbias, float* g3x3, float* gb3x3,
                      float* g1x3, float* gb1x3, float* g3x1, float* gb3x1, float* gw1, float* gw2, float* gw3,
                      void* stream) {
    using namespace onr;
    ONR_REQUIRE(pl && pl->train, "fold bwd: plan was not created for training");
    cudaStream_t st = (cudaStream_t)stream;
    const FoldDims& d = pl->d;
    PrepBwdArgs a{dKt, dbias, pl->dKq_h, pl->dKq_l, pl->dKtt_h, pl->dKtt_l, g3x3, gb3x3, g1x3, gb1x3, g3x1, gb3x1, d};
    fold_prep_bwd_kernel<<<ft_grid(2 * (size_t)d.co * d.X), 256, 0, st>>>(a);
    ONR_LAUNCH_CHECK();
    FtGemm b1 = pl->B1;
    b1.p.out[0].hi = gw3;
    int rc = run_gemm(b1, st);
    if (rc) return rc;
    rc = run_gemm(pl->B2, st);
    if (rc) return rc;
    FtGemm b3 = pl->B3;
    b3.p.out[0].hi = gw2;
    rc = run_gemm(b3, st);
    if (rc) return rc;
    FtGemm b4 = pl->B4;
    b4.p.out[0].hi = gw1;
    return run_gemm(b4, st);
}

int onr_tapmajor_permute(const float* src, float* dst, int Cin, int Cout, int to_oihw, void* stream) {
    using namespace onr;
    tapmajor_permute_kernel<<<ft_grid((size_t)Cout * Cin * 9), 256, 0, (cudaStream_t)stream>>>(src, dst, Cin, Cout,
                                                                                              to_oihw);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_pack_weights_t(const float* Kt, const float* bias, int Cin, int Cnew, int s, int Npad, int Cpi_rows,
                       void* wf, void* wd, float* bias_p, void* stream) {
    using namespace onr;
    const int Cpi = pad32(Cin), Cpo = pad32(Cnew);
    ONR_REQUIRE(Npad >= s * s * Cpo && Cpi_rows >= Cpi, "pack_weights_t: padded sizes too small");
    const size_t total = (size_t)9 * Npad * Cpi + (size_t)9 * Cpi_rows * s * s * Cpo + Npad;
    pack_weights_t_kernel<<<ft_grid(total), 256, 0, (cudaStream_t)stream>>>(
        Kt, bias, Cin, Cnew, s, Npad, Cpi_rows, Cpi, Cpo, reinterpret_cast<__nv_bfloat16*>(wf),
        reinterpret_cast<__nv_bfloat16*>(wd), bias_p);
    ONR_LAUNCH_CHECK();
    return 0;
}

int onr_unpack_wgrad_t(const float* dKp, const float* dbias_p, int Cin, int Cnew, int s, float* dKt, float* dbias,
                       void* stream) {
    using namespace onr;
    const int Cpi = pad32(Cin), Cpo = pad32(Cnew);
    unpack_wgrad_t_kernel<<<ft_grid((size_t)Cnew * s * s * Cin * 9), 256, 0, (cudaStream_t)stream>>>(
        dKp, dbias_p, Cin, Cnew, s, Cpi, Cpo, dKt, dbias);
    ONR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
