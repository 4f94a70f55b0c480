// mma_bench.cu — issue-rate microbenchmark for tcgen05.mma (test infrastructure, selftest only).
// One CTA per SM; operands are whatever happens to be in shared memory (values are irrelevant); one thread
// issues `iters` groups of `per_commit` MMAs (M=128, N, K=16, bf16) round-robin over `nacc` accumulators and
// commits after each group, waiting for the commit every `depth` groups (so up to `depth` groups are in flight).
// Reports cycles per MMA.  Answers: what does the single-CTA tensor pipe sustain, and what does it depend on?
#include "onr_common.cuh"
#include "selftest_kernels.h"
#include "onr_ptx.cuh"

namespace onr {

struct MmaBenchParams {
    int N, nacc, per_commit, iters, depth;
    int layout;      // 0: K-major SW128, 1: K-major SW64, 2: MN-major SW64, 3: MN-major SW128
    int uniform;     // 1: whole warp runs the loop, issue predicated by elect_one; 0: everything under lane==0
    int a_stride;    // bytes between consecutive A tiles used by successive MMAs (0 = same tile)
    long long* out;  // [grid] cycles
};

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(const MmaBenchParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[16];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if ((warp == 3 || (warp == 2 && p.uniform == 2)) && p.uniform) {
        const int who = warp == 3 ? 0 : 1;
        const bool mn = p.layout >= 2;
        const uint32_t lay = (p.layout == 0 || p.layout == 3) ? SWZ_128B : SWZ_64B;
        const uint32_t sbo = (p.layout == 0 || p.layout == 3) ? 1024u : 512u;
        const uint32_t lbo = mn ? 8192u : 16u;
        const uint32_t idesc = make_idesc_bf16(128, p.N, mn ? 1 : 0, mn ? 1 : 0);
        const uint32_t a0 = base + who * 32768, b0 = base + 96 * 1024 + who * 32768;
        const uint32_t tm = tmem + who * 256;
        uint64_t* mybars = bars + who * 8;
        const long long t0 = clock64();
        int grp = 0;
        for (int it = 0; it < p.iters; ++it) {
            if (elect_one()) {
                for (int j = 0; j < p.per_commit; ++j) {
                    const int m = it * p.per_commit + j;
                    const uint32_t koff = mn ? (uint32_t)(m & 3) * 1024u : (uint32_t)(m & 3) * 32u;
                    const uint64_t ad = make_smem_desc(a0 + (m % 4) * p.a_stride + koff, lbo, sbo, lay);
                    const uint64_t bd = make_smem_desc(b0 + koff, lbo, sbo, lay);
                    umma_bf16(tm + (m % p.nacc) * p.N, ad, bd, idesc, 1u);
                }
                umma_commit(smem_u32(&mybars[grp % 8]));
            }
            __syncwarp();
            ++grp;
            if (grp >= p.depth) {
                const int w = grp - p.depth;
                mbar_wait(smem_u32(&mybars[w % 8]), (uint32_t)(w / 8) & 1u);
            }
        }
        for (int w = grp - p.depth + 1; w < grp; ++w)
            if (w >= 0) mbar_wait(smem_u32(&mybars[w % 8]), (uint32_t)(w / 8) & 1u);
        if (lane == 0 && who == 0) p.out[blockIdx.x] = clock64() - t0;
    } else if (warp == 3 && lane == 0) {
        const bool mn = p.layout >= 2;
        const uint32_t lay = (p.layout == 0 || p.layout == 3) ? SWZ_128B : SWZ_64B;
        const uint32_t sbo = (p.layout == 0 || p.layout == 3) ? 1024u : 512u;
        const uint32_t lbo = mn ? 8192u : 16u;
        const uint32_t idesc = make_idesc_bf16(128, p.N, mn ? 1 : 0, mn ? 1 : 0);
        const uint32_t a0 = base, b0 = base + 96 * 1024;
        const long long t0 = clock64();
        int grp = 0;
        for (int it = 0; it < p.iters; ++it) {
            for (int j = 0; j < p.per_commit; ++j) {
                const int m = it * p.per_commit + j;
                const uint32_t koff = mn ? (uint32_t)(m & 3) * 1024u : (uint32_t)(m & 3) * 32u;
                const uint64_t ad = make_smem_desc(a0 + (m % 4) * p.a_stride + koff, lbo, sbo, lay);
                const uint64_t bd = make_smem_desc(b0 + koff, lbo, sbo, lay);
                umma_bf16(tmem + (m % p.nacc) * p.N, ad, bd, idesc, 1u);
            }
            umma_commit(smem_u32(&bars[grp % 16]));
            ++grp;
            if (grp >= p.depth) {   // wait for the group issued `depth` groups ago
                const int w = grp - p.depth;
                mbar_wait(smem_u32(&bars[w % 16]), (uint32_t)(w / 16) & 1u);
            }
        }
        for (int w = grp - p.depth + 1; w < grp; ++w)
            if (w >= 0) mbar_wait(smem_u32(&bars[w % 16]), (uint32_t)(w / 16) & 1u);
        p.out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace onr

extern "C" int onr_mma_bench(int N, int nacc, int per_commit, int iters, int depth, int layout, int a_stride,
                             int uniform, long long* out_dev, int grid, void* stream) {
    using namespace onr;
    ONR_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && nacc >= 1 && nacc * N <= 512 && depth >= 1 && depth <= 7,
                "mma_bench: bad parameters");
    static bool attr = false;
    if (!attr) {
        ONR_CUDA(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    MmaBenchParams p{N, nacc, per_commit, iters, depth, layout, uniform, a_stride, out_dev};
    mma_bench_kernel<<<grid, 128, 200 * 1024, (cudaStream_t)stream>>>(p);
    ONR_LAUNCH_CHECK();
    return 0;
}
