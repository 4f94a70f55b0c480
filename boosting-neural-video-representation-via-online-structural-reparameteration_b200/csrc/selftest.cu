// selftest.cu — device-side cross-check of the tcgen05 kernels against the SIMT kernels on the same
// bf16 inputs.  Usage: onr_selftest <case> [reps].  One case per process so that a hang in one
// configuration cannot block the others (the runner wraps every call in `timeout`).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/orepnerv.h"
#include "selftest_kernels.h"

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)
#define OK(x)                                                           \
    do {                                                                \
        int r_ = (x);                                                   \
        if (r_ != 0) {                                                  \
            printf("onr error %d: %s (%s:%d)\n", r_, onr_last_error(), __FILE__, __LINE__); \
            exit(3);                                                    \
        }                                                               \
    } while (0)

static uint32_t rng_state = 12345u;
static float frand() {  // uniform in [-1, 1)
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xffff) / 32768.0f - 1.0f;
}
static __nv_bfloat16* dev_bf16_random(size_t n, float scale) {
    std::vector<__nv_bfloat16> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16(frand() * scale);
    __nv_bfloat16* d;
    CK(cudaMalloc(&d, n * 2));
    CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
    return d;
}
static float* dev_f32_random(size_t n, float scale) {
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = frand() * scale;
    float* d;
    CK(cudaMalloc(&d, n * 4));
    CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    return d;
}
static bool compare_bf16(const char* what, const __nv_bfloat16* a, const __nv_bfloat16* b, size_t n,
                         float atol, float rtol) {
    std::vector<__nv_bfloat16> ha(n), hb(n);
    CK(cudaMemcpy(ha.data(), a, n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), b, n * 2, cudaMemcpyDeviceToHost));
    double max_abs = 0, sum_sq = 0, ref_sq = 0;
    size_t bad = 0, first_bad = (size_t)-1;
    for (size_t i = 0; i < n; ++i) {
        const float x = __bfloat162float(ha[i]), y = __bfloat162float(hb[i]);
        const float diff = fabsf(x - y);
        if (diff > max_abs) max_abs = diff;
        sum_sq += (double)diff * diff;
        ref_sq += (double)y * y;
        if (!(diff <= atol + rtol * fabsf(y))) {
            if (first_bad == (size_t)-1) first_bad = i;
            ++bad;
        }
    }
    printf("  %-10s n=%zu max_abs=%.4g rel_l2=%.4g ref_rms=%.4g bad=%zu", what, n, max_abs,
           sqrt(sum_sq / (ref_sq + 1e-30)), sqrt(ref_sq / n), bad);
    if (bad) printf(" first_bad=%zu (got %g want %g)", first_bad, __bfloat162float(ha[first_bad]),
                    __bfloat162float(hb[first_bad]));
    printf("\n");
    return bad == 0;
}
static bool compare_f32(const char* what, const float* a, const float* b, size_t n, float atol, float rtol) {
    std::vector<float> ha(n), hb(n);
    CK(cudaMemcpy(ha.data(), a, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), b, n * 4, cudaMemcpyDeviceToHost));
    double max_abs = 0, sum_sq = 0, ref_sq = 0;
    size_t bad = 0, first_bad = (size_t)-1;
    for (size_t i = 0; i < n; ++i) {
        const float diff = fabsf(ha[i] - hb[i]);
        if (diff > max_abs) max_abs = diff;
        sum_sq += (double)diff * diff;
        ref_sq += (double)hb[i] * hb[i];
        if (!(diff <= atol + rtol * fabsf(hb[i]))) {
            if (first_bad == (size_t)-1) first_bad = i;
            ++bad;
        }
    }
    printf("  %-10s n=%zu max_abs=%.4g rel_l2=%.4g ref_rms=%.4g bad=%zu", what, n, max_abs,
           sqrt(sum_sq / (ref_sq + 1e-30)), sqrt(ref_sq / n), bad);
    if (bad) printf(" first_bad=%zu (got %g want %g)", first_bad, ha[first_bad], hb[first_bad]);
    printf("\n");
    return bad == 0;
}

struct Shape {
    const char* name;
    int B, H, W, cin_p, cout_p, s;
};
// cin_p: padded input channels; cout_p: padded channels after the shuffle; s: PixelShuffle factor.
static const Shape kShapes[] = {
    {"tiny", 1, 8, 16, 32, 32, 1},        // one tile, N = 32
    {"l0", 1, 9, 16, 32, 32, 5},          // N = 800 -> 3 n-tiles of 288
    {"l1", 1, 45, 80, 32, 96, 2},         // N = 384, K = 9*32
    {"l2s", 1, 24, 40, 96, 96, 2},        // N = 384, K = 9*96, ragged tiles
    {"b2", 2, 19, 35, 96, 96, 2},         // batch 2, odd sizes
    {"u3", 1, 16, 32, 96, 96, 3},         // stride 3 (UVG 1080p block 1): N = 864
    {"wide", 1, 9, 16, 128, 128, 5},      // NeRV-L width block 0: N = 3200
    {"xl", 1, 12, 20, 160, 96, 2},        // more than 128 input channels: wgrad channel chunks (128 + 32)
    {"l3", 1, 180, 320, 96, 96, 2},
    {"l4", 1, 360, 640, 96, 96, 2},       // Bunny 720p last block (timing)
};

static int run_conv(const Shape& sh, int kind, int reps, bool check) {
    const int n_pre = sh.s * sh.s * sh.cout_p;
    const bool dgrad = kind == ONR_CONV_DGRAD;
    const int n_total = dgrad ? sh.cin_p : n_pre;
    int block_n, n_tiles;
    OK(onr_conv_tile_n(n_total, &block_n, &n_tiles));
    const int n_rows = block_n * n_tiles;
    const int k_tap = dgrad ? n_pre : sh.cin_p;
    const size_t px = (size_t)sh.B * sh.H * sh.W;
    const size_t a_elems = dgrad ? px * n_pre : px * sh.cin_p;
    const size_t out_elems = dgrad ? px * sh.cin_p : px * n_pre;
    __nv_bfloat16* a = dev_bf16_random(a_elems, 1.0f);
    // weights [9][n_rows][k_tap], rows >= n_total zero
    std::vector<__nv_bfloat16> hw((size_t)9 * n_rows * k_tap);
    const float wscale = 1.0f / sqrtf(9.0f * k_tap) * 2.0f;
    for (int t = 0; t < 9; ++t)
        for (int n = 0; n < n_rows; ++n)
            for (int k = 0; k < k_tap; ++k)
                hw[((size_t)t * n_rows + n) * k_tap + k] = __float2bfloat16(n < n_total ? frand() * wscale : 0.0f);
    __nv_bfloat16* w;
    CK(cudaMalloc(&w, hw.size() * 2));
    CK(cudaMemcpy(w, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    float* bias = dev_f32_random(n_rows, 0.5f);
    __nv_bfloat16* dmul = dev_bf16_random(px * sh.cin_p, 1.0f);
    __nv_bfloat16 *o1, *o2, *d1, *d2;
    CK(cudaMalloc(&o1, out_elems * 2));
    CK(cudaMalloc(&o2, out_elems * 2));
    CK(cudaMalloc(&d1, out_elems * 2));
    CK(cudaMalloc(&d2, out_elems * 2));
    CK(cudaMemset(o1, 0x7f, out_elems * 2));
    CK(cudaMemset(d1, 0x7f, out_elems * 2));

    onr_conv_desc d = {};
    memset(&d, 0, sizeof(d));
    d.kind = kind;
    d.B = sh.B; d.H = sh.H; d.W = sh.W;
    d.a = a;
    d.a_cp = dgrad ? sh.cout_p : sh.cin_p;
    d.a_s = dgrad ? sh.s : 1;
    d.w = w;
    d.n_rows = n_rows;
    d.n_total = n_total;
    d.out = o1;
    d.out_cp = dgrad ? sh.cin_p : sh.cout_p;
    d.out_s = dgrad ? 1 : sh.s;
    d.out_d = d1;
    d.bias_p = bias;
    d.dmul = dmul;
    onr_conv_plan* plan = nullptr;
    OK(onr_conv_plan_create(&plan, &d));
    OK(onr_conv_plan_run(plan, 0));
    CK(cudaDeviceSynchronize());
    bool ok = true;
    if (check) {
        onr_conv_desc dr = d;
        dr.out = o2;
        dr.out_d = d2;
        OK(onr_simt_conv(&dr, 0));
        CK(cudaDeviceSynchronize());
        ok &= compare_bf16("out", o1, o2, out_elems, 2e-2f, 2e-2f);
        if (kind == ONR_CONV_FPROP_TRAIN) ok &= compare_bf16("dsilu", d1, d2, out_elems, 2e-2f, 2e-2f);
    }
    if (reps > 0) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) OK(onr_conv_plan_run(plan, 0));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) OK(onr_conv_plan_run(plan, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        const double flop = 2.0 * px * (double)n_total * 9.0 * k_tap;
        if (getenv("ONR_PROF")) {
            int grid = 0;
            OK(onr_conv_plan_set_prof(plan, nullptr, &grid));
            long long* pd;
            CK(cudaMalloc(&pd, (size_t)grid * 8 * sizeof(long long)));
            CK(cudaMemset(pd, 0, (size_t)grid * 8 * sizeof(long long)));
            OK(onr_conv_plan_set_prof(plan, pd, &grid));
            OK(onr_conv_plan_run(plan, 0));
            CK(cudaDeviceSynchronize());
            std::vector<long long> hp((size_t)grid * 8);
            CK(cudaMemcpy(hp.data(), pd, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            double acc[8] = {0};
            for (int c = 0; c < grid; ++c)
                for (int j = 0; j < 8; ++j) acc[j] += (double)hp[(size_t)c * 8 + j] / grid;
            printf("  prof (avg cycles per CTA over %d CTAs, %.1f tiles each): total %.0f | MMA waits: A %.0f  W %.0f  acc %.0f"
                   " | epilogue: wait acc %.0f  wait store %.0f  busy %.0f\n", grid, acc[7], acc[0], acc[1], acc[2], acc[3],
                   acc[4], acc[5], acc[6]);
            OK(onr_conv_plan_set_prof(plan, nullptr, &grid));
        }
        int pbn, pnt, pms, pna, pnb;
        OK(onr_conv_plan_info(plan, &pbn, &pnt, &pms, &pna, &pnb));
        printf("  time %.3f ms  %.1f TFLOP/s (block_n %d x %d n-tiles, %d sub-tiles, rings A%d B%d)\n", ms,
               flop / ms * 1e-9, pbn, pnt, pms, pna, pnb);
    }
    onr_conv_plan_destroy(plan);
    return ok ? 0 : 1;
}

static int run_wgrad(const Shape& sh, int reps, bool check) {
    const int n_pre = sh.s * sh.s * sh.cout_p;
    const size_t px = (size_t)sh.B * sh.H * sh.W;
    __nv_bfloat16* x = dev_bf16_random(px * sh.cin_p, 1.0f);
    __nv_bfloat16* dz = dev_bf16_random(px * n_pre, 1.0f);
    const size_t kn = (size_t)n_pre * 9 * sh.cin_p;
    float *k1, *k2, *b1, *b2;
    CK(cudaMalloc(&k1, kn * 4));
    CK(cudaMalloc(&k2, kn * 4));
    CK(cudaMalloc(&b1, n_pre * 4));
    CK(cudaMalloc(&b2, n_pre * 4));
    CK(cudaMemset(k1, 0, kn * 4));
    CK(cudaMemset(k2, 0, kn * 4));
    CK(cudaMemset(b1, 0, n_pre * 4));
    CK(cudaMemset(b2, 0, n_pre * 4));
    onr_wgrad_desc d;
    memset(&d, 0, sizeof(d));
    d.B = sh.B; d.H = sh.H; d.W = sh.W;
    d.x = x; d.x_cp = sh.cin_p;
    d.dz = dz; d.dz_cp = sh.cout_p; d.s = sh.s;
    d.dKp = k1; d.dbias_p = b1;
    onr_wgrad_plan* plan = nullptr;
    OK(onr_wgrad_plan_create(&plan, &d));
    OK(onr_wgrad_plan_run(plan, 0));
    CK(cudaDeviceSynchronize());
    bool ok = true;
    if (check) {
        onr_wgrad_desc dr = d;
        dr.dKp = k2; dr.dbias_p = b2;
        OK(onr_simt_wgrad(&dr, 0));
        CK(cudaDeviceSynchronize());
        const float scale = sqrtf((float)px) * 0.33f;
        ok &= compare_f32("dKp", k1, k2, kn, 2e-3f * scale, 2e-3f);
        ok &= compare_f32("dbias", b1, b2, n_pre, 2e-3f * scale, 2e-3f);
    }
    if (reps > 0) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) OK(onr_wgrad_plan_run(plan, 0));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) OK(onr_wgrad_plan_run(plan, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        const double flop = 2.0 * px * (double)n_pre * 9.0 * sh.cin_p;
        printf("  time %.3f ms  %.1f TFLOP/s (incl. bias column sums)\n", ms, flop / ms * 1e-9);
    }
    onr_wgrad_plan_destroy(plan);
    return ok ? 0 : 1;
}

static int run_mma_bench() {
    const int grid = 148;
    long long* out;
    CK(cudaMalloc(&out, grid * sizeof(long long)));
    const char* lname[4] = {"K-major SW128", "K-major SW64", "MN-major SW64", "MN-major SW128"};
    printf("%-15s %4s %4s %6s %5s %7s | cycles/MMA  ideal  MAC/clk/SM\n", "layout", "N", "nacc", "percmt", "depth", "uniform");
    const int Ns[] = {96, 192, 256};
    const int a_stride = 16384;
    for (int layout = 0; layout < 1; ++layout)
        for (int n : Ns)
            for (int nacc : {1})
                for (int per_commit : {2, 4, 8, 16, 64})
                    for (int depth : {1, 6})
                        for (int uniform : {1, 2}) {
                            if (uniform == 2 && n > 256) continue;
                            const int iters = 2048 / per_commit;
                            OK(onr_mma_bench(n, nacc, per_commit, iters, depth, layout, a_stride, uniform, out, grid, 0));
                            CK(cudaDeviceSynchronize());
                            std::vector<long long> h(grid);
                            CK(cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                            double avg = 0;
                            for (long long v : h) avg += (double)v / grid;
                            const double per = avg / 2048.0;
                            printf("%-15s %4d %4d %6d %5d %7d | %9.1f  %5.0f  %8.0f\n", lname[layout], n, nacc, per_commit,
                                   depth, uniform, per, 128.0 * n / 256.0, 128.0 * n * 16.0 / per * (uniform == 2 ? 2 : 1));
                        }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "mmabench")) {
        OK(onr_check_device());
        return run_mma_bench();
    }
    if (argc < 3) {
        printf("usage: %s <fprop|infer|dgrad|wgrad> <shape> [reps] [nocheck]\nshapes:", argv[0]);
        for (const Shape& s : kShapes) printf(" %s", s.name);
        printf("\n");
        return 64;
    }
    const int reps = argc > 3 ? atoi(argv[3]) : 0;
    const bool check = !(argc > 4 && !strcmp(argv[4], "nocheck"));
    OK(onr_check_device());
    const Shape* sh = nullptr;
    for (const Shape& s : kShapes)
        if (!strcmp(s.name, argv[2])) sh = &s;
    if (!sh) { printf("unknown shape %s\n", argv[2]); return 64; }
    printf("[%s %s] B%d H%d W%d cin_p%d cout_p%d s%d\n", argv[1], sh->name, sh->B, sh->H, sh->W, sh->cin_p,
           sh->cout_p, sh->s);
    int rc;
    if (!strcmp(argv[1], "fprop")) rc = run_conv(*sh, ONR_CONV_FPROP_TRAIN, reps, check);
    else if (!strcmp(argv[1], "infer")) rc = run_conv(*sh, ONR_CONV_FPROP_INFER, reps, check);
    else if (!strcmp(argv[1], "dgrad")) rc = run_conv(*sh, ONR_CONV_DGRAD, reps, check);
    else if (!strcmp(argv[1], "wgrad")) rc = run_wgrad(*sh, reps, check);
    else { printf("unknown op %s\n", argv[1]); return 64; }
    printf("[%s %s] %s\n", argv[1], sh->name, rc == 0 ? "PASS" : "FAIL");
    return rc;
}
