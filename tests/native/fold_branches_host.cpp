// TEST INFRASTRUCTURE: compiles the element arithmetic of csrc/fold_branches.cuh — the very functions the CUDA kernels
// of fold_branches.cu call per thread — with g++ and drives them with plain loops, so that tests/ can check the
// index math of the branch-set fold and of its backward against the oracle on the CPU (tests/test_branch_fold_host.py).
// Nothing in the product links or loads this file.
#include "../../boosting-neural-video-representation-via-online-structural-reparameteration_b200/csrc/fold_branches.cuh"

extern "C" {

void host_branch_fold_fwd(const onr_branch_set* s, float* K, float* bias) {
    for (int o = 0; o < s->cout; ++o) {
        for (int i = 0; i < s->cin; ++i)
            for (int t = 0; t < 9; ++t) K[((size_t)o * s->cin + i) * 9 + t] = onr::branch_k_elem(*s, o, i, t);
        bias[o] = onr::branch_b_elem(*s, o);
    }
}

void host_branch_fold_bwd(const onr_branch_set* s, const float* dK, const float* db, const onr_branch_set* g) {
    for (int o = 0; o < s->cout; ++o) {
        for (int i = 0; i < s->cin; ++i) onr::branch_bwd_oi(*s, *g, dK, o, i);
        float dot[3] = {0.0f, 0.0f, 0.0f};
        for (int e = 0; e < 3; ++e)
            if (s->edge_k0[e])
                for (int l = 0; l < 32; ++l) dot[e] += onr::branch_bwd_scale_partial(*s, dK, e, o, l, 32);
        onr::branch_bwd_o(*s, *g, db, o, dot);
    }
    if (s->seq_w1) {
        const int cm = 2 * s->cin;
        if (g->seq_w2)
            for (int o = 0; o < s->cout; ++o)
                for (int m = 0; m < cm; ++m)
                    for (int t = 0; t < 9; ++t)
                        g->seq_w2[((size_t)o * cm + m) * 9 + t] = onr::branch_bwd_w2_elem(*s, dK, o, m, t);
        if (g->seq_w1)
            for (int m = 0; m < cm; ++m)
                for (int i = 0; i < s->cin; ++i) {
                    // the kernel gives lane l the partial (start l, stride 32) and adds the 32 partials
                    float acc = 0.0f;
                    for (int l = 0; l < 32; ++l) acc += onr::branch_bwd_w1_partial(*s, dK, m, i, l, 32);
                    g->seq_w1[(size_t)m * s->cin + i] = acc;
                }
    }
}

}  // extern "C"
