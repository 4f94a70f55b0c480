// selftest_kernels.h — TEST INFRASTRUCTURE, linked into the onr_selftest binary only (never into liborepnerv.so):
// plain SIMT versions of the three convolution passes for on-device cross-checks at sizes the CPU oracle cannot
// reach, and the tcgen05.mma issue-rate microbenchmark.
#pragma once
#include "../../include/orepnerv.h"

extern "C" {
int onr_simt_conv(const onr_conv_desc* desc, void* stream);
int onr_simt_wgrad(const onr_wgrad_desc* desc, void* stream);
/* cycles for iters*per_commit MMAs per CTA */
int onr_mma_bench(int N, int nacc, int per_commit, int iters, int depth, int layout, int a_stride,
                  int uniform, long long* out_dev, int grid, void* stream);
}
