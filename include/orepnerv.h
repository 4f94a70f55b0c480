/* orepnerv.h — C ABI of liborepnerv.so: the B200 (sm_100a) kernels behind the Online-RepNeRV
 * frame-fitting hot path.
 *
 * The reference (maoqingyu1996/Boosting-Neural-Video-Representation-via-Online-Structural-
 * Reparameteration) has no FFI: its seam is the Python API in model.py / utils.py.  Every entry
 * point below names the reference call site (file:line under /root/reference) whose arithmetic
 * it replaces.  The Python host layer (package dir, `model.py` / `utils.py`) binds these with
 * ctypes; see INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *  - every function is stream-ordered and asynchronous on `stream` (a cudaStream_t passed as void*),
 *    never allocates or frees device memory, never synchronises, and is CUDA-graph capturable;
 *  - return value: 0 = OK, <0 = bad argument / unsupported shape / wrong architecture,
 *    >0 = cudaError_t;  onr_last_error() returns a thread-local message;
 *  - activations between kernels are NHWC bf16 with the channel count padded to a multiple of 32
 *    ("Cp"); parameters, gradients, optimizer state, images and losses are fp32 in the reference's
 *    own layouts (OIHW / NCHW), so checkpoints stay byte-compatible.
 */
#ifndef OREPNERV_H
#define OREPNERV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONR_ABI_VERSION 1

/* ------------------------------------------------------------------ library / errors */
int onr_abi_version(void);
const char* onr_last_error(void);
/* 0 when the current device is compute capability 10.x (B200); <0 otherwise.  The library has no
 * other code path: callers must fail loudly when this fails. */
int onr_check_device(void);
/* Number of kernels this library has launched in the calling process so far (bench.py's gpu_launches). */
unsigned long long onr_launch_count(void);

/* ------------------------------------------------------------------ A1+A2: positional encoding + stem
 * utils.py:121-129 (PositionalEncoding.forward), model.py:174-188, 612-613 (stem MLP + view).
 * t_norm[B] (fp32 frame index i/N).  Produces embed[B,2L] (fp32, sin/cos interleaved, rounding order
 * (t*lbase^i)*pi; the host passes freqs[i] = fp32(lbase**i) evaluated in double like the reference),
 * h1[B,hid] = SiLU(W1 e + b1), and the stem output written as NHWC bf16 x0[B,fh,fw,Cp]
 * (value of reference out[b, c*fh*fw + h*fw + w]), plus dstem[B,fh,fw,Cp] = SiLU'(pre-activation).
 * t_norm may be NULL: then embed[B,2L] is an INPUT (the reference's Generator.forward(embed) entry). */
int onr_pe_stem_fwd(const float* t_norm, int B, const float* freqs /* [levels] = fp32(lbase**i) */, int levels,
                    const float* W1, const float* b1, int hid,
                    const float* W2, const float* b2, int fc_dim, int fh, int fw, int Cp,
                    float* embed, float* pre1, float* h1,
                    void* x0_bf16, void* dstem_bf16, void* stream);
/* The same with the stem's activation given as a code (see onr_act_map): model.py:184-188 builds the MLP with the
 * --act of the run.  onr_pe_stem_fwd == act 0 (swish). */
int onr_pe_stem_fwd_act(const float* t_norm, int B, const float* freqs, int levels,
                        const float* W1, const float* b1, int hid,
                        const float* W2, const float* b2, int fc_dim, int fh, int fw, int Cp,
                        float* embed, float* pre1, float* h1,
                        void* x0_bf16, void* dstem_bf16, int act, void* stream);
/* utils.py:121-129 alone: embed[B,2L] from t_norm[B] (what PositionalEncoding.forward returns). */
int onr_pos_encoding(const float* t_norm, int B, const float* freqs, int levels, float* embed, void* stream);
/* Backward of the stem (autograd of model.py:612 through main_train.py:249).
 * g0[B,fh,fw,Cp] bf16 is dL/d(pre-activation of the 2nd Linear) (SiLU' already applied by the
 * block-0 dgrad epilogue).  Accumulates (+=) into gW1,gb1,gW2,gb2. */
int onr_stem_bwd(const void* g0_bf16, int B, const float* embed, int emb_len,
                 const float* pre1, const float* h1, int hid,
                 const float* W2, int fc_dim, int fh, int fw, int Cp,
                 float* gW1, float* gb1, float* gW2, float* gb2,
                 float* scratch_dh1 /* [B,hid] */, void* stream);
int onr_stem_bwd_act(const void* g0_bf16, int B, const float* embed, int emb_len,
                     const float* pre1, const float* h1, int hid,
                     const float* W2, int fc_dim, int fh, int fw, int Cp,
                     float* gW1, float* gb1, float* gW2, float* gb2,
                     float* scratch_dh1 /* [B,hid] */, int act, void* stream);

/* Data-parallel form of the stem backward (batch 1 per rank): the stem gradients of one frame are rank-1, so ranks
 * exchange the FACTORS (onr_stem_factor_floats() fp32 per rank: g | h1 | dpre1 | embed) with an all-gather instead of
 * all-reducing the 7.7 .. 33 MB matrices, and each rank forms the sums over the K gathered slots in rank order
 * (bit-identical on every rank).  onr_stem_grads_from_factors OVERWRITES gW1, gb1, gW2, gb2 with the SUM over ranks. */
size_t onr_stem_factor_floats(int fc_dim, int fh, int fw, int hid, int emb_len);
int onr_stem_bwd_factors(const void* g0_bf16, const float* embed, int emb_len, const float* pre1, const float* h1,
                         int hid, const float* W2, int fc_dim, int fh, int fw, int Cp,
                         float* slot /* this rank's slot */, float* scratch_dh1 /* [hid] */, void* stream);
int onr_stem_bwd_factors_act(const void* g0_bf16, const float* embed, int emb_len, const float* pre1, const float* h1,
                             int hid, const float* W2, int fc_dim, int fh, int fw, int Cp,
                             float* slot, float* scratch_dh1, int act, void* stream);
int onr_stem_grads_from_factors(const float* slots /* [K][onr_stem_factor_floats()] */, int K, int emb_len, int hid,
                                int fc_dim, int fh, int fw, float* gW1, float* gb1, float* gW2, float* gb2,
                                void* stream);

/* ------------------------------------------------------------------ A3: ERB online fold
 * model.py:450-516 (get_equivalent_kernel_bias, _fuse_1x3_3x1_branch, _fuse_1x1_3x3_1x1_branch).
 * Inputs are the nine branch tensors in the reference's OIHW layouts.  Outputs:
 *   K[Cout,Cin,3,3], bias[Cout]  fp32 (reference layout; what switch_to_deploy stores),
 *   T[Cout,Cin,3,3] fp32 = W2 (x) W1 contraction, kept for the backward,
 * `Cout` = Cnew*s*s in the reference's PixelShuffle order (c*s*s + i*s + j). */
int onr_erb_fold_fwd(const float* w3x3, const float* b3x3, const float* w1x3, const float* b1x3,
                     const float* w3x1, const float* b3x1, const float* w1 /*[2Cin,Cin]*/,
                     const float* w2 /*[Cout,2Cin,3,3]*/, const float* w3 /*[Cout,Cout]*/,
                     int Cin, int Cout, float* K, float* bias, float* T, void* stream);
/* Backward of the fold: scatters dK[Cout,Cin,3,3], dbias[Cout] to the nine branch gradients (+=).
 * dT is scratch [Cout,Cin,3,3]. */
int onr_erb_fold_bwd(const float* dK, const float* dbias, const float* w1, const float* w2,
                     const float* w3, const float* T, int Cin, int Cout,
                     float* g3x3, float* gb3x3, float* g1x3, float* gb1x3, float* g3x1, float* gb3x1,
                     float* gw1, float* gw2, float* gw3, float* dT, void* stream);

/* Tensor-core fold (csrc/fold_tc.cu): the same arithmetic as onr_erb_fold_fwd / _bwd with the six contractions as
 * tcgen05 GEMMs in 3xTF32 split precision (fp32-accurate, bit-reproducible).  The folded kernel and its gradient
 * travel "tap-major": Kt[Cout][9][Cin] (= K[o][i][kh][kw] with the tap index ahead of the input channel), which is
 * the operand order of the convolution kernels.  A plan owns the TMA descriptors of one (Cin, Cout) block and a
 * caller-provided workspace of onr_fold_workspace_bytes() bytes (1024-byte aligned); train = 0 plans only fold. */
typedef struct onr_fold_plan onr_fold_plan;
size_t onr_fold_workspace_bytes(int Cin, int Cout, int train);
int onr_fold_plan_create(onr_fold_plan** out, int Cin, int Cout, void* workspace, int train);
void onr_fold_plan_destroy(onr_fold_plan* plan);
int onr_fold_plan_fwd(onr_fold_plan* plan, const float* w3x3, const float* b3x3, const float* w1x3,
                      const float* b1x3, const float* w3x1, const float* b3x1, const float* w1, const float* w2,
                      const float* w3, float* Kt /*[Cout][9][Cin]*/, float* bias /*[Cout]*/, void* stream);
/* dKt[Cout][9][Cin], dbias[Cout] -> the nine branch gradients (reference OIHW layouts), OVERWRITTEN: the fold is the
 * only consumer of the branch tensors, so nothing else contributes to these gradients within a backward pass. */
int onr_fold_plan_bwd(onr_fold_plan* plan, const float* dKt, const float* dbias,
                      float* g3x3, float* gb3x3, float* g1x3, float* gb1x3, float* g3x1, float* gb3x1,
                      float* gw1, float* gw2, float* gw3, void* stream);
/* ------------------------------------------------------------------ the other linear branch sets
 * model.py:345-393 (ACB, RepVGG, DBB, ECB; SeqConv3x3 :191-300).  The reference runs them as an explicit multi-branch
 * forward (:541-565) and has no fold for them; every branch is linear, so their sum is ONE 3x3 convolution.  The fold
 * below builds that kernel on the device every step (fp32), the block then runs the same tcgen05 convolution as ERB,
 * and the backward scatters dK / dbias to the branch parameters.  Unused branches are NULL.  All tensors in the
 * reference's layouts: w3x3[Cout,Cin,3,3], w1x3[Cout,Cin,1,3], w3x1[Cout,Cin,3,1], w1x1[Cout,Cin,1,1] (+ biases [Cout]);
 * seq_w1[2Cin,Cin,1,1] -> seq_w2[Cout,2Cin,3,3] (the 1x1 -> 3x3 branch, no biases); avg_w[Cout,Cin,1,1] (1x1 ->
 * AvgPool2d(3,1,1)); edge_*[e] = SeqConv3x3 e (sobel-x, sobel-y, laplacian): k0[Cout,Cin,1,1], b0[Cout],
 * scale[Cout,1,1,1], bias[Cout], mask[Cout,1,3,3] (mask is a constant: it never receives a gradient). */
typedef struct onr_branch_set {
    int cin, cout;
    float *w3x3, *b3x3, *w1x3, *b1x3, *w3x1, *b3x1, *w1x1, *b1x1;
    float *seq_w1, *seq_w2, *avg_w;
    float *edge_k0[3], *edge_b0[3], *edge_scale[3], *edge_bias[3], *edge_mask[3];
} onr_branch_set;
/* K[Cout,Cin,3,3], bias[Cout] (OVERWRITTEN) = the single-convolution equivalent of the branch set. */
int onr_branch_fold_fwd(const onr_branch_set* w, float* K, float* bias, void* stream);
/* g holds the gradient pointers in the same slots (NULL: not wanted); every present gradient is OVERWRITTEN. */
int onr_branch_fold_bwd(const onr_branch_set* w, const float* dK, const float* dbias, const onr_branch_set* g,
                        void* stream);

/* Kt[Cout][9][Cin] <-> K[Cout][Cin][3][3] (to_oihw != 0: tap-major -> reference layout). */
int onr_tapmajor_permute(const float* src, float* dst, int Cin, int Cout, int to_oihw, void* stream);
/* onr_pack_weights / onr_unpack_wgrad for tap-major kernels. */
int onr_pack_weights_t(const float* Kt, const float* bias, int Cin, int Cnew, int s, int Npad, int Cpi_rows,
                       void* wf_bf16, void* wd_bf16, float* bias_p, void* stream);
int onr_unpack_wgrad_t(const float* dKp, const float* dbias_p, int Cin, int Cnew, int s,
                       float* dKt, float* dbias, void* stream);

/* Pack a folded (or vanilla / deploy) OIHW fp32 kernel into the two bf16 implicit-GEMM operand
 * layouts the convolution kernels consume.  n' = (i*s+j)*Cpo + c  for reference channel
 * o = c*s*s + i*s + j (nn.PixelShuffle order, model.py:310), Cpo = pad32(Cnew), Cpi = pad32(Cin):
 *   wf[9][Npad][Cpi]     (fprop:  B operand, K = input channel)        Npad = n_tiles*block_n
 *   wd[9][Cpi_pad][Nk]   (dgrad:  B operand, K = n', taps flipped)      Nk = s*s*Cpo
 *   bias_p[Npad]         fp32 bias in n' order (0 in padding)                                   */
int onr_pack_weights(const float* K, const float* bias, int Cin, int Cnew, int s,
                     int Npad, int Cpi_rows,
                     void* wf_bf16, void* wd_bf16, float* bias_p, void* stream);
/* Inverse for gradients: dKp[Nk][9][Cpi] fp32 (wgrad output, n' order) and dbias_p[Nk] ->
 * dK[Cout,Cin,3,3], dbias[Cout] in reference layout (overwrites). */
int onr_unpack_wgrad(const float* dKp, const float* dbias_p, int Cin, int Cnew, int s,
                     float* dK, float* dbias, void* stream);

/* ------------------------------------------------------------------ A4+A5: block conv3x3 + PixelShuffle + SiLU
 * model.py:539 / :523 / :520 (F.conv2d 3x3 pad 1) and :567 (up_scale, act).  tcgen05/TMEM implicit
 * GEMM fed by TMA.  A plan owns only TMA descriptors for fixed buffers; create once, run per step. */
typedef struct onr_conv_plan onr_conv_plan;

/* ONR_CONV_FPROP_Z: the epilogue stores the bf16 PRE-activation z = conv + bias only (no SiLU, no SiLU' map): for the
 * last block of a training model, whose output feeds nothing but the RGB head — onr_head_fwd_z / onr_head_bwd_z
 * evaluate SiLU(z) and SiLU'(z) on the fly, which saves one 177 MB map per 720p frame and most of the epilogue.
 * ONR_CONV_FPROP_HEAD: decode of the LAST block with the RGB head (model.py:620-623) fused into the epilogue:
 * SiLU, the 1x1 conv C -> 3, bias and (tanh+1)/2 (or sigmoid) are applied to the accumulator tile and only the fp32
 * image is written — the block's 177 MB activation (720p) never exists.  Needs out_cp <= 256 (one N tile per
 * PixelShuffle sub-pixel). */
enum { ONR_CONV_FPROP_TRAIN = 0, ONR_CONV_FPROP_INFER = 1, ONR_CONV_DGRAD = 2, ONR_CONV_FPROP_Z = 3,
       ONR_CONV_FPROP_HEAD = 4 };

typedef struct {
    int kind;        /* ONR_CONV_* */
    int B, H, W;     /* GEMM pixel grid = the conv's input grid */
    /* A operand: NHWC bf16 [B][H*a_s][W*a_s][a_cp], read through the "un-shuffled" view
       (a_s = 1: plain activation for fprop; a_s = s: dZ of this block for dgrad). */
    const void* a; int a_cp; int a_s;
    /* B operand: packed weights [9][n_rows][k_tap] bf16, k_tap = a_s*a_s*a_cp */
    const void* w; int n_rows;
    int n_total;     /* valid GEMM N (multiple of 32) = out_s*out_s*out_cp */
    /* outputs: NHWC bf16 [B][H*out_s][W*out_s][out_cp] written through the shuffled view */
    void* out; int out_cp; int out_s;
    void* out_d;     /* FPROP_TRAIN: SiLU'(z) in the same layout as out */
    const float* bias_p;  /* FPROP_*: [n_total] in n' order */
    const void* dmul;     /* DGRAD: NHWC bf16 [B][H][W][n_total] multiplied into the result */
    /* FPROP_HEAD only: head weight [3][head_c] and bias [3] (fp32, reference layout of head_layers.N), activation
       flag, and the image fp32 NCHW [B][3][H*out_s][W*out_s] (re-bindable with onr_conv_plan_set_image) */
    const float* head_w; const float* head_b; int head_c; int use_sigmoid; float* img;
} onr_conv_desc;

int onr_conv_plan_create(onr_conv_plan** plan, const onr_conv_desc* desc);
int onr_conv_plan_run(const onr_conv_plan* plan, void* stream);
/* FPROP_HEAD plans: bind the image buffer (a fresh tensor per decoded frame on the reference API) and the head tensors
 * of the next runs. */
int onr_conv_plan_set_head(onr_conv_plan* plan, float* img, const float* head_w, const float* head_b);
void onr_conv_plan_destroy(onr_conv_plan* plan);
/* Tiling a plan chose: N tile, number of N tiles, 128-pixel sub-tiles per CTA tile, A / weight ring depths. */
int onr_conv_plan_info(const onr_conv_plan* plan, int* block_n, int* n_tiles, int* ms, int* na, int* nb);
/* Test/profiling hook: per-CTA cycle counters (8 x int64 per CTA: total, MMA wait A, MMA wait weights, MMA wait
 * accumulator, epilogue wait accumulator, epilogue wait store, epilogue busy, tiles); NULL switches it off. */
int onr_conv_plan_set_prof(onr_conv_plan* plan, long long* prof_dev, int* grid);
/* block_n / n_tiles the plan will use for a given N, so callers can size packed weights. */
int onr_conv_tile_n(int n_total, int* block_n, int* n_tiles);

/* A8 (wgrad): dKp[Nk][9][Cpi] (fp32, += via red.global) from x NHWC [B][H][W][Cpi] and
 * dz NHWC [B][H*s][W*s][Cpo] (read through the un-shuffled view).  tcgen05 with MN-major operands. */
typedef struct onr_wgrad_plan onr_wgrad_plan;
typedef struct {
    int B, H, W;
    const void* x; int x_cp;
    const void* dz; int dz_cp; int s;
    float* dKp;      /* [s*s*dz_cp][9][x_cp], must be zeroed by the caller before the run */
    float* dbias_p;  /* [s*s*dz_cp], zeroed by the caller; column sums of dz */
} onr_wgrad_desc;
int onr_wgrad_plan_create(onr_wgrad_plan** plan, const onr_wgrad_desc* desc);
int onr_wgrad_plan_run(const onr_wgrad_plan* plan, void* stream);
void onr_wgrad_plan_destroy(onr_wgrad_plan* plan);

/* Layout converters at the module boundary (NeRVBlock.forward takes/returns NCHW fp32). */
int onr_nchw_to_nhwc_bf16(const float* src, int B, int C, int H, int W, int Cp, void* dst, void* stream);
int onr_nhwc_bf16_to_nchw(const void* src, int B, int C, int H, int W, int Cp, float* dst, void* stream);

/* ------------------------------------------------------------------ A5 for the other activations (8f-4)
 * model.py:86-117 (ActivationLayer) + :567.  Activation codes: 0 swish, 1 relu, 2 leaky (0.01), 3 leaky01 (0.1),
 * 4 relu6, 5 gelu (erf), 6 softplus, 7 hardswish, 8 sin.  swish is fused into the convolution epilogue
 * (ONR_CONV_FPROP_TRAIN / _INFER); for the others the block runs in ONR_CONV_FPROP_Z mode (pre-activation z) and this
 * kernel turns z[pixels][Cp] (NHWC bf16) into y = act(z) IN PLACE and d = act'(z) (d may be NULL: decode). */
int onr_act_map(void* zy_bf16, void* d_bf16, size_t pixels, int C, int Cp, int act, void* stream);

/* ------------------------------------------------------------------ multi-resolution heads (8f-4; sin_res=False)
 * model.py:598-608, :615-623: a 1x1 RGB head on EVERY stage; main_train.py:239-244: the frame is pooled to each head's
 * resolution (F.adaptive_avg_pool2d) and the per-stage losses are summed with weight --lw for all but the last.
 * The per-stage heads reuse onr_head_fwd / onr_head_bwd; the two helpers below are what is new:
 * dst += src over n bf16 elements (n % 8 == 0) — the head's gradient joins the dgrad output of the next block; and
 * adaptive average pooling of NCHW fp32 planes [planes][H][W] -> [planes][Ho][Wo] (PyTorch's window rule). */
int onr_add_bf16(void* dst_bf16, const void* src_bf16, size_t n, void* stream);
int onr_adaptive_avg_pool(const float* src, int planes, int H, int W, int Ho, int Wo, float* dst, void* stream);

/* ------------------------------------------------------------------ A6: RGB head
 * model.py:601, :620-623: 1x1 conv C->3 + bias, then (tanh+1)/2 or sigmoid.
 * y NHWC bf16 [B][H][W][Cp] -> img NCHW fp32 [B][3][H][W]. */
int onr_head_fwd(const void* y_bf16, int B, int H, int W, int C, int Cp,
                 const float* Wh /*[3,C]*/, const float* bh /*[3]*/, int use_sigmoid,
                 float* img, void* stream);
/* Backward: gimg NCHW fp32, img (saved output) -> gWh,gbh (+=) and
 * dz[B][H][W][Cp] bf16 = (Wh^T g_pre) * dsilu  (dsilu = SiLU' of the last block, same layout). */
int onr_head_bwd(const float* gimg, const float* img, const void* y_bf16, const void* dsilu_bf16,
                 int B, int H, int W, int C, int Cp, const float* Wh, int use_sigmoid,
                 float* gWh, float* gbh, void* dz_bf16, void* stream);

/* The same head on the last block's bf16 PRE-activation z (written by an ONR_CONV_FPROP_Z plan): y = SiLU(z) and
 * SiLU'(z) are evaluated inside the kernels.  Single image only (B == 1, H*W % 4 == 0: the bulk-copy streaming path). */
int onr_head_fwd_z(const void* z_bf16, int B, int H, int W, int C, int Cp, const float* Wh, const float* bh,
                   int use_sigmoid, float* img, void* stream);
int onr_head_bwd_z(const float* gimg, const float* img, const void* z_bf16, int B, int H, int W, int C, int Cp,
                   const float* Wh, int use_sigmoid, float* gWh, float* gbh, void* dz_bf16, void* stream);

/* The two halves of onr_head_bwd as separate launches (dz is on the critical path of the backward; the weight /
 * bias gradient reduction is not and can run on another stream). */
int onr_head_bwd_dz(const float* gimg, const float* img, const void* dsilu_bf16, int B, int H, int W, int C, int Cp,
                    const float* Wh, int use_sigmoid, void* dz_bf16, void* stream);
int onr_head_bwd_gw(const float* gimg, const float* img, const void* y_bf16, int B, int H, int W, int C, int Cp,
                    int use_sigmoid, float* gWh, float* gbh, void* stream);

/* ------------------------------------------------------------------ A7+A10: Fusion6 loss, PSNR, MS-SSIM
 * utils.py:159-160 with pytorch_msssim 0.2.1 ssim (11-tap sigma 1.5 valid windows, C1=1e-4, C2=9e-4),
 * utils.py:191-199 (psnr_fn).  pred/target NCHW fp32 [B][3][H][W].
 * out[0]=loss, out[1]=L1 mean, out[2]=SSIM mean, out[3]=MSE, out[4]=PSNR (over the whole batch).
 * grad_pred (may be NULL) receives dloss/dpred * grad_scale.  work: onr_loss_workspace_bytes(). */
size_t onr_loss_workspace_bytes(int B, int H, int W);
int onr_fusion6_fwd_bwd(const float* pred, const float* target, int B, int H, int W,
                        float w_l1, float w_ssim, float grad_scale,
                        float* out5, float* grad_pred, void* work, void* stream);
/* The general member of the family utils.py:142-166 spells out (L2, L1, SSIM, Fusion1..9):
 * loss = w_l1 * mean|p-t| + w_mse * mean (p-t)^2 + w_ssim * (1 - SSIM); same outputs as above. */
int onr_fusion_loss(const float* pred, const float* target, int B, int H, int W,
                    float w_l1, float w_mse, float w_ssim, float grad_scale,
                    float* out5, float* grad_pred, void* work, void* stream);
/* x[0..n) *= *scalar_dev (autograd's upstream gradient of the loss, main_train.py:243-244, without a host sync). */
int onr_scale_by_device_scalar(float* x, size_t n, const float* scalar_dev, void* stream);
/* utils.py:201-211 / pytorch_msssim.ms_ssim: 5 scales, avg_pool2d(2, padding = dim%2). out[1]. */
size_t onr_msssim_workspace_bytes(int B, int H, int W);
int onr_msssim(const float* pred, const float* target, int B, int H, int W, float* out1,
               void* work, const void* loss_work /* NULL, or the workspace of an onr_fusion6_fwd_bwd call on
               the same (pred, target) ordered before this call: its scale-0 statistics are reused */,
               void* stream);

/* ------------------------------------------------------------------ A9: fused multi-tensor Adam
 * torch.optim.Adam as used at main_train.py:196, :248-250 (betas (beta,0.999), eps 1e-8, no wd).
 * table: n_tensors entries {param, grad, m, v, numel} as 5 x uint64 each (device array).
 * lr_dev, step_dev: device scalars (lr from adjust_lr, utils.py:240-259; step count t >= 1).
 * grad_scale multiplies grads (1/world_size); grads are zeroed after use when zero_grad != 0.
 * One CTA updates one onr_adam_block_elems()-element piece of one tensor: total_blocks is the sum over the
 * table of ceil(numel / onr_adam_block_elems()); at most onr_adam_max_tensors() entries per call.
 * hyp_scratch: 2 floats of device memory the call overwrites (bias corrections, evaluated once in double). */
size_t onr_adam_block_elems(void);
int onr_adam_max_tensors(void);
int onr_adam_multi(const uint64_t* table, int n_tensors, size_t total_blocks,
                   const float* lr_dev, const int* step_dev, float* hyp_scratch, float beta1, float beta2,
                   float eps, float grad_scale, int zero_grad, void* stream);

/* adjust_lr (utils.py:240-259) on the device: step_dev += 1, lr_dev = lr0 * multiplier(step) with
 * cur_epoch = epoch + iter/data_size.  lr_type 0 = cosine, 1 = const; warm-up 0.1 -> 1 over `warmup` epochs. */
int onr_sched_tick(int* step_dev, float* lr_dev, double lr0, int steps_per_epoch, int data_size,
                   int warmup, int epochs, int lr_type, void* stream);
/* The same tick for the prune-then-finetune loop (main_eval.py:446-466): the optimizer restarts (step count from 1)
 * while the epoch numbering continues at the checkpoint's epoch; adjust_lr is evaluated at
 * (epoch_offset + step / steps_per_epoch) % epoch_mod against the original --epochs / warm-up. */
int onr_sched_tick_ex(int* step_dev, float* lr_dev, double lr0, int steps_per_epoch, int data_size,
                      int warmup, int epochs, int lr_type, int epoch_offset, int epoch_mod, void* stream);

/* ------------------------------------------------------------------ A12/A13: eval-side weight transforms
 * main_eval.py:572-587 (prune.global_unstructured, L1Unstructured): the global k-th smallest |w| is found
 * by an exact 4-pass radix select over the fp32 bit patterns: hist[256] += histogram of byte
 * (bits(|w|) >> shift) & 255 over the elements whose bits match `prefix` under `prefix_mask`.
 * Call once per tensor per pass (hist accumulates); the host picks the bucket between passes. */
int onr_abs_radix_hist(const float* w, size_t n, uint32_t prefix, uint32_t prefix_mask, int shift,
                       unsigned long long* hist256, void* stream);
/* mask[i] = |w[i]| > thr ? 1 : 0 (fp32, like torch's weight_mask); w_out = w * mask. Either may be NULL. */
int onr_apply_magnitude_mask(const float* w, size_t n, float thr, float* mask, float* w_out, void* stream);
/* g[i] *= m[i] over n floats (both 16-byte aligned): the gradient side of prune's weight = weight_orig * weight_mask
 * (main_eval.py:213-545, prune-then-finetune; torch.nn.utils.prune's forward-pre-hook + autograd in the reference). */
int onr_mul_inplace_f32(float* g, const float* m, size_t n, void* stream);
/* utils.py:11-67 quantize_per_tensor with axis 0 over rows of a [rows, cols] view (axis -1: rows=1):
 * min/max over non-zero entries, scale=(max-min)/2^bit, q=round((t-min)/(scale+1e-19)),
 * new=min+scale*q.  q_out and new_out may alias nothing; either may be NULL. */
int onr_quant_rows(const float* t, int rows, size_t cols, int bit, float* q_out, float* new_out,
                   void* scratch /* rows * 8 bytes */, void* stream);

/* uint8 frame (PNG sample) -> fp32 in [0,1], bit-identical to torchvision ToTensor (model.py:64-65). */
int onr_frame_u8_to_f32(const void* src_u8, size_t n, float* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OREPNERV_H */
