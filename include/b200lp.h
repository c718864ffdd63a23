/*
 * b200lp.h — C ABI of libb200lp.so, the sm_100a arithmetic behind the latent-pose-reenactment hot path.
 *
 * The reference (shrubb/latent-pose-reenactment) has NO native boundary: every FLOP on the path is a stock
 * torch/torchvision operator called from Python plugin modules.  This header is therefore the
 * boundary a maintainer would bind *instead of* those torch operators; each entry point names the reference
 * call site(s) it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B200LP_E* code on failure; b200lp_last_error() gives text.
 *   - no exceptions, no torch types: raw device pointers, int32/int64 sizes, a cudaStream_t passed as void*.
 *   - the library never allocates device memory and never synchronises; it launches only on the given stream.
 *   - activations are NHWC fp32 ("pixels x channels"), weights are "packed" [Cout][tap][Cin] (see b200lp_pack_*).
 *   - tensor-core operands are TF32 (10-bit mantissa), accumulation FP32 in TMEM.
 */
#ifndef B200LP_H_
#define B200LP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LP_ABI_VERSION 18

#define B200LP_OK 0
#define B200LP_EINVAL (-1)   /* bad shape / unsupported configuration */
#define B200LP_ECUDA (-2)    /* a CUDA runtime / driver call failed   */
#define B200LP_ENODEV (-3)   /* no sm_100 device                      */

int32_t b200lp_abi_version(void);
const char* b200lp_last_error(void);
/* compute capability major*10+minor of the current device, or a negative error */
int32_t b200lp_device_cc(void);
/* number of kernels this library has launched so far in this process (a monotone counter) */
int64_t b200lp_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TF32 x TF32 -> FP32), stride 1, "same" zero padding.
 * Replaces nn.Conv2d forward at  generators/common/blocks.py:78-80,86-88,98-100 (G ResBlock convs + 1x1 skip),
 * discriminators/no_landmarks.py:54-66 and blocks.py ResBlock(norm='none') (D convs),
 * criterions/common/perceptual_loss.py:104 (VGG19 / VGG16 feature convs);
 * and, with weights packed by b200lp_pack_conv_weight(..., transpose=1), the data-gradient of the same convs
 * (torch autograd's conv backward-data).
 *
 *   y[n,h,w,co] = epilogue( sum_{kh,kw,ci} x[n,h+kh-p,w+kw-p,ci] * wp[co][kh*k+kw][ci] )
 *   epilogue: (* out_scale) (+ bias[co]) (+ residual) (relu) (round to tf32)
 * Requirements: Cin % 32 == 0, Cout % 32 == 0, H and W powers of two >= 2, ksize in {1,3}.
 */
typedef struct {
    const void* x;          /* precision 0: float [N,H,W,Cin] NHWC (tf32-rounded values)
                               precision 1: bf16  [2][N,H,W,Cin] = (hi, lo) planes, hi + lo == value            */
    const void* wp;         /* packed weights [Cout][ksize*ksize][Cin]: float (tf32) or bf16 [2][...] (hi, lo)   */
    const float* out_scale; /* NULL, or device pointer to ONE float s: y = s * conv(x, wp) (+bias...).  With s = 1/sigma
                               the spectral-norm division costs nothing and `wp` is packed once per weight update
                               instead of once per forward call (3 discriminator passes share one packing).       */
    const float* bias;      /* [Cout] or NULL                                                                    */
    const float* residual;  /* NULL, or [N,H,W,Cout] (mode 1), or [N,H/2,W/2,Cout] (mode 2: nearest-2x source)    */
    float* y;               /* [N,H,W,Cout] fp32                                                                 */
    void* y_split;          /* NULL, or bf16 [2][N,H,W,Cout]: the same result as (hi, lo) planes (operand of a
                               following bf16x3 convolution)                                                    */
    int32_t N, H, W, Cin, Cout;
    int32_t ksize;          /* 1 or 3                                                  */
    int32_t residual_mode;  /* 0 none, 1 same resolution, 2 half resolution            */
    int32_t relu;           /* 1: y = max(y, 0)                                        */
    int32_t round_tf32;     /* 1: round y to tf32 (y only feeds further tf32 MMAs)     */
    int32_t block_n;        /* 0 = auto; else 32 / 64 / 128 / 256 (256: tf32 only)     */
    int32_t precision;      /* 0 = tf32 (1 MMA / K-step), 1 = bf16x3 (3 MMAs: Ah*Bh + Ah*Bl + Al*Bh, ~fp32 accuracy) */
    int32_t stages;         /* 0 = auto; else depth of the shared-memory operand ring (tuning knob)                 */
    int32_t ctas_per_sm;    /* 0 = auto; else 1 / 2 persistent CTAs per SM (tuning knob)                            */
    int32_t splits;         /* 0 = auto split-K for layers with too few tiles to fill the GPU, 1 = never, >1 forced  */
    int32_t variant;        /* 0 = auto; -1 = per-tap kernel (one activation box per filter tap);
                               1 / 2 / 4 = halo kernel (3x3, H >= 16*variant, W >= 8): 8 x 16*variant pixel tiles, three
                               column-shifted activation slabs per channel block shared by the 3 row taps, `variant`
                               accumulators sharing every weight tile.  In the halo kernel `stages` is the depth of the
                               weight-tile ring and `a_stages` that of the slab ring.                               */
    int32_t a_stages;       /* 0 = auto (halo kernel only)                                                          */
    int32_t grouped;        /* 1 = block-diagonal (grouped) 3x3 convolution, Cin == Cout: `wp` is
                               [Cout][9][B] (B = 32 tf32 / 64 bf16x3, b200lp_pack_gconv_weight) and output channels
                               [nB, (n+1)B) read input channels [nB, (n+1)B) only — torchvision ResNeXt's
                               Conv2d(groups=32) with 4..32 channels per group as dense B x B blocks                 */
    int32_t reserved0;
    float* workspace;       /* split-K partial sums; NULL = never split.  Size: b200lp_conv_fwd_workspace(args)      */
    int64_t workspace_bytes;
} b200lp_conv_args;

/* bytes of split-K workspace the call described by `a` would use (0 if it will not split; pointers in `a` are ignored) */
int64_t b200lp_conv_fwd_workspace(const b200lp_conv_args* a);
int32_t b200lp_conv_fwd(const b200lp_conv_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Weight packing (every step: weights change with every optimizer update, and spectral norm rescales them).
 * Replaces the `weight = weight_orig / sigma` materialisation of torch.nn.utils.spectral_norm
 * (SpectralNorm.compute_weight; call sites blocks.py:78-100, generator :84-86, discriminator :54-66).
 *   transpose = 0:  wp[co][tap][ci]      = tf32( w[co][ci][kh][kw] * (*scale) )          (forward)
 *   transpose = 1:  wp[ci][T-1-tap][co]  = tf32( w[co][ci][kh][kw] * (*scale) )          (data-gradient)
 * `scale` is a device pointer to one float (1/sigma) or NULL for 1.0.
 * precision 0: `wp` is float (tf32-rounded); precision 1: `wp` is bf16 [2][...] = (hi, lo) planes (bf16x3 operand).
 */
int32_t b200lp_pack_conv_weight(const float* w_oihw, const float* scale, void* wp, int32_t Cout, int32_t Cin,
                                int32_t ksize, int32_t transpose, int32_t precision, void* stream);

/* Multi-tensor form: ONE launch re-packs every (weight, transpose, precision) copy of a network after its optimizer
 * update (scale = 1: the spectral-norm 1/sigma is applied by the conv epilogue).  `table` (device): one row of
 * 8 int64 per copy = { w_oihw pointer, wp pointer, Cout, Cin, taps (1 or 9), transpose, precision, Cout*Cin*taps };
 * work is cut into chunks of `chunk_elems` packed elements: chunk c covers row chunk_item[c], elements
 * [chunk_off[c], chunk_off[c] + chunk_elems). */
int32_t b200lp_pack_conv_weight_multi(const void* table, const int32_t* chunk_item, const int64_t* chunk_off,
                                      int32_t n_chunks, int64_t chunk_elems, void* stream);

/* Tiled form of the multi-tensor packing (same `table` rows; taps <= 9): one block per 32 (co) x 32 (ci) tile, tile t of
 * row tile_item[t] is tile_index[t] = co_tile * ceil(Cin/32) + ci_tile.  Source rows are read coalesced through shared
 * memory; results are bit-identical to b200lp_pack_conv_weight. */
int32_t b200lp_pack_conv_weight_tiles(const void* table, const int32_t* tile_item, const int32_t* tile_index,
                                      int32_t n_tiles, void* stream);

/* Grouped 3x3 weight w[C][cpg][3][3] (groups of cpg channels, cpg divides `block`; block = 32 for tf32, 64 for bf16x3)
 * as block-diagonal dense tiles for b200lp_conv_fwd(grouped = 1):
 *   transpose = 0:  wp[co][tap][j]     = w[co][ci - g*cpg][tap]    if ci = (co/block)*block + j lies in co's group g, else 0
 *   transpose = 1:  wp[ci][8-tap][j]   = w[co][ci - g*cpg][tap]    with co = (ci/block)*block + j           (data-gradient)
 * precision as in b200lp_pack_conv_weight. */
int32_t b200lp_pack_gconv_weight(const float* w, void* wp, int32_t C, int32_t cpg, int32_t transpose, int32_t precision,
                                 void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Spectral normalisation, batched over all weights of a network pass (3 launches instead of ~14 per weight).
 * Replaces torch.nn.utils.spectral_norm's SpectralNorm.compute_weight (one power iteration per forward call in
 * training mode; call sites generators/common/blocks.py:78-100, discriminators/no_landmarks.py:54-66):
 *   training: v <- normalize(W^T u), u <- normalize(W v), sigma = u^T W v   (u, v updated in place)
 *   eval:     sigma = u^T W v with the stored u, v
 * Writes 1/sigma (consumed by b200lp_conv_fwd as `out_scale`) and, if snap_u / snap_v are given, a copy of the
 * vectors this sigma was computed with (the backward pass needs them after later passes overwrote the buffers).
 * `scratch` must hold b200lp_sn_scratch_floats(rows, cols) floats per item.  count <= b200lp_sn_max_tensors().
 */
typedef struct {
    const float* w;      /* weight_orig viewed as [rows = Cout][cols = Cin*k*k]        */
    float* u;            /* weight_u [rows]                                            */
    float* v;            /* weight_v [cols]                                            */
    float* snap_u;       /* [rows] or NULL                                             */
    float* snap_v;       /* [cols] or NULL                                             */
    float* scratch;      /* b200lp_sn_scratch_floats(rows, cols) floats                */
    float* inv_sigma;    /* [1] out                                                    */
    int32_t rows, cols;
    float eps;           /* torch's normalize eps (1e-4 on the hot path)               */
    int32_t reserved;
} b200lp_sn_item;

int32_t b200lp_sn_max_tensors(void);
int64_t b200lp_sn_scratch_floats(int32_t rows, int32_t cols);
int32_t b200lp_sn_sigma_multi(const b200lp_sn_item* items /* host array */, int32_t count, int32_t training,
                              void* stream);
/* Spectral-norm correction of a weight gradient (autograd through `weight_orig / sigma`):
 *   dw = s*g - s^2 <g, w> u v^T,  s = *inv_sigma, g = gradient w.r.t. the normalised weight (b200lp_conv_wgrad output),
 *   all [rows][cols].  workspace: b200lp_sn_wgrad_fix_workspace(rows*cols) bytes. */
int64_t b200lp_sn_wgrad_fix_workspace(int64_t n);
int32_t b200lp_sn_wgrad_fix(const float* g, const float* w, const float* inv_sigma, const float* u, const float* v,
                            float* dw, float* workspace, int32_t rows, int32_t cols, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Weight gradient of the same convolutions on tcgen05 (torch autograd's conv backward-filter).
 *   dw[co][ci][kh][kw] (OIHW, fp32) = sum_{n,h,w} dy[n,h,w,co] * x[n,h+kh-p,w+kw-p,ci]
 * `workspace` holds split-K partial sums; query its size with b200lp_conv_wgrad_workspace().
 */
typedef struct {
    const float* x;   /* [N,H,W,Cin]  NHWC (the conv's forward input, tf32-rounded)   */
    const float* dy;  /* [N,H,W,Cout] NHWC                                            */
    float* dw;        /* [Cout][Cin][k][k] OIHW                                       */
    float* workspace; /* >= b200lp_conv_wgrad_workspace() bytes                       */
    int64_t workspace_bytes;
    int32_t N, H, W, Cin, Cout;
    int32_t ksize;
    float scale;      /* dw *= scale (e.g. 1/sigma of the spectral norm)              */
    int32_t kstep;    /* tuning: pixels per pipeline stage, 0 = auto, else 32 / 64    */
    int32_t stages;   /* tuning: smem ring depth, 0 = auto                            */
    int32_t splits;   /* tuning: split-K factor, 0 = auto (workspace must hold splits * |dw| floats) */
    int32_t grouped;  /* 0 = dense; cpg > 0: grouped 3x3 conv (Cin == Cout, groups of cpg channels) — only
                         b200lp_gconv3x3_wgrad_tc accepts it */
} b200lp_wgrad_args;

int64_t b200lp_conv_wgrad_workspace(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize);
int32_t b200lp_conv_wgrad(const b200lp_wgrad_args* a, void* stream);

/* The same weight gradient delivered straight into the parameter's gradient buffer (the `.grad` view inside the
 * data-parallel bucket, runners/holycow.py GradBucket), with the spectral-norm rank-1 term fused:
 *   G = wgrad(x, dy);   grad (+)= s*G - s^2 <G, w> u v^T,   s = *inv_sigma       (a->dw = grad, OIHW; a->scale unused)
 * inv_sigma == NULL: grad (+)= G (w, u, v ignored).  Three launches: tensor-core split-K kernel, tiled reduction
 * (+ accumulate + <G, w> partials; coalesced on both sides), rank-1 update.  Replaces conv backward-filter + the autograd
 * of `weight_orig / sigma` + AccumulateGrad's add_.  Tuning knobs of `a` must be 0.
 * Workspace: b200lp_conv_wgrad_sn_acc_workspace() bytes. */
int64_t b200lp_conv_wgrad_sn_acc_workspace(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize);
int32_t b200lp_conv_wgrad_sn_acc(const b200lp_wgrad_args* a, const float* w, const float* inv_sigma, const float* u,
                                 const float* v, int32_t accumulate, void* stream);

/* Weight gradient of a GROUPED 3x3 convolution (stride 1, padding 1) on tcgen05: per 32-channel block one
 * [4 taps x 32 ci] x [32 co] TF32 GEMM over the pixels (the off-group products of the block are computed and dropped),
 * split-K partial sums, then a deterministic reduction that keeps the block-diagonal entries:
 *   dw[co][cig][tap] (+)= sum_{n,h,w} dy[n,h,w,co] * x[n,h+kh-1,w+kw-1, g(co)*cpg + cig]        (a->dw: [C][cpg][3][3])
 * a->grouped = cpg, a->Cin == a->Cout == C, a->ksize == 3; tuning knobs 0.  Replaces autograd's backward-filter of
 * torchvision ResNeXt's conv2 (embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:27). */
int64_t b200lp_gconv3x3_wgrad_tc_workspace(int32_t N, int32_t H, int32_t W, int32_t C);
int32_t b200lp_gconv3x3_wgrad_tc(const b200lp_wgrad_args* a, int32_t accumulate, void* stream);

/* out[n, 2i, 2j, c] = x[n, i, j, c], zero elsewhere (out is [N, 2H, 2W, C]): the gradient of a stride-2 convolution's
 * output placed on the stride-1 grid, so that its data / weight gradients run through the stride-1 tensor-core kernels. */
int32_t b200lp_zero_stuff2(const float* x, float* out, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Instance-norm statistics, AdaIN affine + ReLU (+ nearest 2x upsample)  — HBM-bound.
 * Replaces nn.InstanceNorm2d(eps, affine=False) + `out*gamma+beta` (generators/common/blocks.py:18-26),
 * the following nn.ReLU(inplace) and nn.Upsample(scale_factor=2) (blocks.py:73-75).
 *   mean[n,c], rstd[n,c] over the H*W plane (biased variance, rstd = (var+eps)^-1/2)
 *   y = tf32( relu( (x-mean)*rstd*gamma[n,c] + beta[n,c] ) ),   optionally written 2x nearest-upsampled.
 * gamma/beta are rows of the projector output: element (n,c) at gamma[n*affine_stride + c].
 */
int64_t b200lp_in_stats_workspace(int32_t N, int32_t HW, int32_t C);
int32_t b200lp_in_stats(const float* x, float* mean, float* rstd, float* workspace, int64_t workspace_bytes,
                        int32_t N, int32_t HW, int32_t C, float eps, void* stream);
/* `y` (fp32, tf32-rounded if round_tf32) and/or `y_split` (bf16 [2][...]: (hi, lo) planes of the UNROUNDED result, the
 * operand of a bf16x3 convolution) may be given; at least one. */
int32_t b200lp_adain_relu(const float* x, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, int64_t affine_stride, float* y, void* y_split, int32_t N, int32_t H,
                          int32_t W, int32_t C, int32_t upsample2, int32_t round_tf32, void* stream);
/* The AdaIN site as ONE launch (b200lp_in_stats + b200lp_adain_relu): statistics partials, a barrier among the CTAs of a
 * sample, merge, apply — the apply pass re-reads the pixels the same CTA streamed for the statistics (L2-resident unless
 * the tensor exceeds the L2).  mean / rstd [N][C] are outputs (the backward pass needs them).  workspace:
 * b200lp_in_stats_workspace bytes.  sync: 4*N uint32 counters, zero before the FIRST use (the kernel leaves them zero); one
 * buffer per stream.  Fails (no launch) if the grid cannot be made co-resident — call the two-kernel form then. */
int32_t b200lp_adain_relu_fused(const float* x, const float* gamma, const float* beta, int64_t affine_stride, float* y,
                                void* y_split, float* mean, float* rstd, float* workspace, int64_t workspace_bytes,
                                uint32_t* sync, int32_t N, int32_t H, int32_t W, int32_t C, float eps, int32_t upsample2,
                                int32_t round_tf32, void* stream);
/* backward of the above: given dy (w.r.t. the post-ReLU, possibly 2x-upsampled output) produce
 * dx, dgamma[n,c], dbeta[n,c].  The ReLU mask is recomputed from x (no saved activation needed).
 * add (optional, shaped like x): a second gradient of x (the residual block's skip branch) merged into dx;
 * round_tf32: dx stored rounded to tf32 (it only feeds the previous layer's tf32 gradient MMAs). */
int64_t b200lp_adain_relu_bwd_workspace(int32_t N, int32_t HW, int32_t C);
int32_t b200lp_adain_relu_bwd(const float* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, int64_t affine_stride, const float* dy, float* dx, float* dgamma,
                              float* dbeta, float* workspace, int64_t workspace_bytes, int32_t N, int32_t H,
                              int32_t W, int32_t C, int32_t upsample2, const float* add, int32_t round_tf32,
                              void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Small elementwise / reduction kernels (all NHWC unless stated).
 */
/* NCHW <-> NHWC (image boundary of the plugins: data_dict tensors are NCHW, reference layout) */
int32_t b200lp_nchw_to_nhwc(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, void* stream);
int32_t b200lp_nhwc_to_nchw(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, void* stream);
/* y = tf32(relu(x)) (D's in-place ReLU, blocks.py:73 with norm='none'); mask variant for backward: dx = dy*[y>0] */
int32_t b200lp_relu_round(const float* x, float* y, int64_t n, void* stream);
int32_t b200lp_relu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream);
/* ReLU backward fused with its neighbours in the discriminator's backward chain (discriminators/no_landmarks.py:90-108,
 * generators/common/blocks.py:70-103 with norm_layer='none'):
 *   dx = [y > 0] * (dy + add)            add (optional): a second gradient of the same tensor (feature-matching term)
 *   dq = 0.25 * dx                       (optional) the gradient behind an AvgPool2d(2), kept at the low resolution
 *   db_a[c] += sum_p dx[p][c], db_b[c] += ...   (optional) bias gradients of the convs that produced the tensor
 * round_tf32: dx (and dq) are stored rounded to nearest tf32 — they only feed tf32 MMAs, which would otherwise truncate;
 * the bias sums use the unrounded values.   y, dy, add, dx, dq: [pixels][C]; C % 4 == 0, C <= 1024. */
int32_t b200lp_relu_bwd_fused(const float* y, const float* dy, const float* add, float* dx, float* dq, float* db_a,
                              float* db_b, int64_t pixels, int32_t C, int32_t round_tf32, void* stream);
/* 2x2 average pool / its backward (nn.AvgPool2d(2): blocks.py:89-90,101-102; perceptual_loss.py:77) */
int32_t b200lp_avgpool2(const float* x, const float* addend, float* y, int32_t N, int32_t H, int32_t W, int32_t C,
                        int32_t round_tf32, void* stream);
int32_t b200lp_avgpool2_bwd(const float* dy, float* dx, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
/* backward of nearest 2x upsample: dx[n,h,w,c] = sum of the 2x2 block of dy */
int32_t b200lp_upsample2_bwd(const float* dy, float* dx, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
/* sum |a-b| over n elements -> out[0] += result * scale (L1 feature losses: perceptual_loss.py:108, featmat.py:18-20)
 * and its gradient wrt a: da = sign(a-b) * gscale[0]*scale2 (accumulated into da if accumulate) */
int32_t b200lp_l1_sum(const float* a, const float* b, float* out, int64_t n, float scale, void* stream);
int32_t b200lp_l1_bwd(const float* a, const float* b, const float* gscale, float scale2, float* da, int64_t n,
                      int32_t accumulate, void* stream);
/* one VGG backward tap in one pass (perceptual_loss.py:104-108 backward): `a` is a post-ReLU feature, so
 *   d_out = [a > 0] * ( (d_in or 0) + sign(a-b) * gscale[0]*scale2 )      = l1_bwd followed by relu_bwd */
int32_t b200lp_l1_relu_bwd(const float* a, const float* b, const float* gscale, float scale2, const float* d_in,
                           float* d_out, int64_t n, void* stream);
/* The same tap split so that the backward pass needs neither feature map: the forward pass adds scale * sum|a-b| to out[0]
 * and writes one code byte per 4 elements (2 bits each: 0 = a <= 0, 1 / 2 / 3 = a > 0 and sign(a-b) = -1 / 0 / +1);
 *   d_out = code ? tf32( (d_in or 0) + (code - 2) * gscale[0]*scale2 ) : 0
 * code: [n/4] bytes.  24 -> 16.5 bytes of HBM traffic per element and tap, and the real-image branch's activations are
 * not kept for the backward pass (criterions/common/perceptual_loss.py:104-110). */
int32_t b200lp_l1_sum_code(const float* a, const float* b, float* out, uint8_t* code, int64_t n, float scale,
                           void* stream);
int32_t b200lp_l1_code_bwd(const uint8_t* code, const float* gscale, float scale2, const float* d_in, float* d_out,
                           int64_t n, void* stream);
/* The tap in front of an AvgPool2d(2) (a, b: [N,H,W,C], H and W even): the forward pass also writes the pooled, tf32-rounded
 * maps a_pool, b_pool [N,H/2,W/2,C] (= b200lp_avgpool2 of each); the backward pass takes the gradient of the POOLED map,
 *   d_out[n,h,w,c] = code ? tf32( 0.25 * d_low[n,h/2,w/2,c] + (code - 2) * gscale[0]*scale2 ) : 0 . */
int32_t b200lp_l1_sum_code_pool(const float* a, const float* b, float* out, uint8_t* code, float* a_pool, float* b_pool,
                                int32_t N, int32_t H, int32_t W, int32_t C, float scale, void* stream);
int32_t b200lp_l1_code_bwd_unpool(const uint8_t* code, const float* gscale, float scale2, const float* d_low, float* d_out,
                                  int32_t N, int32_t H, int32_t W, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Direct (CUDA-core) convolutions for the two degenerate, HBM-bound shapes (SURVEY §7 "Degenerate GEMM shapes").
 */
/* 3x3, Cin=3 (NCHW image in) -> Cout (NHWC out) (+bias)(relu)(round): D stem discriminators/no_landmarks.py:54,
 * VGG features.0 perceptual_loss.py:104.  `pre_scale/pre_shift[3]`: x' = x*pre_scale[c] + pre_shift[c] is applied
 * on the fly before zero padding (VGG input normalisation perceptual_loss.py:88-98). May be NULL. */
int32_t b200lp_conv3x3_c3_fwd(const float* x_nchw, const float* w_oihw, const float* wscale, const float* bias,
                              const float* pre_scale, const float* pre_shift, float* y_nhwc, int32_t N, int32_t H,
                              int32_t W, int32_t Cout, int32_t relu, int32_t round_tf32, void* stream);
/* The same layer (Cout = 64) on tcgen05: M = 128 pixels x N = 64 x K = 27 (+5) implicit GEMM whose A tile is built in shared
 * memory by the pixel threads (swizzled K-major rows, input normalisation + tf32 rounding applied on the way); operands
 * are tf32 like every other layer of the discriminator / VGG networks. */
int32_t b200lp_conv3x3_c3_fwd_tc(const float* x_nchw, const float* w_oihw, const float* wscale, const float* bias,
                                 const float* pre_scale, const float* pre_shift, float* y_nhwc, int32_t N, int32_t H,
                                 int32_t W, int32_t Cout, int32_t relu, int32_t round_tf32, void* stream);
/* its data gradient (NHWC dy -> NCHW dx, 3 channels) */
int32_t b200lp_conv3x3_c3_dgrad(const float* dy_nhwc, const float* w_oihw, const float* wscale,
                                const float* pre_scale, float* dx_nchw, int32_t N, int32_t H, int32_t W,
                                int32_t Cout, void* stream);
/* its weight gradient: dw[co][3][3][3] (+= if accumulate) and dbias */
int32_t b200lp_conv3x3_c3_wgrad(const float* x_nchw, const float* dy_nhwc, float* dw_oihw, float wscale_host,
                                int32_t N, int32_t H, int32_t W, int32_t Cout, void* stream);

/* The gradients of those Cin=3 convs as tensor-core GEMMs over an explicit 27(+5 zero)-column patch matrix:
 *   col[n,h,w, c*9+kh*3+kw] = tf32(x[n,c,h+kh-1,w+kw-1]) (0 outside the image; columns 27..31 are 0)
 *   weight gradient = b200lp_conv_wgrad(col, dy, ksize=1)[:, :27];
 *   data gradient   = col2im( b200lp_conv_fwd(dy, W^T as a 1x1 conv 64 -> 32) ):
 *   dx[n,c,h,w] = pre_scale[c] * sum_{kh,kw} dcol[n,h-kh+1,w-kw+1, c*9+kh*3+kw] */
int32_t b200lp_im2col3x3_c3(const float* x_nchw, float* col_nhwc32, int32_t N, int32_t H, int32_t W, void* stream);
int32_t b200lp_col2im3x3_c3(const float* dcol_nhwc32, const float* pre_scale, float* dx_nchw, int32_t N, int32_t H,
                            int32_t W, void* stream);

/* Generator tail (generators/vector_pose_unsupervised_segmentation_noBottleneck.py:84-88,165-181):
 * a[n,h,w,0:4] = conv3x3(x[N,H,W,64] ; w[4][64][3][3]*(*wscale)) + bias ; t = tanh(a);
 * rgb = t[0:3]*0.75+0.5 ; segm = t[3]*0.5+0.5 ; fake_rgbs = rgb*segm (NCHW [N,3,H,W]) ; fake_segm (NCHW [N,1,H,W]).
 * `t_out` ([N,H,W,4]) keeps tanh for the backward. */
int32_t b200lp_gen_tail_fwd(const float* x_nhwc, const float* w_oihw, const float* wscale, const float* bias,
                            float* fake_rgbs_nchw, float* fake_segm_nchw, float* t_out, int32_t N, int32_t H,
                            int32_t W, int32_t Cin, void* stream);
/* Tensor-core form of the same tail: the 3x3 conv runs as b200lp_conv_fwd (bf16x3) on a weight zero-padded to 32 output
 * channels; this entry point is the remaining composition  t = tanh(a[..., 0:4] + bias) -> fake_rgbs, fake_segm, t_out.
 * `a_nhwc` has `a_stride` floats per pixel (32). */
int32_t b200lp_gen_tail_compose(const float* a_nhwc, const float* bias, float* fake_rgbs_nchw, float* fake_segm_nchw,
                                float* t_out, int32_t N, int32_t H, int32_t W, int32_t a_stride, void* stream);
/* backward: from d(fake_rgbs) [N,3,H,W], d(fake_segm) [N,1,H,W] (either may be NULL) and saved t:
 * da [N,H,W,4] (pre-tanh gradient), then dx = conv-transpose, dw, dbias */
/* da has `da_stride` (4 or 32) floats per pixel; with 32 the 28 extra channels are written as zeros so that `da` is a
 * 32-channel NHWC tensor the tensor-core weight- / data-gradient kernels accept (values rounded to tf32 in that form). */
int32_t b200lp_gen_tail_bwd_act(const float* t, const float* d_rgbs, const float* d_segm, float* da, int32_t N,
                                int32_t H, int32_t W, int32_t da_stride, void* stream);
int32_t b200lp_gen_tail_bwd_data(const float* da, const float* w_oihw, const float* wscale, float* dx_nhwc,
                                 int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t da_stride, void* stream);
int32_t b200lp_gen_tail_bwd_weight(const float* x_nhwc, const float* da, float* dw_oihw, float* dbias, int32_t N,
                                   int32_t H, int32_t W, int32_t Cin, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused multi-tensor optimizer step + weight running average (HBM-bound; graph-capturable: the step counter and the
 * step-dependent scalars live on the device).  Replaces torch.optim.Adam.step / the vendored RAdam.step
 * (utils/radam.py:29-95; runners/holycow.py:244,252) and TrainingModule.update_running_average's parameter loop
 * (runners/holycow.py:99-105).
 *   table_dev        : device array of {float* p; const float* g; float* m; float* v; float* ema (or NULL); int64 n}
 *   chunk_tensor_dev / chunk_off_dev : work list, chunk i covers table[chunk_tensor[i]] elements
 *                      [chunk_off[i], chunk_off[i] + chunk_elems)
 *   state_dev        : 8 floats {step, step_size, rectified flag, 1/sqrt(bias_correction2), lr, ema_alpha, -, -};
 *                      step advances by 1 per call; lr and ema_alpha are READ from the device vector (the caller
 *                      rewrites them between CUDA-graph replays when a schedule changes them)
 *   mode 0 = torch.optim.Adam semantics, 1 = RAdam (utils/radam.py) semantics;  ema = ema*ema_alpha + p*(1-ema_alpha)
 */
int32_t b200lp_adam_ema_multi(const void* table_dev, const int32_t* chunk_tensor_dev, const int64_t* chunk_off_dev,
                              int32_t n_chunks, int64_t chunk_elems, float* state_dev, float beta1, float beta2,
                              float eps, int32_t mode, int32_t degenerated_to_sgd, void* stream);
/* ema = ema*ema_alpha + p*(1-ema_alpha) over the same kind of table (entries with ema == NULL are skipped) */
int32_t b200lp_ema_multi(const void* table_dev, const int32_t* chunk_tensor_dev, const int64_t* chunk_off_dev,
                         int32_t n_chunks, int64_t chunk_elems, float ema_alpha, void* stream);

/* per-channel sum over pixels (bias gradients): db[c] = scale * sum_{n,h,w} dy[n,h,w,c] */
int32_t b200lp_bias_grad(const float* dy, float* db, int64_t pixels, int32_t C, void* stream);

/* the same, accumulated into db (db += ...): db is the bias parameter's gradient buffer */
int32_t b200lp_bias_grad_acc(const float* dy, float* db, int64_t pixels, int32_t C, void* stream);

/* dst <- src for `count` small buffers in ONE launch.  table_dev: device array of {void* dst; const void* src;
 * int64 nbytes}.  Replaces the per-buffer copies of TrainingModule.update_running_average (runners/holycow.py:106-109:
 * BatchNorm statistics and spectral-norm vectors of E and G, ~370 tensors per step). */
int32_t b200lp_copy_multi(const void* table_dev, int32_t count, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Pose encoder (torchvision MobileNetV2) forward — FP32 CUDA-core kernels (csrc/mobilenet.cu).
 * Replaces Embedder.get_pose_embedding (embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:56-58):
 * Conv2d(1x1 | depthwise 3x3 | 3x3 stride-2 stem) + BatchNorm2d(eps 1e-5, momentum 0.1, train-mode batch statistics or
 * eval-mode running statistics) + ReLU6, inverted-residual skips, adaptive average pool, Linear.
 * Activations NHWC fp32.  A conv kernel applies the PRODUCER layer's BatchNorm (+ReLU6) on load through per-channel
 * (scale, shift), writes its own raw output, and (if `part` != NULL) per-channel [part][2][C] sum / sum-of-squares
 * partials of that output for b200lp_bn_finalize.  All channel counts must be multiples of 4.
 */
/* 1x1 conv / linear: y[M][Cout] = f(x)[M][Cin] . w[Cout][Cin]^T (+ bias);  f(v) = v (in_scale NULL), v*scale+shift,
 * or relu6(v*scale+shift) (in_relu6).  `part`: b200lp_pw_conv_parts(M, Cout) x 2 x Cout floats, or NULL. */
int32_t b200lp_pw_conv_parts(int64_t M, int32_t Cout);
int32_t b200lp_pw_conv(const float* x, const float* in_scale, const float* in_shift, int32_t in_relu6, const float* w,
                       const float* bias, float* y, float* part, int64_t M, int32_t Cin, int32_t Cout, void* stream);
/* The same product with split-K for layers whose (row tile x column tile) grid leaves most SMs idle (small planes of a
 * small batch, the classifiers): partial sums in `workspace` (b200lp_pw_conv_workspace bytes; 0 = the layer does not split),
 * then one reduction pass that also adds the bias and emits the statistics partials.  Deterministic (fixed split order). */
int64_t b200lp_pw_conv_workspace(int64_t M, int32_t Cin, int32_t Cout);
int32_t b200lp_pw_conv_ws(const float* x, const float* in_scale, const float* in_shift, int32_t in_relu6, const float* w,
                          const float* bias, float* y, float* part, int64_t M, int32_t Cin, int32_t Cout,
                          float* workspace, int64_t workspace_bytes, void* stream);
/* depthwise 3x3, padding 1, stride 1 or 2, on relu6(x*scale+shift); w [C][1][3][3]; y [N,Ho,Wo,C] raw;
 * `part`: b200lp_dw_conv3x3_parts(N,H,W,stride) x 2 x C floats, or NULL. */
int32_t b200lp_dw_conv3x3_parts(int32_t N, int32_t H, int32_t W, int32_t stride);
int32_t b200lp_dw_conv3x3(const float* x, const float* in_scale, const float* in_shift, const float* w, float* y,
                          float* part, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride, void* stream);
/* stem: 3x3 stride-2 conv 3 -> 32 on the NCHW image; y [N,H/2,W/2,32] raw; part: b200lp_mbv2_stem_parts x 2 x 32 */
int32_t b200lp_mbv2_stem_parts(int32_t N, int32_t H, int32_t W);
int32_t b200lp_mbv2_stem(const float* x_nchw, const float* w, float* y_nhwc, float* part, int32_t N, int32_t H,
                         int32_t W, void* stream);
/* training != 0: batch mean / biased variance over `count` samples from the partials (fp64 merge, fixed order) ->
 * scale = gamma*rstd, shift = beta - mean*scale; running_mean / running_var (unbiased) / num_batches_tracked updated
 * like nn.BatchNorm2d (any of the three may be NULL).  training == 0: scale / shift from the running statistics.
 * mean_out / rstd_out (both or neither, may be NULL): the statistics the layer normalised with (for b200lp_bn_bwd). */
int32_t b200lp_bn_finalize(const float* part, int32_t nparts, int64_t count, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                           float eps, float* scale, float* shift, float* mean_out, float* rstd_out, int32_t C,
                           int32_t training, void* stream);
/* y = x*scale[c] + shift[c] (+ residual) (relu6): materialises an inverted-residual block output */
int32_t b200lp_bn_apply(const float* x, const float* scale, const float* shift, const float* residual, float* y,
                        int64_t M, int32_t C, int32_t relu6, void* stream);
/* y[n][c] = mean_p relu6(x[n,p,c]*scale[c] + shift[c])   (features.18 BN + ReLU6 + adaptive_avg_pool2d(1)) */
int32_t b200lp_bn_relu6_avgpool(const float* x, const float* scale, const float* shift, float* y, int32_t N, int32_t HW,
                                int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Identity encoder (torchvision ResNeXt50-32x4d, train-mode BatchNorm) forward + backward, and the shared BatchNorm
 * backward of both encoders — csrc/encoder.cu.  Replaces Embedder.get_identity_embedding
 * (embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:26-27,37-54) and its autograd backward: cuDNN
 * BatchNorm forward / backward, grouped 3x3 convolutions (groups = 32), max-pool, ReLU / residual adds.  The 1x1
 * convolutions and the 7x7 stem (through b200lp_im2col7x7_s2) run on b200lp_conv_fwd / b200lp_conv_wgrad.
 * Activations NHWC fp32, a layer is an [M][C] matrix, C % 4 == 0.
 */
/* per-channel sum / sum-of-squares partials of x [M][C] -> part [b200lp_col_stats_parts(M)][2][C] (for bn_finalize) */
int32_t b200lp_col_stats_parts(int64_t M);
int32_t b200lp_col_stats(const float* x, float* part, int64_t M, int32_t C, void* stream);
/* y = act(x*scale + shift (+ res [*res_scale + res_shift]));  act 0 none / 1 relu / 2 relu6.  scale/shift may be NULL
 * (identity).  Outputs (either or both): y fp32 (tf32-rounded when round_tf32) and y_split = (hi, lo) bf16 planes
 * [2][M][C] of the unrounded value (operand of the bf16x3 tensor-core GEMM).  mask_out (optional): one byte per 4
 * elements, bit k = element k's activation passed (> 0) — what b200lp_bn_bwd's mask_mode 4 reads instead of y. */
int32_t b200lp_bn_act(const float* x, const float* scale, const float* shift, const float* res, const float* res_scale,
                      const float* res_shift, float* y, void* y_split, int64_t M, int32_t C, int32_t act,
                      int32_t round_tf32, uint8_t* mask_out, void* stream);
/* BatchNorm (+ activation) backward over [M][C].  dz = dy * mask, mask_mode 0: none; 1: mask_src > 0 (a materialised
 * activation output, fp32 [M][C]); 2: relu(x_raw*scale+shift) > 0; 3: 0 < x_raw*scale+shift < 6 (ReLU6); 4: mask_src is the
 * uint8 [M*C/4] bit mask written by b200lp_bn_act (0.25 instead of 4 bytes per element in both passes).
 *   dgamma (+)= sum dz*xhat, dbeta (+)= sum dz   (accumulate != 0: added to the buffers; either may be NULL)
 *   dx = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat))  [batch_stats != 0]   or   gamma*rstd*dz  [running statistics]
 *   dz_out (optional) = dz.   xhat = (x_raw - mean)*rstd.   dx tf32-rounded when round_tf32 (it feeds a TF32 GEMM).
 * Three launches (partial sums, fp64 fixed-order merge, apply); workspace >= b200lp_bn_bwd_workspace(M, C) bytes. */
int64_t b200lp_bn_bwd_workspace(int64_t M, int32_t C);
int32_t b200lp_bn_bwd(const float* dy, const void* mask_src, const float* x_raw, const float* mean, const float* rstd,
                      const float* scale, const float* shift, const float* gamma, float* dgamma, float* dbeta,
                      int32_t accumulate, float* dx, float* dz_out, float* workspace, int64_t workspace_bytes, int64_t M,
                      int32_t C, int32_t mask_mode, int32_t batch_stats, int32_t round_tf32, void* stream);
/* grouped 3x3 convolution, padding 1, groups = C / cpg (cpg in {4,8,16,32}, C a multiple of 8*cpg), w [C][cpg][3][3],
 * FP32 on the CUDA cores.  Input = relu(x*in_scale+in_shift) (producer BatchNorm + ReLU on load; NULL: x itself).
 * transposed != 0 (stride 1 only): y = data gradient of the stride-1 convolution for output gradient x.
 * part (optional): [b200lp_gconv3x3_parts(...)][2][C] statistics partials of y. */
int32_t b200lp_gconv3x3_parts(int32_t N, int32_t H, int32_t W, int32_t C, int32_t cpg, int32_t stride);
int32_t b200lp_gconv3x3_fwd(const float* x, const float* in_scale, const float* in_shift, const float* w, float* y,
                            float* part, int32_t N, int32_t H, int32_t W, int32_t C, int32_t cpg, int32_t stride,
                            int32_t transposed, void* stream);
/* data gradient: dy [N,Ho,Wo,C] -> dx [N,H,W,C] (stride 1 or 2) */
int32_t b200lp_gconv3x3_dgrad(const float* dy, const float* w, float* dx, int32_t N, int32_t H, int32_t W, int32_t C,
                              int32_t cpg, int32_t stride, void* stream);
/* weight gradient: dw [C][cpg][3][3] (+)= sum_pixels dy (x) relu(x*in_scale+in_shift); deterministic two-stage sum */
int64_t b200lp_gconv3x3_wgrad_workspace(int32_t N, int32_t H, int32_t W, int32_t C, int32_t cpg, int32_t stride);
int32_t b200lp_gconv3x3_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dw,
                              int32_t accumulate, float* workspace, int64_t workspace_bytes, int32_t N, int32_t H,
                              int32_t W, int32_t C, int32_t cpg, int32_t stride, void* stream);
/* 7x7 stride-2 padding-3 patch matrix of an NCHW image: col [N*Ho*Wo][KP], column c*49+kh*7+kw (147 real, zero padded to
 * KP); col fp32 tf32-rounded and / or col_split (hi, lo) bf16 planes */
int32_t b200lp_im2col7x7_s2(const float* x_nchw, float* col, void* col_split, int32_t N, int32_t H, int32_t W, int32_t KP,
                            void* stream);
/* y = maxpool3x3/s2/p1(relu(x*scale+shift)); idx [N,Ho,Wo,C] uint8 = tap of the maximum (for the backward gather) */
int32_t b200lp_maxpool3x3s2_fwd(const float* x, const float* scale, const float* shift, float* y, void* y_split,
                                uint8_t* idx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t round_tf32, void* stream);
int32_t b200lp_maxpool3x3s2_bwd(const float* dy, const uint8_t* idx, float* dx, int32_t N, int32_t H, int32_t W, int32_t C,
                                void* stream);
/* y[n,h,w,:] = x[n,2h,2w,:] (fp32 and / or (hi, lo) planes): operand of a stride-2 1x1 convolution; and its adjoint
 * dx[n,2h,2w,:] += dsub[n,h,w,:] */
int32_t b200lp_subsample2(const float* x, const void* x_split, float* y, void* y_split, int32_t N, int32_t Ho, int32_t Wo,
                          int32_t C, void* stream);
int32_t b200lp_scatter_add2(const float* dsub, float* dx, int32_t N, int32_t Ho, int32_t Wo, int32_t C, void* stream);
/* global average pool over the HW pixels of [N][HW][C] and its backward */
int32_t b200lp_avgpool_fwd(const float* x, float* y, int32_t N, int32_t HW, int32_t C, void* stream);
int32_t b200lp_avgpool_bwd(const float* dy, float* dx, int32_t N, int32_t HW, int32_t C, void* stream);
/* C[M][N] (+)= alpha * sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj] (+ bias[j]); alpha = *alpha_dev (device scalar) or 1
 * (fp32; the classifier / projector layers (generators/...noBottleneck.py:97-101) and their gradients) */
int64_t b200lp_sgemm_strided_workspace(int32_t M, int32_t N, int32_t K);    /* split-K partial sums (0: none needed) */
int32_t b200lp_sgemm_strided(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj, float* C,
                             const float* alpha_dev, const float* bias, int32_t M, int32_t N, int32_t K,
                             int32_t accumulate, float* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Scalar losses and the small dense pieces around them — csrc/losses.cu (two-stage fixed-order reductions, gather-form
 * backward kernels: bit-reproducible).
 */
/* dice (criterions/dice.py:30-34): fake (B, HW), real (B, CR, HW) -> sums[3] = (sum f*r, sum f^2, sum r^2) with f
 * broadcast over real's CR channels, loss[0] = -log(2 sums0 / (sums1 + sums2)) * weight; backward: d_fake (B, HW) */
int64_t b200lp_dice_workspace(int32_t B, int32_t HW);
int32_t b200lp_dice_fwd(const float* fake, const float* real, float weight, float* sums, float* loss, float* workspace,
                        int64_t workspace_bytes, int32_t B, int32_t CR, int32_t HW, void* stream);
int32_t b200lp_dice_bwd(const float* fake, const float* real, const float* sums, const float* grad, float weight,
                        float* d_fake, int32_t B, int32_t CR, int32_t HW, void* stream);
/* adversarial losses (criterions/adversarial.py:20-47) on (B,) score vectors: out2 = (loss_G, loss_D);
 * relativistic 0 = gan (hinge D, -mean G), 1 = rgan, 2 = ragan.  Backward for gan: any output / gradient pointer may be NULL */
int32_t b200lp_adversarial_fwd(const float* fake_g, const float* fake_d, const float* real, float* out2, int32_t B,
                               int32_t relativistic, void* stream);
int32_t b200lp_adversarial_bwd(const float* fake_d, const float* real, const float* grad_g, const float* grad_d,
                               float* d_fake_g, float* d_fake_d, float* d_real, int32_t B, void* stream);
/* box crop + resize (criterions/idt_embed.py:62-83): boxes (B, 4) = [t, b, l, r] in pixels (device), x (B, C, H, W) ->
 * y (B, C, OH, OW) with torch's affine_grid(align_corners=False) + grid_sample(bilinear, reflection) semantics.
 * Backward (gather form) requires boxes whose sampling positions stay inside the image (no reflection active). */
int32_t b200lp_crop_bilinear_fwd(const float* x, const float* boxes, float* y, int32_t B, int32_t C, int32_t H, int32_t W,
                                 int32_t OH, int32_t OW, void* stream);
int32_t b200lp_crop_bilinear_bwd(const float* dy, const float* boxes, float* dx, int32_t B, int32_t C, int32_t H, int32_t W,
                                 int32_t OH, int32_t OW, void* stream);
/* discriminator head (discriminators/no_landmarks.py:101-105): feat (B, P, C) NHWC raw -> o (B, C) = sum_p relu(feat),
 * score (B,) = inv_sigma * <o, w> + bias + <o, embed> (embed (B, C) or NULL).  Backward: d_feat, d_embed, and the linear
 * layer's dw (C) (+)= inv_sigma * sum_b g o, ds = sum_b g <o, w> (gradient w.r.t. inv_sigma), dbias (+)= sum_b g. */
int32_t b200lp_disc_head_fwd(const float* feat, const float* embed, const float* w, const float* inv_sigma,
                             const float* bias, float* o, float* score, int32_t B, int32_t P, int32_t C, void* stream);
int32_t b200lp_disc_head_bwd(const float* feat, const float* embed, const float* w, const float* inv_sigma, const float* o,
                             const float* grad, float* d_feat, float* d_embed, float* dw, float* ds, float* dbias,
                             int32_t accumulate, int32_t B, int32_t P, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Pose encoder (MobileNetV2) BACKWARD — csrc/mobilenet_bwd.cu (FP32 CUDA cores; BatchNorm backward = b200lp_bn_bwd with
 * the ReLU6 mask mode).  Replaces autograd's backward of Embedder.get_pose_embedding
 * (embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:56-58) in meta-training.
 */
/* dst [C][R] = src [R][C]^T (the 1x1 data gradient runs as b200lp_pw_conv on the transposed weight) */
int32_t b200lp_transpose2d(const float* src, float* dst, int32_t R, int32_t C, void* stream);
/* 1x1 conv / linear weight gradient: dw [Cout][Cin] (+)= sum_m dy[m][co] * f(x[m][ci]), f as in b200lp_pw_conv */
int64_t b200lp_pw_wgrad_workspace(int64_t M, int32_t Cin, int32_t Cout);
int32_t b200lp_pw_wgrad(const float* dy, const float* x, const float* in_scale, const float* in_shift, int32_t in_relu6,
                        float* dw, int32_t accumulate, float* workspace, int64_t workspace_bytes, int64_t M, int32_t Cin,
                        int32_t Cout, void* stream);
/* depthwise 3x3 (padding 1, stride 1 | 2) backward: data gradient dy [N,Ho,Wo,C] -> dx [N,H,W,C]; weight gradient
 * dw [C][9] (+)= sum_p dy[p][c] * relu6(x[p (+) tap][c]*in_scale[c] + in_shift[c]) */
int32_t b200lp_dw_dgrad(const float* dy, const float* w, float* dx, int32_t N, int32_t H, int32_t W, int32_t C,
                        int32_t stride, void* stream);
int64_t b200lp_dw_wgrad_workspace(int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride);
int32_t b200lp_dw_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dw,
                        int32_t accumulate, float* workspace, int64_t workspace_bytes, int32_t N, int32_t H, int32_t W,
                        int32_t C, int32_t stride, void* stream);
/* stem (3x3 stride-2 conv on the NCHW image) weight gradient: dw [32][27] (+)= sum_p dy[p][co] * patch(x)[p][27] */
int64_t b200lp_mbv2_stem_wgrad_workspace(int32_t N, int32_t H, int32_t W);
int32_t b200lp_mbv2_stem_wgrad(const float* x_nchw, const float* dy, float* dw, int32_t accumulate, float* workspace,
                               int64_t workspace_bytes, int32_t N, int32_t H, int32_t W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LP_H_ */
