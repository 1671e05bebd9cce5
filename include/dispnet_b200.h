/*
 * dispnet_b200.h -- C ABI of libdispnet_b200.so (sm_100a).
 *
 * Drop-in boundary for the data-parallel training hot path of zenithfang/supervised_dispnet:
 * the library kernels that the reference reaches through PyTorch (ATen/cuDNN) from
 *   models/Disp_vgg_BN.py:136-191, models/DispNetS.py:93-140, models/PoseExpNet.py:58-95,
 *   models/Disp_res_50.py:139-198            (conv / convT / BN / pool / act / cat / upsample)
 *   loss_functions.py:104-129, :317-386, :401-448 and inverse_warp.py:160-193  (per-pixel losses)
 * are replaced one-for-one by the entry points below.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function is stream-ordered on `stream` (a cudaStream_t passed as void*), allocates
 *     nothing, keeps no global state (except cached function attributes) and is re-entrant;
 *   - return value: 0 = ok, >0 = cudaError_t of the launch, <0 = DN_E_* argument error;
 *   - activations live in HBM as NHWC "views": channel stride 1, arbitrary element strides for
 *     N/H/W so that concat slices, 2x2 phase sub-lattices and crops are views, not copies;
 *   - dtype codes: DN_F32 = 0, DN_F16 = 1, DN_BF16 = 2.
 */
#ifndef DISPNET_B200_H
#define DISPNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DN_F32 0
#define DN_F16 1
#define DN_BF16 2
/* pack-only code: the bf16 residual  bf16(v - float(bf16(v)))  of a split-precision operand (precision 'tc32') */
#define DN_BF16_LO 3

#define DN_ACT_NONE 0
#define DN_ACT_RELU 1
#define DN_ACT_LRELU 2 /* LeakyReLU(0.1): models/Disp_vgg_BN.py:40-45 */

#define DN_E_ARG (-1)
#define DN_E_UNSUPPORTED (-2)

/* 7x7 taps x 3 split-precision terms; 4 stride-2 phases x (hi, lo) planes */
#define DN_MAX_TAPS 160
#define DN_MAX_SRC 8

/* NHWC view of an activation tensor (strides in ELEMENTS of `dtype`; channel stride is 1).
 * c_ext (0 = C): channels readable from `ptr` inside one pixel record, c_ext >= C.  Channels [C, c_ext) hold FINITE values
 * (zero padding of the buffer or the neighbouring slices of a concatenation buffer) that a gather-convolution may fetch as
 * padding -- they meet zero rows of the packed weights.  Lets the TMA maps of 17- / 97- / 193-channel tensors declare whole
 * 32- / 64-channel rows instead of partially out-of-bounds ones (which the TMA unit fills at a fraction of its row rate). */
typedef struct dn_view {
  void* ptr;
  int32_t dtype;
  int32_t N, H, W, C;
  int32_t c_ext;
  int64_t sN, sH, sW;
} dn_view;

/* One filter tap of a gather-convolution: reads source view `src` at (ho*stride+dh, wo*stride+dw)
 * and multiplies by packed weight matrix number `wt`. */
typedef struct dn_tap {
  int32_t src, dh, dw, wt;
} dn_tap;

/*
 * Gather-convolution ("implicit GEMM") problem:
 *   out[n,ho,wo,co] = act( bias[co] + sum_t sum_ci in[t.src][n, ho*stride+t.dh, wo*stride+t.dw, ci]
 *                                                * w[t.wt][co][ci] )  (+ out if accumulate)
 * with zero contribution outside a source view.  nn.Conv2d forward, each 2x2 output phase of
 * nn.ConvTranspose2d forward, and the data-gradients of both are instances.
 * `w` is packed [ntaps_w][Cout_pad][Cin_pad] (Cin contiguous) in dtype `w_dtype` by dn_pack_weight.
 */
typedef struct dn_igemm {
  dn_view in[DN_MAX_SRC];
  int32_t nsrc;
  dn_view out;
  const void* w;
  int32_t w_dtype;
  int32_t cin_pad, cout_pad; /* packed weight matrix dims (multiples of 8) */
  const float* bias;         /* [Cout] fp32 or NULL */
  int32_t act;               /* DN_ACT_* applied after bias */
  int32_t accumulate;        /* 1: out += result (act must be NONE) */
  int32_t stride;            /* input stride of the gather (1 or 2) */
  int32_t ntaps;
  dn_tap taps[DN_MAX_TAPS];
  float out_scale;           /* result multiplied by this before bias/act (1.0 normally) */
  int32_t out_pad_ok;        /* 1: channels [out.C, roundup(out.C, 8)) of every output pixel are padding that the
                                kernel may overwrite (lets the 16-byte vector epilogue serve odd channel counts) */
  void* out2;                /* optional second copy of the result (same N/H/W/C and strides as `out`) ... */
  int32_t out2_dtype;        /* ... in this 16-bit dtype: the bf16 image a later weight-gradient GEMM consumes */
  int32_t nphase;            /* 0 / 1: one output view.  n > 1 (the output phases of a stride-2 nn.ConvTranspose2d,
                                models/Disp_vgg_BN.py:92-104, as ONE launch): n sub-problems that share `in`, the weights and the
                                geometry of `out`; sub-problem i uses taps [i * ntaps / n, (i + 1) * ntaps / n) and writes at
                                out.ptr (and out2) + phase_off[i] elements */
  int64_t phase_off[4];
  int32_t phase_cout;        /* > 0 (with nphase > 1): the phases are stacked along the OUTPUT CHANNELS instead: `taps` is the union of the
                                phases' taps over one input neighbourhood, the packed weight matrices have nphase * phase_cout rows
                                (rows [i * phase_cout, (i + 1) * phase_cout) = phase i, zero where a phase does not use the tap),
                                cout_pad = nphase * phase_cout, and column block i of every output pixel goes to out + phase_off[i].
                                One fetch of the input tile serves all phases (thin transposed convolutions). */
  int32_t pad2_;
} dn_igemm;

/*
 * Weight-gradient problem:
 *   dw[t.wt][cp][cq] (+)= scale * sum_{n,h,w} p[t.src][n,h,w,cp] * q[n, h*stride+t.dh, w*stride+t.dw, cq]
 * p = output-gradient view(s) (one per phase), q = saved input activation.  dw is fp32
 * [ntaps_w][cp_pad][cq_pad] and must be zeroed by the caller (split-K partial sums are added atomically).
 */
typedef struct dn_wgrad {
  dn_view p[DN_MAX_SRC];
  int32_t nsrc;
  dn_view q;
  float* dw;
  int32_t cp_pad, cq_pad;
  int32_t stride;
  int32_t ntaps;
  dn_tap taps[DN_MAX_TAPS];
  float scale;
} dn_wgrad;

/* ---- library ------------------------------------------------------------------------------ */
int dn_version(void);
const char* dn_error_string(int code);
/* 1 when the running device is sm_100 and the tcgen05 path can be used. */
int dn_tc_available(void);
/* Profiling aid: 8 x uint64 device counters that igemm launches add per-role cycle counts to (NULL switches it off). */
int dn_tc_set_debug(void* device_counters);
/* 1 (default): 3x3 stride-1 problems with many pixels use the shared-memory halo variant of the tcgen05 kernel. */
int dn_tc_set_halo(int enabled);

/* ---- layout / packing (replaces ATen copies: torch.cat, .contiguous(), weight re-layout) ---- */
/* NCHW fp32 [N,C,H,W] -> channels [c0, c0+C) of NHWC view `dst` (channels >= c0+C left untouched). */
int dn_pack_input(const float* src, int N, int C, int H, int W, const dn_view* dst, int c0, void* stream);
/* Device input pipeline (custom_transforms.py:25-70; train.py:137-142 composes RandomHorizontalFlip, ArrayToTensor, Normalize):
 * src uint8 [B,H,W,C] frames as the loader reads them -> dst fp32 [B,C,H,W] = ((src/255) - mean[c]) / std[c], sample b mirrored
 * left-right when flip[b] != 0 (flip may be NULL; mean / std are HOST arrays of C <= 4 floats).  Bit-identical to the reference's
 * per-sample numpy/torch chain; the host sends 1 byte per value instead of 4. */
int dn_input_transform(const uint8_t* src, int B, int H, int W, int C, const int32_t* flip, const float* mean,
                       const float* std, float* dst, void* stream);
/* the same mirror for the ground-truth depth (custom_transforms.py:64): src, dst fp32 [B, rows, W] */
int dn_flip_rows(const float* src, int B, int64_t rows, int W, const int32_t* flip, float* dst, void* stream);
/* dst[t][r][c] = src[r*s_r + c*s_c + kh[t]*s_kh + kw[t]*s_kw] for r<R, c<Cc, else 0.
 * dst is [T][R_pad][C_pad] of dtype `dst_dtype`; src is the fp32 torch parameter. */
int dn_pack_weight(const float* src, void* dst, int dst_dtype, int T, int R, int Cc, int R_pad, int C_pad,
                   const int32_t* kh, const int32_t* kw, int64_t s_r, int64_t s_c, int64_t s_kh, int64_t s_kw,
                   void* stream);
/* inverse of dn_pack_weight for fp32 gradients: dst[r*s_r + c*s_c + kh[t]*s_kh + kw[t]*s_kw] = scale*src[t][r][c] */
int dn_unpack_wgrad(const float* src, float* dst, int T, int R, int Cc, int R_pad, int C_pad,
                    const int32_t* kh, const int32_t* kw, int64_t s_r, int64_t s_c, int64_t s_kh, int64_t s_kw,
                    float scale, void* stream);

/* Batched form of the two calls above: one launch serves every layer of a network (the per-layer calls are a few
 * microseconds of fixed cost each, 80+ of them per step).  `jobs` is a DEVICE array. */
typedef struct dn_pack_job {
  const void* src;      /* pack: fp32 torch parameter;        unpack: packed fp32 gradient [T][R_pad][C_pad] */
  void* dst;            /* pack: packed [T][R_pad][C_pad];    unpack: fp32 torch-layout gradient */
  int32_t dst_dtype;    /* pack only */
  int32_t unpack;       /* 0 = pack, 1 = unpack */
  int32_t T, R, Cc, R_pad, C_pad, k;   /* tap t = (kh, kw) = (t / k, t % k) */
  int64_t s_r, s_c, s_kh, s_kw;
  float scale;          /* unpack only */
  int32_t pad_;
  const float* row_scale; /* pack only, may be NULL: row r is multiplied by row_scale[r] -- the per-output-channel
                             gamma/sqrt(var+eps) of an eval-mode BatchNorm folded into the convolution (validate_with_gt,
                             train.py:642-723: BN uses running statistics there, so conv+BN is one affine map) */
} dn_pack_job;
int dn_pack_jobs(const dn_pack_job* jobs, int njobs, void* stream);

/* ---- first-layer convolutions (input = the image: 3 or 3*(1+R) channels; no data gradient) -------------------------
 * The k x k convolution (models/DispNetS.py:17-19 conv1 7x7/2, models/Disp_res_50.py:46 stem, vgg16_bn features.0 3x3) is
 * run as k taps over k*C channels of a row-expanded image
 *   out[n][h][wo][kw*C + c] = x[n][h][stride*wo + kw - pad][c]   (zero outside), out: [N, H, Wo, k*C], out2: optional second
 * copy in another dtype (the weight-gradient operand).  Weights: dst[kh][Cout_pad][Cx_pad], column kw*Cin + c. */
int dn_rowx_expand(const dn_view* x, int k, int stride, int pad, const dn_view* out, const dn_view* out2, void* stream);
int dn_rowx_pack_weight(const float* w /* [Cout][Cin][k][k] */, int Cout, int Cin, int k, void* dst, int dst_dtype, int cout_pad,
                        int cx_pad, const float* row_scale /* [Cout] or NULL, see dn_pack_job */, void* stream);
int dn_rowx_unpack_wgrad(const float* dwp /* [k][cout_pad][cx_pad] */, float* grad /* [Cout][Cin][k][k] */, int Cout, int Cin, int k,
                         int cout_pad, int cx_pad, float scale, void* stream);

/* ---- convolutions (nn.Conv2d / nn.ConvTranspose2d fwd, dgrad, wgrad) ------------------------- */
/* backend: 0 = CUDA-core tiled kernel (any shape/dtype), 1 = tcgen05/TMA kernel (fp16/bf16, stride 1). */
int dn_igemm_run(const dn_igemm* p, int backend, void* stream);
int dn_wgrad_run(const dn_wgrad* p, int backend, void* stream);
/* returns 1 if the tcgen05 backend accepts this problem */
int dn_igemm_tc_supported(const dn_igemm* p);
int dn_wgrad_tc_supported(const dn_wgrad* p);

/* ---- BatchNorm2d (training) + ReLU + MaxPool2d(2,2)  (models/Disp_vgg_BN.py:137-141) ---------- */
/* Per-channel reductions write per-block partial rows into `ws`; the last block of each channel slab to arrive folds the
 * rows in a fixed order (deterministic, double accumulation) inside the same launch - no contended atomics, no second
 * kernel.  `ws` must hold dn_reduce_ws_floats(C) floats and be ZERO before its first use (it starts with arrival
 * counters that every launch leaves at zero again); one workspace serves any number of launches on one stream. */
int64_t dn_reduce_ws_floats(int C);
/* sums[2*C] (double) = per-channel sum and sum of squares of y (overwritten). */
int dn_bn_stats(const dn_view* y, double* sums, float* ws, void* stream);
/* mean/invstd from sums; running stats update (momentum, unbiased var); scale_shift[2C] = (g*invstd, b-mean*g*invstd).
 * count = N*H*W.  If training == 0 uses running stats instead. */
int dn_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, int training, int update_running,
                   float* mean_invstd, float* scale_shift, int C, void* stream);
/* Eval-mode conv+BN folding: bias_out[c] = (conv_bias ? conv_bias[c] : 0) * scale_shift[c] + scale_shift[C + c], with
 * scale_shift from dn_bn_finalize(training = 0); the scale half goes into the packed weights through dn_pack_job.row_scale. */
int dn_bn_fold_bias(const float* conv_bias, const float* scale_shift, int C, float* bias_out, void* stream);
/* Training-mode statistics in ONE launch: dn_bn_stats + dn_bn_finalize(training = 1) + `num_batches_tracked += 1`
 * (nn.BatchNorm2d.forward, torch/nn/modules/batchnorm.py; reference call sites models/Disp_vgg_BN.py:137-141).
 * sums (double[2C]) and num_batches_tracked (int64[1]) may be NULL. */
int dn_bn_train_stats(const dn_view* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, float momentum, float eps, int update_running, double* sums,
                      float* mean_invstd, float* scale_shift, float* ws, void* stream);
/* out = pool?( act( y*scale+shift (+ residual) ) );  pool in {0,1}: 1 = 2x2/2 max pool (out is H/2 x W/2).
 * out2 (optional, same shape, may have another 16-bit dtype) receives a second copy of the result: the bf16 image of the
 * activation that the weight-gradient GEMM consumes (tcgen05 kind::f16 cannot mix fp16 and bf16 operands). */
int dn_bn_apply(const dn_view* y, const float* scale_shift, const dn_view* residual, int act, int pool,
                const dn_view* out, const dn_view* out2, void* stream);
/* backward: pass 1 writes red[2C] (double): sum(g), sum(g*y), where g is the gradient routed back through pool/act
 * (raw sums: pass 2 derives sum(g*xhat) = invstd*(sum(g*y) - mean*sum(g)) from them); pass 2 writes dy (and dres). */
int dn_bn_bwd_reduce(const dn_view* dout, const dn_view* y, const dn_view* residual, const float* mean_invstd,
                     const float* gamma, const float* beta, int act, int pool, double* red /* overwritten */, float* ws,
                     void* stream);
int dn_bn_bwd_apply(const dn_view* dout, const dn_view* y, const dn_view* residual, const float* mean_invstd,
                    const float* gamma, const float* beta, int act, int pool, const double* red, double count,
                    float gscale, float* dgamma, float* dbeta, const dn_view* dy, const dn_view* dres,
                    int dres_accumulate, void* stream);

/* ---- pointwise / pooling ------------------------------------------------------------------ */
/* dy = dout * act'(out) in place on `dout`; dbias[c] = gscale * sum(dy) (fp32, overwritten) if dbias != NULL */
int dn_act_bwd(const dn_view* dout, const dn_view* out, int act, float* dbias, float gscale, float* ws, void* stream);
int dn_maxpool_fwd(const dn_view* x, const dn_view* out, int k, int stride, int pad, void* stream);
int dn_maxpool_bwd(const dn_view* dout, const dn_view* x, const dn_view* dx, int k, int stride, int pad,
                   int accumulate, void* stream);
/* out = act(a + b); bwd: d = dout*act'(out) added (accumulate flags) to da and db */
int dn_add_act_fwd(const dn_view* a, const dn_view* b, int act, const dn_view* out, void* stream);
int dn_add_act_bwd(const dn_view* dout, const dn_view* out, int act, const dn_view* da, int da_acc,
                   const dn_view* db, int db_acc, void* stream);
/* out (same shape) = act(x)  and its backward (not in place) */
int dn_act_fwd(const dn_view* x, int act, const dn_view* out, void* stream);
int dn_copy_view(const dn_view* src, const dn_view* dst, int accumulate, void* stream);
/* Split-precision operands (precision 'tc32': fp32 storage, tcgen05 arithmetic): hi = bf16(x), lo = bf16(x - hi), so that
 * x = hi + lo to 2^-18 relative.  A convolution then runs as the three tensor-core terms hi*w_hi + lo*w_hi + hi*w_lo of
 * nn.Conv2d's fp32 product (models/Disp_vgg_BN.py:136-191 runs fp32 end to end).  x: fp32 view; hi, lo: bf16 views of the
 * same geometry. */
int dn_split_bf16(const dn_view* x, const dn_view* hi, const dn_view* lo, void* stream);

/* ---- disparity heads (alpha*sigmoid(conv)+beta, models/Disp_vgg_BN.py:168) -------------------- */
/* predict_disp's nn.Conv2d(C, 1, 3, padding=1) (models/Disp_vgg_BN.py:66-70): w is the fp32 torch parameter [1,C,3,3],
 * z a 1-channel view.  bwd: gx (+)= dz * w; gw[C*9] and gb[1] (fp32, torch layout, overwritten) = gscale * sums;
 * ws needs dn_reduce_ws_floats(C) floats. */
int dn_head_conv_fwd(const dn_view* x, const float* w, const float* bias, const dn_view* z, void* stream);
int dn_head_conv_bwd(const dn_view* x, const float* w, const dn_view* dz, const dn_view* gx, int gx_accumulate, float* gw,
                     float* gb, float gscale, float* ws, void* stream);
/* z: 1-channel conv output view; disp: fp32 [N,1,H,W]; optional `up` = 1-channel view of size (upH,upW) that
 * receives the x2-upsampled disparity (mode 0 nearest, 1 bilinear align_corners=False), cropped to the view. */
int dn_head_fwd(const dn_view* z, float alpha, float beta, float* disp, const dn_view* up, int up_mode,
                void* stream);
/* the same with a second up-sampled copy `up2` (same geometry, another 16-bit dtype: the weight-gradient operand of the next
 * iconv) written in the same pass when up_mode == 0.  up_mode bit 4 (value 16): `up` / `up2` are the LAST channel slice of their
 * buffer and what follows them inside the pixel record is zero padding (dn_view.c_ext >= 16) -- the slot is then written as whole
 * 32-byte sectors (value + zeros) */
int dn_head_fwd2(const dn_view* z, float alpha, float beta, float* disp, const dn_view* up, const dn_view* up2, int up_mode,
                 void* stream);
/* dz = (gscale*gdisp + upsample^T(dup)) * alpha*s*(1-s), s = sigmoid(z) recomputed from the saved conv output.
 * gdisp (fp32, from autograd, may be NULL) is unscaled; dup already carries the gradient scale. */
int dn_head_bwd(const float* gdisp, const dn_view* dup, int up_mode, const dn_view* z, float alpha, float gscale,
                const dn_view* dz, void* stream);
/* sigmoid mask heads of PoseExpNet (models/PoseExpNet.py:82-85): mask fp32 NCHW [N,C,H,W] = sigmoid(z) */
int dn_sigmoid_nchw_fwd(const dn_view* z, float* out, void* stream);
int dn_sigmoid_nchw_bwd(const float* gout, const float* out, float gscale, const dn_view* dz, void* stream);
/* pose = scale * mean_{h,w} z  -> fp32 [N,C]  (models/PoseExpNet.py:71-73) */
int dn_spatial_mean_fwd(const dn_view* z, float scale, float* out, void* stream);
int dn_spatial_mean_bwd(const float* gout, float scale, const dn_view* dz, void* stream);

/* ---- per-pixel losses ------------------------------------------------------------------------ */
/* l1_loss (loss_functions.py:104-129): pred/gt fp32 [B,H,W] (pred row-stride = W).  ws: [2B] floats
 * (per-sample sum, count).  loss[0] = sum_b (sum_b/count_b) / B  (NaN when a sample has no valid pixel). */
int dn_l1_fwd(const float* gt, const float* pred, int B, int HW, float max_depth, float* ws, float* loss,
              void* stream);
int dn_l1_bwd(const float* gt, const float* pred, int B, int HW, float max_depth, const float* ws,
              const float* gout, float* gpred, void* stream);
/* The supervised depth losses behind train.py's `--loss` switch (train.py:449-470), one masked-reduce kernel family:
 *   kind 0 l1_loss (:104-129)  1 l2_loss (:77-102)  2 berhu_loss (:131-160)  3 Scale_invariant_loss (:162-187)
 * per sample: valid = (gt > 0) & (gt < max_depth), pred clamped to [1e-3, max_depth], value_b from the sample's masked sums,
 * loss[0] (+)= weight * sum_b value_b / B.  joint = 1: one mask over the whole batch and no division by B -- the
 * Multiscale_{L1,L2,berhu,scale_inv}_loss forms (:217-315; weight = 1/2^scale, max_depth = 80).  up_factor > 1: pred is
 * [B, H/f, W/f] and is up-sampled on the fly (mode 0 nearest, 1 bilinear align_corners=False) as Multiscale_FULL_L1_loss does
 * (:224-241).  Reductions are two-stage and deterministic (no float atomics); berhu runs a second pass with c = 0.2 max|d|.
 * ws: dn_depth_loss_ws_floats(B) floats, written by fwd and read by bwd.  bwd: gpred [B, H/f, W/f] overwritten. */
int64_t dn_depth_loss_ws_floats(int B);
int dn_depth_loss_fwd(const float* gt, const float* pred, int B, int H, int W, int up_factor, int up_mode, float max_depth,
                      int kind, int joint, float weight, int accumulate, float* ws, float* loss, void* stream);
int dn_depth_loss_bwd(const float* gt, const float* pred, int B, int H, int W, int up_factor, int up_mode, float max_depth,
                      int kind, int joint, float weight, const float* ws, const float* gout, float* gpred, void* stream);
/* 2x2 / stride-2 pooling of the ground-truth pyramid (loss_functions.py:185-215): mode 0 = average (= F.interpolate(
 * scale_factor=0.5, 'bilinear', align_corners=False) and F.avg_pool2d), 1 = F.max_pool2d; src [NC,H,W] -> dst [NC,H/2,W/2] */
int dn_pool2(const float* src, int64_t NC, int H, int W, int mode, float* dst, void* stream);
/* smooth_loss (loss_functions.py:367-386) for one scale: loss[0] += weight * (4 abs-means).  p fp32 [B,H,W]. */
int dn_smooth_fwd(const float* p, int B, int H, int W, float weight, float* loss, void* stream);
int dn_smooth_bwd(const float* p, int B, int H, int W, float weight, const float* gout, float* gp, void* stream);
/* compute_errors (loss_functions.py:401-448): counters[B][4] int32 = n_valid, n<1.25, n<1.25^2, n<1.25^3;
 * sums[B][5] double = sum|d|, sum|d|/gt, sum d^2/gt, sum d^2, sum (ln gt - ln p)^2.  Caller zeroes both.
 * crop window [y1,y2)x[x1,x2) (whole image when crop==0).  scale[b] (optional) multiplies pred (median scaling). */
int dn_depth_errors(const float* gt, const float* pred, int B, int H, int W, float max_depth, int crop, int y1,
                    int y2, int x1, int x2, const float* scale, int32_t* counters, double* sums, void* stream);
/* nn.UpsamplingBilinear2d(size=(H, W)) (align_corners=True) of fp32 [N,h,w] -> [N,H,W]: the NYU branch of validate_with_gt
 * (train.py:696-700) brings the prediction to the ground truth's resolution before compute_errors. */
int dn_resize_bilinear_ac(const float* src, int N, int h, int w, int H, int W, float* dst, void* stream);
/* F.interpolate(mode='area') by integer factor f (loss_functions.py:326-327): NCHW fp32 */
int dn_area_down(const float* src, int NC, int H, int W, int f, float* dst, void* stream);
/*
 * Fused inverse warp + photometric term for one (scale, ref) pair (inverse_warp.py:160-193,
 * loss_functions.py:331-342):  tgt/ref fp32 [B,3,h,w]; depth fp32 [B,h,w]; pose [B,6] with batch stride
 * pose_stride; K, Kinv [B,3,3] already scaled for this pyramid level; mask [B,h,w] with batch stride or NULL.
 * rot_mode 0 euler / 1 quat; pad_mode 0 zeros / 1 border; align_corners 0/1.
 * fwd: loss[0] += sum|diff| / (B*3*h*w); nanflag[0] |= 1 if any diff is NaN; optionally writes `warped`.
 * bwd: gdepth [B,h,w] (+=, caller zeroes; summed over reference frames), gpose [B,6] (+=, batch stride pose_stride),
 *      gmask [B,h,w] (overwritten).
 */
int dn_warp_photo_fwd(const float* tgt, const float* ref, const float* depth, const float* pose, int pose_stride,
                      const float* K, const float* Kinv, const float* mask, int64_t mask_bstride, int B, int h,
                      int w, int rot_mode, int pad_mode, int align_corners, float* warped, float* loss,
                      int32_t* nanflag, void* stream);
int dn_warp_photo_bwd(const float* tgt, const float* ref, const float* depth, const float* pose, int pose_stride,
                      const float* K, const float* Kinv, const float* mask, int64_t mask_bstride, int B, int h,
                      int w, int rot_mode, int pad_mode, int align_corners, const float* gout, float* gdepth,
                      float* gpose, float* gmask, int64_t gmask_bstride, float* ws /* [12*B] scratch */, void* stream);
/*
 * The whole photometric_reconstruction_loss (loss_functions.py:317-354) in three launches instead of one launch per
 * (scale, reference frame) pair plus ~20 ATen launches for the pyramids and the scaled intrinsics:
 *   dn_area_pyramid     the /2, /4, /8 area pyramids (F.interpolate(mode='area'), :326-327) of the target and all reference
 *                       frames in ONE pass over the full-size images (each thread owns an 8x8 block: float4 loads / stores);
 *   dn_photo_batch_fwd  every scale and every reference frame in one grid: a thread handles four neighbouring pixels
 *                       (float4 loads of target / depth / mask), loops over the reference frames, scales K / K^-1 for its
 *                       pyramid level itself (:329-330); block partials + a fixed-order fold: the scalar is deterministic;
 *   dn_photo_batch_bwd  the analytic backward likewise; the depth gradient of a pixel is summed over the reference frames in
 *                       registers and written once (no read-modify-write), the pose accumulators go through partial rows.
 */
#define DN_PHOTO_MAX_SCALES 4
#define DN_PHOTO_MAX_REFS 4
typedef struct dn_photo_scale {
  const float* tgt;                      /* [B,3,h,w] target at this pyramid level */
  const float* ref[DN_PHOTO_MAX_REFS];   /* [B,3,h,w] reference frames at this level */
  const float* depth;                    /* [B,h,w] */
  const float* mask;                     /* [B,R,h,w] explainability mask or NULL */
  float* gdepth;                         /* bwd: [B,h,w], overwritten */
  float* gmask;                          /* bwd: [B,R,h,w], overwritten (NULL without mask) */
  int32_t h, w;
  float downscale;                       /* H_full / h (:328) */
  int32_t pad_;
} dn_photo_scale;
typedef struct dn_photo_batch {
  dn_photo_scale sc[DN_PHOTO_MAX_SCALES];
  int32_t nscales, nrefs, B;
  int32_t rot_mode, pad_mode, align_corners;
  const float* K;                        /* [B,3,3] full-resolution intrinsics and inverse */
  const float* Kinv;
  const float* pose;                     /* [B,R,6] */
} dn_photo_batch;
typedef struct dn_pyr_job { const float* src; float* l1; float* l2; float* l3; } dn_pyr_job;  /* [NC,H,W] -> /2, /4, /8 */
int dn_area_pyramid(const dn_pyr_job* jobs /* host array */, int njobs /* <= 8 */, int64_t NC, int H, int W, void* stream);
int64_t dn_photo_ws_floats(const dn_photo_batch* p);
/* loss[0] += sum over scales and reference frames of mean|diff|; nanflag[0] |= 1 on a NaN difference */
int dn_photo_batch_fwd(const dn_photo_batch* p, float* ws, float* loss, int32_t* nanflag, void* stream);
/* gpose [B,R,6] overwritten; gdepth / gmask of every scale overwritten */
int dn_photo_batch_bwd(const dn_photo_batch* p, const float* gout, float* ws, float* gpose, void* stream);
/* plain inverse_warp forward/backward on its own (inverse_warp.py:160-193): out [B,C,h,w]. */
int dn_inverse_warp_fwd(const float* img, const float* depth, const float* pose, const float* K, const float* Kinv,
                        int B, int C, int h, int w, int rot_mode, int pad_mode, int align_corners, float* out,
                        void* stream);
int dn_inverse_warp_bwd(const float* img, const float* depth, const float* pose, const float* K, const float* Kinv,
                        int B, int C, int h, int w, int rot_mode, int pad_mode, int align_corners, const float* gout,
                        float* gimg /* += or NULL */, float* gdepth, float* gpose /* += or NULL */,
                        float* ws /* [12*B] scratch */, void* stream);
/* explainability_loss (loss_functions.py:357-364): loss[0] += -mean(max(log m, -100)) */
int dn_explain_fwd(const float* mask, int64_t n, float* loss, void* stream);
int dn_explain_bwd(const float* mask, int64_t n, const float* gout, float* gmask, void* stream);

/* ---- monodepth2-style optional terms (reference layers.py:199-266; named by the north star, uncalled in the reference) */
/* layers.SSIM.forward(x, y) (:231-245): reflection pad 1, 3x3 average pools, clamp((1-SSIM)/2, 0, 1); x, y, out fp32 [NC,h,w] */
int dn_ssim_fwd(const float* x, const float* y, int NC, int h, int w, float* out, void* stream);
int dn_ssim_bwd(const float* x, const float* y, const float* gout, int NC, int h, int w, float* gx /* or NULL */,
                float* gy /* or NULL */, void* stream);
/* layers.get_smooth_loss(disp, img) (:199-212): loss[0] += mean|dx disp| e^{-mean_c|dx img|} + same in y; disp [B,1,h,w] */
int dn_edge_smooth_fwd(const float* disp, const float* img, int B, int C, int h, int w, float* loss, void* stream);
int dn_edge_smooth_bwd(const float* disp, const float* img, int B, int C, int h, int w, const float* gout, float* gdisp,
                       void* stream);
/* layers.compute_depth_errors(gt, pred) (:248-266) on pre-masked 1-D tensors: counters[3] (n<1.25^k), sums[4] double =
 * sum|d|/gt, sum d^2/gt, sum d^2, sum (ln gt - ln p)^2; caller zeroes both. */
int dn_depth_errors_raw(const float* gt, const float* pred, int64_t n, int32_t* counters, double* sums, void* stream);

/* ---- misc ---------------------------------------------------------------------------------- */
int dn_fill_f32(float* p, int64_t n, float v, void* stream);
int dn_axpy_f32(const float* x, float a, float* y, int64_t n, void* stream); /* y += a*x */

#ifdef __cplusplus
}
#endif
#endif /* DISPNET_B200_H */
