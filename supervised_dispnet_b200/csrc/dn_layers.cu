// Layer kernels of the DispNet-family encoder/decoder stacks: packing, the CUDA-core gather-convolution
// (backend 0 of dn_igemm_run / dn_wgrad_run: any shape, stride, dtype), BatchNorm(+ReLU+MaxPool2) training
// forward/backward, pooling, pointwise and the disparity heads.  All HBM-bound kernels walk NHWC views with
// 16-byte channel vectors when the view allows it and size their grids in multiples of the SM count.
//
// Reference semantics restated: nn.Conv2d / nn.ConvTranspose2d (cross-correlation,
// zero padding), nn.BatchNorm2d training statistics (biased var to normalise, unbiased to running_var,
// momentum .1, eps 1e-5), MaxPool2d first-max tie rule, LeakyReLU(0.1), alpha*sigmoid+beta heads.
#include "dn_common.cuh"

// =================================================================================================
// packing
// =================================================================================================
__global__ void pack_input_kernel(const float* __restrict__ src, int N, int C, int H, int W, dn_view dst, int c0) {
  dn_pdl_trigger();
  dn_pdl_wait();
  long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    long long r = i / W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    long long o = dn_off(dst, n, h, w) + c0;
    for (int c = 0; c < C; ++c) dn_st(dst.ptr, dst.dtype, o + c, src[(((long long)n * C + c) * H + h) * W + w]);
  }
}

DN_EXPORT int dn_pack_input(const float* src, int N, int C, int H, int W, const dn_view* dst, int c0, void* stream) {
  if (!src || !dst || dst->N != N || dst->H != H || dst->W != W || c0 + C > dst->C) return DN_E_ARG;
  long long total = (long long)N * H * W;
  int blocks = (int)((total + 255) / 256);
  int cap = dn_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  dn_launch(pack_input_kernel, dim3(blocks), dim3(256), 0, dn_stream(stream), src, N, C, H, W, *dst, c0);
  DN_CHECK_LAUNCH();
  return 0;
}

// ---- device input pipeline (custom_transforms.py:25-70): uint8 HWC frames -> normalised fp32 NCHW, optional horizontal flip --
// RandomHorizontalFlip (:56-72) + ArrayToTensor (:41-53: HWC -> CHW, float, /255) + Normalize (:25-38: (x - mean) / std) for a whole
// batch in one pass; the flip decision per sample comes from the host (the reference draws random.random() per sample).
struct Norm4 { float mean[4], std[4]; };
__global__ void __launch_bounds__(256) input_transform_kernel(const uint8_t* __restrict__ src, int B, int H, int W, int C,
                                                              const int32_t* __restrict__ flip, Norm4 nm, float* __restrict__ dst) {
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long q = i / W;
    const int h = (int)(q % H);
    const int b = (int)(q / H);
    const int ws = (flip && flip[b]) ? W - 1 - w : w;
    const uint8_t* p = src + (((long long)b * H + h) * W + ws) * C;
    for (int c = 0; c < C; ++c) {
      const float v = __fdiv_rn((float)p[c], 255.f);
      dst[(((long long)b * C + c) * H + h) * W + w] = __fdiv_rn(__fsub_rn(v, nm.mean[c]), nm.std[c]);
    }
  }
}
__global__ void __launch_bounds__(256) flip_rows_kernel(const float* __restrict__ src, int B, long long rows, int W,
                                                        const int32_t* __restrict__ flip, float* __restrict__ dst) {
  const long long total = (long long)B * rows * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long q = i / W;
    const int b = (int)(q / rows);
    dst[i] = src[q * W + ((flip && flip[b]) ? W - 1 - w : w)];
  }
}
DN_EXPORT int dn_input_transform(const uint8_t* src, int B, int H, int W, int C, const int32_t* flip, const float* mean,
                                 const float* std, float* dst, void* stream) {
  if (!src || !dst || !mean || !std || B < 1 || H < 1 || W < 1 || C < 1 || C > 4) return DN_E_ARG;
  Norm4 nm;
  for (int c = 0; c < 4; ++c) { nm.mean[c] = c < C ? mean[c] : 0.f; nm.std[c] = c < C ? std[c] : 1.f; }
  long long total = (long long)B * H * W;
  int blocks = (int)((total + 255) / 256), cap = dn_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  input_transform_kernel<<<blocks, 256, 0, dn_stream(stream)>>>(src, B, H, W, C, flip, nm, dst);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_flip_rows(const float* src, int B, int64_t rows, int W, const int32_t* flip, float* dst, void* stream) {
  if (!src || !dst || B < 1 || rows < 1 || W < 1) return DN_E_ARG;
  long long total = (long long)B * rows * W;
  int blocks = (int)((total + 255) / 256), cap = dn_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  flip_rows_kernel<<<blocks, 256, 0, dn_stream(stream)>>>(src, B, rows, W, flip, dst);
  DN_CHECK_LAUNCH();
  return 0;
}

struct TapList {
  int32_t kh[DN_MAX_TAPS];
  int32_t kw[DN_MAX_TAPS];
};

__global__ void pack_weight_kernel(const float* __restrict__ src, void* dst, int dt, int T, int R, int Cc, int R_pad,
                                   int C_pad, TapList taps, long long s_r, long long s_c, long long s_kh, long long s_kw) {
  const unsigned total = (unsigned)T * R_pad * C_pad;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned q = i / (unsigned)C_pad;
    const int c = (int)(i - q * (unsigned)C_pad);
    const int t = (int)(q / (unsigned)R_pad);
    const int r = (int)(q - (unsigned)t * (unsigned)R_pad);
    float v = 0.f;
    if (r < R && c < Cc) v = src[r * s_r + c * s_c + taps.kh[t] * s_kh + taps.kw[t] * s_kw];
    dn_st(dst, dt, i, v);
  }
}

DN_EXPORT int dn_pack_weight(const float* src, void* dst, int dst_dtype, int T, int R, int Cc, int R_pad, int C_pad,
                             const int32_t* kh, const int32_t* kw, int64_t s_r, int64_t s_c, int64_t s_kh, int64_t s_kw,
                             void* stream) {
  if (T < 1 || T > DN_MAX_TAPS || R_pad < R || C_pad < Cc) return DN_E_ARG;
  TapList tl;
  for (int i = 0; i < T; ++i) { tl.kh[i] = kh[i]; tl.kw[i] = kw[i]; }
  long long total = (long long)T * R_pad * C_pad;
  int blocks = (int)((total + 255) / 256);
  int cap = dn_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  pack_weight_kernel<<<blocks, 256, 0, dn_stream(stream)>>>(src, dst, dst_dtype, T, R, Cc, R_pad, C_pad, tl, s_r, s_c, s_kh, s_kw);
  DN_CHECK_LAUNCH();
  return 0;
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int R, int Cc, int R_pad,
                                    int C_pad, TapList taps, long long s_r, long long s_c, long long s_kh, long long s_kw,
                                    float scale) {
  const unsigned total = (unsigned)T * R * Cc;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned q = i / (unsigned)Cc;
    const int c = (int)(i - q * (unsigned)Cc);
    const int t = (int)(q / (unsigned)R);
    const int r = (int)(q - (unsigned)t * (unsigned)R);
    dst[r * s_r + c * s_c + taps.kh[t] * s_kh + taps.kw[t] * s_kw] = scale * src[((long long)t * R_pad + r) * C_pad + c];
  }
}

DN_EXPORT int dn_unpack_wgrad(const float* src, float* dst, int T, int R, int Cc, int R_pad, int C_pad, const int32_t* kh,
                              const int32_t* kw, int64_t s_r, int64_t s_c, int64_t s_kh, int64_t s_kw, float scale,
                              void* stream) {
  if (T < 1 || T > DN_MAX_TAPS) return DN_E_ARG;
  TapList tl;
  for (int i = 0; i < T; ++i) { tl.kh[i] = kh[i]; tl.kw[i] = kw[i]; }
  long long total = (long long)T * R * Cc;
  int blocks = (int)((total + 255) / 256);
  int cap = dn_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  unpack_wgrad_kernel<<<blocks, 256, 0, dn_stream(stream)>>>(src, dst, T, R, Cc, R_pad, C_pad, tl, s_r, s_c, s_kh, s_kw, scale);
  DN_CHECK_LAUNCH();
  return 0;
}

// ---- first-layer convolutions: kernel columns folded into the channel dimension -------------------------------------------------
// The input image has 3 (PoseExpNet: 3 * (1 + R)) channels; a k x k convolution over it as a gather-convolution spends one
// 64-channel K chunk per tap on 3 real channels (49 taps for the 7x7 stems of DispNetS / PoseExpNet / Disp_res_50,
// models/DispNetS.py:17-19, models/Disp_res_50.py:46).  Instead the image is expanded once per step to
//   xr[n][h][wo][kw * C + c] = x[n][h][stride * wo + kw - pad][c]          (zero outside the image)
// and the convolution becomes k taps (kernel rows only) over k * C channels, stride 1 in wo.  The weights are packed
// to [kh][Cout_pad][Cx_pad] with column kw * C + c, the weight gradient comes back in that layout.
__global__ void rowx_expand_kernel(dn_view x, int k, int stride, int pad, dn_view out, dn_view out2, int has_out2) {
  dn_pdl_trigger();
  dn_pdl_wait();
  const int C = x.C, Cx = k * C;
  const long long total = (long long)out.N * out.H * out.W * Cx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % Cx);
    long long q = i / Cx;
    const int wo = (int)(q % out.W); q /= out.W;
    const int h = (int)(q % out.H);
    const int n = (int)(q / out.H);
    const int kw = j / C, c = j - kw * C;
    const int w = stride * wo + kw - pad;
    const float v = (w >= 0 && w < x.W) ? dn_ld(x.ptr, x.dtype, dn_off(x, n, h, w) + c) : 0.f;
    const long long o = dn_off(out, n, h, wo) + j;
    dn_st(out.ptr, out.dtype, o, v);
    if (has_out2) dn_st(out2.ptr, out2.dtype, dn_off(out2, n, h, wo) + j, v);
  }
}

// 16-bit outputs whose pixel pitch is a multiple of 8: one thread per (pixel, 8-channel group), 16-byte stores; the channels
// between k * C and the pitch are written as zeros (they are operand padding of the convolution)
__global__ void __launch_bounds__(256) rowx_expand_vec_kernel(dn_view x, int k, int stride, int pad, dn_view out, dn_view out2, int has_out2) {
  dn_pdl_trigger();
  dn_pdl_wait();
  const int C = x.C, Cx = k * C;
  const unsigned groups = (unsigned)out.sW / 8;                 // out.sW == pixel pitch (dense buffer)
  const unsigned total = (unsigned)out.N * out.H * out.W * groups;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / groups, gidx = i - pix * groups;
    const unsigned q = pix / (unsigned)out.W;
    const int wo = (int)(pix - q * (unsigned)out.W);
    const int n = (int)(q / (unsigned)out.H);
    const int h = (int)(q - (unsigned)n * (unsigned)out.H);
    float v[8];
    const long long xrow = dn_off(x, n, h, 0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = (int)gidx * 8 + e;
      v[e] = 0.f;
      if (j < Cx) {
        const int kw = j / C, c = j - kw * C;
        const int w = stride * wo + kw - pad;
        if (w >= 0 && w < x.W) v[e] = dn_ld(x.ptr, x.dtype, xrow + (long long)w * x.sW + c);
      }
    }
    const long long o = (long long)pix * out.sW + gidx * 8;
    if (out.dtype == DN_F16) Vec8<__half>::store((__half*)out.ptr + o, v);
    else Vec8<__nv_bfloat16>::store((__nv_bfloat16*)out.ptr + o, v);
    if (has_out2) {
      if (out2.dtype == DN_F16) Vec8<__half>::store((__half*)out2.ptr + o, v);
      else Vec8<__nv_bfloat16>::store((__nv_bfloat16*)out2.ptr + o, v);
    }
  }
}

DN_EXPORT int dn_rowx_expand(const dn_view* x, int k, int stride, int pad, const dn_view* out, const dn_view* out2, void* stream) {
  if (!x || !out || k < 1 || stride < 1 || out->C != k * x->C || out->N != x->N || out->H != x->H) return DN_E_ARG;
  if (out2 && (out2->C != out->C || out2->N != out->N || out2->H != out->H || out2->W != out->W)) return DN_E_ARG;
  auto dense16 = [](const dn_view* v) {
    return v->dtype != DN_F32 && (v->sW % 8) == 0 && v->sH == (long long)v->W * v->sW && v->sN == (long long)v->H * v->sH &&
           ((uintptr_t)v->ptr % 16) == 0 && v->sW >= v->C;
  };
  const bool vec = dense16(out) && (!out2 || (dense16(out2) && out2->sW == out->sW)) &&
                   (long long)out->N * out->H * out->W * (out->sW / 8) < (1ll << 31);
  const int cap = dn_num_sms() * 16;
  if (vec) {
    const long long total = (long long)out->N * out->H * out->W * (out->sW / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > cap) blocks = cap;
    dn_launch(rowx_expand_vec_kernel, dim3(blocks), dim3(256), 0, dn_stream(stream), *x, k, stride, pad, *out, out2 ? *out2 : *out, out2 != nullptr);
  } else {
    const long long total = (long long)out->N * out->H * out->W * out->C;
    int blocks = (int)((total + 255) / 256);
    if (blocks > cap) blocks = cap;
    dn_launch(rowx_expand_kernel, dim3(blocks), dim3(256), 0, dn_stream(stream), *x, k, stride, pad, *out, out2 ? *out2 : *out, out2 != nullptr);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

// unpack = 0: dst[kh][co][kw * Cin + c] = w[co][c][kh][kw] (16-bit or fp32 dst, padding untouched);  unpack = 1: the reverse, fp32, scaled
__global__ void rowx_weight_kernel(const float* __restrict__ src, void* __restrict__ dst, int dst_dtype, int Cout, int Cin, int k,
                                   int cout_pad, int cx_pad, int unpack, float scale, const float* __restrict__ row_scale) {
  dn_pdl_trigger();
  dn_pdl_wait();
  const int total = Cout * Cin * k * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kw = i % k;
    int q = i / k;
    const int kh = q % k; q /= k;
    const int c = q % Cin;
    const int co = q / Cin;
    const long long pk = ((long long)kh * cout_pad + co) * cx_pad + kw * Cin + c;
    if (!unpack) dn_st(dst, dst_dtype, pk, row_scale ? src[i] * row_scale[co] : src[i]);
    else ((float*)dst)[i] = scale * src[pk];
  }
}

DN_EXPORT int dn_rowx_pack_weight(const float* w, int Cout, int Cin, int k, void* dst, int dst_dtype, int cout_pad, int cx_pad,
                                  const float* row_scale, void* stream) {
  if (!w || !dst || Cout < 1 || Cin < 1 || k < 1 || cout_pad < Cout || cx_pad < k * Cin) return DN_E_ARG;
  const int total = Cout * Cin * k * k;
  dn_launch(rowx_weight_kernel, dim3((total + 255) / 256), dim3(256), 0, dn_stream(stream), w, dst, dst_dtype, Cout, Cin, k, cout_pad, cx_pad, 0, 1.f, row_scale);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_rowx_unpack_wgrad(const float* dwp, float* grad, int Cout, int Cin, int k, int cout_pad, int cx_pad, float scale,
                                   void* stream) {
  if (!dwp || !grad || Cout < 1 || Cin < 1 || k < 1 || cout_pad < Cout || cx_pad < k * Cin) return DN_E_ARG;
  const int total = Cout * Cin * k * k;
  dn_launch(rowx_weight_kernel, dim3((total + 255) / 256), dim3(256), 0, dn_stream(stream), dwp, (void*)grad, (int)DN_F32, Cout, Cin, k, cout_pad, cx_pad, 1, scale, (const float*)nullptr);
  DN_CHECK_LAUNCH();
  return 0;
}

// one thread per (row, column) of the packed matrices; it walks all taps, so the fp32 side is touched in contiguous
// k*k-float runs ([..][kh][kw] is innermost in both nn.Conv2d and nn.ConvTranspose2d weights) and every packed plane is
// written (read) with unit stride across the warp
__global__ void __launch_bounds__(256) pack_jobs_kernel(const dn_pack_job* __restrict__ jobs) {
  dn_pdl_trigger();
  dn_pdl_wait();
  __shared__ float tsm[32][33];
  __shared__ long long toffs[DN_MAX_TAPS];          // torch-side offset of tap t (the divisions by k happen once per block)
  const dn_pack_job j = jobs[blockIdx.y];
  const unsigned k = (unsigned)j.k;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  if (threadIdx.x < (unsigned)j.T) toffs[threadIdx.x] = (long long)(threadIdx.x / k) * j.s_kh + (long long)(threadIdx.x % k) * j.s_kw;
  __syncthreads();
  if (!j.unpack && j.s_r < j.s_c) {
    // Transposing pack (data-gradient layout [tap][Cin][Cout] of a [Cout][Cin][kh][kw] parameter): the fp32 side is
    // contiguous along the packed ROW index, so a 32 x 32 (row, column) tile is read with the lanes along rows, turned in
    // shared memory and written with the lanes along columns - both sides unit-stride across the warp.
    const unsigned plane = (unsigned)j.R_pad * j.C_pad;
    const unsigned tr = ((unsigned)j.R_pad + 31) / 32, tc = ((unsigned)j.C_pad + 31) / 32;
    const float* src = (const float*)j.src;
    for (unsigned tile = blockIdx.x; tile < tr * tc; tile += gridDim.x) {
      const unsigned r0 = (tile / tc) * 32, c0 = (tile - (tile / tc) * tc) * 32;
      for (unsigned t = 0; t < (unsigned)j.T; ++t) {
        const long long toff = toffs[t];
        for (int cc = wrp; cc < 32; cc += 8) {
          const unsigned r = r0 + lane, c = c0 + cc;
          tsm[cc][lane] = ((int)r < j.R && (int)c < j.Cc) ? src[r * j.s_r + c * j.s_c + toff] * (j.row_scale ? j.row_scale[r] : 1.f) : 0.f;
        }
        __syncthreads();
        for (int rr = wrp; rr < 32; rr += 8) {
          const unsigned r = r0 + rr, c = c0 + lane;
          if (r < (unsigned)j.R_pad && c < (unsigned)j.C_pad) dn_st(j.dst, j.dst_dtype, (long long)t * plane + (long long)r * j.C_pad + c, tsm[lane][rr]);
        }
        __syncthreads();
      }
    }
  } else if (!j.unpack) {
    const float* src = (const float*)j.src;
    const unsigned plane = (unsigned)j.R_pad * j.C_pad;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
      const unsigned r = i / (unsigned)j.C_pad;
      const unsigned c = i - r * (unsigned)j.C_pad;
      const bool in = (int)r < j.R && (int)c < j.Cc;
      const float* sp = src + r * j.s_r + c * j.s_c;
      const float rs = (in && j.row_scale) ? j.row_scale[r] : 1.f;
      if (j.dst_dtype == DN_F16) {
        __half* d = (__half*)j.dst + i;
        for (unsigned t = 0; t < (unsigned)j.T; ++t) d[(long long)t * plane] = __float2half_rn(in ? sp[toffs[t]] * rs : 0.f);
      } else if (j.dst_dtype == DN_BF16) {
        __nv_bfloat16* d = (__nv_bfloat16*)j.dst + i;
        for (unsigned t = 0; t < (unsigned)j.T; ++t) d[(long long)t * plane] = __float2bfloat16_rn(in ? sp[toffs[t]] * rs : 0.f);
      } else if (j.dst_dtype == DN_BF16_LO) {
        for (unsigned t = 0; t < (unsigned)j.T; ++t) dn_st(j.dst, DN_BF16_LO, (long long)t * plane + i, in ? sp[toffs[t]] * rs : 0.f);
      } else {
        float* d = (float*)j.dst + i;
        for (unsigned t = 0; t < (unsigned)j.T; ++t) d[(long long)t * plane] = in ? sp[toffs[t]] * rs : 0.f;
      }
    }
  } else if (j.T <= 16 && j.s_c == j.T && j.s_kh == j.k && j.s_kw == 1) {
    // Unpack into an [R][Cc][kh][kw] gradient (nn.Conv2d): for one row r, 32 columns x T taps are 32*T contiguous floats on
    // the torch side.  A warp gathers them tap by tap (unit stride along the columns of the packed planes), stages them in
    // shared memory and writes the run with unit stride.
    __shared__ float usm[8][32 * 17];
    const float* src = (const float*)j.src;
    float* dst = (float*)j.dst;
    const unsigned T = (unsigned)j.T;
    const unsigned ctiles = ((unsigned)j.Cc + 31) / 32;
    const unsigned units = (unsigned)j.R * ctiles;
    const long long planeP = (long long)j.R_pad * j.C_pad;
    for (unsigned u = blockIdx.x * 8 + wrp; u < units; u += gridDim.x * 8) {
      const unsigned r = u / ctiles, c0 = (u - r * ctiles) * 32;
      const unsigned nc = (unsigned)j.Cc - c0 < 32 ? (unsigned)j.Cc - c0 : 32;
      const float* sp = src + (long long)r * j.C_pad + c0 + lane;
      for (unsigned t = 0; t < T; ++t) usm[wrp][lane * T + t] = lane < (int)nc ? j.scale * sp[(long long)t * planeP] : 0.f;
      __syncwarp();
      float* dp = dst + r * j.s_r + (long long)c0 * T;
      for (unsigned q = lane; q < nc * T; q += 32) dp[q] = usm[wrp][q];
      __syncwarp();
    }
  } else {
    const float* src = (const float*)j.src;
    float* dst = (float*)j.dst;
    const unsigned plane = (unsigned)j.R * j.Cc;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
      const unsigned r = i / (unsigned)j.Cc;
      const unsigned c = i - r * (unsigned)j.Cc;
      float* dp = dst + r * j.s_r + c * j.s_c;
      const float* sp = src + (long long)r * j.C_pad + c;
      for (unsigned t = 0; t < (unsigned)j.T; ++t) dp[toffs[t]] = j.scale * sp[(long long)t * j.R_pad * j.C_pad];
    }
  }
}

DN_EXPORT int dn_pack_jobs(const dn_pack_job* jobs, int njobs, void* stream) {
  if (!jobs || njobs < 1) return DN_E_ARG;
  dn_launch(pack_jobs_kernel, dim3(512, njobs), dim3(256), 0, dn_stream(stream), jobs);
  DN_CHECK_LAUNCH();
  return 0;
}

// =================================================================================================
// CUDA-core gather-convolution (backend 0): 256 threads, BMxBN output tile, 4x4 per thread, K chunks of 16
// =================================================================================================
template <int BM, int BN>
__global__ void __launch_bounds__(256) igemm_generic_kernel(const __grid_constant__ dn_igemm p) {
  constexpr int BK = 16;
  constexpr int TXN = BN / 4;
  static_assert((BM / 4) * (BN / 4) == 256, "tile/threads mismatch");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int rn[BM], rh[BM], rw[BM];
  const int tid = threadIdx.x;
  const long long M = (long long)p.out.N * p.out.H * p.out.W;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  for (int r = tid; r < BM; r += 256) {
    long long m = m0 + r;
    if (m < M) {
      int wo = (int)(m % p.out.W);
      long long q = m / p.out.W;
      rw[r] = wo; rh[r] = (int)(q % p.out.H); rn[r] = (int)(q / p.out.H);
    } else { rn[r] = -1; rh[r] = 0; rw[r] = 0; }
  }
  __syncthreads();
  const int ty = tid / TXN, tx = tid % TXN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < p.ntaps; ++t) {
    const dn_tap tap = p.taps[t];
    const dn_view& src = p.in[tap.src];
    const int Cin = src.C;
    const long long wbase = (long long)tap.wt * p.cout_pad * p.cin_pad;
    for (int c0 = 0; c0 < Cin; c0 += BK) {
      for (int i = tid; i < BM * BK; i += 256) {
        int r = i / BK, k = i % BK;
        float v = 0.f;
        int n = rn[r];
        int ci = c0 + k;
        if (n >= 0 && ci < Cin) {
          int hi = rh[r] * p.stride + tap.dh, wi = rw[r] * p.stride + tap.dw;
          if (hi >= 0 && hi < src.H && wi >= 0 && wi < src.W) v = dn_ld(src.ptr, src.dtype, dn_off(src, n, hi, wi) + ci);
        }
        As[k][r] = v;
      }
      for (int i = tid; i < BN * BK; i += 256) {
        int c = i / BK, k = i % BK;
        int co = n0 + c, ci = c0 + k;
        float v = 0.f;
        if (co < p.cout_pad && ci < p.cin_pad) v = dn_ld(p.w, p.w_dtype, wbase + (long long)co * p.cin_pad + ci);
        Bs[k][c] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = ty * 4 + i;
    int n = rn[r];
    if (n < 0) continue;
    long long o = dn_off(p.out, n, rh[r], rw[r]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= p.out.C) continue;
      float v = acc[i][j] * p.out_scale;
      if (p.bias) v += p.bias[co];
      v = dn_act(v, p.act);
      if (p.accumulate) v += dn_ld(p.out.ptr, p.out.dtype, o + co);
      dn_st(p.out.ptr, p.out.dtype, o + co, v);
      if (p.out2) dn_st(p.out2, p.out2_dtype, o + co, v);
    }
  }
}

static int igemm_check(const dn_igemm* p) {
  if (!p || p->nsrc < 1 || p->nsrc > DN_MAX_SRC || p->ntaps < 1 || p->ntaps > DN_MAX_TAPS) return DN_E_ARG;
  if (p->stride < 1 || !p->w || !p->out.ptr) return DN_E_ARG;
  if (p->accumulate && p->act != DN_ACT_NONE) return DN_E_ARG;
  for (int t = 0; t < p->ntaps; ++t)
    if (p->taps[t].src < 0 || p->taps[t].src >= p->nsrc) return DN_E_ARG;
  for (int s = 0; s < p->nsrc; ++s)
    if (!p->in[s].ptr || p->in[s].C > p->cin_pad || p->in[s].N != p->out.N) return DN_E_ARG;
  if (p->out.C > p->cout_pad) return DN_E_ARG;
  if (p->phase_cout > 0 && (p->nphase < 2 || p->cout_pad != p->nphase * p->phase_cout || p->out.C > p->phase_cout)) return DN_E_ARG;
  if (p->nphase < 0 || p->nphase > 4 || (p->nphase > 1 && p->phase_cout <= 0 && p->ntaps % p->nphase != 0)) return DN_E_ARG;
  return 0;
}

int dn_igemm_generic(const dn_igemm* p, cudaStream_t st) {
  long long M = (long long)p->out.N * p->out.H * p->out.W;
  if (M == 0) return 0;
  if (p->nphase > 1 && p->phase_cout > 0) return DN_E_UNSUPPORTED;      // (channel-stacked phases exist on the tensor-core path only)
  if (p->nphase > 1) {      // merged output phases: one single-view problem per phase
    const int tph = p->ntaps / p->nphase;
    const int esz = dn_esize(p->out.dtype);
    for (int i = 0; i < p->nphase; ++i) {
      dn_igemm q = *p;
      q.nphase = 0;
      q.ntaps = tph;
      for (int t = 0; t < tph; ++t) q.taps[t] = p->taps[i * tph + t];
      q.out.ptr = (char*)p->out.ptr + p->phase_off[i] * esz;
      if (p->out2) q.out2 = (char*)p->out2 + p->phase_off[i] * 2;
      int e = dn_igemm_generic(&q, st);
      if (e) return e;
    }
    return 0;
  }
  if (p->out.C <= 16) {
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)((p->out.C + 15) / 16));
    igemm_generic_kernel<256, 16><<<grid, 256, 0, st>>>(*p);
  } else {
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((p->out.C + 63) / 64));
    igemm_generic_kernel<64, 64><<<grid, 256, 0, st>>>(*p);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

// weight gradient on CUDA cores: one 64x64 (cp x cq) tile per block per tap per pixel split
__global__ void __launch_bounds__(256) wgrad_generic_kernel(const __grid_constant__ dn_wgrad p, int splits, long long chunk) {
  constexpr int BK = 16;
  __shared__ float As[BK][64 + 4];
  __shared__ float Bs[BK][64 + 4];
  __shared__ long long offP[BK], offQ[BK];
  const int tid = threadIdx.x;
  const int t = blockIdx.z / splits, split = blockIdx.z % splits;
  const dn_tap tap = p.taps[t];
  const dn_view& P = p.p[tap.src];
  const dn_view& Q = p.q;
  const long long M = (long long)P.N * P.H * P.W;
  const long long mbeg = split * chunk;
  long long mend = mbeg + chunk;
  if (mend > M) mend = M;
  const int cp0 = blockIdx.x * 64, cq0 = blockIdx.y * 64;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long m0 = mbeg; m0 < mend; m0 += BK) {
    if (tid < BK) {
      long long m = m0 + tid;
      long long op = -1, oq = -1;
      if (m < mend) {
        int w = (int)(m % P.W);
        long long q = m / P.W;
        int h = (int)(q % P.H), n = (int)(q / P.H);
        op = dn_off(P, n, h, w);
        int hi = h * p.stride + tap.dh, wi = w * p.stride + tap.dw;
        if (hi >= 0 && hi < Q.H && wi >= 0 && wi < Q.W) oq = dn_off(Q, n, hi, wi);
      }
      offP[tid] = op; offQ[tid] = oq;
    }
    __syncthreads();
    for (int i = tid; i < BK * 64; i += 256) {
      int k = i / 64, c = i % 64;
      float a = 0.f, b = 0.f;
      if (offP[k] >= 0 && offQ[k] >= 0) {
        if (cp0 + c < P.C) a = dn_ld(P.ptr, P.dtype, offP[k] + cp0 + c);
        if (cq0 + c < Q.C) b = dn_ld(Q.ptr, Q.dtype, offQ[k] + cq0 + c);
      }
      As[k][c] = a; Bs[k][c] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dw = p.dw + (long long)tap.wt * p.cp_pad * p.cq_pad;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int cp = cp0 + ty * 4 + i;
    if (cp >= P.C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int cq = cq0 + tx * 4 + j;
      if (cq >= Q.C) continue;
      atomicAdd(dw + (long long)cp * p.cq_pad + cq, acc[i][j] * p.scale);
    }
  }
}

int dn_wgrad_generic(const dn_wgrad* p, cudaStream_t st) {
  const dn_view& P = p->p[0];
  long long M = (long long)P.N * P.H * P.W;
  if (M == 0) return 0;
  int tiles = ((P.C + 63) / 64) * ((p->q.C + 63) / 64) * p->ntaps;
  int target = dn_num_sms() * 4;
  int splits = (target + tiles - 1) / tiles;
  long long maxsplits = (M + 255) / 256;
  if (splits > maxsplits) splits = (int)maxsplits;
  if (splits < 1) splits = 1;
  while ((long long)splits * p->ntaps > 65535) splits = (splits + 1) / 2;
  long long chunk = (M + splits - 1) / splits;
  chunk = (chunk + 15) / 16 * 16;
  dim3 grid((P.C + 63) / 64, (p->q.C + 63) / 64, p->ntaps * splits);
  wgrad_generic_kernel<<<grid, 256, 0, st>>>(*p, splits, chunk);
  DN_CHECK_LAUNCH();
  return 0;
}

static int wgrad_check(const dn_wgrad* p) {
  if (!p || p->nsrc < 1 || p->nsrc > DN_MAX_SRC || p->ntaps < 1 || p->ntaps > DN_MAX_TAPS || !p->dw || !p->q.ptr) return DN_E_ARG;
  for (int s = 0; s < p->nsrc; ++s) {
    if (!p->p[s].ptr || p->p[s].C > p->cp_pad) return DN_E_ARG;
    if (p->p[s].N != p->p[0].N || p->p[s].H != p->p[0].H || p->p[s].W != p->p[0].W || p->p[s].C != p->p[0].C) return DN_E_ARG;
  }
  if (p->q.C > p->cq_pad || p->q.N != p->p[0].N) return DN_E_ARG;
  for (int t = 0; t < p->ntaps; ++t)
    if (p->taps[t].src < 0 || p->taps[t].src >= p->nsrc) return DN_E_ARG;
  return 0;
}

int dn_igemm_tc(const dn_igemm* p, cudaStream_t st);   // dn_tc.cu
int dn_wgrad_tc(const dn_wgrad* p, cudaStream_t st);   // dn_tc.cu

DN_EXPORT int dn_igemm_run(const dn_igemm* p, int backend, void* stream) {
  int e = igemm_check(p);
  if (e) return e;
  if (backend == 1) {
    if (!dn_igemm_tc_supported(p)) return DN_E_UNSUPPORTED;
    return dn_igemm_tc(p, dn_stream(stream));
  }
  return dn_igemm_generic(p, dn_stream(stream));
}

DN_EXPORT int dn_wgrad_run(const dn_wgrad* p, int backend, void* stream) {
  int e = wgrad_check(p);
  if (e) return e;
  if (backend == 1) {
    if (!dn_wgrad_tc_supported(p)) return DN_E_UNSUPPORTED;
    return dn_wgrad_tc(p, dn_stream(stream));
  }
  return dn_wgrad_generic(p, dn_stream(stream));
}

// =================================================================================================
// channel-group walkers: a thread owns CH consecutive channels (8 via 16-byte vectors, or 1 scalar)
// =================================================================================================
template <int CH>
__device__ __forceinline__ void ldc(const dn_view& v, long long off, float* f) {
  if (CH == 8) {
    if (v.dtype == DN_F16) Vec8<__half>::load((const __half*)v.ptr + off, f);
    else if (v.dtype == DN_BF16) Vec8<__nv_bfloat16>::load((const __nv_bfloat16*)v.ptr + off, f);
    else {
      const float4 a = *reinterpret_cast<const float4*>((const float*)v.ptr + off), b = *reinterpret_cast<const float4*>((const float*)v.ptr + off + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
  } else {
    f[0] = dn_ld(v.ptr, v.dtype, off);
  }
}
template <int CH>
__device__ __forceinline__ void stc(const dn_view& v, long long off, const float* f) {
  if (CH == 8) {
    if (v.dtype == DN_F16) Vec8<__half>::store((__half*)v.ptr + off, f);
    else if (v.dtype == DN_BF16) Vec8<__nv_bfloat16>::store((__nv_bfloat16*)v.ptr + off, f);
    else {
      *reinterpret_cast<float4*>((float*)v.ptr + off) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>((float*)v.ptr + off + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
  } else {
    dn_st(v.ptr, v.dtype, off, f[0]);
  }
}

struct CgGeom {
  int CG;    // channel groups in the tensor
  int CGb;   // channel groups per block (power of two <= 256)
  int PL;    // pixel lanes per block = 256 / CGb
  dim3 grid;
};

// max_cgb: channel groups per block (slab width); blocks_per_sm: cap of the grid in resident-block units
static CgGeom cg_geom(int C, int ch, long long pixels, int max_cgb = 256, int blocks_per_sm = 8) {
  CgGeom g;
  g.CG = (C + ch - 1) / ch;
  int b = 1;
  while (b < g.CG && b < max_cgb) b <<= 1;
  g.CGb = b;
  g.PL = 256 / b;
  int gy = (g.CG + b - 1) / b;
  // aim at >= 8 pixels per thread (the per-block prologue / partial-sum epilogue is a fixed cost), but never fewer blocks
  // than SMs while there is at least one pixel per thread
  long long bx_max = (pixels + g.PL - 1) / g.PL;
  long long bx = (pixels + (long long)g.PL * 8 - 1) / ((long long)g.PL * 8);
  long long floor_blocks = bx_max < dn_num_sms() ? bx_max : dn_num_sms();
  if (bx < floor_blocks) bx = floor_blocks;
  long long cap = (long long)dn_num_sms() * blocks_per_sm / gy;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx > 2048) bx = 2048;
  if (bx < 1) bx = 1;
  g.grid = dim3((unsigned)bx, (unsigned)gy);
  return g;
}

#define CG_PROLOGUE(view_for_dims)                                            \
  const int cgl = threadIdx.x % CGb;                                          \
  const int pl = threadIdx.x / CGb;                                           \
  const int cg = blockIdx.y * CGb + cgl;                                      \
  const int c0 = cg * CH;                                                     \
  const bool cvalid = c0 < (view_for_dims).C;                                 \
  const int PLn = 256 / CGb;                                                  \
  const long long npix = (long long)(view_for_dims).N * (view_for_dims).H * (view_for_dims).W;

// block-level reduction of NV per-thread float values over the pixel lanes; lane 0 threads get the result
template <int NV>
__device__ __forceinline__ void cg_block_reduce(float* vals, int CGb) {
  __shared__ float red[256 * NV];
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; ++i) red[i * 256 + tid] = vals[i];
  __syncthreads();
  // CGb * NV column sums spread over all threads (a column is read and overwritten by its owner only); fixed summation order
  const int PLn = 256 / CGb;
  for (int o = tid; o < CGb * NV; o += 256) {
    const int i = o / CGb, c = o - i * CGb;
    float s = 0.f;
    for (int l = 0; l < PLn; ++l) s += red[i * 256 + l * CGb + c];
    red[i * 256 + c] = s;
  }
  __syncthreads();
  if (tid < CGb) {
#pragma unroll
    for (int i = 0; i < NV; ++i) vals[i] = red[i * 256 + tid];
  }
  __syncthreads();
}

// second stage of the per-channel reductions: out[j] = scale * sum_blk ws[blk][j]  (double accumulation)
constexpr int kMaxReduceBlocks = 2048;
// 1024 threads = 32 columns x 32 row groups; each group strides over the partial rows, then a shared-memory combine
template <typename TO>
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ ws, int nblk, int n, TO* __restrict__ out, double scale,
                                                               int nout = -1) {
  dn_pdl_trigger();
  dn_pdl_wait();
  __shared__ double part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  if (nout < 0) nout = n;            // row stride n, first nout columns reduced
  double s = 0.0;
  if (j < nout) {
    int b = ty;
    for (; b + 96 < nblk; b += 128) {      // 4 independent loads in flight
      float v0 = ws[(long long)b * n + j], v1 = ws[(long long)(b + 32) * n + j];
      float v2 = ws[(long long)(b + 64) * n + j], v3 = ws[(long long)(b + 96) * n + j];
      s += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; b < nblk; b += 32) s += (double)ws[(long long)b * n + j];
  }
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < nout) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += part[k][tx];
    out[j] = (TO)(t * scale);
  }
}
static inline int reduce_blocks(int n) { return (n + 31) / 32; }

// ---- single-launch reductions -------------------------------------------------------------------------------
// A reduction workspace starts with kWsCounters uint32 arrival counters (zero before first use; the last block resets its
// counter, so the buffer can be reused launch after launch on one stream), followed by the partial rows
// [gridDim.x][ncols].  The grid's y index selects a channel slab; the last block of a slab to arrive folds the slab's
// columns over all rows in a fixed order (deterministic, double accumulation) - no second launch.
constexpr int kWsCounters = 256;
DN_EXPORT int64_t dn_reduce_ws_floats(int C) { return (int64_t)kMaxReduceBlocks * 2 * (C + 8) + kWsCounters; }

__device__ __forceinline__ bool dn_slab_last_block(float* ws) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ctr = reinterpret_cast<unsigned*>(ws) + blockIdx.y;
    const unsigned t = atomicAdd(ctr, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
    if (s_last) *ctr = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// fin[seg * cw + i] = sum over rows of rows[r][seg * seglen + cbeg + i], seg < nseg, i < cw.  All 256 threads of the block.
__device__ __noinline__ void dn_slab_fold(const float* rows, int nrows, int ncols, int nseg, int seglen, int cbeg, int cw, double* fin) {
  __shared__ double sm4[256 * 4];
  const int tid = threadIdx.x;
  if (((cw | cbeg | seglen | ncols) & 3) == 0) {
    const int n4 = nseg * cw / 4;
    const int rs4 = ncols / 4;
    for (int base = 0; base < n4; base += 256) {
      const int cnt = n4 - base < 256 ? n4 - base : 256;
      const int lanes = 256 / cnt;
      const int i4 = tid % cnt, lane = tid / cnt;
      const int idx = (base + i4) * 4;
      const int seg = idx / cw;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (lane < lanes) {
        const float4* p = reinterpret_cast<const float4*>(rows + seg * seglen + cbeg + (idx - seg * cw));
        int r = lane;
        for (; r + 7 * lanes < nrows; r += 8 * lanes) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldcg(p + (long long)(r + u * lanes) * rs4);
#pragma unroll
          for (int u = 0; u < 8; ++u) { a0 += (double)v[u].x; a1 += (double)v[u].y; a2 += (double)v[u].z; a3 += (double)v[u].w; }
        }
        for (; r < nrows; r += lanes) {
          const float4 v = __ldcg(p + (long long)r * rs4);
          a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
      }
      sm4[tid * 4 + 0] = a0; sm4[tid * 4 + 1] = a1; sm4[tid * 4 + 2] = a2; sm4[tid * 4 + 3] = a3;
      __syncthreads();
      if (tid < cnt) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double t = 0.0;
          for (int l = 0; l < lanes; ++l) t += sm4[(l * cnt + tid) * 4 + k];
          fin[idx + k] = t;
        }
      }
      __syncthreads();
    }
  } else {
    const int n = nseg * cw;
    for (int base = 0; base < n; base += 256) {
      const int cnt = n - base < 256 ? n - base : 256;
      const int lanes = 256 / cnt;
      const int i = tid % cnt, lane = tid / cnt;
      const int idx = base + i;
      const int seg = idx / cw;
      double a = 0.0;
      if (lane < lanes) {
        const float* p = rows + seg * seglen + cbeg + (idx - seg * cw);
        int r = lane;
        for (; r + 3 * lanes < nrows; r += 4 * lanes) {
          const float v0 = __ldcg(p + (long long)r * ncols), v1 = __ldcg(p + (long long)(r + lanes) * ncols);
          const float v2 = __ldcg(p + (long long)(r + 2 * lanes) * ncols), v3 = __ldcg(p + (long long)(r + 3 * lanes) * ncols);
          a += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
        }
        for (; r < nrows; r += lanes) a += (double)__ldcg(p + (long long)r * ncols);
      }
      sm4[tid] = a;
      __syncthreads();
      if (tid < cnt) {
        double t = 0.0;
        for (int l = 0; l < lanes; ++l) t += sm4[l * cnt + tid];
        fin[idx] = t;
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void bn_scale_shift(float gamma, float beta, float mean, float invstd, float& sc, float& sh) {
  sc = gamma * invstd;
  sh = beta - mean * sc;
}

// DN_BN_FAST=0 routes everything through the generic CG walkers (A/B comparisons)
static const bool g_bn_fast = []() { const char* e = getenv("DN_BN_FAST"); return !(e && e[0] == '0'); }();

struct BnFinalize {        // optional tail of the statistics kernel: what bn_finalize_kernel does, per channel
  int enabled;
  int update_running;
  float momentum, eps;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  float* mean_invstd;
  float* scale_shift;
};

// everything after the per-thread accumulation of the statistics kernels: block reduce, partial row, slab fold, finalize
template <int CH>
__device__ __forceinline__ void bn_stats_tail(float* acc, int C, double count, float* __restrict__ ws, int CGb, double* __restrict__ sums,
                                              const BnFinalize& fz, bool cvalid, int c0) {
  cg_block_reduce<2 * CH>(acc, CGb);
  float* rows = ws + kWsCounters;
  if (threadIdx.x < CGb && cvalid) {
    float* w = rows + (long long)blockIdx.x * 2 * C;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (c0 + i < C) {
        w[c0 + i] = acc[i];
        w[C + c0 + i] = acc[CH + i];
      }
    }
  }
  if (!dn_slab_last_block(ws)) return;
  __shared__ double fin[512];
  const int cbeg = blockIdx.y * CGb * CH;
  const int cw = C - cbeg < CGb * CH ? C - cbeg : CGb * CH;
  dn_slab_fold(rows, gridDim.x, 2 * C, 2, C, cbeg, cw, fin);
  for (int i = threadIdx.x; i < cw; i += blockDim.x) {
    const int c = cbeg + i;
    const double s = fin[i], ss = fin[cw + i];
    if (sums) { sums[c] = s; sums[C + c] = ss; }
    if (fz.enabled) {
      const double m = s / count;
      double var = ss / count - m * m;
      if (var < 0) var = 0;
      const float mean = (float)m;
      const float invstd = (float)(1.0 / sqrt(var + (double)fz.eps));
      if (fz.update_running) {
        const double unb = count > 1 ? var * count / (count - 1) : var;
        fz.running_mean[c] = (1.f - fz.momentum) * fz.running_mean[c] + fz.momentum * mean;
        fz.running_var[c] = (1.f - fz.momentum) * fz.running_var[c] + fz.momentum * (float)unb;
      }
      fz.mean_invstd[c] = mean;
      fz.mean_invstd[C + c] = invstd;
      float sc, sh;
      bn_scale_shift(fz.gamma ? fz.gamma[c] : 1.f, fz.beta ? fz.beta[c] : 0.f, mean, invstd, sc, sh);
      fz.scale_shift[c] = sc;
      fz.scale_shift[C + c] = sh;
    }
  }
  if (fz.enabled && fz.num_batches_tracked && blockIdx.y == 0 && threadIdx.x == 0) *fz.num_batches_tracked += 1;
}

// ---- BN statistics --------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) bn_stats_kernel(dn_view y, float* __restrict__ ws, int CGb, double* __restrict__ sums,
                                                       BnFinalize fz) {
  CG_PROLOGUE(y)
  float acc[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) acc[i] = 0.f;
  if (cvalid) {
    // 4 independent 16-byte loads in flight per thread (HBM latency hiding), tail handled by the validity flags
    const unsigned stride = gridDim.x * PLn;
    for (unsigned px0 = blockIdx.x * PLn + pl; px0 < (unsigned)npix; px0 += 4 * stride) {
      float f[4][CH];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned px = px0 + u * stride;
        ok[u] = px < (unsigned)npix;
        const unsigned pc = ok[u] ? px : px0;
        const unsigned q = pc / (unsigned)y.W;
        const int w = (int)(pc - q * (unsigned)y.W);
        const int n = (int)(q / (unsigned)y.H);
        const int h = (int)(q - (unsigned)n * (unsigned)y.H);
        ldc<CH>(y, dn_off(y, n, h, w) + c0, f[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (ok[u]) {
#pragma unroll
          for (int i = 0; i < CH; ++i) { acc[i] += f[u][i]; acc[CH + i] = fmaf(f[u][i], f[u][i], acc[CH + i]); }
        }
    }
  }
  bn_stats_tail<CH>(acc, y.C, (double)npix, ws, CGb, sums, fz, cvalid, c0);
}

// forward declaration of the tail used by the fast backward-reduce kernel (defined with the generic kernel below)
template <int CH>
__device__ __forceinline__ void bn_bwd_reduce_tail(float* acc, int C, float* __restrict__ ws, int CGb, double* __restrict__ red,
                                                   bool cvalid, int c0);
template <int CH>
__device__ __forceinline__ void act_bwd_tail(float* acc, int C, float* __restrict__ ws, int CGb, float* __restrict__ dbias, float gscale,
                                             bool cvalid, int c0);
#include "dn_bn_fast.cuh"

static int bn_stats_launch(const dn_view* y, double* sums, const BnFinalize& fz, float* ws, void* stream) {
  long long npix = (long long)y->N * y->H * y->W;
  if (dn_vec8_any(y)) {
    CgGeom g = cg_geom(y->C, 8, npix, 8, 4);
    if (g.grid.y > kWsCounters) return DN_E_UNSUPPORTED;
    if (dn_vec8_ok(y) && dn_lin(y) && npix < (1ll << 31) && g_bn_fast) dn_launch(bnf_stats_kernel, g.grid, dim3(256), 0, dn_stream(stream), *y, ws, g.CGb, sums, fz);
    else bn_stats_kernel<8><<<g.grid, 256, 0, dn_stream(stream)>>>(*y, ws, g.CGb, sums, fz);
  } else {
    CgGeom g = cg_geom(y->C, 1, npix, 256, 3);
    if (g.grid.y > kWsCounters) return DN_E_UNSUPPORTED;
    bn_stats_kernel<1><<<g.grid, 256, 0, dn_stream(stream)>>>(*y, ws, g.CGb, sums, fz);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_bn_stats(const dn_view* y, double* sums, float* ws, void* stream) {
  if (!y || !sums || !ws) return DN_E_ARG;
  BnFinalize fz;
  memset(&fz, 0, sizeof(fz));
  return bn_stats_launch(y, sums, fz, ws, stream);
}

// statistics + finalize (+ num_batches_tracked += 1) of a training-mode BatchNorm in ONE launch
DN_EXPORT int dn_bn_train_stats(const dn_view* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, float momentum, float eps, int update_running, double* sums,
                                float* mean_invstd, float* scale_shift, float* ws, void* stream) {
  if (!y || !ws || !mean_invstd || !scale_shift) return DN_E_ARG;
  if (update_running && (!running_mean || !running_var)) return DN_E_ARG;
  BnFinalize fz;
  fz.enabled = 1; fz.update_running = update_running; fz.momentum = momentum; fz.eps = eps;
  fz.gamma = gamma; fz.beta = beta; fz.running_mean = running_mean; fz.running_var = running_var;
  fz.num_batches_tracked = (long long*)num_batches_tracked; fz.mean_invstd = mean_invstd; fz.scale_shift = scale_shift;
  return bn_stats_launch(y, sums, fz, ws, stream);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var, float momentum,
                                   float eps, int training, int update_running, float* mean_invstd, float* scale_shift, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (update_running) {
      double unb = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + eps);
  }
  mean_invstd[c] = mean;
  mean_invstd[C + c] = invstd;
  float sc, sh;
  bn_scale_shift(gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, mean, invstd, sc, sh);
  scale_shift[c] = sc;
  scale_shift[C + c] = sh;
}

__global__ void bn_fold_bias_kernel(const float* __restrict__ cb, const float* __restrict__ ss, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (cb ? cb[c] : 0.f) * ss[c] + ss[C + c];
}
DN_EXPORT int dn_bn_fold_bias(const float* conv_bias, const float* scale_shift, int C, float* bias_out, void* stream) {
  if (!scale_shift || !bias_out || C < 1) return DN_E_ARG;
  bn_fold_bias_kernel<<<(C + 127) / 128, 128, 0, dn_stream(stream)>>>(conv_bias, scale_shift, C, bias_out);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, float momentum, float eps, int training, int update_running,
                             float* mean_invstd, float* scale_shift, int C, void* stream) {
  if (!mean_invstd || !scale_shift || C < 1) return DN_E_ARG;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, dn_stream(stream)>>>(sums, count, gamma, beta, running_mean, running_var,
                                                                      momentum, eps, training, update_running, mean_invstd,
                                                                      scale_shift, C);
  DN_CHECK_LAUNCH();
  return 0;
}

// ---- BN apply (+residual) + act (+2x2 max pool) ------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) bn_apply_kernel(dn_view y, const float* __restrict__ scale_shift, dn_view res, int has_res,
                                                       int act, int pool, dn_view out, dn_view out2, int has_out2, int CGb) {
  CG_PROLOGUE(out)
  if (!cvalid) return;
  float sc[CH], sh[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    int c = c0 + i < out.C ? c0 + i : out.C - 1;
    sc[i] = scale_shift[c];
    sh[i] = scale_shift[out.C + c];
  }
  for (unsigned px = blockIdx.x * PLn + pl; px < (unsigned)npix; px += gridDim.x * PLn) {
    const unsigned q = px / (unsigned)out.W;
    const int w = (int)(px - q * (unsigned)out.W);
    const int n = (int)(q / (unsigned)out.H);
    const int h = (int)(q - (unsigned)n * (unsigned)out.H);
    float o[CH];
    if (pool) {
#pragma unroll
      for (int i = 0; i < CH; ++i) o[i] = -INFINITY;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float f[CH];
          ldc<CH>(y, dn_off(y, n, 2 * h + a, 2 * w + b) + c0, f);
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            float v = dn_act(fmaf(f[i], sc[i], sh[i]), act);
            o[i] = v > o[i] ? v : o[i];
          }
        }
    } else {
      float f[CH], r[CH];
      ldc<CH>(y, dn_off(y, n, h, w) + c0, f);
      if (has_res) ldc<CH>(res, dn_off(res, n, h, w) + c0, r);
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float v = fmaf(f[i], sc[i], sh[i]);
        if (has_res) v += r[i];
        o[i] = dn_act(v, act);
      }
    }
    stc<CH>(out, dn_off(out, n, h, w) + c0, o);
    if (has_out2) stc<CH>(out2, dn_off(out2, n, h, w) + c0, o);
  }
}

DN_EXPORT int dn_bn_apply(const dn_view* y, const float* scale_shift, const dn_view* residual, int act, int pool,
                          const dn_view* out, const dn_view* out2, void* stream) {
  if (!y || !out || !scale_shift) return DN_E_ARG;
  if (pool && (residual || out->H * 2 > y->H || out->W * 2 > y->W)) return DN_E_ARG;
  if (!pool && (out->H != y->H || out->W != y->W)) return DN_E_ARG;
  if (out->C != y->C || out->N != y->N) return DN_E_ARG;
  long long npix = (long long)out->N * out->H * out->W;
  dn_view r = residual ? *residual : *y;
  dn_view o2 = out2 ? *out2 : *out;
  if (out2 && (out2->C != out->C || out2->H != out->H || out2->W != out->W || out2->N != out->N)) return DN_E_ARG;
  bool vec = dn_vec8_any(y) && dn_vec8_any(out) && (!residual || dn_vec8_any(residual)) && (!out2 || dn_vec8_any(out2));
  const bool all16 = dn_vec8_ok(y) && dn_vec8_ok(out) && (!residual || dn_vec8_ok(residual)) && (!out2 || dn_vec8_ok(out2));
  const bool fast = vec && all16 && g_bn_fast && (!residual || (!pool && dn_lin(residual))) && dn_lin(y) && dn_lin(out) && (!out2 || dn_lin(out2)) &&
                    npix < (1ll << 31) && (!pool || (y->H == 2 * out->H && y->W >= 2 * out->W));
  if (fast) {
    CgGeom g = cg_geom(out->C, 8, npix, 256, 4);
    if (pool) dn_launch(bnf_apply_kernel<true, false>, g.grid, dim3(256), 0, dn_stream(stream), *y, scale_shift, act, *out, o2, out2 != nullptr, g.CGb, r);
    else if (residual) dn_launch(bnf_apply_kernel<false, true>, g.grid, dim3(256), 0, dn_stream(stream), *y, scale_shift, act, *out, o2, out2 != nullptr, g.CGb, r);
    else dn_launch(bnf_apply_kernel<false, false>, g.grid, dim3(256), 0, dn_stream(stream), *y, scale_shift, act, *out, o2, out2 != nullptr, g.CGb, r);
  } else if (vec) {
    CgGeom g = cg_geom(out->C, 8, npix);
    bn_apply_kernel<8><<<g.grid, 256, 0, dn_stream(stream)>>>(*y, scale_shift, r, residual != nullptr, act, pool, *out, o2, out2 != nullptr, g.CGb);
  } else {
    CgGeom g = cg_geom(out->C, 1, npix);
    bn_apply_kernel<1><<<g.grid, 256, 0, dn_stream(stream)>>>(*y, scale_shift, r, residual != nullptr, act, pool, *out, o2, out2 != nullptr, g.CGb);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

// Backward of BatchNorm(+residual)+act(+2x2 max pool).
//
// With g = dout * act'(.) routed to the selected position, the two per-channel reductions BN backward needs are
//   S1 = sum g      and      sum g * xhat = invstd * (S2 - mean * S1),  S2 = sum g * y,
// so the reduce pass accumulates the *raw* S1, S2 and needs only the forward scale/shift per channel (to redo the
// activation / arg-max decision with exactly the forward arithmetic).  The apply pass folds everything else into three
// per-channel constants:  dy = A*g - B - Cc*y,  A = gamma*invstd, Cc = A*invstd*m2, B = A*m1 - Cc*mean, m1 = S1/M,
// m2 = invstd*(S2 - mean*S1)/M.  Few per-thread constants = few registers = enough resident warps to hide HBM latency.
template <int CH, bool POOL>
struct BnBwdPix {
  float yv[POOL ? 4 : 1][CH];
  float g[CH];          // dout * act'(.) of the selected position (pooled) / of the pixel
  unsigned sel;         // pooled: 2 bits per channel, which of the 4 window positions receives g

  __device__ __forceinline__ void load(const dn_view& y, const dn_view& res, int has_res, const dn_view& dout, int n, int h,
                                       int w, int c0, const float* sc, const float* sh, int act) {
    ldc<CH>(dout, dn_off(dout, n, h, w) + c0, g);
    sel = 0;
    if (POOL) {
#pragma unroll
      for (int k = 0; k < 4; ++k) ldc<CH>(y, dn_off(y, n, 2 * h + (k >> 1), 2 * w + (k & 1)) + c0, yv[POOL ? k : 0]);
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float best = -INFINITY;
        unsigned bk = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v = dn_act(fmaf(yv[POOL ? k : 0][i], sc[i], sh[i]), act);
          if (v > best) { best = v; bk = k; }
        }
        g[i] *= dn_act_grad(best, act);
        sel |= bk << (2 * i);
      }
    } else {
      ldc<CH>(y, dn_off(y, n, h, w) + c0, yv[0]);
      float r[CH];
      if (has_res) ldc<CH>(res, dn_off(res, n, h, w) + c0, r);
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float v = fmaf(yv[0][i], sc[i], sh[i]);
        if (has_res) v += r[i];
        g[i] *= dn_act_grad(dn_act(v, act), act);
      }
    }
  }
  __device__ __forceinline__ int which(int i) const { return POOL ? (int)((sel >> (2 * i)) & 3u) : 0; }
  __device__ __forceinline__ float dyhat(int k, int i) const { return (!POOL || which(i) == k) ? g[i] : 0.f; }
  __device__ __forceinline__ float ysel(int i) const {
    float v = yv[0][i];
    if (POOL) {
      const int k = which(i);
      v = k == 1 ? yv[POOL ? 1 : 0][i] : v;
      v = k == 2 ? yv[POOL ? 2 : 0][i] : v;
      v = k == 3 ? yv[POOL ? 3 : 0][i] : v;
    }
    return v;
  }
};

template <int CH>
__device__ __forceinline__ void bn_bwd_reduce_tail(float* acc, int C, float* __restrict__ ws, int CGb, double* __restrict__ red,
                                                   bool cvalid, int c0) {
  cg_block_reduce<2 * CH>(acc, CGb);
  float* rows = ws + kWsCounters;
  if (threadIdx.x < CGb && cvalid) {
    float* wsp = rows + (long long)blockIdx.x * 2 * C;
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (c0 + i < C) {
        wsp[c0 + i] = acc[i];
        wsp[C + c0 + i] = acc[CH + i];
      }
  }
  if (!dn_slab_last_block(ws)) return;
  __shared__ double fin[512];
  const int cbeg = blockIdx.y * CGb * CH;
  const int cw = C - cbeg < CGb * CH ? C - cbeg : CGb * CH;
  dn_slab_fold(rows, gridDim.x, 2 * C, 2, C, cbeg, cw, fin);
  for (int i = threadIdx.x; i < cw; i += blockDim.x) {
    red[cbeg + i] = fin[i];
    red[C + cbeg + i] = fin[cw + i];
  }
}

template <int CH, bool POOL>
__global__ void __launch_bounds__(256, POOL ? 2 : 3) bn_bwd_reduce_kernel(dn_view dout, dn_view y, dn_view res, int has_res,
                                                            const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int act, float* __restrict__ ws, int CGb,
                                                            double* __restrict__ red) {
  CG_PROLOGUE(dout)
  const int C = dout.C;
  float acc[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) acc[i] = 0.f;
  if (cvalid) {
    float sc[CH], sh[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      int c = c0 + i < C ? c0 + i : C - 1;
      bn_scale_shift(gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, mean_invstd[c], mean_invstd[C + c], sc[i], sh[i]);
    }
    if (!POOL && !has_res) {
      // 2 pixels per iteration, loads issued up front
      const unsigned stride = gridDim.x * PLn;
      for (unsigned px0 = blockIdx.x * PLn + pl; px0 < (unsigned)npix; px0 += 2 * stride) {
        float gg[2][CH], yy[2][CH];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const unsigned px = px0 + u * stride;
          ok[u] = px < (unsigned)npix;
          const unsigned pc = ok[u] ? px : px0;
          const unsigned q = pc / (unsigned)dout.W;
          const int w = (int)(pc - q * (unsigned)dout.W);
          const int n = (int)(q / (unsigned)dout.H);
          const int h = (int)(q - (unsigned)n * (unsigned)dout.H);
          ldc<CH>(dout, dn_off(dout, n, h, w) + c0, gg[u]);
          ldc<CH>(y, dn_off(y, n, h, w) + c0, yy[u]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (ok[u]) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const float g = gg[u][i] * dn_act_grad(dn_act(fmaf(yy[u][i], sc[i], sh[i]), act), act);
              acc[i] += g;
              acc[CH + i] = fmaf(g, yy[u][i], acc[CH + i]);
            }
          }
      }
    } else {
      for (unsigned px = blockIdx.x * PLn + pl; px < (unsigned)npix; px += gridDim.x * PLn) {
        const unsigned q = px / (unsigned)dout.W;
        const int w = (int)(px - q * (unsigned)dout.W);
        const int n = (int)(q / (unsigned)dout.H);
        const int h = (int)(q - (unsigned)n * (unsigned)dout.H);
        BnBwdPix<CH, POOL> P;
        P.load(y, res, has_res, dout, n, h, w, c0, sc, sh, act);
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          acc[i] += P.g[i];
          acc[CH + i] = fmaf(P.g[i], P.ysel(i), acc[CH + i]);
        }
      }
    }
  }
  bn_bwd_reduce_tail<CH>(acc, C, ws, CGb, red, cvalid, c0);
}

DN_EXPORT int dn_bn_bwd_reduce(const dn_view* dout, const dn_view* y, const dn_view* residual, const float* mean_invstd,
                               const float* gamma, const float* beta, int act, int pool, double* red, float* ws, void* stream) {
  if (!dout || !y || !mean_invstd || !red || !ws) return DN_E_ARG;
  long long npix = (long long)dout->N * dout->H * dout->W;
  dn_view r = residual ? *residual : *y;
  bool vec = dn_vec8_any(y) && dn_vec8_any(dout) && (!residual || dn_vec8_any(residual));
  const bool all16 = dn_vec8_ok(y) && dn_vec8_ok(dout) && (!residual || dn_vec8_ok(residual));
  cudaStream_t st = dn_stream(stream);
  const int ch = vec ? 8 : 1;
  CgGeom g = cg_geom(dout->C, ch, npix, vec ? 8 : 256, 3);
  if (g.grid.y > kWsCounters) return DN_E_UNSUPPORTED;
  const int hr = residual != nullptr;
  const bool fast = vec && all16 && g_bn_fast && (!residual || (!pool && dn_lin(residual))) && dn_lin(dout) && dn_lin(y) && npix < (1ll << 31) &&
                    (!pool || (y->H == 2 * dout->H && y->W >= 2 * dout->W));
  if (fast && pool) dn_launch(bnf_bwd_reduce_kernel<true, false>, g.grid, dim3(256), 0, st, *dout, *y, mean_invstd, gamma, beta, act, ws, g.CGb, red, r);
  else if (fast && residual) dn_launch(bnf_bwd_reduce_kernel<false, true>, g.grid, dim3(256), 0, st, *dout, *y, mean_invstd, gamma, beta, act, ws, g.CGb, red, r);
  else if (fast) dn_launch(bnf_bwd_reduce_kernel<false, false>, g.grid, dim3(256), 0, st, *dout, *y, mean_invstd, gamma, beta, act, ws, g.CGb, red, r);
  else if (vec && pool) bn_bwd_reduce_kernel<8, true><<<g.grid, 256, 0, st>>>(*dout, *y, r, hr, mean_invstd, gamma, beta, act, ws, g.CGb, red);
  else if (vec) bn_bwd_reduce_kernel<8, false><<<g.grid, 256, 0, st>>>(*dout, *y, r, hr, mean_invstd, gamma, beta, act, ws, g.CGb, red);
  else if (pool) bn_bwd_reduce_kernel<1, true><<<g.grid, 256, 0, st>>>(*dout, *y, r, hr, mean_invstd, gamma, beta, act, ws, g.CGb, red);
  else bn_bwd_reduce_kernel<1, false><<<g.grid, 256, 0, st>>>(*dout, *y, r, hr, mean_invstd, gamma, beta, act, ws, g.CGb, red);
  DN_CHECK_LAUNCH();
  return 0;
}

template <int CH, bool POOL>
__global__ void __launch_bounds__(256, POOL ? 2 : 3) bn_bwd_apply_kernel(dn_view dout, dn_view y, dn_view res, int has_res,
                                                           const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int act,
                                                           const double* __restrict__ red, double count, float gscale,
                                                           float* dgamma, float* dbeta, dn_view dy, dn_view dres, int has_dres,
                                                           int dres_acc, int CGb) {
  CG_PROLOGUE(dout)
  const int C = dout.C;
  if (!cvalid) return;
  float sc[CH], sh[CH], cB[CH], cC[CH];      // A == sc
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    int c = c0 + i < C ? c0 + i : C - 1;
    const float mean = mean_invstd[c], istd = mean_invstd[C + c];
    bn_scale_shift(gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, mean, istd, sc[i], sh[i]);
    const double s1 = red[c], s2 = red[C + c];
    const double sgx = (double)istd * (s2 - (double)mean * s1);      // sum g * xhat
    const float m1 = (float)(s1 / count), m2 = (float)(sgx / count);
    cC[i] = sc[i] * istd * m2;
    cB[i] = sc[i] * m1 - cC[i] * mean;
    if (blockIdx.x == 0 && pl == 0 && c0 + i < C) {
      if (dgamma) dgamma[c] = (float)(sgx * (double)gscale);
      if (dbeta) dbeta[c] = (float)(s1 * (double)gscale);
    }
  }
  if (!POOL && !has_res && !has_dres) {
    const unsigned stride = gridDim.x * PLn;
    for (unsigned px0 = blockIdx.x * PLn + pl; px0 < (unsigned)npix; px0 += 2 * stride) {
      float gg[2][CH], yy[2][CH];
      long long offo[2];
      bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const unsigned px = px0 + u * stride;
        ok[u] = px < (unsigned)npix;
        const unsigned pc = ok[u] ? px : px0;
        const unsigned q = pc / (unsigned)dout.W;
        const int w = (int)(pc - q * (unsigned)dout.W);
        const int n = (int)(q / (unsigned)dout.H);
        const int h = (int)(q - (unsigned)n * (unsigned)dout.H);
        ldc<CH>(dout, dn_off(dout, n, h, w) + c0, gg[u]);
        ldc<CH>(y, dn_off(y, n, h, w) + c0, yy[u]);
        offo[u] = dn_off(dy, n, h, w) + c0;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (ok[u]) {
          float o[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const float g = gg[u][i] * dn_act_grad(dn_act(fmaf(yy[u][i], sc[i], sh[i]), act), act);
            o[i] = fmaf(sc[i], g, -fmaf(cC[i], yy[u][i], cB[i]));
          }
          stc<CH>(dy, offo[u], o);
        }
    }
    return;
  }
  for (unsigned px = blockIdx.x * PLn + pl; px < (unsigned)npix; px += gridDim.x * PLn) {
    const unsigned q = px / (unsigned)dout.W;
    const int w = (int)(px - q * (unsigned)dout.W);
    const int n = (int)(q / (unsigned)dout.H);
    const int h = (int)(q - (unsigned)n * (unsigned)dout.H);
    BnBwdPix<CH, POOL> P;
    P.load(y, res, has_res, dout, n, h, w, c0, sc, sh, act);
#pragma unroll
    for (int k = 0; k < (POOL ? 4 : 1); ++k) {
      float o[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) o[i] = fmaf(sc[i], P.dyhat(k, i), -fmaf(cC[i], P.yv[POOL ? k : 0][i], cB[i]));
      int hh = POOL ? 2 * h + (k >> 1) : h, ww = POOL ? 2 * w + (k & 1) : w;
      stc<CH>(dy, dn_off(dy, n, hh, ww) + c0, o);
    }
    if (has_dres) {
      float o[CH];
      long long off = dn_off(dres, n, h, w) + c0;
      if (dres_acc) ldc<CH>(dres, off, o);
#pragma unroll
      for (int i = 0; i < CH; ++i) o[i] = dres_acc ? o[i] + P.g[i] : P.g[i];
      stc<CH>(dres, off, o);
    }
  }
}

DN_EXPORT int dn_bn_bwd_apply(const dn_view* dout, const dn_view* y, const dn_view* residual, const float* mean_invstd,
                              const float* gamma, const float* beta, int act, int pool, const double* red, double count,
                              float gscale, float* dgamma, float* dbeta, const dn_view* dy, const dn_view* dres,
                              int dres_accumulate, void* stream) {
  if (!dout || !y || !mean_invstd || !red || !dy) return DN_E_ARG;
  long long npix = (long long)dout->N * dout->H * dout->W;
  dn_view r = residual ? *residual : *y;
  dn_view dr = dres ? *dres : *dy;
  bool vec = dn_vec8_any(y) && dn_vec8_any(dout) && dn_vec8_any(dy) && (!residual || dn_vec8_any(residual)) && (!dres || dn_vec8_any(dres));
  const bool all16 = dn_vec8_ok(y) && dn_vec8_ok(dout) && dn_vec8_ok(dy) && (!residual || dn_vec8_ok(residual)) && (!dres || dn_vec8_ok(dres));
  cudaStream_t st = dn_stream(stream);
  const int ch = vec ? 8 : 1;
  CgGeom g = cg_geom(dout->C, ch, npix);
  const int hr = residual != nullptr, hd = dres != nullptr;
#define BN_BWD_APPLY(CHV, PV)                                                                                              \
  bn_bwd_apply_kernel<CHV, PV><<<g.grid, 256, 0, st>>>(*dout, *y, r, hr, mean_invstd, gamma, beta, act, red, count, gscale, \
                                                       dgamma, dbeta, *dy, dr, hd, dres_accumulate, g.CGb)
  const bool res_ok = (!residual && !dres) || (residual && !pool && dn_lin(residual) && (!dres || dn_lin(dres)));
  const bool fast = vec && all16 && g_bn_fast && res_ok && dn_lin(dout) && dn_lin(y) && dn_lin(dy) && npix < (1ll << 31) &&
                    (!pool || (y->H == 2 * dout->H && y->W >= 2 * dout->W));
  if (fast) {
    CgGeom gf = cg_geom(dout->C, 8, npix, 256, 3);
    if (pool) dn_launch(bnf_bwd_apply_kernel<true, false>, gf.grid, dim3(256), 0, st, *dout, *y, mean_invstd, gamma, beta, act, red, count, gscale, dgamma, dbeta, *dy, gf.CGb, r, dr, hd, dres_accumulate);
    else if (residual) dn_launch(bnf_bwd_apply_kernel<false, true>, gf.grid, dim3(256), 0, st, *dout, *y, mean_invstd, gamma, beta, act, red, count, gscale, dgamma, dbeta, *dy, gf.CGb, r, dr, hd, dres_accumulate);
    else dn_launch(bnf_bwd_apply_kernel<false, false>, gf.grid, dim3(256), 0, st, *dout, *y, mean_invstd, gamma, beta, act, red, count, gscale, dgamma, dbeta, *dy, gf.CGb, r, dr, hd, dres_accumulate);
  } else if (vec && pool) BN_BWD_APPLY(8, true);
  else if (vec) BN_BWD_APPLY(8, false);
  else if (pool) BN_BWD_APPLY(1, true);
  else BN_BWD_APPLY(1, false);
#undef BN_BWD_APPLY
  DN_CHECK_LAUNCH();
  return 0;
}

template <int CH>
__device__ __forceinline__ void act_bwd_tail(float* acc, int C, float* __restrict__ ws, int CGb, float* __restrict__ dbias, float gscale,
                                             bool cvalid, int c0) {
  cg_block_reduce<CH>(acc, CGb);
  float* rows = ws + kWsCounters;
  const int ncols = (C + 3) & ~3;          // padded row pitch keeps the vector fold aligned
  if (threadIdx.x < CGb && cvalid) {
    float* w = rows + (long long)blockIdx.x * ncols;
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (c0 + i < C) w[c0 + i] = acc[i];
  }
  if (!dn_slab_last_block(ws)) return;
  __shared__ double fin[512];
  const int cbeg = blockIdx.y * CGb * CH;
  const int cw = C - cbeg < CGb * CH ? C - cbeg : CGb * CH;
  dn_slab_fold(rows, gridDim.x, ncols, 1, ncols, cbeg, cw, fin);
  for (int i = threadIdx.x; i < cw; i += blockDim.x) dbias[cbeg + i] = (float)(fin[i] * (double)gscale);
}

// ---- activation backward (in place) + bias gradient ---------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) act_bwd_kernel(dn_view dout, dn_view out, int act, float* ws, int CGb, float* dbias, float gscale) {
  CG_PROLOGUE(dout)
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = 0.f;
  if (cvalid) {
    for (unsigned px = blockIdx.x * PLn + pl; px < (unsigned)npix; px += gridDim.x * PLn) {
      const unsigned q = px / (unsigned)dout.W;
      const int w = (int)(px - q * (unsigned)dout.W);
      const int n = (int)(q / (unsigned)dout.H);
      const int h = (int)(q - (unsigned)n * (unsigned)dout.H);
      float g[CH], o[CH];
      long long off = dn_off(dout, n, h, w) + c0;
      ldc<CH>(dout, off, g);
      if (act != DN_ACT_NONE) {
        ldc<CH>(out, dn_off(out, n, h, w) + c0, o);
#pragma unroll
        for (int i = 0; i < CH; ++i) g[i] *= dn_act_grad(o[i], act);
        stc<CH>(dout, off, g);
      }
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] += g[i];
    }
  }
  if (!ws) return;
  act_bwd_tail<CH>(acc, dout.C, ws, CGb, dbias, gscale, cvalid, c0);
}

DN_EXPORT int dn_act_bwd(const dn_view* dout, const dn_view* out, int act, float* dbias, float gscale, float* ws, void* stream) {
  if (!dout || (!out && act != DN_ACT_NONE) || (dbias && !ws)) return DN_E_ARG;
  if (act == DN_ACT_NONE && !dbias) return 0;
  long long npix = (long long)dout->N * dout->H * dout->W;
  dn_view o = out ? *out : *dout;
  bool vec = dn_vec8_any(dout) && (!out || dn_vec8_any(out));
  const bool all16 = dn_vec8_ok(dout) && (!out || dn_vec8_ok(out));
  CgGeom g;
  if (vec) {
    g = cg_geom(dout->C, 8, npix, 8, 4);
    if (g.grid.y > kWsCounters) return DN_E_UNSUPPORTED;
    if (all16 && g_bn_fast && dn_lin(dout) && (!out || dn_lin(out)) && npix < (1ll << 31))
      dn_launch(actf_bwd_kernel, g.grid, dim3(256), 0, dn_stream(stream), *dout, o, act, dbias ? ws : nullptr, g.CGb, dbias, gscale);
    else
      act_bwd_kernel<8><<<g.grid, 256, 0, dn_stream(stream)>>>(*dout, o, act, dbias ? ws : nullptr, g.CGb, dbias, gscale);
  } else {
    g = cg_geom(dout->C, 1, npix, 256, 4);
    if (g.grid.y > kWsCounters) return DN_E_UNSUPPORTED;
    act_bwd_kernel<1><<<g.grid, 256, 0, dn_stream(stream)>>>(*dout, o, act, dbias ? ws : nullptr, g.CGb, dbias, gscale);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

// ---- generic max pool (ResNet stem 3/2/1) -----------------------------------------------------------------
__global__ void maxpool_fwd_kernel(dn_view x, dn_view out, int k, int s, int p) {
  long long total = (long long)out.N * out.H * out.W * out.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % out.C);
    long long q = i / out.C;
    int w = (int)(q % out.W); q /= out.W;
    int h = (int)(q % out.H);
    int n = (int)(q / out.H);
    float best = -INFINITY;
    for (int a = 0; a < k; ++a) {
      int hi = h * s - p + a;
      if (hi < 0 || hi >= x.H) continue;
      for (int b = 0; b < k; ++b) {
        int wi = w * s - p + b;
        if (wi < 0 || wi >= x.W) continue;
        float v = dn_ld(x.ptr, x.dtype, dn_off(x, n, hi, wi) + c);
        if (v > best) best = v;
      }
    }
    dn_st(out.ptr, out.dtype, dn_off(out, n, h, w) + c, best);
  }
}

__global__ void maxpool_bwd_kernel(dn_view dout, dn_view x, dn_view dx, int k, int s, int p, int accumulate) {
  long long total = (long long)x.N * x.H * x.W * x.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % x.C);
    long long q = i / x.C;
    int w = (int)(q % x.W); q /= x.W;
    int h = (int)(q % x.H);
    int n = (int)(q / x.H);
    float g = 0.f;
    // windows (oh, ow) that contain (h, w)
    int oh_lo = (h + p - k + 1 + s - 1) / s; if (h + p - k + 1 < 0) oh_lo = 0;
    int oh_hi = (h + p) / s; if (oh_hi > dout.H - 1) oh_hi = dout.H - 1;
    int ow_lo = (w + p - k + 1 + s - 1) / s; if (w + p - k + 1 < 0) ow_lo = 0;
    int ow_hi = (w + p) / s; if (ow_hi > dout.W - 1) ow_hi = dout.W - 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        float best = -INFINITY;
        int bh = -1, bw = -1;
        for (int a = 0; a < k; ++a) {
          int hi = oh * s - p + a;
          if (hi < 0 || hi >= x.H) continue;
          for (int b = 0; b < k; ++b) {
            int wi = ow * s - p + b;
            if (wi < 0 || wi >= x.W) continue;
            float v = dn_ld(x.ptr, x.dtype, dn_off(x, n, hi, wi) + c);
            if (v > best) { best = v; bh = hi; bw = wi; }
          }
        }
        if (bh == h && bw == w) g += dn_ld(dout.ptr, dout.dtype, dn_off(dout, n, oh, ow) + c);
      }
    long long o = dn_off(dx, n, h, w) + c;
    if (accumulate) g += dn_ld(dx.ptr, dx.dtype, o);
    dn_st(dx.ptr, dx.dtype, o, g);
  }
}

static int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  long long cap = (long long)dn_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// 16-bit views, 8 channels per thread (ResNet stem pool 3/2/1 on [N, 128, 160, 64]: the scalar kernels above took 0.2 / 2.2 ms)
__global__ void __launch_bounds__(256) maxpool_fwd_vec_kernel(dn_view x, dn_view out, int k, int s, int p) {
  dn_pdl_trigger();
  dn_pdl_wait();
  const unsigned CG = (unsigned)out.C / 8;
  const unsigned total = (unsigned)out.N * out.H * out.W * CG;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned cg = i % CG;
    unsigned q = i / CG;
    const int w = (int)(q % (unsigned)out.W); q /= (unsigned)out.W;
    const int h = (int)(q % (unsigned)out.H);
    const int n = (int)(q / (unsigned)out.H);
    float best[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = -INFINITY;
    for (int a = 0; a < k; ++a) {
      const int hi = h * s - p + a;
      if (hi < 0 || hi >= x.H) continue;
      for (int b = 0; b < k; ++b) {
        const int wi = w * s - p + b;
        if (wi < 0 || wi >= x.W) continue;
        float v[8];
        ldc<8>(x, dn_off(x, n, hi, wi) + cg * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) best[e] = v[e] > best[e] ? v[e] : best[e];
      }
    }
    stc<8>(out, dn_off(out, n, h, w) + cg * 8, best);
  }
}

// gather form (deterministic): an input pixel collects dout of every window that covers it and whose FIRST maximum
// (scan order kh, kw; strict >) it is - the same rule as the forward kernel
__global__ void __launch_bounds__(256) maxpool_bwd_vec_kernel(dn_view dout, dn_view x, dn_view dx, int k, int s, int p, int accumulate) {
  dn_pdl_trigger();
  dn_pdl_wait();
  const unsigned CG = (unsigned)x.C / 8;
  const unsigned total = (unsigned)x.N * x.H * x.W * CG;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned cg = i % CG;
    unsigned q = i / CG;
    const int w = (int)(q % (unsigned)x.W); q /= (unsigned)x.W;
    const int h = (int)(q % (unsigned)x.H);
    const int n = (int)(q / (unsigned)x.H);
    float mine[8], g[8];
    ldc<8>(x, dn_off(x, n, h, w) + cg * 8, mine);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = 0.f;
    int oh_lo = (h + p - k + 1 + s - 1) / s; if (h + p - k + 1 < 0) oh_lo = 0;
    int oh_hi = (h + p) / s; if (oh_hi > dout.H - 1) oh_hi = dout.H - 1;
    int ow_lo = (w + p - k + 1 + s - 1) / s; if (w + p - k + 1 < 0) ow_lo = 0;
    int ow_hi = (w + p) / s; if (ow_hi > dout.W - 1) ow_hi = dout.W - 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        // this pixel wins channel e of the window iff no earlier element is >= it and no later element is > it
        unsigned win = 0xffu;
        const int a0 = h - (oh * s - p), b0 = w - (ow * s - p);       // my position inside the window
        for (int a = 0; a < k; ++a) {
          const int hi = oh * s - p + a;
          if (hi < 0 || hi >= x.H) continue;
          for (int b = 0; b < k; ++b) {
            const int wi = ow * s - p + b;
            if (wi < 0 || wi >= x.W || (a == a0 && b == b0)) continue;
            float v[8];
            ldc<8>(x, dn_off(x, n, hi, wi) + cg * 8, v);
            const bool earlier = a < a0 || (a == a0 && b < b0);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const bool beats = earlier ? (v[e] >= mine[e]) : (v[e] > mine[e]);
              if (beats) win &= ~(1u << e);
            }
          }
        }
        if (win) {
          float d[8];
          ldc<8>(dout, dn_off(dout, n, oh, ow) + cg * 8, d);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] += ((win >> e) & 1u) ? d[e] : 0.f;
        }
      }
    const long long o = dn_off(dx, n, h, w) + cg * 8;
    if (accumulate) {
      float old[8];
      ldc<8>(dx, o, old);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] += old[e];
    }
    stc<8>(dx, o, g);
  }
}

DN_EXPORT int dn_maxpool_fwd(const dn_view* x, const dn_view* out, int k, int stride, int pad, void* stream) {
  if (!x || !out || x->C != out->C) return DN_E_ARG;
  long long total = (long long)out->N * out->H * out->W * out->C;
  if (dn_vec8_ok(x) && dn_vec8_ok(out) && total / 8 < (1ll << 31)) {
    dn_launch(maxpool_fwd_vec_kernel, dim3(ew_blocks(total / 8)), dim3(256), 0, dn_stream(stream), *x, *out, k, stride, pad);
    DN_CHECK_LAUNCH();
    return 0;
  }
  maxpool_fwd_kernel<<<ew_blocks(total), 256, 0, dn_stream(stream)>>>(*x, *out, k, stride, pad);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_maxpool_bwd(const dn_view* dout, const dn_view* x, const dn_view* dx, int k, int stride, int pad,
                             int accumulate, void* stream) {
  if (!x || !dout || !dx) return DN_E_ARG;
  long long total = (long long)x->N * x->H * x->W * x->C;
  if (dn_vec8_ok(x) && dn_vec8_ok(dout) && dn_vec8_ok(dx) && total / 8 < (1ll << 31)) {
    dn_launch(maxpool_bwd_vec_kernel, dim3(ew_blocks(total / 8)), dim3(256), 0, dn_stream(stream), *dout, *x, *dx, k, stride, pad, accumulate);
    DN_CHECK_LAUNCH();
    return 0;
  }
  maxpool_bwd_kernel<<<ew_blocks(total), 256, 0, dn_stream(stream)>>>(*dout, *x, *dx, k, stride, pad, accumulate);
  DN_CHECK_LAUNCH();
  return 0;
}

// ---- pointwise over views ------------------------------------------------------------------------------------
// mode 0: out = act(a + b?)   mode 1: copy/accumulate   mode 2: add_act backward
template <int CH>
__global__ void __launch_bounds__(256) ew_fwd_kernel(dn_view a, dn_view b, int has_b, int act, dn_view out, int accumulate) {
  dn_pdl_trigger();
  dn_pdl_wait();
  const unsigned CG = (out.C + CH - 1) / CH;
  const unsigned total = (unsigned)out.N * out.H * out.W * CG;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned q = i / CG;
    const int c0 = (int)(i - q * CG) * CH;
    unsigned q2 = q / (unsigned)out.W;
    const int w = (int)(q - q2 * (unsigned)out.W);
    const int n = (int)(q2 / (unsigned)out.H);
    const int h = (int)(q2 - (unsigned)n * (unsigned)out.H);
    float fa[CH], fb[CH], fo[CH];
    ldc<CH>(a, dn_off(a, n, h, w) + c0, fa);
    if (has_b) ldc<CH>(b, dn_off(b, n, h, w) + c0, fb);
    long long oo = dn_off(out, n, h, w) + c0;
    if (accumulate) ldc<CH>(out, oo, fo);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      float v = dn_act(has_b ? fa[k] + fb[k] : fa[k], act);
      fo[k] = accumulate ? fo[k] + v : v;
    }
    stc<CH>(out, oo, fo);
  }
}

static int ew_launch(const dn_view* a, const dn_view* b, int act, const dn_view* out, int accumulate, void* stream) {
  if (!a || !out || a->C != out->C || a->N != out->N || a->H != out->H || a->W != out->W) return DN_E_ARG;
  bool vec = dn_vec8_any(a) && dn_vec8_any(out) && (!b || dn_vec8_any(b));
  dn_view bb = b ? *b : *a;
  if (vec) {
    long long total = (long long)out->N * out->H * out->W * (out->C / 8);
    dn_launch(ew_fwd_kernel<8>, dim3(ew_blocks(total)), dim3(256), 0, dn_stream(stream), *a, bb, b != nullptr, act, *out, accumulate);
  } else {
    long long total = (long long)out->N * out->H * out->W * out->C;
    dn_launch(ew_fwd_kernel<1>, dim3(ew_blocks(total)), dim3(256), 0, dn_stream(stream), *a, bb, b != nullptr, act, *out, accumulate);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_add_act_fwd(const dn_view* a, const dn_view* b, int act, const dn_view* out, void* stream) {
  return ew_launch(a, b, act, out, 0, stream);
}
DN_EXPORT int dn_act_fwd(const dn_view* x, int act, const dn_view* out, void* stream) {
  return ew_launch(x, nullptr, act, out, 0, stream);
}
DN_EXPORT int dn_copy_view(const dn_view* src, const dn_view* dst, int accumulate, void* stream) {
  return ew_launch(src, nullptr, DN_ACT_NONE, dst, accumulate, stream);
}

// fp32 view -> bf16 (hi, lo) planes of the same geometry: hi = bf16(x), lo = bf16(x - hi)   (precision 'tc32')
template <int CH>
__global__ void __launch_bounds__(256) split_bf16_kernel(dn_view x, dn_view hi, dn_view lo) {
  dn_pdl_trigger();
  dn_pdl_wait();
  const unsigned CG = (x.C + CH - 1) / CH;
  const unsigned total = (unsigned)x.N * x.H * x.W * CG;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned q = i / CG;
    const int c0 = (int)(i - q * CG) * CH;
    unsigned q2 = q / (unsigned)x.W;
    const int w = (int)(q - q2 * (unsigned)x.W);
    const int n = (int)(q2 / (unsigned)x.H);
    const int h = (int)(q2 - (unsigned)n * (unsigned)x.H);
    const float* xp = (const float*)x.ptr + dn_off(x, n, h, w) + c0;
    __nv_bfloat16* hp = (__nv_bfloat16*)hi.ptr + dn_off(hi, n, h, w) + c0;
    __nv_bfloat16* lp = (__nv_bfloat16*)lo.ptr + dn_off(lo, n, h, w) + c0;
    if (CH == 8) {
      const float4 a = *reinterpret_cast<const float4*>(xp), b = *reinterpret_cast<const float4*>(xp + 4);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float r[8];
      uint4 uh;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&uh);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        h2[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        const float2 f = __bfloat1622float2(h2[k]);
        r[2 * k] = v[2 * k] - f.x;
        r[2 * k + 1] = v[2 * k + 1] - f.y;
      }
      *reinterpret_cast<uint4*>(hp) = uh;
      Vec8<__nv_bfloat16>::store(lp, r);
    } else {
      const float v = xp[0];
      const __nv_bfloat16 hb = __float2bfloat16_rn(v);
      hp[0] = hb;
      lp[0] = __float2bfloat16_rn(v - __bfloat162float(hb));
    }
  }
}

DN_EXPORT int dn_split_bf16(const dn_view* x, const dn_view* hi, const dn_view* lo, void* stream) {
  if (!x || !hi || !lo || x->dtype != DN_F32 || hi->dtype != DN_BF16 || lo->dtype != DN_BF16) return DN_E_ARG;
  if (x->C != hi->C || x->N != hi->N || x->H != hi->H || x->W != hi->W) return DN_E_ARG;
  if (x->C != lo->C || x->N != lo->N || x->H != lo->H || x->W != lo->W) return DN_E_ARG;
  const bool vec = (x->C % 8) == 0 && ((uintptr_t)x->ptr % 16) == 0 && (x->sN % 4) == 0 && (x->sH % 4) == 0 && (x->sW % 4) == 0 &&
                   dn_vec8_ok(hi) && dn_vec8_ok(lo);
  if (vec) {
    const long long total = (long long)x->N * x->H * x->W * (x->C / 8);
    dn_launch(split_bf16_kernel<8>, dim3(ew_blocks(total)), dim3(256), 0, dn_stream(stream), *x, *hi, *lo);
  } else {
    const long long total = (long long)x->N * x->H * x->W * x->C;
    dn_launch(split_bf16_kernel<1>, dim3(ew_blocks(total)), dim3(256), 0, dn_stream(stream), *x, *hi, *lo);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

template <int CH>
__global__ void __launch_bounds__(256) add_act_bwd_kernel(dn_view dout, dn_view out, int act, dn_view da, int da_acc, int has_da,
                                                          dn_view db, int db_acc, int has_db) {
  const unsigned CG = (dout.C + CH - 1) / CH;
  const unsigned total = (unsigned)dout.N * dout.H * dout.W * CG;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned q = i / CG;
    const int c0 = (int)(i - q * CG) * CH;
    unsigned q2 = q / (unsigned)dout.W;
    const int w = (int)(q - q2 * (unsigned)dout.W);
    const int n = (int)(q2 / (unsigned)dout.H);
    const int h = (int)(q2 - (unsigned)n * (unsigned)dout.H);
    float g[CH], o[CH], t[CH];
    ldc<CH>(dout, dn_off(dout, n, h, w) + c0, g);
    if (act != DN_ACT_NONE) {
      ldc<CH>(out, dn_off(out, n, h, w) + c0, o);
#pragma unroll
      for (int k = 0; k < CH; ++k) g[k] *= dn_act_grad(o[k], act);
    }
    if (has_da) {
      long long off = dn_off(da, n, h, w) + c0;
      if (da_acc) {
        ldc<CH>(da, off, t);
#pragma unroll
        for (int k = 0; k < CH; ++k) t[k] += g[k];
        stc<CH>(da, off, t);
      } else stc<CH>(da, off, g);
    }
    if (has_db) {
      long long off = dn_off(db, n, h, w) + c0;
      if (db_acc) {
        ldc<CH>(db, off, t);
#pragma unroll
        for (int k = 0; k < CH; ++k) t[k] += g[k];
        stc<CH>(db, off, t);
      } else stc<CH>(db, off, g);
    }
  }
}

DN_EXPORT int dn_add_act_bwd(const dn_view* dout, const dn_view* out, int act, const dn_view* da, int da_acc,
                             const dn_view* db, int db_acc, void* stream) {
  if (!dout || (!out && act != DN_ACT_NONE)) return DN_E_ARG;
  dn_view o = out ? *out : *dout;
  dn_view a = da ? *da : *dout, b = db ? *db : *dout;
  bool vec = dn_vec8_any(dout) && (!out || dn_vec8_any(out)) && (!da || dn_vec8_any(da)) && (!db || dn_vec8_any(db));
  if (vec) {
    long long total = (long long)dout->N * dout->H * dout->W * (dout->C / 8);
    add_act_bwd_kernel<8><<<ew_blocks(total), 256, 0, dn_stream(stream)>>>(*dout, o, act, a, da_acc, da != nullptr, b, db_acc, db != nullptr);
  } else {
    long long total = (long long)dout->N * dout->H * dout->W * dout->C;
    add_act_bwd_kernel<1><<<ew_blocks(total), 256, 0, dn_stream(stream)>>>(*dout, o, act, a, da_acc, da != nullptr, b, db_acc, db != nullptr);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

// ---- disparity heads ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dn_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void head_fwd_kernel(dn_view z, float alpha, float beta, float* __restrict__ disp) {
  dn_pdl_trigger();
  dn_pdl_wait();
  long long total = (long long)z.N * z.H * z.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % z.W);
    long long q = i / z.W;
    int h = (int)(q % z.H), n = (int)(q / z.H);
    disp[i] = alpha * dn_sigmoid(dn_ld(z.ptr, z.dtype, dn_off(z, n, h, w))) + beta;
  }
}

// bilinear x2, align_corners=False source index (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void bil_src(int o, int insz, int& i0, int& i1, float& l0, float& l1) {
  float s = 0.5f * ((float)o + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > insz - 1) i0 = insz - 1;
  i1 = i0 + ((i0 < insz - 1) ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__global__ void head_up_kernel(const float* __restrict__ disp, int N, int H, int W, dn_view up, int mode) {
  dn_pdl_trigger();
  dn_pdl_wait();
  long long total = (long long)up.N * up.H * up.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % up.W);
    long long q = i / up.W;
    int y = (int)(q % up.H), n = (int)(q / up.H);
    const float* d = disp + (long long)n * H * W;
    float v;
    if (mode == 0) {
      v = d[(y >> 1) * W + (x >> 1)];
    } else {
      int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
      bil_src(y, H, y0, y1, ly0, ly1);
      bil_src(x, W, x0, x1, lx0, lx1);
      v = ly0 * (lx0 * d[y0 * W + x0] + lx1 * d[y0 * W + x1]) + ly1 * (lx0 * d[y1 * W + x0] + lx1 * d[y1 * W + x1]);
    }
    dn_st(up.ptr, up.dtype, dn_off(up, n, y, x), v);
  }
}

DN_EXPORT int dn_head_fwd(const dn_view* z, float alpha, float beta, float* disp, const dn_view* up, int up_mode, void* stream) {
  if (!z || !disp) return DN_E_ARG;
  long long total = (long long)z->N * z->H * z->W;
  dn_launch(head_fwd_kernel, dim3(ew_blocks(total)), dim3(256), 0, dn_stream(stream), *z, alpha, beta, disp);
  DN_CHECK_LAUNCH();
  if (up) {
    if (up->H > 2 * z->H || up->W > 2 * z->W || up->N != z->N) return DN_E_ARG;
    long long t2 = (long long)up->N * up->H * up->W;
    dn_launch(head_up_kernel, dim3(ew_blocks(t2)), dim3(256), 0, dn_stream(stream), disp, z->N, z->H, z->W, *up, up_mode);
    DN_CHECK_LAUNCH();
  }
  return 0;
}

// disparity + nearest x2 up-sampled copies (activation dtype and, optionally, its gradient-dtype twin) in one pass
// 16-bit value followed by 15 zeros as one 32-byte (full DRAM sector) store
__device__ __forceinline__ void st_sector16(void* p, int dtype, float v) {
  uint4 a = make_uint4(0, 0, 0, 0);
  if (dtype == DN_F16) a.x = (uint32_t)__half_as_ushort(__float2half_rn(v));
  else a.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
  reinterpret_cast<uint4*>(p)[0] = a;
  reinterpret_cast<uint4*>(p)[1] = make_uint4(0, 0, 0, 0);
}

// `sector`: the up-sampled slot is the LAST channel slice of its concatenation buffer and at least 15 zero padding channels follow
// it inside the 32-byte aligned pixel record -- the slot is then written as a whole 32-byte sector (value + zeros) instead of a
// 2-byte piece of one, which spares DRAM the read-modify-write of 1.7 M partially written sectors
__global__ void head_fwd_nearest_kernel(dn_view z, float alpha, float beta, float* __restrict__ disp, dn_view up, dn_view up2, int has_up2,
                                        int sector) {
  dn_pdl_trigger();
  dn_pdl_wait();
  long long total = (long long)z.N * z.H * z.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % z.W);
    long long q = i / z.W;
    int h = (int)(q % z.H), n = (int)(q / z.H);
    const float v = alpha * dn_sigmoid(dn_ld(z.ptr, z.dtype, dn_off(z, n, h, w))) + beta;
    disp[i] = v;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int y = 2 * h + a, x = 2 * w + b;
        if (y < up.H && x < up.W) {
          if (sector) {
            st_sector16((char*)up.ptr + dn_off(up, n, y, x) * 2, up.dtype, v);
            if (has_up2) st_sector16((char*)up2.ptr + dn_off(up2, n, y, x) * 2, up2.dtype, v);
          } else {
            dn_st(up.ptr, up.dtype, dn_off(up, n, y, x), v);
            if (has_up2) dn_st(up2.ptr, up2.dtype, dn_off(up2, n, y, x), v);
          }
        }
      }
  }
}

/* dn_head_fwd with a second up-sampled copy (`up2`, same geometry as `up`, another dtype): saves the separate copy pass */
DN_EXPORT int dn_head_fwd2(const dn_view* z, float alpha, float beta, float* disp, const dn_view* up, const dn_view* up2, int up_mode,
                           void* stream) {
  if (!z || !disp) return DN_E_ARG;
  // up_mode bit 4 (16): everything behind `up` (and `up2`) inside the pixel record is zero padding that may be rewritten with zeros
  const int tail_is_pad = up_mode & 16;
  up_mode &= 15;
  if (up && up_mode == 0 && up->H <= 2 * z->H && up->W <= 2 * z->W && up->N == z->N &&
      (!up2 || (up2->H == up->H && up2->W == up->W && up2->N == up->N))) {
    long long total = (long long)z->N * z->H * z->W;
    static const bool g_sector = []() { const char* e = getenv("DN_HEAD_SECTOR"); return !(e && e[0] == '0'); }();
    auto sector_ok = [](const dn_view* v) {
      return v->dtype != DN_F32 && v->C == 1 && v->c_ext >= 16 && ((uintptr_t)v->ptr % 32) == 0 && (v->sW % 16) == 0 && (v->sH % 16) == 0 &&
             (v->sN % 16) == 0;
    };
    const int sector = (g_sector && tail_is_pad && sector_ok(up) && (!up2 || sector_ok(up2))) ? 1 : 0;
    dn_launch(head_fwd_nearest_kernel, dim3(ew_blocks(total)), dim3(256), 0, dn_stream(stream), *z, alpha, beta, disp, *up, up2 ? *up2 : *up,
              up2 != nullptr, sector);
    DN_CHECK_LAUNCH();
    return 0;
  }
  int e = dn_head_fwd(z, alpha, beta, disp, up, up_mode, stream);
  if (e) return e;
  if (up && up2) return dn_copy_view(up, up2, 0, stream);
  return 0;
}

__global__ void head_bwd_kernel(const float* __restrict__ gdisp, dn_view dup, int has_up, int mode, dn_view z, float alpha,
                                float gscale, dn_view dz) {
  dn_pdl_trigger();
  dn_pdl_wait();
  long long total = (long long)z.N * z.H * z.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % z.W);
    long long q = i / z.W;
    int h = (int)(q % z.H), n = (int)(q / z.H);
    float g = gdisp ? gdisp[i] * gscale : 0.f;   // gdisp is unscaled fp32 from autograd; dup is already scaled
    if (has_up) {
      if (mode == 0) {
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            int y = 2 * h + a, x = 2 * w + b;
            if (y < dup.H && x < dup.W) g += dn_ld(dup.ptr, dup.dtype, dn_off(dup, n, y, x));
          }
      } else {
        for (int y = 2 * h - 2; y <= 2 * h + 2; ++y) {
          if (y < 0 || y >= dup.H) continue;
          int y0, y1; float ly0, ly1;
          bil_src(y, z.H, y0, y1, ly0, ly1);
          float wy = (y0 == h ? ly0 : 0.f) + (y1 == h ? ly1 : 0.f);
          if (wy == 0.f) continue;
          for (int x = 2 * w - 2; x <= 2 * w + 2; ++x) {
            if (x < 0 || x >= dup.W) continue;
            int x0, x1; float lx0, lx1;
            bil_src(x, z.W, x0, x1, lx0, lx1);
            float wx = (x0 == w ? lx0 : 0.f) + (x1 == w ? lx1 : 0.f);
            if (wx == 0.f) continue;
            g += wy * wx * dn_ld(dup.ptr, dup.dtype, dn_off(dup, n, y, x));
          }
        }
      }
    }
    float s = dn_sigmoid(dn_ld(z.ptr, z.dtype, dn_off(z, n, h, w)));
    dn_st(dz.ptr, dz.dtype, dn_off(dz, n, h, w), g * alpha * s * (1.f - s));
  }
}

DN_EXPORT int dn_head_bwd(const float* gdisp, const dn_view* dup, int up_mode, const dn_view* z, float alpha, float gscale,
                          const dn_view* dz, void* stream) {
  if (!z || !dz) return DN_E_ARG;
  long long total = (long long)z->N * z->H * z->W;
  dn_view d = dup ? *dup : *z;
  dn_launch(head_bwd_kernel, dim3(ew_blocks(total)), dim3(256), 0, dn_stream(stream), gdisp, d, dup != nullptr, up_mode, *z, alpha, gscale, *dz);
  DN_CHECK_LAUNCH();
  return 0;
}

__global__ void sigmoid_nchw_fwd_kernel(dn_view z, float* __restrict__ out) {
  long long total = (long long)z.N * z.C * z.H * z.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % z.W);
    long long q = i / z.W;
    int h = (int)(q % z.H); q /= z.H;
    int c = (int)(q % z.C), n = (int)(q / z.C);
    out[i] = dn_sigmoid(dn_ld(z.ptr, z.dtype, dn_off(z, n, h, w) + c));
  }
}
__global__ void sigmoid_nchw_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out, float gscale, dn_view dz) {
  long long total = (long long)dz.N * dz.C * dz.H * dz.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % dz.W);
    long long q = i / dz.W;
    int h = (int)(q % dz.H); q /= dz.H;
    int c = (int)(q % dz.C), n = (int)(q / dz.C);
    float s = out[i];
    dn_st(dz.ptr, dz.dtype, dn_off(dz, n, h, w) + c, (gout ? gout[i] : 0.f) * gscale * s * (1.f - s));
  }
}
DN_EXPORT int dn_sigmoid_nchw_fwd(const dn_view* z, float* out, void* stream) {
  if (!z || !out) return DN_E_ARG;
  long long total = (long long)z->N * z->C * z->H * z->W;
  sigmoid_nchw_fwd_kernel<<<ew_blocks(total), 256, 0, dn_stream(stream)>>>(*z, out);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_sigmoid_nchw_bwd(const float* gout, const float* out, float gscale, const dn_view* dz, void* stream) {
  if (!dz || !out) return DN_E_ARG;
  long long total = (long long)dz->N * dz->C * dz->H * dz->W;
  sigmoid_nchw_bwd_kernel<<<ew_blocks(total), 256, 0, dn_stream(stream)>>>(gout, out, gscale, *dz);
  DN_CHECK_LAUNCH();
  return 0;
}

__global__ void spatial_mean_fwd_kernel(dn_view z, float scale, float* __restrict__ out) {
  int n = blockIdx.x / z.C, c = blockIdx.x % z.C;
  float s = 0.f;
  int HW = z.H * z.W;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s += dn_ld(z.ptr, z.dtype, dn_off(z, n, i / z.W, i % z.W) + c);
  __shared__ float red[32];
  s = dn_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    out[blockIdx.x] = scale * (t / (float)HW);
  }
}
__global__ void spatial_mean_bwd_kernel(const float* __restrict__ gout, float scale, dn_view dz) {
  long long total = (long long)dz.N * dz.H * dz.W * dz.C;
  float inv = scale / (float)(dz.H * dz.W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dz.C);
    long long q = i / dz.C;
    int w = (int)(q % dz.W); q /= dz.W;
    int h = (int)(q % dz.H);
    int n = (int)(q / dz.H);
    dn_st(dz.ptr, dz.dtype, dn_off(dz, n, h, w) + c, gout[n * dz.C + c] * inv);
  }
}
DN_EXPORT int dn_spatial_mean_fwd(const dn_view* z, float scale, float* out, void* stream) {
  if (!z || !out) return DN_E_ARG;
  spatial_mean_fwd_kernel<<<z->N * z->C, 128, 0, dn_stream(stream)>>>(*z, scale, out);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_spatial_mean_bwd(const float* gout, float scale, const dn_view* dz, void* stream) {
  if (!dz || !gout) return DN_E_ARG;
  long long total = (long long)dz->N * dz->H * dz->W * dz->C;
  spatial_mean_bwd_kernel<<<ew_blocks(total), 256, 0, dn_stream(stream)>>>(gout, scale, *dz);
  DN_CHECK_LAUNCH();
  return 0;
}

// ---- misc -------------------------------------------------------------------------------------------------------------
__global__ void fill_kernel(float* p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void axpy_kernel(const float* __restrict__ x, float a, float* y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a * x[i];
}
DN_EXPORT int dn_fill_f32(float* p, int64_t n, float v, void* stream) {
  if (n <= 0) return 0;
  fill_kernel<<<ew_blocks(n), 256, 0, dn_stream(stream)>>>(p, n, v);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_axpy_f32(const float* x, float a, float* y, int64_t n, void* stream) {
  if (n <= 0) return 0;
  axpy_kernel<<<ew_blocks(n), 256, 0, dn_stream(stream)>>>(x, a, y, n);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_version(void) { return 100; }
DN_EXPORT const char* dn_error_string(int code) {
  if (code == 0) return "ok";
  if (code == DN_E_ARG) return "dispnet_b200: invalid argument";
  if (code == DN_E_UNSUPPORTED) return "dispnet_b200: problem not supported by the requested backend";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "dispnet_b200: unknown error";
}

// =================================================================================================
// disparity-head convolution: nn.Conv2d(C, 1, 3, padding=1) (models/Disp_vgg_BN.py:66-70) on CUDA cores
//
// One output channel makes this a 9*C-term dot product per pixel: HBM-bound (read x once), a poor fit for a 128-wide MMA
// tile.  Forward: thread per pixel.  Backward: thread per (pixel, 8-channel group); one pass produces the data gradient
// (accumulated into the gradient arena of x), the weight gradient (per-block partial sums, reduced by a second stage) and
// the bias gradient.
// =================================================================================================
template <int CH>
__global__ void __launch_bounds__(256) head_conv_fwd_kernel(dn_view x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            dn_view z) {
  dn_pdl_trigger();
  dn_pdl_wait();
  extern __shared__ float ws[];          // [9][C]
  const int C = x.C;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    ws[i] = w[c * 9 + t];                 // torch layout [1][C][3][3]
  }
  __syncthreads();
  // 16x16-pixel tiles: the 3x3 neighbourhoods of a block's pixels overlap inside its own L1
  const unsigned tilesW = (x.W + 15) / 16, tilesH = (x.H + 15) / 16;
  const unsigned ntiles = tilesW * tilesH * (unsigned)x.N;
  const float b0 = bias ? bias[0] : 0.f;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const unsigned tq = tile / tilesW;
    const int wv = (int)(tile - tq * tilesW) * 16 + (int)(threadIdx.x & 15);
    const int n = (int)(tq / tilesH);
    const int h = (int)(tq - (unsigned)n * tilesH) * 16 + (int)(threadIdx.x >> 4);
    if (wv >= x.W || h >= x.H) continue;
    float acc = b0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = wv + t % 3 - 1;
      if (hh < 0 || hh >= x.H || ww < 0 || ww >= x.W) continue;
      const long long off = dn_off(x, n, hh, ww);
      for (int c0 = 0; c0 < C; c0 += CH) {
        float f[CH];
        ldc<CH>(x, off + c0, f);
#pragma unroll
        for (int i = 0; i < CH; ++i) acc = fmaf(f[i], ws[t * C + c0 + i], acc);
      }
    }
    dn_st(z.ptr, z.dtype, dn_off(z, n, h, wv), acc);
  }
}

template <int CH>
__global__ void __launch_bounds__(256, 2) head_conv_bwd_kernel(dn_view x, const float* __restrict__ w, dn_view dz, dn_view gx,
                                                            int gx_acc, float* __restrict__ wsp, int CGb) {
  dn_pdl_trigger();
  dn_pdl_wait();
  extern __shared__ float ws[];          // [9][C] weights, then per-block accumulators [9*C + 1]
  const int C = x.C;
  float* accs = ws + 9 * C;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    ws[i] = w[c * 9 + t];
  }
  for (int i = threadIdx.x; i < 9 * C + 1; i += blockDim.x) accs[i] = 0.f;
  __syncthreads();
  CG_PROLOGUE(x)
  float acc[9][CH];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[t][i] = 0.f;
  float accb = 0.f;
  if (cvalid) {
    const unsigned tw = PLn >= 16 ? 16 : PLn, th = PLn >= 16 ? PLn / 16 : 1;      // tile of pixel lanes (PLn is a power of two)
    const unsigned tilesW = (x.W + tw - 1) / tw, tilesH = (x.H + th - 1) / th;
    const unsigned ntiles = tilesW * tilesH * (unsigned)x.N;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const unsigned tq = tile / tilesW;
      const int wv = (int)((tile - tq * tilesW) * tw + ((unsigned)pl % tw));
      const int n = (int)(tq / tilesH);
      const int h = (int)((tq - (unsigned)n * tilesH) * th + ((unsigned)pl / tw));
      if (wv >= x.W || h >= x.H) continue;
      float xv[CH], g[CH];
      ldc<CH>(x, dn_off(x, n, h, wv) + c0, xv);
#pragma unroll
      for (int i = 0; i < CH; ++i) g[i] = 0.f;
      // forward: z[o] = sum_t x[o + d_t] w[t]  =>  this pixel p = o + d_t feeds output o = p - d_t through tap t
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = h - (t / 3 - 1), ww = wv - (t % 3 - 1);
        float d = 0.f;
        if (hh >= 0 && hh < x.H && ww >= 0 && ww < x.W) d = dn_ld(dz.ptr, dz.dtype, dn_off(dz, n, hh, ww));
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          acc[t][i] = fmaf(d, xv[i], acc[t][i]);
          g[i] = fmaf(d, ws[t * C + (c0 + i < C ? c0 + i : C - 1)], g[i]);
        }
        if (t == 4 && cg == 0) accb += d;
      }
      const long long go = dn_off(gx, n, h, wv) + c0;
      if (gx_acc) {
        float o[CH];
        ldc<CH>(gx, go, o);
#pragma unroll
        for (int i = 0; i < CH; ++i) g[i] += o[i];
      }
      stc<CH>(gx, go, g);
    }
  }
  // reduce over the lanes of a warp that own the same channel group (lane bits >= log2(CGb)), then shared atomics
  for (int o = 16; o >= CGb && o >= 1; o >>= 1) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[t][i] += __shfl_xor_sync(0xffffffffu, acc[t][i], o);
    accb += __shfl_xor_sync(0xffffffffu, accb, o);
  }
  const int lane = threadIdx.x & 31;
  if (cvalid && (CGb >= 32 || lane < CGb)) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (c0 + i < C) atomicAdd(&accs[(c0 + i) * 9 + t], acc[t][i]);     // torch layout [C][9]
    if (cg == 0) atomicAdd(&accs[9 * C], accb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * C + 1; i += blockDim.x) wsp[(long long)blockIdx.x * (9 * C + 1) + i] = accs[i];
}

#include "dn_head_mma.cuh"

// DN_HEAD_MMA=0 keeps the CUDA-core kernels (A/B comparisons)
static const bool g_head_mma = []() { const char* e = getenv("DN_HEAD_MMA"); return !(e && e[0] == '0'); }();
static const bool g_head_bulk = []() { const char* e = getenv("DN_HEAD_BULK"); return !(e && e[0] == '0'); }();

template <typename K>
static int hc_set_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

DN_EXPORT int dn_head_conv_fwd(const dn_view* x, const float* w, const float* bias, const dn_view* z, void* stream) {
  if (!x || !w || !z || z->C != 1 || z->N != x->N || z->H != x->H || z->W != x->W) return DN_E_ARG;
  if (g_head_mma && hc::eligible(x) && x->C <= 32 && x->sW == x->C && g_head_bulk) {
    // rows of whole pixel vectors are contiguous: bulk-copy, double-buffered variant
    constexpr int S = 4;                  // C = 16: 4 x 10.6 KB, C = 32: 4 x 21.3 KB per block
    const size_t sm = hc::fwd_bulk_smem(x->C, S);
    const int ntiles = ((x->W + hc::TW - 1) / hc::TW) * ((x->H + hc::TH - 1) / hc::TH) * x->N;
    int per_sm = (int)((220 * 1024) / (sm + 1024));
    if (per_sm > 4) per_sm = 4;
    int grid = dn_num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    int e = 0;
#define HC_FWD_BULK(BF, CC)                                                     \
  do {                                                                          \
    if ((e = hc_set_smem(hc::fwd_bulk_kernel<BF, CC, S>, sm))) return e;        \
    dn_launch(hc::fwd_bulk_kernel<BF, CC, S>, dim3(grid), dim3(256), sm, dn_stream(stream), *x, w, bias, *z); \
  } while (0)
    if (x->dtype == DN_BF16) { if (x->C == 16) HC_FWD_BULK(true, 16); else HC_FWD_BULK(true, 32); }
    else { if (x->C == 16) HC_FWD_BULK(false, 16); else HC_FWD_BULK(false, 32); }
#undef HC_FWD_BULK
    DN_CHECK_LAUNCH();
    return 0;
  }
  if (g_head_mma && hc::eligible(x)) {
    const size_t sm = hc::fwd_smem(x->C);
    const int ntiles = ((x->W + hc::TW - 1) / hc::TW) * ((x->H + hc::TH - 1) / hc::TH) * x->N;
    int per_sm = (int)((200 * 1024) / (sm + 1024));
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    int grid = dn_num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    int e;
    if (x->dtype == DN_BF16) {
      if ((e = hc_set_smem(hc::fwd_kernel<true>, sm))) return e;
      dn_launch(hc::fwd_kernel<true>, dim3(grid), dim3(256), sm, dn_stream(stream), *x, w, bias, *z);
    } else {
      if ((e = hc_set_smem(hc::fwd_kernel<false>, sm))) return e;
      dn_launch(hc::fwd_kernel<false>, dim3(grid), dim3(256), sm, dn_stream(stream), *x, w, bias, *z);
    }
    DN_CHECK_LAUNCH();
    return 0;
  }
  const long long npix = (long long)x->N * ((x->H + 15) / 16) * ((x->W + 15) / 16) * 256;
  const size_t sm = sizeof(float) * 9 * x->C;
  int blocks = ew_blocks(npix);
  if (dn_vec8_any(x)) dn_launch(head_conv_fwd_kernel<8>, dim3(blocks), dim3(256), sm, dn_stream(stream), *x, w, bias, *z);
  else dn_launch(head_conv_fwd_kernel<1>, dim3(blocks), dim3(256), sm, dn_stream(stream), *x, w, bias, *z);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_head_conv_bwd(const dn_view* x, const float* w, const dn_view* dz, const dn_view* gx, int gx_accumulate,
                               float* gw, float* gb, float gscale, float* ws, void* stream) {
  if (!x || !w || !dz || !gx || !gw || !ws || gx->C != x->C) return DN_E_ARG;
  const long long npix = (long long)x->N * x->H * x->W;
  const bool vec = dn_vec8_any(x) && dn_vec8_any(gx);
  const int ch = vec ? 8 : 1;
  CgGeom g = cg_geom(x->C, ch, npix);
  if (g.grid.y != 1) return DN_E_UNSUPPORTED;
  const int n = 9 * x->C + 1;
  float* rows = ws + kWsCounters;        // the head of the workspace holds the arrival counters of the single-launch reductions
  const long long cap = dn_reduce_ws_floats(x->C) - kWsCounters;
  if ((long long)g.grid.x * n > cap) g.grid.x = (unsigned)(cap / n);
  const size_t sm = sizeof(float) * (18 * x->C + 1);
  cudaStream_t st = dn_stream(stream);
  if (g_head_mma && vec && hc::eligible(x) && (gx->dtype == DN_BF16 || gx->dtype == DN_F16)) {
    const size_t smm = hc::bwd_smem(x->C);
    const int ntiles = ((x->W + hc::TW - 1) / hc::TW) * ((x->H + hc::TH - 1) / hc::TH) * x->N;
    int per_sm = x->C >= 128 ? 1 : 2;
    long long grid = (long long)dn_num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (grid * n > cap) grid = cap / n;
    g.grid = dim3((unsigned)grid);
    int e = 0;
    const bool bulk = g_head_bulk && x->C <= 32 && x->sW == x->C;
    constexpr int SB = 4;
    const size_t smb = bulk ? hc::bwd_bulk_smem(x->C, SB) : 0;
    switch (bulk ? -x->C : x->C) {
      case -16: if ((e = hc_set_smem(hc::bwd_bulk_kernel<2, SB>, smb))) return e; dn_launch(hc::bwd_bulk_kernel<2, SB>, g.grid, dim3(256), smb, st, *x, w, *dz, *gx, gx_accumulate, rows); break;
      case -32: if ((e = hc_set_smem(hc::bwd_bulk_kernel<4, SB>, smb))) return e; dn_launch(hc::bwd_bulk_kernel<4, SB>, g.grid, dim3(256), smb, st, *x, w, *dz, *gx, gx_accumulate, rows); break;
      case 16: if ((e = hc_set_smem(hc::bwd_kernel<2>, smm))) return e; dn_launch(hc::bwd_kernel<2>, g.grid, dim3(256), smm, st, *x, w, *dz, *gx, gx_accumulate, rows); break;
      case 32: if ((e = hc_set_smem(hc::bwd_kernel<4>, smm))) return e; dn_launch(hc::bwd_kernel<4>, g.grid, dim3(256), smm, st, *x, w, *dz, *gx, gx_accumulate, rows); break;
      case 64: if ((e = hc_set_smem(hc::bwd_kernel<8>, smm))) return e; dn_launch(hc::bwd_kernel<8>, g.grid, dim3(256), smm, st, *x, w, *dz, *gx, gx_accumulate, rows); break;
      default: if ((e = hc_set_smem(hc::bwd_kernel<16>, smm))) return e; dn_launch(hc::bwd_kernel<16>, g.grid, dim3(256), smm, st, *x, w, *dz, *gx, gx_accumulate, rows); break;
    }
  } else if (vec) dn_launch(head_conv_bwd_kernel<8>, dim3(g.grid), dim3(256), sm, st, *x, w, *dz, *gx, gx_accumulate, rows, g.CGb);
  else dn_launch(head_conv_bwd_kernel<1>, dim3(g.grid), dim3(256), sm, st, *x, w, *dz, *gx, gx_accumulate, rows, g.CGb);
  DN_CHECK_LAUNCH();
  // second stage: [blocks][9C+1] -> gw (torch layout [1][C][3][3]) and gb; both scaled by gscale
  dn_launch(reduce_partials_kernel<float>, dim3(reduce_blocks(n - 1)), dim3(1024), 0, st, rows, g.grid.x, n, gw, (double)gscale, n - 1);
  DN_CHECK_LAUNCH();
  if (gb) {
    dn_launch(reduce_partials_kernel<float>, dim3(1), dim3(1024), 0, st, rows + (n - 1), g.grid.x, n, gb, (double)gscale, 1);
    DN_CHECK_LAUNCH();
  }
  return 0;
}
