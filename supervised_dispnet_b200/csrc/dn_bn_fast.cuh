// BatchNorm2d(train) + ReLU (+ MaxPool2d(2,2)) fast paths  (reference call sites: models/Disp_vgg_BN.py:137-141 and the
// torchvision vgg16_bn feature stack they wrap).  Included by dn_layers.cu after the generic kernels, whose results these
// reproduce bit for bit.
//
// The generic CG walkers spend ~80 issue slots per 16-byte load (two 32-bit divisions to turn a pixel index into (n, h, w),
// run-time activation switches per element, scalar fp32 math); at 23 B/clk/SM of HBM bandwidth that is an issue-bound
// kernel, not a memory-bound one (measured: 2.2-3.5 TB/s of 6.5).  The fast paths apply when every view is a 16-bit,
// 8-channel-vector view whose pixels are linearly addressable (off = pixel * sW - true for whole buffers and for channel
// slices of concat buffers, false for crops) and there is no residual operand:
//   * no divisions (one per pooled output pixel in the pooled variants),
//   * activation handled once per 8-channel vector with a warp-uniform branch,
//   * packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2; same rounding as the scalar ops),
//   * 4 pixels (8 x 16 bytes) in flight per thread in the reduction kernels.
#pragma once

static inline bool dn_lin(const dn_view* v) {
  return v->sH == (long long)v->W * v->sW && v->sN == (long long)v->H * v->sH;
}

__device__ __forceinline__ void cvt8(const uint4& u, int dtype, float2* f) {
  if (dtype == DN_F16) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = __half22float2(h[i]);
  } else {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = __bfloat1622float2(h[i]);
  }
}
__device__ __forceinline__ uint4 pack8(const float2* f, int dtype) {
  uint4 u;
  if (dtype == DN_F16) {
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[i].x, f[i].y);
  } else {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[i].x, f[i].y);
  }
  return u;
}
__device__ __forceinline__ uint4 ld16(const char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void st16(char* p, const uint4& u) { *reinterpret_cast<uint4*>(p) = u; }

// v = act(v) for one 8-channel vector
__device__ __forceinline__ void act8(float2* v, int act) {
  if (act == DN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); }
  } else if (act == DN_ACT_LRELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i].x = v[i].x > 0.f ? v[i].x : 0.1f * v[i].x; v[i].y = v[i].y > 0.f ? v[i].y : 0.1f * v[i].y; }
  }
}
// g *= act'(v), v = pre-activation (sign-preserving activations: act(v) > 0 <=> v > 0)
__device__ __forceinline__ void act_grad8(float2* g, const float2* v, int act) {
  if (act == DN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { g[i].x = v[i].x > 0.f ? g[i].x : 0.f; g[i].y = v[i].y > 0.f ? g[i].y : 0.f; }
  } else if (act == DN_ACT_LRELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { g[i].x = v[i].x > 0.f ? g[i].x : 0.1f * g[i].x; g[i].y = v[i].y > 0.f ? g[i].y : 0.1f * g[i].y; }
  }
}

__device__ __forceinline__ void load_sc_sh(const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                                           const float* __restrict__ beta, int C, int c0, float2* sc, float2* sh) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b, c, d;
    bn_scale_shift(gamma ? gamma[c0 + 2 * i] : 1.f, beta ? beta[c0 + 2 * i] : 0.f, mean_invstd[c0 + 2 * i], mean_invstd[C + c0 + 2 * i], a, b);
    bn_scale_shift(gamma ? gamma[c0 + 2 * i + 1] : 1.f, beta ? beta[c0 + 2 * i + 1] : 0.f, mean_invstd[c0 + 2 * i + 1],
                   mean_invstd[C + c0 + 2 * i + 1], c, d);
    sc[i] = make_float2(a, c);
    sh[i] = make_float2(b, d);
  }
}

#define BNF_PROLOGUE(view_for_dims)                                                         \
  const int cgl = threadIdx.x % CGb;                                                        \
  const int pl = threadIdx.x / CGb;                                                         \
  const int PLn = 256 / CGb;                                                                \
  const int c0 = (blockIdx.y * CGb + cgl) * 8;                                              \
  const int C = (view_for_dims).C;                                                          \
  const bool cvalid = c0 < C;                                                               \
  const unsigned npix = (unsigned)(view_for_dims).N * (view_for_dims).H * (view_for_dims).W; \
  const unsigned stride = gridDim.x * PLn;                                                  \
  dn_pdl_trigger();                                                                         \
  dn_pdl_wait();

// ---- statistics ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) bnf_stats_kernel(dn_view y, float* __restrict__ ws, int CGb, double* __restrict__ sums,
                                                           BnFinalize fz) {
  BNF_PROLOGUE(y)
  float2 s[4], q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { s[i] = make_float2(0.f, 0.f); q[i] = make_float2(0.f, 0.f); }
  if (cvalid) {
    const char* base = (const char*)y.ptr + (long long)c0 * 2;
    const long long pitch = y.sW * 2;
    unsigned px = blockIdx.x * PLn + pl;
    for (; px + 7 * stride < npix && px + 7 * stride >= px; px += 8 * stride) {      // 8 x 16 bytes in flight per thread
      uint4 r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = ld16(base + (long long)(px + u * stride) * pitch);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float2 f[4];
        cvt8(r[u], y.dtype, f);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i] = __fadd2_rn(s[i], f[i]); q[i] = __ffma2_rn(f[i], f[i], q[i]); }
      }
    }
    for (; px < npix; px += stride) {
      float2 f[4];
      cvt8(ld16(base + (long long)px * pitch), y.dtype, f);
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[i] = __fadd2_rn(s[i], f[i]); q[i] = __ffma2_rn(f[i], f[i], q[i]); }
    }
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[2 * i] = s[i].x; acc[2 * i + 1] = s[i].y; acc[8 + 2 * i] = q[i].x; acc[8 + 2 * i + 1] = q[i].y; }
  bn_stats_tail<8>(acc, C, (double)npix, ws, CGb, sums, fz, cvalid, c0);
}

// ---- apply: out = pool?(act(y * sc + sh)), optional second copy in another 16-bit type ------------------------------
template <bool POOL, bool RES>      // RES: out = act(y * sc + sh + residual) (ResNet bottleneck tails; never together with POOL)
__global__ void __launch_bounds__(256, 4) bnf_apply_kernel(dn_view y, const float* __restrict__ scale_shift, int act, dn_view out,
                                                           dn_view out2, int has_out2, int CGb, dn_view res) {
  BNF_PROLOGUE(out)
  if (!cvalid) return;
  float2 sc[4], sh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sc[i] = make_float2(scale_shift[c0 + 2 * i], scale_shift[c0 + 2 * i + 1]);
    sh[i] = make_float2(scale_shift[C + c0 + 2 * i], scale_shift[C + c0 + 2 * i + 1]);
  }
  const char* yb = (const char*)y.ptr + (long long)c0 * 2;
  char* ob = (char*)out.ptr + (long long)c0 * 2;
  char* o2b = (char*)out2.ptr + (long long)c0 * 2;
  const long long yp = y.sW * 2, op = out.sW * 2, o2p = out2.sW * 2;
  const char* rb = (const char*)res.ptr + (long long)c0 * 2;
  const long long rp = res.sW * 2;
  if (!POOL) {
    unsigned px = blockIdx.x * PLn + pl;
    for (; px < npix; px += 2 * stride) {
      const bool ok1 = px + stride < npix && px + stride > px;
      const unsigned px1 = ok1 ? px + stride : px;
      const uint4 r0 = ld16(yb + (long long)px * yp), r1 = ld16(yb + (long long)px1 * yp);
      uint4 q0 = r0, q1 = r1;
      if (RES) { q0 = ld16(rb + (long long)px * rp); q1 = ld16(rb + (long long)px1 * rp); }
      float2 f[4];
      cvt8(r0, y.dtype, f);
#pragma unroll
      for (int i = 0; i < 4; ++i) f[i] = __ffma2_rn(f[i], sc[i], sh[i]);
      if (RES) {
        float2 r[4];
        cvt8(q0, res.dtype, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = __fadd2_rn(f[i], r[i]);
      }
      act8(f, act);
      st16(ob + (long long)px * op, pack8(f, out.dtype));
      if (has_out2) st16(o2b + (long long)px * o2p, pack8(f, out2.dtype));
      if (ok1) {
        cvt8(r1, y.dtype, f);
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = __ffma2_rn(f[i], sc[i], sh[i]);
        if (RES) {
          float2 r[4];
          cvt8(q1, res.dtype, r);
#pragma unroll
          for (int i = 0; i < 4; ++i) f[i] = __fadd2_rn(f[i], r[i]);
        }
        act8(f, act);
        st16(ob + (long long)px1 * op, pack8(f, out.dtype));
        if (has_out2) st16(o2b + (long long)px1 * o2p, pack8(f, out2.dtype));
      }
    }
  } else {
    // y.H == 2 * out.H: input pixel row index n * H + 2 h = 2 * (n * Ho + h)
    const unsigned Wo = out.W, Wi = y.W;
    for (unsigned px = blockIdx.x * PLn + pl; px < npix; px += stride) {
      const unsigned q = px / Wo;
      const unsigned w = px - q * Wo;
      const long long p00 = ((long long)(2 * q) * Wi + 2 * w) * yp;
      uint4 r[4];
      r[0] = ld16(yb + p00);
      r[1] = ld16(yb + p00 + yp);
      r[2] = ld16(yb + p00 + (long long)Wi * yp);
      r[3] = ld16(yb + p00 + (long long)Wi * yp + yp);
      float2 o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f[4];
        cvt8(r[k], y.dtype, f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f[i] = __ffma2_rn(f[i], sc[i], sh[i]);
          if (k == 0) o[i] = f[i];
          else { o[i].x = fmaxf(o[i].x, f[i].x); o[i].y = fmaxf(o[i].y, f[i].y); }
        }
      }
      act8(o, act);          // monotone activation: act(max) == max(act)
      st16(ob + (long long)px * op, pack8(o, out.dtype));
      if (has_out2) st16(o2b + (long long)px * o2p, pack8(o, out2.dtype));
    }
  }
}

// ---- backward, pass 1: S1 = sum g, S2 = sum g * y  (g = dout * act'(.), routed to the arg-max position when pooled) ----------
// pooled: picks the FIRST maximum of the pre-activation in window order (00, 01, 10, 11) - the same position the generic kernel
// finds on the activated values, except among non-positive ReLU inputs where g is zero anyway
template <bool POOL, bool RES = false>
__device__ __forceinline__ void bnf_bwd_pix(const uint4* ry, const uint4& rg, int ydt, int gdt, const float2* sc, const float2* sh, int act,
                                            float2* g, float2* ysel, unsigned& sel, const uint4* rres = nullptr, int rdt = 0) {
  cvt8(rg, gdt, g);
  if (!POOL) {
    cvt8(ry[0], ydt, ysel);
    float2 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ffma2_rn(ysel[i], sc[i], sh[i]);
    if (RES) {
      float2 r[4];
      cvt8(*rres, rdt, r);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __fadd2_rn(v[i], r[i]);
    }
    act_grad8(g, v, act);
    sel = 0;
  } else {
    float2 best[4];
    sel = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f[4];
      cvt8(ry[k], ydt, f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = __ffma2_rn(f[i], sc[i], sh[i]);
        if (k == 0) { best[i] = v; ysel[i] = f[i]; }
        else {
          const bool bx = v.x > best[i].x, by = v.y > best[i].y;
          best[i].x = bx ? v.x : best[i].x; ysel[i].x = bx ? f[i].x : ysel[i].x;
          best[i].y = by ? v.y : best[i].y; ysel[i].y = by ? f[i].y : ysel[i].y;
          sel = bx ? ((sel & ~(3u << (4 * i))) | ((unsigned)k << (4 * i))) : sel;
          sel = by ? ((sel & ~(3u << (4 * i + 2))) | ((unsigned)k << (4 * i + 2))) : sel;
        }
      }
    }
    act_grad8(g, best, act);
  }
}

template <bool POOL, bool RES>
__global__ void __launch_bounds__(256, 3) bnf_bwd_reduce_kernel(dn_view dout, dn_view y, const float* __restrict__ mean_invstd,
                                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                         int act, float* __restrict__ ws, int CGb, double* __restrict__ red, dn_view res) {
  BNF_PROLOGUE(dout)
  float2 s1[4], s2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { s1[i] = make_float2(0.f, 0.f); s2[i] = make_float2(0.f, 0.f); }
  if (cvalid) {
    float2 sc[4], sh[4];
    load_sc_sh(mean_invstd, gamma, beta, C, c0, sc, sh);
    const char* yb = (const char*)y.ptr + (long long)c0 * 2;
    const char* gb = (const char*)dout.ptr + (long long)c0 * 2;
    const long long yp = y.sW * 2, gp = dout.sW * 2;
    const char* rb = (const char*)res.ptr + (long long)c0 * 2;
    const long long rp = res.sW * 2;
    if (!POOL) {
      unsigned px = blockIdx.x * PLn + pl;
      for (; !RES && px + 3 * stride < npix && px + 3 * stride >= px; px += 4 * stride) {
        uint4 ry[4], rg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ry[u] = ld16(yb + (long long)(px + u * stride) * yp);
          rg[u] = ld16(gb + (long long)(px + u * stride) * gp);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float2 g[4], ys[4];
          unsigned sel;
          bnf_bwd_pix<false>(&ry[u], rg[u], y.dtype, dout.dtype, sc, sh, act, g, ys, sel);
#pragma unroll
          for (int i = 0; i < 4; ++i) { s1[i] = __fadd2_rn(s1[i], g[i]); s2[i] = __ffma2_rn(g[i], ys[i], s2[i]); }
        }
      }
      for (; RES && px + stride < npix && px + stride >= px; px += 2 * stride) {      // residual operand: 2 pixels x 3 loads in flight
        uint4 ry[2], rg[2], rr[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          ry[u] = ld16(yb + (long long)(px + u * stride) * yp);
          rg[u] = ld16(gb + (long long)(px + u * stride) * gp);
          rr[u] = ld16(rb + (long long)(px + u * stride) * rp);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float2 g[4], ys[4];
          unsigned sel;
          bnf_bwd_pix<false, true>(&ry[u], rg[u], y.dtype, dout.dtype, sc, sh, act, g, ys, sel, &rr[u], res.dtype);
#pragma unroll
          for (int i = 0; i < 4; ++i) { s1[i] = __fadd2_rn(s1[i], g[i]); s2[i] = __ffma2_rn(g[i], ys[i], s2[i]); }
        }
      }
      for (; px < npix; px += stride) {
        const uint4 ry = ld16(yb + (long long)px * yp), rg = ld16(gb + (long long)px * gp);
        uint4 rr = ry;
        if (RES) rr = ld16(rb + (long long)px * rp);
        float2 g[4], ys[4];
        unsigned sel;
        bnf_bwd_pix<false, RES>(&ry, rg, y.dtype, dout.dtype, sc, sh, act, g, ys, sel, &rr, res.dtype);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s1[i] = __fadd2_rn(s1[i], g[i]); s2[i] = __ffma2_rn(g[i], ys[i], s2[i]); }
      }
    } else {
      const unsigned Wo = dout.W, Wi = y.W;
      for (unsigned px = blockIdx.x * PLn + pl; px < npix; px += stride) {
        const unsigned q = px / Wo;
        const unsigned w = px - q * Wo;
        const long long p00 = ((long long)(2 * q) * Wi + 2 * w) * yp;
        uint4 ry[4];
        ry[0] = ld16(yb + p00);
        ry[1] = ld16(yb + p00 + yp);
        ry[2] = ld16(yb + p00 + (long long)Wi * yp);
        ry[3] = ld16(yb + p00 + (long long)Wi * yp + yp);
        const uint4 rg = ld16(gb + (long long)px * gp);
        float2 g[4], ys[4];
        unsigned sel;
        bnf_bwd_pix<true>(ry, rg, y.dtype, dout.dtype, sc, sh, act, g, ys, sel);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s1[i] = __fadd2_rn(s1[i], g[i]); s2[i] = __ffma2_rn(g[i], ys[i], s2[i]); }
      }
    }
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[2 * i] = s1[i].x; acc[2 * i + 1] = s1[i].y; acc[8 + 2 * i] = s2[i].x; acc[8 + 2 * i + 1] = s2[i].y; }
  bn_bwd_reduce_tail<8>(acc, C, ws, CGb, red, cvalid, c0);
}

// ---- backward, pass 2: dy = A * g - B - Cc * y ------------------------------------------------------------------------
template <bool POOL, bool RES>      // RES: the residual operand enters the activation; its gradient (= g) goes to dres (= or +=)
__global__ void __launch_bounds__(256, 3) bnf_bwd_apply_kernel(dn_view dout, dn_view y, const float* __restrict__ mean_invstd,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                        int act, const double* __restrict__ red, double count, float gscale,
                                                                        float* dgamma, float* dbeta, dn_view dy, int CGb, dn_view res, dn_view dres,
                                                                        int has_dres, int dres_acc) {
  BNF_PROLOGUE(dout)
  if (!cvalid) return;
  float2 sc[4], sh[4], nB[4], nC[4];      // A == sc;  nB = -B, nC = -Cc:  dy = fma(A, g, fma(nC, y, nB))
  load_sc_sh(mean_invstd, gamma, beta, C, c0, sc, sh);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float cb[2], cc[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = c0 + 2 * i + j;
      const float mean = mean_invstd[c], istd = mean_invstd[C + c];
      const float a = j ? sc[i].y : sc[i].x;
      const double r1 = red[c], r2 = red[C + c];
      const double sgx = (double)istd * (r2 - (double)mean * r1);      // sum g * xhat
      const float m1 = (float)(r1 / count), m2 = (float)(sgx / count);
      cc[j] = a * istd * m2;
      cb[j] = a * m1 - cc[j] * mean;
      if (blockIdx.x == 0 && pl == 0) {
        if (dgamma) dgamma[c] = (float)(sgx * (double)gscale);
        if (dbeta) dbeta[c] = (float)(r1 * (double)gscale);
      }
    }
    nB[i] = make_float2(-cb[0], -cb[1]);
    nC[i] = make_float2(-cc[0], -cc[1]);
  }
  const char* yb = (const char*)y.ptr + (long long)c0 * 2;
  const char* gb = (const char*)dout.ptr + (long long)c0 * 2;
  char* ob = (char*)dy.ptr + (long long)c0 * 2;
  const long long yp = y.sW * 2, gp = dout.sW * 2, op = dy.sW * 2;
  const char* rb = (const char*)res.ptr + (long long)c0 * 2;
  char* db = (char*)dres.ptr + (long long)c0 * 2;
  const long long rp = res.sW * 2, dp = dres.sW * 2;
  if (!POOL) {
    unsigned px = blockIdx.x * PLn + pl;
    for (; px < npix; px += 2 * stride) {
      const bool ok1 = px + stride < npix && px + stride > px;
      const unsigned px1 = ok1 ? px + stride : px;
      const uint4 ry0 = ld16(yb + (long long)px * yp), rg0 = ld16(gb + (long long)px * gp);
      const uint4 ry1 = ld16(yb + (long long)px1 * yp), rg1 = ld16(gb + (long long)px1 * gp);
      uint4 rr0 = ry0, rr1 = ry1;
      if (RES) { rr0 = ld16(rb + (long long)px * rp); rr1 = ld16(rb + (long long)px1 * rp); }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !ok1) break;
        const unsigned pxu = u ? px1 : px;
        float2 g[4], ys[4], o[4];
        unsigned sel;
        bnf_bwd_pix<false, RES>(u ? &ry1 : &ry0, u ? rg1 : rg0, y.dtype, dout.dtype, sc, sh, act, g, ys, sel, u ? &rr1 : &rr0, res.dtype);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __ffma2_rn(sc[i], g[i], __ffma2_rn(nC[i], ys[i], nB[i]));
        st16(ob + (long long)pxu * op, pack8(o, dy.dtype));
        if (RES && has_dres) {
          char* d = db + (long long)pxu * dp;
          if (dres_acc) {
            float2 old[4];
            cvt8(*reinterpret_cast<const uint4*>(d), dres.dtype, old);
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = __fadd2_rn(old[i], g[i]);
          }
          st16(d, pack8(g, dres.dtype));
        }
      }
    }
  } else {
    const unsigned Wo = dout.W, Wi = y.W;
    for (unsigned px = blockIdx.x * PLn + pl; px < npix; px += stride) {
      const unsigned q = px / Wo;
      const unsigned w = px - q * Wo;
      const long long pix00 = (long long)(2 * q) * Wi + 2 * w;
      const long long poff[4] = {pix00, pix00 + 1, pix00 + Wi, pix00 + Wi + 1};
      uint4 ry[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) ry[k] = ld16(yb + poff[k] * yp);
      const uint4 rg = ld16(gb + (long long)px * gp);
      float2 g[4], ys[4];
      unsigned sel;
      bnf_bwd_pix<true>(ry, rg, y.dtype, dout.dtype, sc, sh, act, g, ys, sel);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f[4], o[4];
        cvt8(ry[k], y.dtype, f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 gk;
          gk.x = ((sel >> (4 * i)) & 3u) == (unsigned)k ? g[i].x : 0.f;
          gk.y = ((sel >> (4 * i + 2)) & 3u) == (unsigned)k ? g[i].y : 0.f;
          o[i] = __ffma2_rn(sc[i], gk, __ffma2_rn(nC[i], f[i], nB[i]));
        }
        st16(ob + poff[k] * op, pack8(o, dy.dtype));
      }
    }
  }
}

// ---- activation backward in place (dout *= act'(out)) + bias gradient (decoder convolutions without BatchNorm) --------------
__global__ void __launch_bounds__(256, 4) actf_bwd_kernel(dn_view dout, dn_view out, int act, float* __restrict__ ws, int CGb,
                                                          float* __restrict__ dbias, float gscale) {
  BNF_PROLOGUE(dout)
  float2 s[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = make_float2(0.f, 0.f);
  if (cvalid) {
    char* gb = (char*)dout.ptr + (long long)c0 * 2;
    const char* ob = (const char*)out.ptr + (long long)c0 * 2;
    const long long gp = dout.sW * 2, op = out.sW * 2;
    unsigned px = blockIdx.x * PLn + pl;
    if (act != DN_ACT_NONE) {
      for (; px < npix; px += 2 * stride) {
        const bool ok1 = px + stride < npix && px + stride > px;
        const unsigned px1 = ok1 ? px + stride : px;
        const uint4 rg0 = *reinterpret_cast<const uint4*>(gb + (long long)px * gp), ro0 = ld16(ob + (long long)px * op);
        const uint4 rg1 = *reinterpret_cast<const uint4*>(gb + (long long)px1 * gp), ro1 = ld16(ob + (long long)px1 * op);
        float2 g[4], o[4];
        cvt8(rg0, dout.dtype, g);
        cvt8(ro0, out.dtype, o);
        act_grad8(g, o, act);
        st16(gb + (long long)px * gp, pack8(g, dout.dtype));
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = __fadd2_rn(s[i], g[i]);
        if (ok1) {
          cvt8(rg1, dout.dtype, g);
          cvt8(ro1, out.dtype, o);
          act_grad8(g, o, act);
          st16(gb + (long long)px1 * gp, pack8(g, dout.dtype));
#pragma unroll
          for (int i = 0; i < 4; ++i) s[i] = __fadd2_rn(s[i], g[i]);
        }
      }
    } else {
      for (; px < npix; px += stride) {
        float2 g[4];
        cvt8(*reinterpret_cast<const uint4*>(gb + (long long)px * gp), dout.dtype, g);
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = __fadd2_rn(s[i], g[i]);
      }
    }
  }
  if (!ws) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[2 * i] = s[i].x; acc[2 * i + 1] = s[i].y; }
  act_bwd_tail<8>(acc, C, ws, CGb, dbias, gscale, cvalid, c0);
}
