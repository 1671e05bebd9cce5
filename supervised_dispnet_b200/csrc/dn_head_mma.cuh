// Disparity-head convolution nn.Conv2d(C, 1, 3, padding=1)  (reference: predict_disp, models/Disp_vgg_BN.py:66-70,
// models/DispNetS.py:34-38) for 16-bit NHWC activations with C % 16 == 0.  Included by dn_layers.cu.
//
// One output channel: 2*9*C FLOP per pixel against 2*C bytes read - HBM-bound, and a 128 x N tcgen05 tile would waste
// 15/16 of its columns.  The CUDA-core version needed ~150 issue slots per pixel (fp16 -> fp32 converts + FMAs) and ran at
// 0.3-1.6 TB/s.  Here a block stages a (8+2) x (32+2)-pixel halo tile in shared memory once (cp.async, zero-filled
// borders) and the arithmetic runs on warp-level m16n8k16 MMAs fed by ldmatrix - the 3x3 taps are just shifted row
// addresses of the same tile - so the instruction stream shrinks to ~3 per (16 pixels x 16 channels x tap) and the kernel
// is bound by the tile loads.  (tcgen05 is the wrong tool for N = 1; legacy mma.sync has far more throughput than this
// memory-bound op can use.)
//   forward : z[p]     = b + sum_t sum_c x[p + d_t][c] w[t][c]          A = x tile (pixels x channels), B = w (col 0 of 8)
//   backward: dx[q][c] += sum_t dz[q - d_t] w[t][c]                      A = dz gathers (pixels x taps),  B = w (taps x channels)
//             dW[t][c] = sum_q dz[q - d_t] x[q][c],  db = sum dz         A = dz gathers (taps x pixels),  B = x tile (pixels x channels)
// The backward products run in bf16 (gradient activations are bf16 in every reduced-precision mode; fp16 activations are
// converted while they are staged), the forward in the activation type.
#pragma once

namespace hc {
constexpr int TH = 8, TW = 32, HH = TH + 2, HW = TW + 2;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t a, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t a, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (BF16) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;      // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// ---- forward ------------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256) fwd_kernel(dn_view x, const float* __restrict__ w, const float* __restrict__ bias, dn_view z) {
  dn_pdl_trigger();
  dn_pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  const int C = x.C, KC = C / 16, PS = C * 2 + 16;      // PS: pixel stride in bytes, an odd number of 16-byte units
  uint2* bfr = reinterpret_cast<uint2*>(smem);          // [9][KC][32 lanes]  B fragments: only output column 0 is non-zero
  uint8_t* tile = smem + (size_t)9 * KC * 256;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 9 * KC * 32; i += 256) {
    const int l = i & 31, tk = i >> 5, t = tk / KC, kc = tk - t * KC;
    uint2 v = make_uint2(0u, 0u);
    if ((l >> 2) == 0) {
      const int c = kc * 16 + 2 * (l & 3);
      v.x = pack2<BF16>(w[c * 9 + t], w[(c + 1) * 9 + t]);
      v.y = pack2<BF16>(w[(c + 8) * 9 + t], w[(c + 9) * 9 + t]);
    }
    bfr[i] = v;
  }
  const int tilesW = (x.W + TW - 1) / TW, tilesH = (x.H + TH - 1) / TH;
  const int ntiles = tilesW * tilesH * x.N;
  const float b0 = bias ? bias[0] : 0.f;
  const int nchunk = C / 8;
  const uint32_t tile_s = smem_addr(tile);
  for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const int tq = ti / tilesW;
    const int w0 = (ti - tq * tilesW) * TW;
    const int n = tq / tilesH;
    const int h0 = (tq - n * tilesH) * TH;
    __syncthreads();                                   // the previous tile's fragments are consumed (first pass: bfr is complete)
    for (int i = tid; i < HH * HW * nchunk; i += 256) {
      const int p = i / nchunk, ck = i - p * nchunk;
      const int hy = p / HW, hx = p - hy * HW;
      const int h = h0 + hy - 1, wv = w0 + hx - 1;
      const bool ok = h >= 0 && h < x.H && wv >= 0 && wv < x.W;
      const uint8_t* src = (const uint8_t*)x.ptr + (ok ? (dn_off(x, n, h, wv) + ck * 8) * 2 : 0);
      cp_async16(tile_s + (uint32_t)(p * PS + ck * 16), src, ok);
    }
    cp_async_wait_all();
    __syncthreads();
    const int mi = lane >> 3, rr = lane & 7;
    const int px = rr + 8 * (mi & 1), choff = 8 * (mi >> 1);
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int cb = f * 16;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int kh = t / 3, kw = t - 3 * kh;
        const uint32_t row = tile_s + (uint32_t)(((warp + kh) * HW + cb + px + kw) * PS + choff * 2);
        for (int kc = 0; kc < KC; ++kc) {
          uint32_t a[4], b[2];
          ldsm_x4(row + (uint32_t)(kc * 32), a);
          const uint2 bv = bfr[(t * KC + kc) * 32 + lane];
          b[0] = bv.x; b[1] = bv.y;
          mma16816<BF16>(acc, a, b);
        }
      }
      if ((lane & 3) == 0) {
        const int h = h0 + warp;
        const int g = lane >> 2;
        if (h < x.H) {
          if (w0 + cb + g < x.W) dn_st(z.ptr, z.dtype, dn_off(z, n, h, w0 + cb + g), acc[0] + b0);
          if (w0 + cb + g + 8 < x.W) dn_st(z.ptr, z.dtype, dn_off(z, n, h, w0 + cb + g + 8), acc[2] + b0);
        }
      }
    }
  }
}

// ---- backward: data gradient (+= or =), weight gradient and bias gradient partial sums in one pass ---------------------------
// wsp row of this block: [C][9] weight-gradient partials (torch layout) followed by the bias-gradient partial
template <int NG>      // NG = C / 8
__global__ void __launch_bounds__(256, NG >= 16 ? 1 : 2) bwd_kernel(dn_view x, const float* __restrict__ w, dn_view dz, dn_view gx, int gx_acc,
                                                                   float* __restrict__ wsp) {
  dn_pdl_trigger();
  dn_pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int C = NG * 8, PS = C * 2 + 16;
  uint2* wfr = reinterpret_cast<uint2*>(smem);                         // [NG][32]   B fragments of the dx product (taps x channels)
  float* accs = reinterpret_cast<float*>(smem + NG * 256);             // [9 * C + 1]
  float* dzt = accs + ((9 * C + 1 + 3) & ~3);                          // [HH][HW]   dz halo tile (fp32)
  uint8_t* xt = reinterpret_cast<uint8_t*>(dzt + ((HH * HW + 3) & ~3)); // [TH * TW][PS] x tile in bf16
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, j = lane & 3;
  for (int i = tid; i < NG * 32; i += 256) {
    const int l = i & 31, n8 = i >> 5;
    const int c = n8 * 8 + (l >> 2), t0 = 2 * (l & 3);
    uint2 v;
    v.x = pack2<true>(w[c * 9 + t0], w[c * 9 + t0 + 1]);               // taps 2j, 2j+1 (<= 7)
    v.y = (l & 3) == 0 ? pack2<true>(w[c * 9 + 8], 0.f) : 0u;          // taps 2j+8, 2j+9: only tap 8 exists
    wfr[i] = v;
  }
  for (int i = tid; i < 9 * C + 1; i += 256) accs[i] = 0.f;
  float dw[NG][4];
#pragma unroll
  for (int n8 = 0; n8 < NG; ++n8) { dw[n8][0] = dw[n8][1] = dw[n8][2] = dw[n8][3] = 0.f; }
  float accb = 0.f;
  const int tilesW = (x.W + TW - 1) / TW, tilesH = (x.H + TH - 1) / TH;
  const int ntiles = tilesW * tilesH * x.N;
  const uint32_t xt_s = smem_addr(xt);
  const bool x_bf16 = x.dtype == DN_BF16;
  for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const int tq = ti / tilesW;
    const int w0 = (ti - tq * tilesW) * TW;
    const int n = tq / tilesH;
    const int h0 = (tq - n * tilesH) * TH;
    __syncthreads();
    for (int i = tid; i < HH * HW; i += 256) {
      const int hy = i / HW, hx = i - hy * HW;
      const int h = h0 + hy - 1, wv = w0 + hx - 1;
      float v = 0.f;
      if (h >= 0 && h < x.H && wv >= 0 && wv < x.W) {
        v = dn_ld(dz.ptr, dz.dtype, dn_off(dz, n, h, wv));
        if (hy >= 1 && hy <= TH && hx >= 1 && hx <= TW) accb += v;      // interior pixel: counted once for the bias gradient
      }
      dzt[i] = v;
    }
    for (int i = tid; i < TH * TW * NG; i += 256) {
      const int p = i / NG, ck = i - p * NG;
      const int r = p / TW, cx = p - r * TW;
      const int h = h0 + r, wv = w0 + cx;
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (h < x.H && wv < x.W) {
        u = __ldg(reinterpret_cast<const uint4*>((const uint8_t*)x.ptr + (dn_off(x, n, h, wv) + ck * 8) * 2));
        if (!x_bf16) {
          const __half2* hv = reinterpret_cast<const __half2*>(&u);
          uint4 o;
          uint32_t* ov = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(hv[k]); ov[k] = pack2<true>(f.x, f.y); }
          u = o;
        }
      }
      *reinterpret_cast<uint4*>(xt + p * PS + ck * 16) = u;
    }
    __syncthreads();
    const int h = h0 + warp;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int cb = f * 16;
      // dz[q - d_t] for pixel i of the fragment and tap t sits at halo position (warp + 2 - t / 3, cb + i + 2 - t % 3)
      auto val = [&](int i, int t) -> float { return dzt[(warp + 2 - t / 3) * HW + cb + i + 2 - t % 3]; };
      uint32_t ax[4], aw[4];
      ax[0] = pack2<true>(val(g, 2 * j), val(g, 2 * j + 1));
      ax[1] = pack2<true>(val(g + 8, 2 * j), val(g + 8, 2 * j + 1));
      ax[2] = j == 0 ? pack2<true>(val(g, 8), 0.f) : 0u;
      ax[3] = j == 0 ? pack2<true>(val(g + 8, 8), 0.f) : 0u;
      aw[0] = pack2<true>(val(2 * j, g), val(2 * j + 1, g));
      aw[1] = g == 0 ? pack2<true>(val(2 * j, 8), val(2 * j + 1, 8)) : 0u;
      aw[2] = pack2<true>(val(2 * j + 8, g), val(2 * j + 9, g));
      aw[3] = g == 0 ? pack2<true>(val(2 * j + 8, 8), val(2 * j + 9, 8)) : 0u;
      const uint32_t xrow = xt_s + (uint32_t)((warp * TW + cb + (lane & 15)) * PS);
      const int wa = w0 + cb + g, wb = wa + 8;
      const bool oka = h < x.H && wa < x.W, okb = h < x.H && wb < x.W;
      const long long offa = oka ? dn_off(gx, n, h, wa) : 0, offb = okb ? dn_off(gx, n, h, wb) : 0;
#pragma unroll
      for (int n8 = 0; n8 < NG; ++n8) {
        uint32_t b[2];
        const uint2 bv = wfr[n8 * 32 + lane];
        b[0] = bv.x; b[1] = bv.y;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816<true>(d, ax, b);
        const int c = n8 * 8 + 2 * j;
        if (oka) {
          float lo = d[0], hi = d[1];
          if (gx_acc) { lo += dn_ld(gx.ptr, gx.dtype, offa + c); hi += dn_ld(gx.ptr, gx.dtype, offa + c + 1); }
          if (gx.dtype == DN_BF16) *reinterpret_cast<uint32_t*>((__nv_bfloat16*)gx.ptr + offa + c) = pack2<true>(lo, hi);
          else *reinterpret_cast<uint32_t*>((__half*)gx.ptr + offa + c) = pack2<false>(lo, hi);
        }
        if (okb) {
          float lo = d[2], hi = d[3];
          if (gx_acc) { lo += dn_ld(gx.ptr, gx.dtype, offb + c); hi += dn_ld(gx.ptr, gx.dtype, offb + c + 1); }
          if (gx.dtype == DN_BF16) *reinterpret_cast<uint32_t*>((__nv_bfloat16*)gx.ptr + offb + c) = pack2<true>(lo, hi);
          else *reinterpret_cast<uint32_t*>((__half*)gx.ptr + offb + c) = pack2<false>(lo, hi);
        }
        uint32_t bx[2];
        ldsm_x2_trans(xrow + (uint32_t)(n8 * 16), bx);
        mma16816<true>(dw[n8], aw, bx);
      }
    }
  }
  // dw[n8]: [0],[1] = dW[tap g][c, c+1], [2],[3] = dW[tap g + 8][c, c+1] (tap 8 lives in g == 0), c = n8*8 + 2j
  __syncthreads();
#pragma unroll
  for (int n8 = 0; n8 < NG; ++n8) {
    const int c = n8 * 8 + 2 * j;
    atomicAdd(&accs[c * 9 + g], dw[n8][0]);
    atomicAdd(&accs[(c + 1) * 9 + g], dw[n8][1]);
    if (g == 0) {
      atomicAdd(&accs[c * 9 + 8], dw[n8][2]);
      atomicAdd(&accs[(c + 1) * 9 + 8], dw[n8][3]);
    }
  }
  accb = dn_warp_sum(accb);
  if (lane == 0) atomicAdd(&accs[9 * C], accb);
  __syncthreads();
  for (int i = tid; i < 9 * C + 1; i += 256) wsp[(long long)blockIdx.x * (9 * C + 1) + i] = accs[i];
}

static inline size_t fwd_smem(int C) { return (size_t)9 * (C / 16) * 256 + (size_t)HH * HW * (C * 2 + 16); }
static inline size_t bwd_smem(int C) {
  return (size_t)(C / 8) * 256 + sizeof(float) * (((9 * C + 1 + 3) & ~3) + ((HH * HW + 3) & ~3)) + (size_t)TH * TW * (C * 2 + 16);
}
static inline bool eligible(const dn_view* x) { return dn_vec8_ok(x) && (x->C == 16 || x->C == 32 || x->C == 64 || x->C == 128); }
}  // namespace hc
