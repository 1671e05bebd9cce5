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

// gx[off], gx[off + 1] = (lo, hi) + old (a 16-bit pair fetched earlier; zero when overwriting)
__device__ __forceinline__ void add_store2(const dn_view& gx, long long off, float lo, float hi, uint32_t old, bool ok) {
  if (!ok) return;
  if (gx.dtype == DN_BF16) {
    const float2 o = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old));
    *reinterpret_cast<__nv_bfloat162*>((__nv_bfloat16*)gx.ptr + off) = __floats2bfloat162_rn(lo + o.x, hi + o.y);
  } else {
    const float2 o = __half22float2(*reinterpret_cast<const __half2*>(&old));
    *reinterpret_cast<__half2*>((__half*)gx.ptr + off) = __floats2half2_rn(lo + o.x, hi + o.y);
  }
}

// gx[off], gx[off + 1] (+)= (lo, hi) as one 4-byte access
__device__ __forceinline__ void rmw2(const dn_view& gx, long long off, float lo, float hi, int acc, bool ok) {
  if (!ok) return;
  if (gx.dtype == DN_BF16) {
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>((__nv_bfloat16*)gx.ptr + off);
    if (acc) { const float2 o = __bfloat1622float2(*p); lo += o.x; hi += o.y; }
    *p = __floats2bfloat162_rn(lo, hi);
  } else {
    __half2* p = reinterpret_cast<__half2*>((__half*)gx.ptr + off);
    if (acc) { const float2 o = __half22float2(*p); lo += o.x; hi += o.y; }
    *p = __floats2half2_rn(lo, hi);
  }
}

// ---- bulk-copy variants for the wide heads (C = 16, 32): whole halo rows are contiguous in NHWC, so a tile is fetched
// with one cp.async.bulk per row (10 instructions per tile instead of ~700 address computations), completion on an
// mbarrier, two tile buffers so the next tile streams in while the MMAs run on the current one.
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_addr(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(smem_addr(b))
               : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct TileIdx { int n, h0, w0; };
__device__ __forceinline__ TileIdx tile_idx(int ti, int tilesW, int tilesH) {
  TileIdx t;
  const int tq = ti / tilesW;
  t.w0 = (ti - tq * tilesW) * TW;
  t.n = tq / tilesH;
  t.h0 = (tq - t.n * tilesH) * TH;
  return t;
}

// Fetch rows [r0, r0 + nrows) x columns [c0, c0 + ncols) of image n (pixel vectors of PSB bytes) into a dense [nrows][ncols]
// tile; whatever lies outside the image is zero.  All threads call it; contains __syncthreads() only for border tiles.
template <int PSB>
__device__ __forceinline__ void fetch_tile(const dn_view& x, int n, int r0, int c0, int nrows, int ncols, uint8_t* buf, uint64_t* bar) {
  const bool border = r0 < 0 || c0 < 0 || r0 + nrows > x.H || c0 + ncols > x.W;
  if (border) {          // block-uniform
    uint4* p = reinterpret_cast<uint4*>(buf);
    for (int i = threadIdx.x; i < nrows * ncols * PSB / 16; i += blockDim.x) p[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
  }
  if (threadIdx.x < 32) {      // warp 0: lane 0 posts the byte count, lanes 0 .. nrows-1 each issue one row
    const int lo = c0 < 0 ? 0 : c0, hi = c0 + ncols > x.W ? x.W : c0 + ncols;
    const int rlo = r0 < 0 ? 0 : r0, rhi = r0 + nrows > x.H ? x.H : r0 + nrows;
    const uint32_t row_bytes = (uint32_t)(hi - lo) * PSB;
    if (threadIdx.x == 0) {
      fence_async_proxy();
      mbar_expect_tx(bar, row_bytes * (uint32_t)(rhi - rlo));
    }
    __syncwarp();
    const int r = rlo + (int)threadIdx.x;
    if (r < rhi)
      bulk_g2s(smem_addr(buf) + (uint32_t)(((r - r0) * ncols + (lo - c0)) * PSB), (const uint8_t*)x.ptr + dn_off(x, n, r, lo) * 2, row_bytes, bar);
  }
}

template <bool BF16, int C, int S>      // S tile buffers: S - 1 tiles in flight while one is consumed
__global__ void __launch_bounds__(256) fwd_bulk_kernel(dn_view x, const float* __restrict__ w, const float* __restrict__ bias, dn_view z) {
  // Tap-parallel formulation: the legacy warp MMA path of sm_100 is slow enough (~1 HMMA.16816 per 32 clk per SM
  // sub-partition, measured) that spending one MMA per tap with 7 of 8 output columns empty made this kernel MMA-bound.
  // Instead the 8 MMA columns are 8 TAPS: P[q][t] = sum_c x[q][c] w[t][c] for every halo pixel q (taps 0..7 in one MMA,
  // tap 8 in a second one), un-shifted; P goes to shared memory and each output pixel then sums its nine shifted entries
  // z[p] = b + sum_t P[p + d_t][t].  2 MMAs per 16 halo pixels and 16 channels instead of 9 per 16 output pixels.
  dn_pdl_trigger();
  constexpr int KC = C / 16, PSB = C * 2, NPX = HH * HW, NFRAG = (NPX + 15) / 16, TILE_BYTES = NFRAG * 16 * PSB;
  extern __shared__ __align__(128) uint8_t smem[];
  uint2* bfr = reinterpret_cast<uint2*>(smem);          // [KC][2][32 lanes]
  uint8_t* tiles = smem + KC * 512;                     // S x TILE_BYTES (halo tile + padding up to a whole fragment)
  float* P = reinterpret_cast<float*>(tiles + S * TILE_BYTES);      // [NFRAG * 16][9]
  uint64_t* bars = reinterpret_cast<uint64_t*>(P + NFRAG * 16 * 9);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, j = lane & 3;
  if (tid == 0) {
    for (int k = 0; k < S; ++k) mbar_init(&bars[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  dn_pdl_wait();
  for (int i = tid; i < KC * 64; i += 256) {
    const int l = i & 31, which = (i >> 5) & 1, kc = i >> 6;
    const int t = which ? 8 : (l >> 2);                  // MMA column n = l >> 2: tap n (first MMA) / tap 8 in column 0 (second)
    const int c = kc * 16 + 2 * (l & 3);
    uint2 v = make_uint2(0u, 0u);
    if (!which || (l >> 2) == 0) {
      v.x = pack2<BF16>(w[c * 9 + t], w[(c + 1) * 9 + t]);
      v.y = pack2<BF16>(w[(c + 8) * 9 + t], w[(c + 9) * 9 + t]);
    }
    bfr[i] = v;
  }
  __syncthreads();
  const int tilesW = (x.W + TW - 1) / TW, tilesH = (x.H + TH - 1) / TH;
  const int ntiles = tilesW * tilesH * x.N;
  const float b0 = bias ? bias[0] : 0.f;
  const int mi = lane >> 3, rr = lane & 7;
  const int px = rr + 8 * (mi & 1), choff = 8 * (mi >> 1);
  int ti = blockIdx.x;
  for (int k = 0; k < S - 1; ++k) {
    const int tk = ti + k * (int)gridDim.x;
    if (tk < ntiles) {
      const TileIdx t0 = tile_idx(tk, tilesW, tilesH);
      fetch_tile<PSB>(x, t0.n, t0.h0 - 1, t0.w0 - 1, HH, HW, tiles + k * TILE_BYTES, &bars[k]);
    }
  }
  for (int it = 0; ti < ntiles; ++it, ti += gridDim.x) {
    const int cur = it % S;
    const int tnext = ti + (S - 1) * (int)gridDim.x;
    if (tnext < ntiles) {      // its buffer was released by the __syncthreads() that ended the previous pass
      const int nb = (it + S - 1) % S;
      const TileIdx tn = tile_idx(tnext, tilesW, tilesH);
      fetch_tile<PSB>(x, tn.n, tn.h0 - 1, tn.w0 - 1, HH, HW, tiles + nb * TILE_BYTES, &bars[nb]);
    }
    const TileIdx tc = tile_idx(ti, tilesW, tilesH);
    mbar_wait(&bars[cur], (uint32_t)((it / S) & 1));
    const uint32_t tile_s = smem_addr(tiles + cur * TILE_BYTES);
    for (int f = warp; f < NFRAG; f += 8) {
      float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t row = tile_s + (uint32_t)((f * 16 + px) * PSB + choff * 2);
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        uint32_t a[4], b[2];
        ldsm_x4(row + (uint32_t)(kc * 32), a);
        uint2 bv = bfr[(kc * 2) * 32 + lane];
        b[0] = bv.x; b[1] = bv.y;
        mma16816<BF16>(d1, a, b);
        bv = bfr[(kc * 2 + 1) * 32 + lane];
        b[0] = bv.x; b[1] = bv.y;
        mma16816<BF16>(d2, a, b);
      }
      float* pa = P + (f * 16 + g) * 9;
      pa[2 * j] = d1[0]; pa[2 * j + 1] = d1[1];
      pa[72 + 2 * j] = d1[2]; pa[72 + 2 * j + 1] = d1[3];
      if (j == 0) { pa[8] = d2[0]; pa[72 + 8] = d2[2]; }
    }
    __syncthreads();
    {
      const int r = warp, c = lane;
      float acc = b0;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc += P[((r + t / 3) * HW + c + t % 3) * 9 + t];
      const int h = tc.h0 + r, wv = tc.w0 + c;
      if (h < x.H && wv < x.W) dn_st(z.ptr, z.dtype, dn_off(z, tc.n, h, wv), acc);
    }
    __syncthreads();
  }
}

template <int NG, int S>      // NG = C / 8 in {2, 4}; S x-tile buffers
__global__ void __launch_bounds__(256, 2) bwd_bulk_kernel(dn_view x, const float* __restrict__ w, dn_view dz, dn_view gx, int gx_acc,
                                                          float* __restrict__ wsp) {
  dn_pdl_trigger();
  constexpr int C = NG * 8, PSB = C * 2, XT_BYTES = TH * TW * PSB, NDZ = HH * HW;
  extern __shared__ __align__(128) uint8_t smem[];
  uint2* wfr = reinterpret_cast<uint2*>(smem);                          // [NG][32]
  float* accs = reinterpret_cast<float*>(smem + NG * 256);              // [9 * C + 1]
  float* dzt = accs + ((9 * C + 1 + 3) & ~3);                           // 2 x [HH][HW]
  uint8_t* xt = reinterpret_cast<uint8_t*>(dzt + 2 * ((NDZ + 3) & ~3)); // S x XT_BYTES (dense [TH][TW][C])
  uint64_t* bars = reinterpret_cast<uint64_t*>(xt + S * XT_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, j = lane & 3;
  if (tid == 0) {
    for (int k = 0; k < S; ++k) mbar_init(&bars[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  dn_pdl_wait();
  for (int i = tid; i < NG * 32; i += 256) {
    const int l = i & 31, n8 = i >> 5;
    const int c = n8 * 8 + (l >> 2), t0 = 2 * (l & 3);
    uint2 v;
    v.x = pack2<true>(w[c * 9 + t0], w[c * 9 + t0 + 1]);
    v.y = (l & 3) == 0 ? pack2<true>(w[c * 9 + 8], 0.f) : 0u;
    wfr[i] = v;
  }
  for (int i = tid; i < 9 * C + 1; i += 256) accs[i] = 0.f;
  float dw[NG][4];
#pragma unroll
  for (int n8 = 0; n8 < NG; ++n8) { dw[n8][0] = dw[n8][1] = dw[n8][2] = dw[n8][3] = 0.f; }
  float accb = 0.f;
  const int tilesW = (x.W + TW - 1) / TW, tilesH = (x.H + TH - 1) / TH;
  const int ntiles = tilesW * tilesH * x.N;
  const bool x_bf16 = x.dtype == DN_BF16;
  constexpr int DZ_PER_THREAD = (NDZ + 255) / 256;      // 2
  // dz halo values of a tile, one or two per thread (zero outside the image)
  // raw bits now, conversion when they are stored: nothing waits on the global loads while the current tile is processed
  auto load_dz = [&](const TileIdx& t, uint32_t* v) {
#pragma unroll
    for (int k = 0; k < DZ_PER_THREAD; ++k) {
      const int i = tid + k * 256;
      v[k] = 0u;
      if (i < NDZ) {
        const int hy = i / HW, hx = i - hy * HW;
        const int h = t.h0 + hy - 1, wv = t.w0 + hx - 1;
        if (h >= 0 && h < x.H && wv >= 0 && wv < x.W) {
          const long long o = dn_off(dz, t.n, h, wv);
          v[k] = dz.dtype == DN_F32 ? ((const uint32_t*)dz.ptr)[o] : (uint32_t)((const uint16_t*)dz.ptr)[o];
        }
      }
    }
  };
  auto store_dz = [&](const uint32_t* v, float* dst) {
#pragma unroll
    for (int k = 0; k < DZ_PER_THREAD; ++k) {
      const int i = tid + k * 256;
      if (i < NDZ) {
        const float fv = dz.dtype == DN_F32 ? __uint_as_float(v[k])
                         : dz.dtype == DN_BF16 ? __uint_as_float(v[k] << 16) : __half2float(__ushort_as_half((unsigned short)v[k]));
        dst[i] = fv;
        const int hy = i / HW, hx = i - hy * HW;
        if (hy >= 1 && hy <= TH && hx >= 1 && hx <= TW) accb += fv;      // interior: counted once for the bias gradient
      }
    }
  };
  __syncthreads();
  int ti = blockIdx.x;
  for (int k = 0; k < S - 1; ++k) {
    const int tk = ti + k * (int)gridDim.x;
    if (tk < ntiles) {
      const TileIdx t0 = tile_idx(tk, tilesW, tilesH);
      fetch_tile<PSB>(x, t0.n, t0.h0, t0.w0, TH, TW, xt + k * XT_BYTES, &bars[k]);
    }
  }
  if (ti < ntiles) {
    const TileIdx t0 = tile_idx(ti, tilesW, tilesH);
    uint32_t v[DZ_PER_THREAD];
    load_dz(t0, v);
    store_dz(v, dzt);
  }
  __syncthreads();
  for (int it = 0; ti < ntiles; ++it, ti += gridDim.x) {
    const int cur = it & 1;                             // dz buffers alternate; x buffers rotate through S
    const int xcur = it % S;
    const bool has_next = ti + (int)gridDim.x < ntiles;
    const int tfar = ti + (S - 1) * (int)gridDim.x;
    if (tfar < ntiles) {
      const int nb = (it + S - 1) % S;
      const TileIdx tn = tile_idx(tfar, tilesW, tilesH);
      fetch_tile<PSB>(x, tn.n, tn.h0, tn.w0, TH, TW, xt + nb * XT_BYTES, &bars[nb]);
    }
    uint32_t vn[DZ_PER_THREAD];
    if (has_next) load_dz(tile_idx(ti + gridDim.x, tilesW, tilesH), vn);      // in registers until the current tile is done
    const TileIdx tc = tile_idx(ti, tilesW, tilesH);
    // old values of the gradient slots this lane updates, requested before anything waits on them
    const int hq = tc.h0 + warp;
    uint32_t oldv[2][NG][2];
    long long offq[2][2];
    bool okq[2][2];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int wa = tc.w0 + f * 16 + g;
      okq[f][0] = hq < x.H && wa < x.W;
      okq[f][1] = hq < x.H && wa + 8 < x.W;
      offq[f][0] = okq[f][0] ? dn_off(gx, tc.n, hq, wa) : 0;
      offq[f][1] = okq[f][1] ? dn_off(gx, tc.n, hq, wa + 8) : 0;
#pragma unroll
      for (int n8 = 0; n8 < NG; ++n8)
#pragma unroll
        for (int k = 0; k < 2; ++k)
          oldv[f][n8][k] = (gx_acc && okq[f][k]) ? *reinterpret_cast<const uint32_t*>((const uint8_t*)gx.ptr + (offq[f][k] + n8 * 8 + 2 * j) * 2) : 0u;
    }
    mbar_wait(&bars[xcur], (uint32_t)((it / S) & 1));
    uint8_t* xc = xt + xcur * XT_BYTES;
    if (!x_bf16) {                                      // fp16 activations -> bf16 in place (the gradient products run in bf16)
      for (int i = tid; i < XT_BYTES / 16; i += 256) {
        uint4 u = reinterpret_cast<uint4*>(xc)[i];
        const __half2* hv = reinterpret_cast<const __half2*>(&u);
        uint4 o;
        uint32_t* ov = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(hv[k]); ov[k] = pack2<true>(f.x, f.y); }
        reinterpret_cast<uint4*>(xc)[i] = o;
      }
      __syncthreads();
    }
    const float* dzc = dzt + cur * ((NDZ + 3) & ~3);
    const uint32_t xt_s = smem_addr(xc);
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int cb = f * 16;
      auto val = [&](int i, int t) -> float { return dzc[(warp + 2 - t / 3) * HW + cb + i + 2 - t % 3]; };
      uint32_t ax[4], aw[4];
      ax[0] = pack2<true>(val(g, 2 * j), val(g, 2 * j + 1));
      ax[1] = pack2<true>(val(g + 8, 2 * j), val(g + 8, 2 * j + 1));
      ax[2] = j == 0 ? pack2<true>(val(g, 8), 0.f) : 0u;
      ax[3] = j == 0 ? pack2<true>(val(g + 8, 8), 0.f) : 0u;
      aw[0] = pack2<true>(val(2 * j, g), val(2 * j + 1, g));
      aw[1] = g == 0 ? pack2<true>(val(2 * j, 8), val(2 * j + 1, 8)) : 0u;
      aw[2] = pack2<true>(val(2 * j + 8, g), val(2 * j + 9, g));
      aw[3] = g == 0 ? pack2<true>(val(2 * j + 8, 8), val(2 * j + 9, 8)) : 0u;
      const uint32_t xrow = xt_s + (uint32_t)((warp * TW + cb + (lane & 15)) * PSB);
#pragma unroll
      for (int n8 = 0; n8 < NG; ++n8) {
        uint32_t b[2];
        const uint2 bv = wfr[n8 * 32 + lane];
        b[0] = bv.x; b[1] = bv.y;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816<true>(d, ax, b);
        const int c = n8 * 8 + 2 * j;
        add_store2(gx, offq[f][0] + c, d[0], d[1], oldv[f][n8][0], okq[f][0]);
        add_store2(gx, offq[f][1] + c, d[2], d[3], oldv[f][n8][1], okq[f][1]);
        uint32_t bx[2];
        ldsm_x2_trans(xrow + (uint32_t)(n8 * 16), bx);
        mma16816<true>(dw[n8], aw, bx);
      }
    }
    if (has_next) store_dz(vn, dzt + (cur ^ 1) * ((NDZ + 3) & ~3));
    __syncthreads();
  }
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

static inline size_t fwd_bulk_smem(int C, int S) {
  const size_t nfrag = (HH * HW + 15) / 16;
  return (size_t)(C / 16) * 512 + (size_t)S * nfrag * 16 * C * 2 + nfrag * 16 * 9 * sizeof(float) + 8 * S;
}
static inline size_t bwd_bulk_smem(int C, int S) {
  return (size_t)(C / 8) * 256 + sizeof(float) * (((9 * C + 1 + 3) & ~3) + 2 * ((HH * HW + 3) & ~3)) + (size_t)S * TH * TW * C * 2 + 8 * S;
}
static inline size_t fwd_smem(int C) { return (size_t)9 * (C / 16) * 256 + (size_t)HH * HW * (C * 2 + 16); }
static inline size_t bwd_smem(int C) {
  return (size_t)(C / 8) * 256 + sizeof(float) * (((9 * C + 1 + 3) & ~3) + ((HH * HW + 3) & ~3)) + (size_t)TH * TW * (C * 2 + 16);
}
static inline bool eligible(const dn_view* x) { return dn_vec8_ok(x) && (x->C == 16 || x->C == 32 || x->C == 64 || x->C == 128); }
}  // namespace hc
