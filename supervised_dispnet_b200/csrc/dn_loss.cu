// Per-pixel loss / geometry kernels (HBM-bound, fp32): masked L1 depth loss, second-order smoothness,
// depth-error counters, area pyramid, explainability BCE, and the fused inverse-warp + photometric term.
// Each replaces a chain of ATen launches in loss_functions.py / inverse_warp.py with one pass; reductions are
// warp-shuffle -> shared -> one atomic per block.  Arithmetic follows SURVEY.md appendix A.1-A.5.
#include "dn_common.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  v = dn_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = dn_warp_sum(t);
  }
  return t;  // valid in warp 0
}
__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double redd[32];
  v = dn_warp_sum_d(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redd[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? redd[threadIdx.x] : 0.0;
    t = dn_warp_sum_d(t);
  }
  return t;
}

inline int blocks_for(long long n, int per_block, int cap_mult = 8) {
  long long b = (n + per_block - 1) / per_block;
  long long cap = (long long)dn_num_sms() * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------------
// l1_loss (loss_functions.py:104-129)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l1_fwd_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int HW,
                                                     float maxd, float* ws) {
  const int b = blockIdx.y;
  const float* g = gt + (long long)b * HW;
  const float* p = pred + (long long)b * HW;
  float s = 0.f, c = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float gv = g[i];
    if (gv > 0.f && gv < maxd) {
      float pv = fminf(fmaxf(p[i], 1e-3f), maxd);
      s += fabsf(gv - pv);
      c += 1.f;
    }
  }
  s = block_sum(s);
  c = block_sum(c);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 2 * b, s);
    atomicAdd(ws + 2 * b + 1, c);
  }
}
__global__ void l1_finalize_kernel(const float* ws, int B, float* loss) {
  float t = 0.f;
  for (int b = 0; b < B; ++b) t += ws[2 * b] / ws[2 * b + 1];   // 0/0 -> NaN like mean() of an empty selection
  loss[0] = t / (float)B;
}
__global__ void __launch_bounds__(256) l1_bwd_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int HW,
                                                     float maxd, const float* __restrict__ ws, int B,
                                                     const float* __restrict__ gout, float* __restrict__ gpred) {
  const int b = blockIdx.y;
  const float* g = gt + (long long)b * HW;
  const float* p = pred + (long long)b * HW;
  float* o = gpred + (long long)b * HW;
  const float k = gout[0] / (ws[2 * b + 1] * (float)B);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float gv = g[i], pv = p[i], r = 0.f;
    if (gv > 0.f && gv < maxd && pv >= 1e-3f && pv <= maxd) {
      float d = pv - gv;
      r = d > 0.f ? k : (d < 0.f ? -k : 0.f);
    }
    o[i] = r;
  }
}

// ---------------------------------------------------------------------------------------------------
// smooth_loss (loss_functions.py:367-386), one scale
// ---------------------------------------------------------------------------------------------------
struct SmoothCoef { float c1, c2, c3, c4; };

__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float sm_dx2(const float* p, int W, int v, int u) {
  const float* r = p + v * W + u;
  return (r[2] - r[1]) - (r[1] - r[0]);
}
__device__ __forceinline__ float sm_dy2(const float* p, int W, int v, int u) {
  const float* r = p + v * W + u;
  return (r[2 * W] - r[W]) - (r[W] - r[0]);
}
__device__ __forceinline__ float sm_dxdy(const float* p, int W, int v, int u) {   // dx[v+1,u] - dx[v,u]
  const float* r = p + v * W + u;
  return (r[W + 1] - r[W]) - (r[1] - r[0]);
}
__device__ __forceinline__ float sm_dydx(const float* p, int W, int v, int u) {   // dy[v,u+1] - dy[v,u]
  const float* r = p + v * W + u;
  return (r[W + 1] - r[1]) - (r[W] - r[0]);
}

__global__ void __launch_bounds__(256) smooth_fwd_kernel(const float* __restrict__ p, int H, int W, SmoothCoef k, float* loss) {
  const float* pb = p + (long long)blockIdx.y * H * W;
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    int v = i / W, u = i % W;
    if (u < W - 2) s += k.c1 * fabsf(sm_dx2(pb, W, v, u));
    if (v < H - 2) s += k.c4 * fabsf(sm_dy2(pb, W, v, u));
    if (u < W - 1 && v < H - 1) s += k.c2 * fabsf(sm_dxdy(pb, W, v, u)) + k.c3 * fabsf(sm_dydx(pb, W, v, u));
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(loss, s);
}

__global__ void __launch_bounds__(256) smooth_bwd_kernel(const float* __restrict__ p, int H, int W, SmoothCoef k,
                                                         const float* __restrict__ gout, float* __restrict__ gp) {
  const float* pb = p + (long long)blockIdx.y * H * W;
  float* gb = gp + (long long)blockIdx.y * H * W;
  const float go = gout[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    int v = i / W, u = i % W;
    float g = 0.f;
    // dx2[v,uu] = p[uu+2] - 2 p[uu+1] + p[uu]
    if (u <= W - 3) g += k.c1 * sgn(sm_dx2(pb, W, v, u));
    if (u >= 1 && u - 1 <= W - 3) g -= 2.f * k.c1 * sgn(sm_dx2(pb, W, v, u - 1));
    if (u >= 2) g += k.c1 * sgn(sm_dx2(pb, W, v, u - 2));
    if (v <= H - 3) g += k.c4 * sgn(sm_dy2(pb, W, v, u));
    if (v >= 1 && v - 1 <= H - 3) g -= 2.f * k.c4 * sgn(sm_dy2(pb, W, v - 1, u));
    if (v >= 2) g += k.c4 * sgn(sm_dy2(pb, W, v - 2, u));
    // mixed terms: +p[v,u] - p[v,u+1] - p[v+1,u] + p[v+1,u+1], defined for v<=H-2, u<=W-2
    if (v <= H - 2 && u <= W - 2) g += k.c2 * sgn(sm_dxdy(pb, W, v, u)) + k.c3 * sgn(sm_dydx(pb, W, v, u));
    if (v <= H - 2 && u >= 1) g -= k.c2 * sgn(sm_dxdy(pb, W, v, u - 1)) + k.c3 * sgn(sm_dydx(pb, W, v, u - 1));
    if (v >= 1 && u <= W - 2) g -= k.c2 * sgn(sm_dxdy(pb, W, v - 1, u)) + k.c3 * sgn(sm_dydx(pb, W, v - 1, u));
    if (v >= 1 && u >= 1) g += k.c2 * sgn(sm_dxdy(pb, W, v - 1, u - 1)) + k.c3 * sgn(sm_dydx(pb, W, v - 1, u - 1));
    gb[i] = g * go;
  }
}

SmoothCoef smooth_coef(int B, int H, int W, float weight) {
  SmoothCoef k;
  k.c1 = weight / ((float)B * H * (W - 2));
  k.c2 = weight / ((float)B * (H - 1) * (W - 1));
  k.c3 = k.c2;
  k.c4 = weight / ((float)B * (H - 2) * W);
  return k;
}

// ---------------------------------------------------------------------------------------------------
// compute_errors (loss_functions.py:401-448)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) depth_errors_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int H, int W,
                                                           float maxd, int y1, int y2, int x1, int x2,
                                                           const float* __restrict__ scale, int32_t* counters, double* sums) {
  const int b = blockIdx.y;
  const float* g = gt + (long long)b * H * W;
  const float* p = pred + (long long)b * H * W;
  const float sc = scale ? scale[b] : 1.f;
  int n = 0, a1 = 0, a2 = 0, a3 = 0;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  const float t1 = 1.25f, t2 = (float)(1.25 * 1.25), t3 = (float)(1.25 * 1.25 * 1.25);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    int y = i / W, x = i % W;
    float gv = g[i];
    if (!(gv > 0.f && gv < maxd) || y < y1 || y >= y2 || x < x1 || x >= x2) continue;
    float pv = fminf(fmaxf(p[i], 1e-3f), maxd);
    if (scale) pv = pv * sc;
    float th = fmaxf(gv / pv, pv / gv);
    n += 1;
    a1 += th < t1; a2 += th < t2; a3 += th < t3;
    float d = gv - pv;
    float lg = logf(gv) - logf(pv);
    s0 += (double)fabsf(d);
    s1 += (double)(fabsf(d) / gv);
    s2 += (double)(d * d / gv);
    s3 += (double)(d * d);
    s4 += (double)(lg * lg);
  }
  // integer counters: warp reduce then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if ((threadIdx.x & 31) == 0 && n) {
    atomicAdd(counters + 4 * b, n); atomicAdd(counters + 4 * b + 1, a1);
    atomicAdd(counters + 4 * b + 2, a2); atomicAdd(counters + 4 * b + 3, a3);
  }
  s0 = block_sum_d(s0); s1 = block_sum_d(s1); s2 = block_sum_d(s2); s3 = block_sum_d(s3); s4 = block_sum_d(s4);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 5 * b, s0); atomicAdd(sums + 5 * b + 1, s1); atomicAdd(sums + 5 * b + 2, s2);
    atomicAdd(sums + 5 * b + 3, s3); atomicAdd(sums + 5 * b + 4, s4);
  }
}

__global__ void area_down_kernel(const float* __restrict__ src, int NC, int H, int W, int f, float* __restrict__ dst) {
  const int h = H / f, w = W / f;
  long long total = (long long)NC * h * w;
  const float inv = 1.f / (float)(f * f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % w);
    long long q = i / w;
    int y = (int)(q % h);
    long long nc = q / h;
    const float* s = src + (nc * H + (long long)y * f) * W + (long long)x * f;
    float a = 0.f;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) a += s[dy * W + dx];
    dst[i] = a * inv;
  }
}

__global__ void explain_fwd_kernel(const float* __restrict__ m, long long n, float inv, float* loss) {
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s -= fmaxf(logf(m[i]), -100.f);
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(loss, s * inv);
}
__global__ void explain_bwd_kernel(const float* __restrict__ m, long long n, float inv, const float* __restrict__ gout, float* __restrict__ gm) {
  const float go = gout[0] * inv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = m[i];
    gm[i] = (v - 1.f) / fmaxf((1.f - v) * v, 1e-12f) * go;
  }
}

// ---------------------------------------------------------------------------------------------------
// inverse warp geometry (inverse_warp.py:26-193; grid_sample semantics of ATen GridSampler)
// ---------------------------------------------------------------------------------------------------
struct PoseMats {   // per batch element, built once per block in shared memory
  float Kinv[9];
  float K[9];
  float R[9];
  float t[3];
  float Prot[9];    // K @ R
  float Ptr[3];     // K @ t
};

__device__ void mat3_mul(const float* a, const float* b, float* c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

__device__ void euler_mats(const float* ang, float* X, float* Y, float* Z) {
  float cx = cosf(ang[0]), sx = sinf(ang[0]), cy = cosf(ang[1]), sy = sinf(ang[1]), cz = cosf(ang[2]), sz = sinf(ang[2]);
  float z[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float y[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  float x[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  for (int i = 0; i < 9; ++i) { X[i] = x[i]; Y[i] = y[i]; Z[i] = z[i]; }
}

__device__ void quat_rot(const float* q4, float* R) {   // q4 normalised (w,x,y,z)  (inverse_warp.py:128-137)
  float w = q4[0], x = q4[1], y = q4[2], z = q4[3];
  float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz; R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy; R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy; R[7] = 2 * wx + 2 * yz; R[8] = w2 - x2 - y2 + z2;
}

__device__ void build_pose(const float* pose, const float* K, const float* Kinv, int rot_mode, PoseMats& m) {
  for (int i = 0; i < 9; ++i) { m.K[i] = K[i]; m.Kinv[i] = Kinv[i]; }
  for (int i = 0; i < 3; ++i) m.t[i] = pose[i];
  if (rot_mode == 0) {
    float X[9], Y[9], Z[9], XY[9];
    euler_mats(pose + 3, X, Y, Z);
    mat3_mul(X, Y, XY);
    mat3_mul(XY, Z, m.R);
  } else {
    float q[4] = {1.f, pose[3], pose[4], pose[5]};
    float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= nrm;
    quat_rot(q, m.R);
  }
  mat3_mul(m.K, m.R, m.Prot);
  for (int i = 0; i < 3; ++i) m.Ptr[i] = m.K[i * 3] * m.t[0] + m.K[i * 3 + 1] * m.t[1] + m.K[i * 3 + 2] * m.t[2];
}

struct WarpPix {
  float cam[3];     // camera-frame point
  float ray[3];     // Kinv @ (u,v,1)
  float px, py, pz, Z;
  float ix, iy;     // unnormalised sampling position
  float dix_dxn, diy_dyn;   // multipliers incl. zeros-mode replacement and border clipping (0 when no gradient)
  bool zclamped;
  int x0, y0;
  float wx0, wx1, wy0, wy1;
  bool finite;
};

__device__ __forceinline__ void warp_pixel(const PoseMats& m, int u, int v, float d, int h, int w, int pad_mode, int align,
                                           WarpPix& o) {
  const float fu = (float)u, fv = (float)v;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o.ray[k] = m.Kinv[k * 3] * fu + m.Kinv[k * 3 + 1] * fv + m.Kinv[k * 3 + 2];
    o.cam[k] = o.ray[k] * d;
  }
  o.px = m.Prot[0] * o.cam[0] + m.Prot[1] * o.cam[1] + m.Prot[2] * o.cam[2] + m.Ptr[0];
  o.py = m.Prot[3] * o.cam[0] + m.Prot[4] * o.cam[1] + m.Prot[5] * o.cam[2] + m.Ptr[1];
  o.pz = m.Prot[6] * o.cam[0] + m.Prot[7] * o.cam[1] + m.Prot[8] * o.cam[2] + m.Ptr[2];
  o.zclamped = !(o.pz >= 1e-3f);
  o.Z = fmaxf(o.pz, 1e-3f);
  float xn = 2.f * (o.px / o.Z) / (float)(w - 1) - 1.f;
  float yn = 2.f * (o.py / o.Z) / (float)(h - 1) - 1.f;
  float gx = 1.f, gy = 1.f;
  if (pad_mode == 0) {
    if (xn > 1.f || xn < -1.f) { xn = 2.f; gx = 0.f; }
    if (yn > 1.f || yn < -1.f) { yn = 2.f; gy = 0.f; }
  }
  float ix, iy;
  if (align) {
    ix = ((xn + 1.f) / 2.f) * (float)(w - 1); gx *= 0.5f * (float)(w - 1);
    iy = ((yn + 1.f) / 2.f) * (float)(h - 1); gy *= 0.5f * (float)(h - 1);
  } else {
    ix = ((xn + 1.f) * (float)w - 1.f) / 2.f; gx *= 0.5f * (float)w;
    iy = ((yn + 1.f) * (float)h - 1.f) / 2.f; gy *= 0.5f * (float)h;
  }
  if (pad_mode == 1) {   // border: clip_coordinates(_set_grad)
    if (!(ix > 0.f)) { ix = 0.f; gx = 0.f; } else if (ix >= (float)(w - 1)) { ix = (float)(w - 1); gx = 0.f; }
    if (!(iy > 0.f)) { iy = 0.f; gy = 0.f; } else if (iy >= (float)(h - 1)) { iy = (float)(h - 1); gy = 0.f; }
  }
  o.ix = ix; o.iy = iy; o.dix_dxn = gx; o.diy_dyn = gy;
  o.finite = isfinite(ix) && isfinite(iy) && fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
  float fx = floorf(ix), fy = floorf(iy);
  o.x0 = o.finite ? (int)fx : -10; o.y0 = o.finite ? (int)fy : -10;
  o.wx1 = ix - fx; o.wx0 = (fx + 1.f) - ix;
  o.wy1 = iy - fy; o.wy0 = (fy + 1.f) - iy;
}

// bilinear fetch of C channels (channel stride cs); returns values and, optionally, d/dix, d/diy
template <int C, bool GRAD>
__device__ __forceinline__ void bil_fetch(const float* __restrict__ img, long long cs, int h, int w, const WarpPix& p, float* val,
                                          float* dvx, float* dvy) {
  const int x0 = p.x0, y0 = p.y0, x1 = x0 + 1, y1 = y0 + 1;
  const bool bx0 = x0 >= 0 && x0 < w, bx1 = x1 >= 0 && x1 < w, by0 = y0 >= 0 && y0 < h, by1 = y1 >= 0 && y1 < h;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* ic = img + c * cs;
    float nw = (bx0 && by0) ? ic[y0 * w + x0] : 0.f;
    float ne = (bx1 && by0) ? ic[y0 * w + x1] : 0.f;
    float sw = (bx0 && by1) ? ic[y1 * w + x0] : 0.f;
    float se = (bx1 && by1) ? ic[y1 * w + x1] : 0.f;
    val[c] = nw * (p.wx0 * p.wy0) + ne * (p.wx1 * p.wy0) + sw * (p.wx0 * p.wy1) + se * (p.wx1 * p.wy1);
    if (GRAD) {
      dvx[c] = -nw * p.wy0 + ne * p.wy0 - sw * p.wy1 + se * p.wy1;
      dvy[c] = -nw * p.wx0 - ne * p.wx1 + sw * p.wx0 + se * p.wx1;
    }
  }
}

// chain d(loss)/d(ix,iy) -> depth gradient and the 12 per-batch pose accumulators (q = K^T gp ; G = q cam^T)
__device__ __forceinline__ void warp_chain(const PoseMats& m, const WarpPix& p, int h, int w, float gix, float giy, float& gdepth,
                                           float* acc12) {
  float gxn = gix * p.dix_dxn, gyn = giy * p.diy_dyn;
  float ax = 2.f / (float)(w - 1), ay = 2.f / (float)(h - 1);
  float gpx = gxn * ax / p.Z;
  float gpy = gyn * ay / p.Z;
  float gZ = -(gxn * ax * p.px + gyn * ay * p.py) / (p.Z * p.Z);
  float gpz = p.zclamped ? 0.f : gZ;
  // p = Prot cam + Ptr ; cam = ray * d
  float gc0 = m.Prot[0] * gpx + m.Prot[3] * gpy + m.Prot[6] * gpz;
  float gc1 = m.Prot[1] * gpx + m.Prot[4] * gpy + m.Prot[7] * gpz;
  float gc2 = m.Prot[2] * gpx + m.Prot[5] * gpy + m.Prot[8] * gpz;
  gdepth = gc0 * p.ray[0] + gc1 * p.ray[1] + gc2 * p.ray[2];
  float q0 = m.K[0] * gpx + m.K[3] * gpy + m.K[6] * gpz;
  float q1 = m.K[1] * gpx + m.K[4] * gpy + m.K[7] * gpz;
  float q2 = m.K[2] * gpx + m.K[5] * gpy + m.K[8] * gpz;
  acc12[0] += q0; acc12[1] += q1; acc12[2] += q2;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    acc12[3 + j] += q0 * p.cam[j];
    acc12[6 + j] += q1 * p.cam[j];
    acc12[9 + j] += q2 * p.cam[j];
  }
}

__device__ void block_reduce12(float* acc12, float* dst) {
  __shared__ float red12[12][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float v = dn_warp_sum(acc12[i]);
    if (lane == 0) red12[i][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red12[threadIdx.x][k];
    atomicAdd(dst + threadIdx.x, t);
  }
}

// (q, G) -> pose gradient for one batch element
__device__ void pose_grad_chain(const float* a, const float* ps, int rot_mode, float* g) {
  g[0] += a[0]; g[1] += a[1]; g[2] += a[2];
  const float* G = a + 3;
  if (rot_mode == 0) {
    float X[9], Y[9], Z[9];
    euler_mats(ps + 3, X, Y, Z);
    float cx = X[4], sx = X[7], cy = Y[0], sy = Y[2], cz = Z[0], sz = Z[3];
    float dX[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx};
    float dY[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy};
    float dZ[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
    float T1[9], T2[9];
    float r;
    mat3_mul(dX, Y, T1); mat3_mul(T1, Z, T2);
    r = 0.f; for (int i = 0; i < 9; ++i) r += G[i] * T2[i];
    g[3] += r;
    mat3_mul(X, dY, T1); mat3_mul(T1, Z, T2);
    r = 0.f; for (int i = 0; i < 9; ++i) r += G[i] * T2[i];
    g[4] += r;
    mat3_mul(X, Y, T1); mat3_mul(T1, dZ, T2);
    r = 0.f; for (int i = 0; i < 9; ++i) r += G[i] * T2[i];
    g[5] += r;
  } else {
    float q[4] = {1.f, ps[3], ps[4], ps[5]};
    float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] / nrm, x = q[1] / nrm, y = q[2] / nrm, z = q[3] / nrm;
    // dR/d(w,x,y,z) contracted with G
    float gw = G[0] * 2 * w + G[1] * (-2 * z) + G[2] * (2 * y) + G[3] * (2 * z) + G[4] * 2 * w + G[5] * (-2 * x) + G[6] * (-2 * y) + G[7] * (2 * x) + G[8] * 2 * w;
    float gx = G[0] * 2 * x + G[1] * (2 * y) + G[2] * (2 * z) + G[3] * (2 * y) + G[4] * (-2 * x) + G[5] * (-2 * w) + G[6] * (2 * z) + G[7] * (2 * w) + G[8] * (-2 * x);
    float gy = G[0] * (-2 * y) + G[1] * (2 * x) + G[2] * (2 * w) + G[3] * (2 * x) + G[4] * (2 * y) + G[5] * (2 * z) + G[6] * (-2 * w) + G[7] * (2 * z) + G[8] * (-2 * y);
    float gz = G[0] * (-2 * z) + G[1] * (-2 * w) + G[2] * (2 * x) + G[3] * (2 * w) + G[4] * (-2 * z) + G[5] * (2 * y) + G[6] * (2 * x) + G[7] * (2 * y) + G[8] * (2 * z);
    float dot = gw * w + gx * x + gy * y + gz * z;
    g[3] += (gx - x * dot) / nrm;
    g[4] += (gy - y * dot) / nrm;
    g[5] += (gz - z * dot) / nrm;
  }
}

__global__ void pose_grad_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ pose, int pose_stride, int B,
                                          int rot_mode, float* gpose) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  pose_grad_chain(ws + 12 * b, pose + (long long)b * pose_stride, rot_mode, gpose + (long long)b * pose_stride);
}

__global__ void __launch_bounds__(256) warp_photo_fwd_kernel(const float* __restrict__ tgt, const float* __restrict__ ref,
                                                             const float* __restrict__ depth, const float* __restrict__ pose,
                                                             int pose_stride, const float* __restrict__ K,
                                                             const float* __restrict__ Kinv, const float* __restrict__ mask,
                                                             long long mask_bs, int h, int w, int rot_mode, int pad_mode, int align,
                                                             float* warped, float inv_count, float* loss, int32_t* nanflag) {
  __shared__ PoseMats m;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) build_pose(pose + (long long)b * pose_stride, K + 9 * b, Kinv + 9 * b, rot_mode, m);
  __syncthreads();
  const long long hw = (long long)h * w;
  const float* tb = tgt + 3 * hw * b;
  const float* rb = ref + 3 * hw * b;
  float s = 0.f;
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    int v = i / w, u = i % w;
    WarpPix p;
    warp_pixel(m, u, v, depth[hw * b + i], h, w, pad_mode, align, p);
    float val[3];
    bil_fetch<3, false>(rb, hw, h, w, p, val, nullptr, nullptr);
    if (warped) { warped[3 * hw * b + i] = val[0]; warped[3 * hw * b + hw + i] = val[1]; warped[3 * hw * b + 2 * hw + i] = val[2]; }
    float oob = (val[0] == 0.f && val[1] == 0.f && val[2] == 0.f) ? 0.f : 1.f;
    float mk = mask ? mask[mask_bs * b + i] : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float e = (tb[c * hw + i] - val[c]) * oob;
      if (mask) e *= mk;
      s += fabsf(e);
      bad |= (e != e);
    }
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(loss, s * inv_count);
  if (bad && nanflag) atomicOr(nanflag, 1);
}

__global__ void __launch_bounds__(256) warp_photo_bwd_kernel(const float* __restrict__ tgt, const float* __restrict__ ref,
                                                             const float* __restrict__ depth, const float* __restrict__ pose,
                                                             int pose_stride, const float* __restrict__ K,
                                                             const float* __restrict__ Kinv, const float* __restrict__ mask,
                                                             long long mask_bs, int h, int w, int rot_mode, int pad_mode, int align,
                                                             float inv_count, const float* __restrict__ gout, float* __restrict__ gdepth,
                                                             float* ws, float* gmask, long long gmask_bs) {
  __shared__ PoseMats m;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) build_pose(pose + (long long)b * pose_stride, K + 9 * b, Kinv + 9 * b, rot_mode, m);
  __syncthreads();
  const long long hw = (long long)h * w;
  const float* tb = tgt + 3 * hw * b;
  const float* rb = ref + 3 * hw * b;
  const float go = gout[0] * inv_count;
  float acc12[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc12[i] = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    int v = i / w, u = i % w;
    WarpPix p;
    warp_pixel(m, u, v, depth[hw * b + i], h, w, pad_mode, align, p);
    float val[3], dvx[3], dvy[3];
    bil_fetch<3, true>(rb, hw, h, w, p, val, dvx, dvy);
    float oob = (val[0] == 0.f && val[1] == 0.f && val[2] == 0.f) ? 0.f : 1.f;
    float mk = mask ? mask[mask_bs * b + i] : 1.f;
    float gix = 0.f, giy = 0.f, gm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float e0 = (tb[c * hw + i] - val[c]) * oob;
      float e = mask ? e0 * mk : e0;
      float sg = sgn(e) * go;          // dL/de
      gm += sg * e0;
      float gW = -sg * oob * (mask ? mk : 1.f);
      gix += gW * dvx[c];
      giy += gW * dvy[c];
    }
    float gd;
    warp_chain(m, p, h, w, gix, giy, gd, acc12);
    gdepth[hw * b + i] += gd;   // accumulated over reference frames; caller zeroes
    if (gmask) gmask[gmask_bs * b + i] = gm;
  }
  block_reduce12(acc12, ws + 12 * b);
}


// ---------------------------------------------------------------------------------------------------
// batched photometric loss: area pyramid, all (scale, reference) pairs forward, all backward -- three launches
// ---------------------------------------------------------------------------------------------------
struct PyrJobs { dn_pyr_job j[8]; };
// one thread per 8x8 block of a full-size plane: 16 float4 loads, /2 level as 4 float4 stores, /4 as 2 float2, /8 as one float
__global__ void __launch_bounds__(256) area_pyramid_kernel(PyrJobs jobs, long long NC, int H, int W) {
  const dn_pyr_job jb = jobs.j[blockIdx.y];
  const int bw = W >> 3, bh = H >> 3;
  const long long total = NC * bh * bw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int bx = (int)(i % bw);
    const long long q = i / bw;
    const int by = (int)(q % bh);
    const long long nc = q / bh;
    const float* s = jb.src + (nc * H + (long long)by * 8) * W + bx * 8;
    float l1v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(s + (2 * r) * W), a1 = *reinterpret_cast<const float4*>(s + (2 * r) * W + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s + (2 * r + 1) * W), b1 = *reinterpret_cast<const float4*>(s + (2 * r + 1) * W + 4);
      l1v[r][0] = ((a0.x + a0.y) + (b0.x + b0.y)) * 0.25f;
      l1v[r][1] = ((a0.z + a0.w) + (b0.z + b0.w)) * 0.25f;
      l1v[r][2] = ((a1.x + a1.y) + (b1.x + b1.y)) * 0.25f;
      l1v[r][3] = ((a1.z + a1.w) + (b1.z + b1.w)) * 0.25f;
    }
    const int W1 = W >> 1, W2 = W >> 2, W3 = W >> 3;
    if (jb.l1) {
      float* d = jb.l1 + (nc * (H >> 1) + (long long)by * 4) * W1 + bx * 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(d + r * W1) = make_float4(l1v[r][0], l1v[r][1], l1v[r][2], l1v[r][3]);
    }
    float l2v[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c)
        l2v[r][c] = ((l1v[2 * r][2 * c] + l1v[2 * r][2 * c + 1]) + (l1v[2 * r + 1][2 * c] + l1v[2 * r + 1][2 * c + 1])) * 0.25f;
    if (jb.l2) {
      float* d = jb.l2 + (nc * (H >> 2) + (long long)by * 2) * W2 + bx * 2;
      *reinterpret_cast<float2*>(d) = make_float2(l2v[0][0], l2v[0][1]);
      *reinterpret_cast<float2*>(d + W2) = make_float2(l2v[1][0], l2v[1][1]);
    }
    if (jb.l3) jb.l3[(nc * (H >> 3) + by) * W3 + bx] = ((l2v[0][0] + l2v[0][1]) + (l2v[1][0] + l2v[1][1])) * 0.25f;
  }
}

constexpr int PB_NBX = 52;      // blocks per (scale, sample): 52 x 256 threads x 4 pixels = one 128x416 plane

__device__ __forceinline__ void pb_build_poses(const dn_photo_batch& P, const dn_photo_scale& sc, int b, PoseMats* m) {
  if ((int)threadIdx.x < P.nrefs) {
    float Ks[9], Kis[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float k = P.K[9 * b + i], ki = P.Kinv[9 * b + i];
      Ks[i] = i < 6 ? k / sc.downscale : k;                      // intrinsics[:, 0:2] / downscale (:329)
      Kis[i] = (i % 3) < 2 ? ki * sc.downscale : ki;             // intrinsics_inv[:, :, 0:2] * downscale (:330)
    }
    build_pose(P.pose + ((long long)b * P.nrefs + threadIdx.x) * 6, Ks, Kis, P.rot_mode, m[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(256, 3) photo_batch_fwd_kernel(const __grid_constant__ dn_photo_batch P, float* __restrict__ part,
                                                              int32_t* nanflag) {
  __shared__ PoseMats m[DN_PHOTO_MAX_REFS];
  const int s_ = blockIdx.z, b = blockIdx.y;
  const dn_photo_scale& sc = P.sc[s_];
  const int h = sc.h, w = sc.w;
  const long long hw = (long long)h * w;
  const int nq = (int)((hw + 3) >> 2);
  float acc = 0.f;
  bool bad = false;
  if ((int)(blockIdx.x * blockDim.x) < nq) {          // (uniform per block: coarse levels need only the first few blocks)
    pb_build_poses(P, sc, b, m);
    __syncthreads();
    const float* tb = sc.tgt + 3 * hw * b;
    const float* db = sc.depth + hw * b;
    const bool vec = (w & 3) == 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
      const int i0 = q << 2;
      const int v0 = i0 / w, u0 = i0 - v0 * w;       // w % 4 == 0: the four pixels share the image row
      float t[3][4], d[4];
      if (vec) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float4 v4 = *reinterpret_cast<const float4*>(tb + c * hw + i0);
          t[c][0] = v4.x; t[c][1] = v4.y; t[c][2] = v4.z; t[c][3] = v4.w;
        }
        const float4 d4 = *reinterpret_cast<const float4*>(db + i0);
        d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool in = i0 + k < hw;
          d[k] = in ? db[i0 + k] : 1.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) t[c][k] = in ? tb[c * hw + i0 + k] : 0.f;
        }
      }
      for (int r = 0; r < P.nrefs; ++r) {
        const float* rb = sc.ref[r] + 3 * hw * b;
        float mk[4] = {1.f, 1.f, 1.f, 1.f};
        if (sc.mask) {
          const float* mp = sc.mask + ((long long)b * P.nrefs + r) * hw + i0;
          if (vec) { const float4 m4 = *reinterpret_cast<const float4*>(mp); mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w; }
          else
#pragma unroll
            for (int k = 0; k < 4; ++k) mk[k] = i0 + k < hw ? mp[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k;
          if (i >= hw) break;
          WarpPix p;
          warp_pixel(m[r], vec ? u0 + k : i % w, vec ? v0 : i / w, d[k], h, w, P.pad_mode, P.align_corners, p);
          float val[3];
          bil_fetch<3, false>(rb, hw, h, w, p, val, nullptr, nullptr);
          const float oob = (val[0] == 0.f && val[1] == 0.f && val[2] == 0.f) ? 0.f : 1.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float e = (t[c][k] - val[c]) * oob;
            if (sc.mask) e *= mk[k];
            acc += fabsf(e);
            bad |= (e != e);
          }
        }
      }
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) part[((long long)s_ * P.B + b) * gridDim.x + blockIdx.x] = acc;
  if (bad && nanflag) atomicOr(nanflag, 1);
}
// one warp per scale folds that scale's partials in a fixed order; thread 0 adds the weighted total to the loss
__global__ void photo_fold_kernel(const __grid_constant__ dn_photo_batch P, const float* __restrict__ part, int nbx, float* loss) {
  __shared__ double tot[DN_PHOTO_MAX_SCALES];
  const int s_ = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (s_ < P.nscales) {
    const int n = P.B * nbx;
    double a = 0.0;
    for (int k = lane; k < n; k += 32) a += (double)part[(long long)s_ * n + k];
    a = dn_warp_sum_d(a);
    if (lane == 0) tot[s_] = a / ((double)P.B * 3.0 * P.sc[s_].h * P.sc[s_].w);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < P.nscales; ++k) t += tot[k];
    loss[0] += (float)t;
  }
}

__device__ void block_reduce12_row(float* acc12, float* dst) {     // like block_reduce12, written (not added) to a partial row
  __shared__ float red12r[12][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float v = dn_warp_sum(acc12[i]);
    if (lane == 0) red12r[i][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red12r[threadIdx.x][k];
    dst[threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) photo_batch_bwd_kernel(const __grid_constant__ dn_photo_batch P, const float* __restrict__ gout,
                                                              float* __restrict__ part12) {
  __shared__ PoseMats m[DN_PHOTO_MAX_REFS];
  const int s_ = blockIdx.z, b = blockIdx.y;
  const dn_photo_scale& sc = P.sc[s_];
  const int h = sc.h, w = sc.w;
  const long long hw = (long long)h * w;
  const int nq = (int)((hw + 3) >> 2);
  const bool active = (int)(blockIdx.x * blockDim.x) < nq;
  if (active) pb_build_poses(P, sc, b, m);
  __syncthreads();
  const float go = gout[0] / ((float)P.B * 3.f * (float)h * (float)w);
  const float* tb = sc.tgt + 3 * hw * b;
  const float* db = sc.depth + hw * b;
  const bool vec = (w & 3) == 0;
  for (int r = 0; r < P.nrefs; ++r) {
    // one reference frame at a time keeps 12 pose accumulators live; the depth gradient of the quad is carried across the
    // frames through gdepth itself only when R > 1 (first frame writes, later frames add: same thread, same address)
    float acc12[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc12[i] = 0.f;
    if (active) {
      const float* rb = sc.ref[r] + 3 * hw * b;
      for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const int i0 = q << 2;
        const int v0 = i0 / w, u0 = i0 - v0 * w;
        float t[3][4], d[4], mk[4] = {1.f, 1.f, 1.f, 1.f}, gd[4] = {0.f, 0.f, 0.f, 0.f}, gm[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float4 v4 = *reinterpret_cast<const float4*>(tb + c * hw + i0);
            t[c][0] = v4.x; t[c][1] = v4.y; t[c][2] = v4.z; t[c][3] = v4.w;
          }
          const float4 d4 = *reinterpret_cast<const float4*>(db + i0);
          d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool in = i0 + k < hw;
            d[k] = in ? db[i0 + k] : 1.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) t[c][k] = in ? tb[c * hw + i0 + k] : 0.f;
          }
        }
        const float* mp = sc.mask ? sc.mask + ((long long)b * P.nrefs + r) * hw + i0 : nullptr;
        if (mp) {
          if (vec) { const float4 m4 = *reinterpret_cast<const float4*>(mp); mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w; }
          else
#pragma unroll
            for (int k = 0; k < 4; ++k) mk[k] = i0 + k < hw ? mp[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k;
          if (i >= hw) break;
          WarpPix p;
          warp_pixel(m[r], vec ? u0 + k : i % w, vec ? v0 : i / w, d[k], h, w, P.pad_mode, P.align_corners, p);
          float val[3], dvx[3], dvy[3];
          bil_fetch<3, true>(rb, hw, h, w, p, val, dvx, dvy);
          const float oob = (val[0] == 0.f && val[1] == 0.f && val[2] == 0.f) ? 0.f : 1.f;
          float gix = 0.f, giy = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float e0 = (t[c][k] - val[c]) * oob;
            const float e = mp ? e0 * mk[k] : e0;
            const float sg = sgn(e) * go;
            gm[k] += sg * e0;
            const float gW = -sg * oob * (mp ? mk[k] : 1.f);
            gix += gW * dvx[c];
            giy += gW * dvy[c];
          }
          warp_chain(m[r], p, h, w, gix, giy, gd[k], acc12);
        }
        float* gp = sc.gdepth + hw * b + i0;
        if (vec) {
          float4 o = make_float4(gd[0], gd[1], gd[2], gd[3]);
          if (r > 0) { const float4 pr = *reinterpret_cast<const float4*>(gp); o.x += pr.x; o.y += pr.y; o.z += pr.z; o.w += pr.w; }
          *reinterpret_cast<float4*>(gp) = o;
          if (sc.gmask) *reinterpret_cast<float4*>(sc.gmask + ((long long)b * P.nrefs + r) * hw + i0) = make_float4(gm[0], gm[1], gm[2], gm[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (i0 + k < hw) {
              gp[k] = r > 0 ? gp[k] + gd[k] : gd[k];
              if (sc.gmask) sc.gmask[((long long)b * P.nrefs + r) * hw + i0 + k] = gm[k];
            }
        }
      }
    }
    block_reduce12_row(acc12, part12 + ((((long long)s_ * P.nrefs + r) * P.B + b) * gridDim.x + blockIdx.x) * 12);
  }
}
// one warp per (sample, reference frame): lanes stride over the partial rows of all scales and blocks, a shuffle tree adds them
// (fixed order: deterministic), lane 0 chains the 12 sums to the 6-DoF pose gradient
__global__ void __launch_bounds__(128) photo_pose_finalize_kernel(const __grid_constant__ dn_photo_batch P, const float* __restrict__ part12,
                                                                  int nbx, float* __restrict__ gpose) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= P.B * P.nrefs) return;
  const int b = i / P.nrefs, r = i % P.nrefs;
  float a[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) a[k] = 0.f;
  for (int s_ = 0; s_ < P.nscales; ++s_) {
    const float* row = part12 + ((((long long)s_ * P.nrefs + r) * P.B + b) * nbx) * 12;
    for (int x = lane; x < nbx; x += 32)
#pragma unroll
      for (int k = 0; k < 12; ++k) a[k] += row[x * 12 + k];
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) a[k] = dn_warp_sum(a[k]);
  if (lane == 0) {
    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    pose_grad_chain(a, P.pose + (long long)i * 6, P.rot_mode, g);
    for (int k = 0; k < 6; ++k) gpose[(long long)i * 6 + k] = g[k];
  }
}

template <int CMAX>
__global__ void __launch_bounds__(256) inverse_warp_fwd_kernel(const float* __restrict__ img, const float* __restrict__ depth,
                                                               const float* __restrict__ pose, const float* __restrict__ K,
                                                               const float* __restrict__ Kinv, int C, int h, int w, int rot_mode,
                                                               int pad_mode, int align, float* __restrict__ out) {
  __shared__ PoseMats m;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) build_pose(pose + 6 * b, K + 9 * b, Kinv + 9 * b, rot_mode, m);
  __syncthreads();
  const long long hw = (long long)h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    WarpPix p;
    warp_pixel(m, i % w, i / w, depth[hw * b + i], h, w, pad_mode, align, p);
    for (int c = 0; c < C; ++c) {
      float val[1];
      bil_fetch<1, false>(img + ((long long)b * C + c) * hw, hw, h, w, p, val, nullptr, nullptr);
      out[((long long)b * C + c) * hw + i] = val[0];
    }
  }
}

__global__ void __launch_bounds__(256) inverse_warp_bwd_kernel(const float* __restrict__ img, const float* __restrict__ depth,
                                                               const float* __restrict__ pose, const float* __restrict__ K,
                                                               const float* __restrict__ Kinv, int C, int h, int w, int rot_mode,
                                                               int pad_mode, int align, const float* __restrict__ gout, float* gimg,
                                                               float* __restrict__ gdepth, float* ws) {
  __shared__ PoseMats m;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) build_pose(pose + 6 * b, K + 9 * b, Kinv + 9 * b, rot_mode, m);
  __syncthreads();
  const long long hw = (long long)h * w;
  float acc12[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc12[i] = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    WarpPix p;
    warp_pixel(m, i % w, i / w, depth[hw * b + i], h, w, pad_mode, align, p);
    float gix = 0.f, giy = 0.f;
    const int x0 = p.x0, y0 = p.y0, x1 = x0 + 1, y1 = y0 + 1;
    const bool bx0 = x0 >= 0 && x0 < w, bx1 = x1 >= 0 && x1 < w, by0 = y0 >= 0 && y0 < h, by1 = y1 >= 0 && y1 < h;
    for (int c = 0; c < C; ++c) {
      float val[1], dvx[1], dvy[1];
      const long long base = ((long long)b * C + c) * hw;
      bil_fetch<1, true>(img + base, hw, h, w, p, val, dvx, dvy);
      float g = gout[base + i];
      gix += g * dvx[0];
      giy += g * dvy[0];
      if (gimg) {
        if (bx0 && by0) atomicAdd(gimg + base + y0 * w + x0, g * p.wx0 * p.wy0);
        if (bx1 && by0) atomicAdd(gimg + base + y0 * w + x1, g * p.wx1 * p.wy0);
        if (bx0 && by1) atomicAdd(gimg + base + y1 * w + x0, g * p.wx0 * p.wy1);
        if (bx1 && by1) atomicAdd(gimg + base + y1 * w + x1, g * p.wx1 * p.wy1);
      }
    }
    float gd;
    warp_chain(m, p, h, w, gix, giy, gd, acc12);
    if (gdepth) gdepth[hw * b + i] = gd;
  }
  block_reduce12(acc12, ws + 12 * b);
}

// ---------------------------------------------------------------------------------------------------
// Supervised depth losses on the masked-reduce pattern (loss_functions.py:77-315): l1 / l2 / berhu / Scale_invariant per
// sample, and their Multiscale_* forms (one mask over the whole batch, B = 1 here).  Deterministic: every block writes its
// partial row, one block per sample folds the rows in a fixed order in double (no float atomics).
//   ws layout (floats): [B][DL_NSTAT] folded statistics, then [B][nblk][DL_NPART] partial rows.
//   statistics per sample: 0 count, 1 sum|d|, 2 sum d^2, 3 sum d (d = gt - pred_clamped), 4 max|d|, 5 sum berhu, 6 sum dberhu/dc
// ---------------------------------------------------------------------------------------------------
constexpr int DL_NSTAT = 8, DL_NPART = 6, DL_NBLK = 64;
enum { DL_L1 = 0, DL_L2 = 1, DL_BERHU = 2, DL_SCALE_INV = 3 };

__device__ __forceinline__ float block_max(float v) {
  __shared__ float redm[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? redm[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  }
  return t;
}

// optional x f up-sampling of pred (Multiscale_FULL_L1_loss :224-241): value of F.upsample(pred, scale_factor=f) at (y, x)
struct UpSrc { int f, mode, h, w; };     // mode 0 nearest, 1 bilinear (align_corners=False); f == 1: pred is full size
__device__ __forceinline__ float up_sample(const float* __restrict__ p, const UpSrc u, int y, int x, int* i00, float* wts) {
  if (u.f == 1) { i00[0] = y * u.w + x; i00[1] = i00[2] = i00[3] = -1; wts[0] = 1.f; return p[i00[0]]; }
  if (u.mode == 0) { i00[0] = (y / u.f) * u.w + x / u.f; i00[1] = i00[2] = i00[3] = -1; wts[0] = 1.f; return p[i00[0]]; }
  const float inv = 1.f / (float)u.f;
  float sy = ((float)y + 0.5f) * inv - 0.5f, sx = ((float)x + 0.5f) * inv - 0.5f;
  sy = sy < 0.f ? 0.f : sy; sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < u.h - 1 ? 1 : 0), x1 = x0 + (x0 < u.w - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  i00[0] = y0 * u.w + x0; i00[1] = y0 * u.w + x1; i00[2] = y1 * u.w + x0; i00[3] = y1 * u.w + x1;
  wts[0] = (1.f - ly) * (1.f - lx); wts[1] = (1.f - ly) * lx; wts[2] = ly * (1.f - lx); wts[3] = ly * lx;
  return wts[0] * p[i00[0]] + wts[1] * p[i00[1]] + wts[2] * p[i00[2]] + wts[3] * p[i00[3]];
}

// pass 0: count, sum|d|, sum d^2, sum d, max|d|;  pass 1 (berhu): sum berhu(|d|, c), sum d berhu / dc  with c = 0.2 max|d|
__global__ void __launch_bounds__(256) dl_partial_kernel(const float* __restrict__ gt, const float* __restrict__ pred, long long HW,
                                                         long long pHW, int W, UpSrc up, float maxd, int pass, float* __restrict__ ws, int B,
                                                         int joint) {
  const int b = blockIdx.y;
  const float* g = gt + (long long)b * HW;
  const float* p = pred + (long long)b * pHW;
  const float* stat = ws + (long long)(joint ? 0 : b) * DL_NSTAT;
  float* part = ws + (long long)B * DL_NSTAT + ((long long)b * gridDim.x + blockIdx.x) * DL_NPART;
  const float c = pass ? 0.2f * stat[4] : 0.f;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, mx = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const float gv = g[i];
    if (gv > 0.f && gv < maxd) {
      int idx[4]; float wt[4];
      const float raw = up.f == 1 ? p[i] : up_sample(p, up, (int)(i / W), (int)(i % W), idx, wt);
      const float pv = fminf(fmaxf(raw, 1e-3f), maxd);
      const float d = gv - pv, r = fabsf(d);
      if (!pass) { a0 += 1.f; a1 += r; a2 += d * d; a3 += d; mx = fmaxf(mx, r); }
      else if (r > c) { a0 += (r * r + c * c) / (2.f * c); a1 += 0.5f * (1.f - (r * r) / (c * c)); }
      else a0 += r;
    }
  }
  a0 = block_sum(a0); a1 = block_sum(a1); a2 = block_sum(a2); a3 = block_sum(a3); mx = block_max(mx);
  if (threadIdx.x == 0) { part[0] = a0; part[1] = a1; part[2] = a2; part[3] = a3; part[4] = mx; }
}
// fold + finalize in one launch: warp w folds the partial rows of samples w, w + nwarps, ... (lanes stride over the rows, shuffle
// tree: a fixed order, double) into the sample's statistics; after the last pass warp 0 adds the per-sample values in sample
// order.  (joint: one mask over the whole batch, Multiscale_* losses -- all rows fold into sample 0.)
__device__ __forceinline__ float dl_value(const float* st, int kind) {
  const float n = st[0];
  if (kind == DL_L1) return st[1] / n;                       // 0/0 -> NaN like mean() of an empty selection
  if (kind == DL_L2) return st[2] / n;
  if (kind == DL_BERHU) return st[5] / n;
  return st[2] / n - 0.5f * (st[3] * st[3]) / (n * n);       // Scale_invariant_loss :166
}
__global__ void __launch_bounds__(1024) dl_fold_kernel(float* ws, int B, int nblk, int pass, int joint, int final, int kind, float weight,
                                                       int accumulate, float* loss) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nfold = joint ? 1 : B, rows = joint ? nblk * B : nblk;
  for (int b = warp; b < nfold; b += nwarps) {
    const float* part = ws + (long long)B * DL_NSTAT + (long long)b * nblk * DL_NPART;
    float* stat = ws + (long long)b * DL_NSTAT;
    double s[4] = {0, 0, 0, 0};
    float mx = 0.f;
    for (int k = lane; k < rows; k += 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] += (double)part[k * DL_NPART + j];
      mx = fmaxf(mx, part[k * DL_NPART + 4]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = dn_warp_sum_d(s[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
      if (!pass) { stat[0] = (float)s[0]; stat[1] = (float)s[1]; stat[2] = (float)s[2]; stat[3] = (float)s[3]; stat[4] = mx; }
      else { stat[5] = (float)s[0]; stat[6] = (float)s[1]; }
    }
  }
  if (!final) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int b = 0; b < nfold; ++b) t += dl_value(ws + (long long)b * DL_NSTAT, kind);
    t = weight * t / (float)nfold;
    loss[0] = accumulate ? loss[0] + t : t;
  }
}
__global__ void __launch_bounds__(256) dl_bwd_kernel(const float* __restrict__ gt, const float* __restrict__ pred, long long HW,
                                                     long long pHW, int W, UpSrc up, float maxd, int kind, float weight,
                                                     const float* __restrict__ ws, int B, int joint, const float* __restrict__ gout,
                                                     float* __restrict__ gpred) {
  const int b = blockIdx.y;
  const float* g = gt + (long long)b * HW;
  const float* p = pred + (long long)b * pHW;
  float* o = gpred + (long long)b * pHW;
  const float* st = ws + (long long)(joint ? 0 : b) * DL_NSTAT;
  const float n = st[0], k = gout[0] * weight / (float)(joint ? 1 : B);
  const float mxr = st[4], c = 0.2f * mxr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const float gv = g[i];
    float r = 0.f;
    int idx[4] = {(int)i, -1, -1, -1};
    float wt[4] = {1.f, 0.f, 0.f, 0.f};
    if (gv > 0.f && gv < maxd) {
      const float raw = up.f == 1 ? p[i] : up_sample(p, up, (int)(i / W), (int)(i % W), idx, wt);
      if (raw >= 1e-3f && raw <= maxd) {                       // clamp passes the gradient on [min, max]
        const float d = raw - gv, ad = fabsf(d), sg = sgn(d);  // d(pred - gt)
        if (kind == DL_L1) r = k * sg / n;
        else if (kind == DL_L2) r = k * 2.f * d / n;
        else if (kind == DL_SCALE_INV) r = k * (2.f * d / n + st[3] / (n * n));
        else {
          r = k * sg * (ad > c ? ad / c : 1.f) / n;
          if (ad == mxr) r += k * st[6] * 0.2f * sg / n;       // through c = 0.2 * max|d| (torch.max routes it to the arg-max)
        }
      }
    }
    if (up.f == 1) o[i] = r;
    else if (r != 0.f) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (idx[j] >= 0 && wt[j] != 0.f) atomicAdd(o + idx[j], r * wt[j]);
    }
  }
}
// F.max_pool2d(x, 2, 2) / F.avg_pool2d(x, 2, 2) (= F.interpolate(scale_factor=0.5, 'bilinear', align_corners=False)) of the
// ground-truth pyramid (loss_functions.py:185-215)
__global__ void __launch_bounds__(256) pool2_kernel(const float* __restrict__ src, long long NC, int H, int W, int mode, float* __restrict__ dst) {
  const int h = H / 2, w = W / 2;
  const long long total = NC * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long q = i / w;
    const int y = (int)(q % h);
    const float* r = src + (q / h) * (long long)H * W + (long long)(2 * y) * W + 2 * x;
    const float a = r[0], b2 = r[1], c2 = r[W], d = r[W + 1];
    dst[i] = mode ? fmaxf(fmaxf(a, b2), fmaxf(c2, d)) : 0.5f * (0.5f * a + 0.5f * b2) + 0.5f * (0.5f * c2 + 0.5f * d);
  }
}


}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================

DN_EXPORT int64_t dn_depth_loss_ws_floats(int B) { return (int64_t)B * DL_NSTAT + (int64_t)B * DL_NBLK * DL_NPART; }

static int dl_check(const float* gt, const float* pred, int B, int H, int W, int upf, int kind) {
  if (!gt || !pred || B < 1 || H < 1 || W < 1 || upf < 1 || (H % upf) || (W % upf) || kind < 0 || kind > 3) return DN_E_ARG;
  return 0;
}

DN_EXPORT int dn_depth_loss_fwd(const float* gt, const float* pred, int B, int H, int W, int up_factor, int up_mode, float max_depth,
                                int kind, int joint, float weight, int accumulate, float* ws, float* loss, void* stream) {
  int e = dl_check(gt, pred, B, H, W, up_factor, kind);
  if (e || !ws || !loss) return e ? e : DN_E_ARG;
  cudaStream_t st = dn_stream(stream);
  const long long HW = (long long)H * W, pHW = HW / ((long long)up_factor * up_factor);
  UpSrc up{up_factor, up_mode, H / up_factor, W / up_factor};
  const int npass = kind == DL_BERHU ? 2 : 1;
  for (int pass = 0; pass < npass; ++pass) {
    dl_partial_kernel<<<dim3(DL_NBLK, B), 256, 0, st>>>(gt, pred, HW, pHW, W, up, max_depth, pass, ws, B, joint);
    dl_fold_kernel<<<1, 1024, 0, st>>>(ws, B, DL_NBLK, pass, joint, pass == npass - 1, kind, weight, accumulate, loss);
  }
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_depth_loss_bwd(const float* gt, const float* pred, int B, int H, int W, int up_factor, int up_mode, float max_depth,
                                int kind, int joint, float weight, const float* ws, const float* gout, float* gpred, void* stream) {
  int e = dl_check(gt, pred, B, H, W, up_factor, kind);
  if (e || !ws || !gout || !gpred) return e ? e : DN_E_ARG;
  cudaStream_t st = dn_stream(stream);
  const long long HW = (long long)H * W, pHW = HW / ((long long)up_factor * up_factor);
  UpSrc up{up_factor, up_mode, H / up_factor, W / up_factor};
  if (up_factor > 1) cudaMemsetAsync(gpred, 0, sizeof(float) * (size_t)B * pHW, st);
  dl_bwd_kernel<<<dim3(DL_NBLK * 2, B), 256, 0, st>>>(gt, pred, HW, pHW, W, up, max_depth, kind, weight, ws, B, joint, gout, gpred);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_pool2(const float* src, int64_t NC, int H, int W, int mode, float* dst, void* stream) {
  if (!src || !dst || NC < 1 || H < 2 || W < 2 || mode < 0 || mode > 1) return DN_E_ARG;
  pool2_kernel<<<blocks_for(NC * (H / 2) * (W / 2), 256), 256, 0, dn_stream(stream)>>>(src, NC, H, W, mode, dst);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_l1_fwd(const float* gt, const float* pred, int B, int HW, float max_depth, float* ws, float* loss, void* stream) {
  if (!gt || !pred || !ws || !loss || B < 1 || HW < 1) return DN_E_ARG;
  cudaStream_t st = dn_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * B, st);
  if (e != cudaSuccess) return (int)e;
  int bx = blocks_for(HW, 256 * 4);
  int cap = dn_num_sms() * 4 / B; if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  l1_fwd_kernel<<<dim3(bx, B), 256, 0, st>>>(gt, pred, HW, max_depth, ws);
  DN_CHECK_LAUNCH();
  l1_finalize_kernel<<<1, 1, 0, st>>>(ws, B, loss);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_l1_bwd(const float* gt, const float* pred, int B, int HW, float max_depth, const float* ws, const float* gout,
                        float* gpred, void* stream) {
  if (!gt || !pred || !ws || !gout || !gpred) return DN_E_ARG;
  int bx = blocks_for(HW, 256 * 4);
  int cap = dn_num_sms() * 4 / B; if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  l1_bwd_kernel<<<dim3(bx, B), 256, 0, dn_stream(stream)>>>(gt, pred, HW, max_depth, ws, B, gout, gpred);
  DN_CHECK_LAUNCH();
  return 0;
}

static int grid_bx(int hw, int B) {
  int bx = blocks_for(hw, 256);
  int cap = dn_num_sms() * 8 / B; if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  return bx;
}

DN_EXPORT int dn_smooth_fwd(const float* p, int B, int H, int W, float weight, float* loss, void* stream) {
  if (!p || !loss || H < 3 || W < 3) return DN_E_ARG;
  smooth_fwd_kernel<<<dim3(grid_bx(H * W, B), B), 256, 0, dn_stream(stream)>>>(p, H, W, smooth_coef(B, H, W, weight), loss);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_smooth_bwd(const float* p, int B, int H, int W, float weight, const float* gout, float* gp, void* stream) {
  if (!p || !gout || !gp || H < 3 || W < 3) return DN_E_ARG;
  smooth_bwd_kernel<<<dim3(grid_bx(H * W, B), B), 256, 0, dn_stream(stream)>>>(p, H, W, smooth_coef(B, H, W, weight), gout, gp);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_depth_errors(const float* gt, const float* pred, int B, int H, int W, float max_depth, int crop, int y1, int y2,
                              int x1, int x2, const float* scale, int32_t* counters, double* sums, void* stream) {
  if (!gt || !pred || !counters || !sums) return DN_E_ARG;
  if (!crop) { y1 = 0; y2 = H; x1 = 0; x2 = W; }
  depth_errors_kernel<<<dim3(grid_bx(H * W, B), B), 256, 0, dn_stream(stream)>>>(gt, pred, H, W, max_depth, y1, y2, x1, x2, scale, counters, sums);
  DN_CHECK_LAUNCH();
  return 0;
}

__global__ void __launch_bounds__(256) resize_bilinear_ac_kernel(const float* __restrict__ src, int N, int h, int w, int H, int W,
                                                                 float* __restrict__ dst) {
  const float ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long q = i / W;
    const int y = (int)(q % H);
    const float* p = src + (q / H) * (long long)h * w;
    const float sy = ry * (float)y, sx = rx * (float)x;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    dst[i] = (1.f - ly) * ((1.f - lx) * p[y0 * w + x0] + lx * p[y0 * w + x1]) + ly * ((1.f - lx) * p[y1 * w + x0] + lx * p[y1 * w + x1]);
  }
}
DN_EXPORT int dn_resize_bilinear_ac(const float* src, int N, int h, int w, int H, int W, float* dst, void* stream) {
  if (!src || !dst || N < 1 || h < 1 || w < 1 || H < 1 || W < 1) return DN_E_ARG;
  resize_bilinear_ac_kernel<<<blocks_for((long long)N * H * W, 256), 256, 0, dn_stream(stream)>>>(src, N, h, w, H, W, dst);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_area_down(const float* src, int NC, int H, int W, int f, float* dst, void* stream) {
  if (!src || !dst || f < 1 || H % f || W % f) return DN_E_ARG;
  long long total = (long long)NC * (H / f) * (W / f);
  area_down_kernel<<<blocks_for(total, 256, 16), 256, 0, dn_stream(stream)>>>(src, NC, H, W, f, dst);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_explain_fwd(const float* mask, int64_t n, float* loss, void* stream) {
  if (!mask || !loss || n < 1) return DN_E_ARG;
  explain_fwd_kernel<<<blocks_for(n, 1024), 256, 0, dn_stream(stream)>>>(mask, n, 1.f / (float)n, loss);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_explain_bwd(const float* mask, int64_t n, const float* gout, float* gmask, void* stream) {
  if (!mask || !gout || !gmask || n < 1) return DN_E_ARG;
  explain_bwd_kernel<<<blocks_for(n, 1024), 256, 0, dn_stream(stream)>>>(mask, n, 1.f / (float)n, gout, gmask);
  DN_CHECK_LAUNCH();
  return 0;
}


DN_EXPORT int dn_area_pyramid(const dn_pyr_job* jobs, int njobs, int64_t NC, int H, int W, void* stream) {
  if (!jobs || njobs < 1 || njobs > 8 || NC < 1 || (H & 7) || (W & 7)) return DN_E_ARG;
  PyrJobs pj;
  memset(&pj, 0, sizeof(pj));
  for (int i = 0; i < njobs; ++i) {
    pj.j[i] = jobs[i];
    if (!jobs[i].src || ((uintptr_t)jobs[i].src & 15)) return DN_E_ARG;
  }
  const long long total = NC * (H >> 3) * (W >> 3);
  area_pyramid_kernel<<<dim3(blocks_for(total, 256), njobs), 256, 0, dn_stream(stream)>>>(pj, NC, H, W);
  DN_CHECK_LAUNCH();
  return 0;
}

static int photo_check(const dn_photo_batch* p) {
  if (!p || p->nscales < 1 || p->nscales > DN_PHOTO_MAX_SCALES || p->nrefs < 1 || p->nrefs > DN_PHOTO_MAX_REFS || p->B < 1) return DN_E_ARG;
  if (!p->K || !p->Kinv || !p->pose) return DN_E_ARG;
  for (int s = 0; s < p->nscales; ++s) {
    const dn_photo_scale& sc = p->sc[s];
    if (!sc.tgt || !sc.depth || sc.h < 1 || sc.w < 1) return DN_E_ARG;
    for (int r = 0; r < p->nrefs; ++r)
      if (!sc.ref[r]) return DN_E_ARG;
  }
  return 0;
}
DN_EXPORT int64_t dn_photo_ws_floats(const dn_photo_batch* p) {
  if (photo_check(p)) return 0;
  return (int64_t)p->nscales * p->nrefs * p->B * PB_NBX * 12;      // the backward's partial rows (the forward needs 1/12 of it)
}
DN_EXPORT int dn_photo_batch_fwd(const dn_photo_batch* p, float* ws, float* loss, int32_t* nanflag, void* stream) {
  int e = photo_check(p);
  if (e || !ws || !loss) return e ? e : DN_E_ARG;
  cudaStream_t st = dn_stream(stream);
  photo_batch_fwd_kernel<<<dim3(PB_NBX, p->B, p->nscales), 256, 0, st>>>(*p, ws, nanflag);
  photo_fold_kernel<<<1, 32 * DN_PHOTO_MAX_SCALES, 0, st>>>(*p, ws, PB_NBX, loss);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_photo_batch_bwd(const dn_photo_batch* p, const float* gout, float* ws, float* gpose, void* stream) {
  int e = photo_check(p);
  if (e || !ws || !gout || !gpose) return e ? e : DN_E_ARG;
  for (int s = 0; s < p->nscales; ++s)
    if (!p->sc[s].gdepth || (p->sc[s].mask && !p->sc[s].gmask)) return DN_E_ARG;
  cudaStream_t st = dn_stream(stream);
  photo_batch_bwd_kernel<<<dim3(PB_NBX, p->B, p->nscales), 256, 0, st>>>(*p, gout, ws);
  const int n = p->B * p->nrefs;
  photo_pose_finalize_kernel<<<(n + 3) / 4, 128, 0, st>>>(*p, ws, PB_NBX, gpose);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_warp_photo_fwd(const float* tgt, const float* ref, const float* depth, const float* pose, int pose_stride,
                                const float* K, const float* Kinv, const float* mask, int64_t mask_bstride, int B, int h, int w,
                                int rot_mode, int pad_mode, int align_corners, float* warped, float* loss, int32_t* nanflag,
                                void* stream) {
  if (!tgt || !ref || !depth || !pose || !K || !Kinv || !loss || B < 1 || h < 2 || w < 2) return DN_E_ARG;
  float inv = 1.f / ((float)B * 3.f * (float)h * (float)w);
  warp_photo_fwd_kernel<<<dim3(grid_bx(h * w, B), B), 256, 0, dn_stream(stream)>>>(tgt, ref, depth, pose, pose_stride, K, Kinv, mask,
                                                                                   mask_bstride, h, w, rot_mode, pad_mode,
                                                                                   align_corners, warped, inv, loss, nanflag);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_warp_photo_bwd(const float* tgt, const float* ref, const float* depth, const float* pose, int pose_stride,
                                const float* K, const float* Kinv, const float* mask, int64_t mask_bstride, int B, int h, int w,
                                int rot_mode, int pad_mode, int align_corners, const float* gout, float* gdepth, float* gpose,
                                float* gmask, int64_t gmask_bstride, float* ws, void* stream) {
  if (!tgt || !ref || !depth || !pose || !K || !Kinv || !gout || !gdepth || !gpose || !ws) return DN_E_ARG;
  cudaStream_t st = dn_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 12 * B, st);
  if (e != cudaSuccess) return (int)e;
  float inv = 1.f / ((float)B * 3.f * (float)h * (float)w);
  warp_photo_bwd_kernel<<<dim3(grid_bx(h * w, B), B), 256, 0, st>>>(tgt, ref, depth, pose, pose_stride, K, Kinv, mask, mask_bstride, h, w,
                                                                    rot_mode, pad_mode, align_corners, inv, gout, gdepth, ws, gmask,
                                                                    gmask_bstride);
  DN_CHECK_LAUNCH();
  pose_grad_finalize_kernel<<<(B + 63) / 64, 64, 0, st>>>(ws, pose, pose_stride, B, rot_mode, gpose);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_inverse_warp_fwd(const float* img, const float* depth, const float* pose, const float* K, const float* Kinv, int B,
                                  int C, int h, int w, int rot_mode, int pad_mode, int align_corners, float* out, void* stream) {
  if (!img || !depth || !pose || !K || !Kinv || !out || B < 1 || h < 2 || w < 2) return DN_E_ARG;
  inverse_warp_fwd_kernel<1><<<dim3(grid_bx(h * w, B), B), 256, 0, dn_stream(stream)>>>(img, depth, pose, K, Kinv, C, h, w, rot_mode,
                                                                                        pad_mode, align_corners, out);
  DN_CHECK_LAUNCH();
  return 0;
}

DN_EXPORT int dn_inverse_warp_bwd(const float* img, const float* depth, const float* pose, const float* K, const float* Kinv, int B,
                                  int C, int h, int w, int rot_mode, int pad_mode, int align_corners, const float* gout, float* gimg,
                                  float* gdepth, float* gpose, float* ws, void* stream) {
  if (!img || !depth || !pose || !K || !Kinv || !gout || !ws) return DN_E_ARG;
  cudaStream_t st = dn_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 12 * B, st);
  if (e != cudaSuccess) return (int)e;
  inverse_warp_bwd_kernel<<<dim3(grid_bx(h * w, B), B), 256, 0, st>>>(img, depth, pose, K, Kinv, C, h, w, rot_mode, pad_mode,
                                                                      align_corners, gout, gimg, gdepth, ws);
  DN_CHECK_LAUNCH();
  if (gpose) {
    pose_grad_finalize_kernel<<<(B + 63) / 64, 64, 0, st>>>(ws, pose, 6, B, rot_mode, gpose);
    DN_CHECK_LAUNCH();
  }
  return 0;
}

// =================================================================================================
// monodepth2-style optional terms named by the north star (reference layers.py:199-266, uncalled there)
// =================================================================================================
namespace {

__device__ __forceinline__ int refl_idx(int k, int n) { return k < 0 ? -k : (k >= n ? 2 * (n - 1) - k : k); }

struct SsimStats { float mx, my, ex2, ey2, exy; };

__device__ __forceinline__ SsimStats ssim_stats(const float* __restrict__ x, const float* __restrict__ y, int h, int w, int i, int j) {
  SsimStats s = {0, 0, 0, 0, 0};
#pragma unroll
  for (int a = -1; a <= 1; ++a) {
    const int ii = refl_idx(i + a, h);
#pragma unroll
    for (int b = -1; b <= 1; ++b) {
      const int jj = refl_idx(j + b, w);
      const float xv = x[ii * w + jj], yv = y[ii * w + jj];
      s.mx += xv; s.my += yv; s.ex2 += xv * xv; s.ey2 += yv * yv; s.exy += xv * yv;
    }
  }
  const float k = 1.f / 9.f;
  s.mx *= k; s.my *= k; s.ex2 *= k; s.ey2 *= k; s.exy *= k;
  return s;
}

constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

__device__ __forceinline__ float ssim_value(const SsimStats& s, float& n, float& d) {
  const float sx = s.ex2 - s.mx * s.mx, sy = s.ey2 - s.my * s.my, sxy = s.exy - s.mx * s.my;
  n = (2.f * s.mx * s.my + kC1) * (2.f * sxy + kC2);
  d = (s.mx * s.mx + s.my * s.my + kC1) * (sx + sy + kC2);
  return (1.f - n / d) * 0.5f;
}

__global__ void ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int h, int w, float* __restrict__ out) {
  const long long plane = (long long)blockIdx.y * h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h * w; i += gridDim.x * blockDim.x) {
    float n, d;
    const float v = ssim_value(ssim_stats(x + plane, y + plane, h, w, i / w, i % w), n, d);
    out[plane + i] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

// gradient of sum(gout * ssim) w.r.t. x and y: gather over the (up to 5x5) outputs whose reflected window holds the pixel
__global__ void ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gout, int h, int w,
                                float* __restrict__ gx, float* __restrict__ gy) {
  const long long plane = (long long)blockIdx.y * h * w;
  const float* xp = x + plane;
  const float* yp = y + plane;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < h * w; p += gridDim.x * blockDim.x) {
    const int i = p / w, j = p % w;
    const float xv = xp[p], yv = yp[p];
    float ax = 0.f, ay = 0.f;
    for (int oi = i - 2; oi <= i + 2; ++oi) {
      if (oi < 0 || oi >= h) continue;
      int mi = 0;
      for (int a = -1; a <= 1; ++a) mi += refl_idx(oi + a, h) == i;
      if (!mi) continue;
      for (int oj = j - 2; oj <= j + 2; ++oj) {
        if (oj < 0 || oj >= w) continue;
        int mj = 0;
        for (int b = -1; b <= 1; ++b) mj += refl_idx(oj + b, w) == j;
        if (!mj) continue;
        const SsimStats s = ssim_stats(xp, yp, h, w, oi, oj);
        float n, d;
        const float v = ssim_value(s, n, d);
        if (v < 0.f || v > 1.f) continue;                       // clamp: no gradient outside [0, 1]
        const float go = gout[plane + oi * w + oj] * (-0.5f) * (float)(mi * mj) * (1.f / 9.f);
        // n = N1*N2, d = D1*D2 with N1 = 2 mx my + C1, N2 = 2 sxy + C2, D1 = mx^2 + my^2 + C1, D2 = sx + sy + C2
        const float sx = s.ex2 - s.mx * s.mx, sy = s.ey2 - s.my * s.my, sxy = s.exy - s.mx * s.my;
        const float N1 = 2.f * s.mx * s.my + kC1, N2 = 2.f * sxy + kC2, D1 = s.mx * s.mx + s.my * s.my + kC1, D2 = sx + sy + kC2;
        const float inv_d = 1.f / d, q = n * inv_d * inv_d;
        // d(n/d)/d(window element x_p) = [dn - (n/d) dd] / d, chain through mx, E[x^2], E[xy] (each has weight 1/9 per occurrence)
        // dmx = 1, dsx = 2 x_p - 2 mx, dsxy = y_p - my
        const float dn_x = 2.f * s.my * N2 + N1 * 2.f * (yv - s.my);
        const float dd_x = 2.f * s.mx * D2 + D1 * (2.f * xv - 2.f * s.mx);
        const float dn_y = 2.f * s.mx * N2 + N1 * 2.f * (xv - s.mx);
        const float dd_y = 2.f * s.my * D2 + D1 * (2.f * yv - 2.f * s.my);
        ax += go * (dn_x * inv_d - q * dd_x);
        ay += go * (dn_y * inv_d - q * dd_y);
      }
    }
    if (gx) gx[plane + p] = ax;
    if (gy) gy[plane + p] = ay;
  }
}

__global__ void __launch_bounds__(256) edge_smooth_fwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, int C, int h,
                                                              int w, float cx, float cy, float* loss) {
  const int b = blockIdx.y;
  const float* d = disp + (long long)b * h * w;
  const float* im = img + (long long)b * C * h * w;
  float s = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < h * w; p += gridDim.x * blockDim.x) {
    const int i = p / w, j = p % w;
    if (j < w - 1) {
      float gi = 0.f;
      for (int c = 0; c < C; ++c) gi += fabsf(im[(long long)c * h * w + p] - im[(long long)c * h * w + p + 1]);
      s += cx * fabsf(d[p] - d[p + 1]) * expf(-gi / (float)C);
    }
    if (i < h - 1) {
      float gi = 0.f;
      for (int c = 0; c < C; ++c) gi += fabsf(im[(long long)c * h * w + p] - im[(long long)c * h * w + p + w]);
      s += cy * fabsf(d[p] - d[p + w]) * expf(-gi / (float)C);
    }
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(loss, s);
}

__global__ void __launch_bounds__(256) edge_smooth_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, int C, int h,
                                                              int w, float cx, float cy, const float* __restrict__ gout,
                                                              float* __restrict__ gdisp) {
  const int b = blockIdx.y;
  const float* d = disp + (long long)b * h * w;
  const float* im = img + (long long)b * C * h * w;
  const float go = gout[0];
  auto wx = [&](int p) {   // weight of the pair (p, p+1)
    float gi = 0.f;
    for (int c = 0; c < C; ++c) gi += fabsf(im[(long long)c * h * w + p] - im[(long long)c * h * w + p + 1]);
    return cx * expf(-gi / (float)C) * sgn(d[p] - d[p + 1]);
  };
  auto wy = [&](int p) {   // weight of the pair (p, p+w)
    float gi = 0.f;
    for (int c = 0; c < C; ++c) gi += fabsf(im[(long long)c * h * w + p] - im[(long long)c * h * w + p + w]);
    return cy * expf(-gi / (float)C) * sgn(d[p] - d[p + w]);
  };
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < h * w; p += gridDim.x * blockDim.x) {
    const int i = p / w, j = p % w;
    float g = 0.f;
    if (j < w - 1) g += wx(p);
    if (j > 0) g -= wx(p - 1);
    if (i < h - 1) g += wy(p);
    if (i > 0) g -= wy(p - w);
    gdisp[(long long)b * h * w + p] = g * go;
  }
}

__global__ void __launch_bounds__(256) depth_errors_raw_kernel(const float* __restrict__ gt, const float* __restrict__ pred, long long n,
                                                               int32_t* counters, double* sums) {
  int a1 = 0, a2 = 0, a3 = 0;
  double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  const float t1 = 1.25f, t2 = (float)(1.25 * 1.25), t3 = (float)(1.25 * 1.25 * 1.25);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = gt[i], p = pred[i];
    const float th = fmaxf(g / p, p / g);
    a1 += th < t1; a2 += th < t2; a3 += th < t3;
    const float d = g - p, lg = logf(g) - logf(p);
    s1 += (double)(fabsf(d) / g); s2 += (double)(d * d / g); s3 += (double)(d * d); s4 += (double)(lg * lg);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(counters, a1); atomicAdd(counters + 1, a2); atomicAdd(counters + 2, a3); }
  s1 = block_sum_d(s1); s2 = block_sum_d(s2); s3 = block_sum_d(s3); s4 = block_sum_d(s4);
  if (threadIdx.x == 0) { atomicAdd(sums, s1); atomicAdd(sums + 1, s2); atomicAdd(sums + 2, s3); atomicAdd(sums + 3, s4); }
}

}  // namespace

DN_EXPORT int dn_ssim_fwd(const float* x, const float* y, int NC, int h, int w, float* out, void* stream) {
  if (!x || !y || !out || h < 2 || w < 2) return DN_E_ARG;
  ssim_fwd_kernel<<<dim3(grid_bx(h * w, NC), NC), 256, 0, dn_stream(stream)>>>(x, y, h, w, out);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_ssim_bwd(const float* x, const float* y, const float* gout, int NC, int h, int w, float* gx, float* gy, void* stream) {
  if (!x || !y || !gout || h < 2 || w < 2) return DN_E_ARG;
  ssim_bwd_kernel<<<dim3(grid_bx(h * w, NC), NC), 256, 0, dn_stream(stream)>>>(x, y, gout, h, w, gx, gy);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_edge_smooth_fwd(const float* disp, const float* img, int B, int C, int h, int w, float* loss, void* stream) {
  if (!disp || !img || !loss || h < 2 || w < 2) return DN_E_ARG;
  const float cx = 1.f / ((float)B * h * (w - 1)), cy = 1.f / ((float)B * (h - 1) * w);
  edge_smooth_fwd_kernel<<<dim3(grid_bx(h * w, B), B), 256, 0, dn_stream(stream)>>>(disp, img, C, h, w, cx, cy, loss);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_edge_smooth_bwd(const float* disp, const float* img, int B, int C, int h, int w, const float* gout, float* gdisp,
                                 void* stream) {
  if (!disp || !img || !gout || !gdisp) return DN_E_ARG;
  const float cx = 1.f / ((float)B * h * (w - 1)), cy = 1.f / ((float)B * (h - 1) * w);
  edge_smooth_bwd_kernel<<<dim3(grid_bx(h * w, B), B), 256, 0, dn_stream(stream)>>>(disp, img, C, h, w, cx, cy, gout, gdisp);
  DN_CHECK_LAUNCH();
  return 0;
}
DN_EXPORT int dn_depth_errors_raw(const float* gt, const float* pred, int64_t n, int32_t* counters, double* sums, void* stream) {
  if (!gt || !pred || !counters || !sums || n < 1) return DN_E_ARG;
  depth_errors_raw_kernel<<<blocks_for(n, 1024), 256, 0, dn_stream(stream)>>>(gt, pred, n, counters, sums);
  DN_CHECK_LAUNCH();
  return 0;
}
