// tcgen05 / TMA / TMEM implicit-GEMM convolution kernels for sm_100a (backend 1 of dn_igemm_run / dn_wgrad_run).
//
//   igemm_tc_kernel : out[pixel][co] = sum_tap sum_ci in[pixel + tap][ci] * w[tap][co][ci]
//       A (activations) is fetched by 4-D TMA boxes {64 ch, wb, hb, nb} straight from the NHWC view -- the tap shift is a
//       coordinate offset and zero padding is TMA out-of-bounds fill, so there is no im2col buffer; B (packed weights) by 3-D
//       TMA; both land in 128B-swizzled shared memory and feed tcgen05.mma (M=128, N<=256, K=16) with fp32 accumulators in
//       TMEM.  Persistent CTAs (one per SM), 3 warp roles (TMA producer / MMA issuer / 4 epilogue warps), a multi-stage
//       smem ring and two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
//   wgrad_tc_kernel : dw[tap][cp][cq] += sum_pixel dy[pixel][cp] * x[pixel + tap][cq]
//       both operands are "MN-major" (the reduction index = pixel is the row index of the TMA box), split-K over pixel
//       tiles across CTAs, fp32 partial sums reduced with vector red.global.add.
//
// Forward nn.Conv2d (stride 1), the four output phases of nn.ConvTranspose2d, and the data gradients of both run on
// igemm_tc_kernel with different tap tables / views; weight gradients of both on wgrad_tc_kernel.
#include "dn_common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>
#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one lane of a fully converged warp; keeps the surrounding control flow warp-uniform so that descriptors and barrier
// addresses stay in uniform registers (issuing tcgen05 / TMA from inside `if (lane == 0)` makes the compiler wrap every
// instruction in an elect/broadcast loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "elect.sync %%rx|%%px, %1;\n"
      "@%%px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// shared-memory matrix descriptor, 128B swizzle (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1
// <<46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// second copy of an epilogue chunk (16 fp32 values) in another 16-bit dtype
__device__ __forceinline__ void store_out2(void* base, int dtype, size_t elem_off, const float* v, int c_lo, int c_eff) {
  if (dtype == DN_F16) {
    __half* o = (__half*)base + elem_off;
    if (c_lo < c_eff) Vec8<__half>::store(o, v);
    if (c_lo + 8 < c_eff) Vec8<__half>::store(o + 8, v + 8);
  } else {
    __nv_bfloat16* o = (__nv_bfloat16*)base + elem_off;
    if (c_lo < c_eff) Vec8<__nv_bfloat16>::store(o, v);
    if (c_lo + 8 < c_eff) Vec8<__nv_bfloat16>::store(o + 8, v + 8);
  }
}

constexpr int kMaxTcTaps = DN_MAX_TAPS;
constexpr int kRows = 128;          // pixels per tile = UMMA M
constexpr int kChunk = 64;          // channels per K chunk = one 128-byte swizzled row

struct TcTap { int16_t src, dh, dw, wt; };

struct IgemmTcParams {
  CUtensorMap tmA[DN_MAX_SRC];
  CUtensorMap tmB;
  TcTap taps[kMaxTcTaps];
  int ntaps, nsrc;
  int kchunks;        // ceil(Cin / 64)
  int last_ksteps;    // UMMA_K steps in the last chunk (1..4)
  int wb, hb, nb;     // pixel box (wb*hb*nb == 128)
  int tilesW, tilesH, tilesN, ntile_n, num_tiles;
  int stages;
  int n_mma;          // MMA N actually issued (multiple of 16, <= BN)
  int c_eff;          // output channels the epilogue may write (out.C, or rounded up to 8 when padding may be overwritten)
  uint32_t idesc;
  dn_view out;
  void* out2;
  int out2_dtype;
  const float* bias;
  int act, accumulate;
  float out_scale;
  unsigned long long* dbg;   // optional per-role cycle counters (dn_tc_set_debug)
};

template <int BN>
__global__ void __launch_bounds__(192, 1) igemm_tc_kernel(const __grid_constant__ IgemmTcParams p) {
  dn_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = kRows * 128;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // carve: [stages x (A|B)] | barriers | tmem ptr | bias tile
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tempty_bar + 2);
  float* bias_s = (float*)(tmem_ptr + 2);   // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  dn_pdl_wait();      // set-up above overlaps the previous kernel's tail; global memory only from here on

  const int k_iters = p.ntaps * p.kchunks;

  if (warp == 0) {
    // ================= TMA producer (whole warp runs the loop, one elected lane issues) =================
    int stage = 0; uint32_t phase = 0;
    long long dbg_acc0 = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nt = tile % p.ntile_n;
      int m = tile / p.ntile_n;
      const int tw = m % p.tilesW; m /= p.tilesW;
      const int th = m % p.tilesH;
      const int tn = m / p.tilesH;
      const int w0 = tw * p.wb, h0 = th * p.hb, n0 = tn * p.nb;
      for (int t = 0; t < p.ntaps; ++t) {
        const TcTap tap = p.taps[t];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.dbg) dbg_acc0 += clock64() - t0;
          if (elect_one()) {
            uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_4d(sa, &p.tmA[tap.src], &full_bar[stage], kc * kChunk, w0 + tap.dw, h0 + tap.dh, n0);
            tma_load_3d(sb, &p.tmB, &full_bar[stage], kc * kChunk, nt * BN, tap.wt);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (p.dbg && lane == 0) {
      atomicAdd(p.dbg + 0, (unsigned long long)dbg_acc0);
      atomicAdd(p.dbg + 1, (unsigned long long)(clock64() - tstart));
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp waits on the barriers, one elected lane issues) =================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    long long dbg_full = 0, dbg_te = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (p.dbg) dbg_te += clock64() - t0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      uint32_t first = 1;
      for (int t = 0; t < p.ntaps; ++t) {
        for (int kc = 0; kc < p.kchunks; ++kc) {
          t0 = p.dbg ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (p.dbg) dbg_full += clock64() - t0;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint64_t ad0 = make_desc(sa, 16, 1024);
          const uint64_t bd0 = make_desc(sa + A_BYTES, 16, 1024);
          const int ks = (kc == p.kchunks - 1) ? p.last_ksteps : 4;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < ks) {
                // advance 16 K-elements = 32 bytes inside the 128-byte swizzled row: +2 in the (addr >> 4) field
                umma_f16(d_tmem, ad0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), p.idesc, (first && k == 0) ? 0u : 1u);
              }
            }
            umma_commit(&empty_bar[stage]);
            if (t == p.ntaps - 1 && kc == p.kchunks - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          first = 0;
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      atomicAdd(p.dbg + 2, (unsigned long long)dbg_full); atomicAdd(p.dbg + 3, (unsigned long long)dbg_te);
      atomicAdd(p.dbg + 4, (unsigned long long)(clock64() - tstart));
    }
  } else {
    // ================= epilogue warps (2..5): TMEM -> registers -> bias/act -> HBM =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // tile row = TMEM lane = pixel within the box
    const int et = threadIdx.x - 64;        // 0..127
    int acc = 0; uint32_t acc_phase = 0;
    const int wi = row % p.wb;
    const int hi = (row / p.wb) % p.hb;
    const int ni = row / (p.wb * p.hb);
    const int esz = p.out.dtype == DN_F32 ? 4 : 2;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nt = tile % p.ntile_n;
      int m = tile / p.ntile_n;
      const int tw = m % p.tilesW; m /= p.tilesW;
      const int th = m % p.tilesH;
      const int tn = m / p.tilesH;
      const int w = tw * p.wb + wi, h = th * p.hb + hi, n = tn * p.nb + ni;
      const bool valid = (w < p.out.W) && (h < p.out.H) && (n < p.out.N);
      const int co0 = nt * BN;
      float* bs = bias_s + acc * BN;
      for (int c = et; c < BN; c += 128) bs[c] = (p.bias && co0 + c < p.out.C) ? p.bias[co0 + c] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long te0 = p.dbg ? clock64() : 0;
      mbar_wait(&tfull_bar[acc], acc_phase);
      const long long te1 = p.dbg ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      uint8_t* optr = (uint8_t*)p.out.ptr + (size_t)(dn_off(p.out, n, h, w) + co0) * esz;
#pragma unroll 1
      for (int c0 = 0; c0 < p.n_mma; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (valid && co0 + c0 < p.c_eff) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = dn_act(__uint_as_float(r[j]) * p.out_scale + bs[c0 + j], p.act);
          if (p.out2) store_out2(p.out2, p.out2_dtype, (size_t)(dn_off(p.out, n, h, w) + co0 + c0), v, co0 + c0, p.c_eff);
          if (p.out.dtype == DN_F32) {
            float* o = (float*)optr + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (co0 + c0 + j < p.out.C) o[j] = p.accumulate ? o[j] + v[j] : v[j];
          } else if (p.out.dtype == DN_F16) {
            __half* o = (__half*)optr + c0;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (co0 + c0 + 8 * hh < p.c_eff) {
                if (p.accumulate) {
                  float a[8];
                  Vec8<__half>::load(o + 8 * hh, a);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[8 * hh + j] += a[j];
                }
                Vec8<__half>::store(o + 8 * hh, v + 8 * hh);
              }
            }
          } else {
            __nv_bfloat16* o = (__nv_bfloat16*)optr + c0;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (co0 + c0 + 8 * hh < p.c_eff) {
                if (p.accumulate) {
                  float a[8];
                  Vec8<__nv_bfloat16>::load(o + 8 * hh, a);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[8 * hh + j] += a[j];
                }
                Vec8<__nv_bfloat16>::store(o + 8 * hh, v + 8 * hh);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (p.dbg && et == 0) {
        atomicAdd(p.dbg + 5, (unsigned long long)(te1 - te0));
        atomicAdd(p.dbg + 6, (unsigned long long)(clock64() - te1));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// 3x3 stride-1 gather-convolution with shared-memory halo reuse ("halo" variant of igemm_tc_kernel)
//
// For layers with many pixels and few channels the plain kernel is bound by the TMA row rate / L2 bandwidth: every output
// tile fetches its 128x64 activation box nine times (once per tap).  Here one (16+2) x 16-pixel halo box is fetched per
// 64-channel chunk (the tile is 16 rows x 8 columns of pixels, the box is padded to 16 columns so that an image row is
// 2048 bytes = two swizzle atoms) and the nine taps are nine *views* of it: the A descriptor starts (kh*16 + kw) pixels
// further and keeps SBO = 2048 B between the 8-pixel row groups (the swizzle is address-based, so the shifted start needs
// no base_offset).  The packed weights of all taps stay resident in shared memory for the
// lifetime of the persistent CTA.
// ------------------------------------------------------------------------------------------------
constexpr int kHaloW = 16, kHaloH = 18;
constexpr uint32_t kHaloBytes = kHaloW * kHaloH * 128;     // 36864


struct HaloParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int kchunks, last_ksteps;
  int wt[9];            // packed-weight matrix index of tap (kh, kw), kh = dh + 1, kw = dw + 1
  int tilesW, tilesH, num_tiles;
  int stages;
  int ring, bstages;    // ring = 1: weight tiles stream through a bstages-deep ring instead of staying resident
  int n_mma, c_eff;
  uint32_t idesc;
  dn_view out;
  void* out2;
  int out2_dtype;
  const float* bias;
  int act, accumulate;
  float out_scale;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) igemm_halo_kernel(const __grid_constant__ HaloParams p) {
  dn_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  const int nb_tiles = p.ring ? p.bstages : 9 * p.kchunks;   // weight tiles in shared memory (ring slots or all of them)
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + (size_t)nb_tiles * B_BYTES;
  uint64_t* full_bar = (uint64_t*)(smem_a + (size_t)stages * kHaloBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bfull_bar = tempty_bar + 2;                      // resident: [0]; ring: full[0..7], empty[8..15]
  uint64_t* bempty_bar = bfull_bar + 8;
  uint32_t* tmem_ptr = (uint32_t*)(bfull_bar + 16);
  float* bias_s = (float*)(tmem_ptr + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
    for (int s = 0; s < 8; ++s) { mbar_init(&bfull_bar[s], 1); mbar_init(&bempty_bar[s], 1); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  dn_pdl_wait();      // set-up above overlaps the previous kernel's tail; global memory only from here on

  if (warp == 0 && p.ring) {
    // ---- producer, streamed weights: the halo box of item i+1 is requested before the nine weight tiles of item i
    int stage = 0; uint32_t phase = 0;
    int bs = 0; uint32_t bphase = 0;
    int a_tile = blockIdx.x, a_kc = 0;
    auto load_a = [&]() {
      if (a_tile >= p.num_tiles) return;
      int m = a_tile;
      const int tw = m % p.tilesW; m /= p.tilesW;
      const int th = m % p.tilesH;
      const int n0 = m / p.tilesH;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], kHaloBytes);
        tma_load_4d(smem_a + (size_t)stage * kHaloBytes, &p.tmA, &full_bar[stage], a_kc * kChunk, tw * 8 - 1, th * 16 - 1, n0);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
      if (++a_kc == p.kchunks) { a_kc = 0; a_tile += gridDim.x; }
    };
    load_a();
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < p.kchunks; ++kc) {
        load_a();
        for (int t = 0; t < 9; ++t) {
          mbar_wait(&bempty_bar[bs], bphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&bfull_bar[bs], B_BYTES);
            tma_load_3d(smem_b + (size_t)bs * B_BYTES, &p.tmB, &bfull_bar[bs], kc * kChunk, 0, p.wt[t]);
          }
          __syncwarp();
          if (++bs == p.bstages) { bs = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && p.ring) {
    // ---- MMA issuer, streamed weights
    int stage = 0; uint32_t phase = 0;
    int bs = 0; uint32_t bphase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t sb0 = smem_u32(smem_b);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem_a + (size_t)stage * kHaloBytes);
        const int ks = (kc == p.kchunks - 1) ? p.last_ksteps : 4;
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          mbar_wait(&bfull_bar[bs], bphase);
          tc_fence_after();
          if (elect_one()) {
            const int kh = t / 3, kw = t - 3 * kh;
            const uint64_t ad0 = make_desc(sa + (uint32_t)(kh * kHaloW + kw) * 128, 16, 2048);
            const uint64_t bd0 = make_desc(sb0 + (uint32_t)bs * B_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < ks) umma_f16(d_tmem, ad0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), p.idesc, (kc == 0 && t == 0 && k == 0) ? 0u : 1u);
            umma_commit(&bempty_bar[bs]);
            if (t == 8) {
              umma_commit(&empty_bar[stage]);
              if (kc == p.kchunks - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++bs == p.bstages) { bs = 0; bphase ^= 1; }
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp == 0) {
    // ---- producer: weights once, then one halo box per (tile, chunk)
    if (elect_one()) {
      int present = 0;
      for (int t = 0; t < 9; ++t) present += p.wt[t] >= 0 ? 1 : 0;
      mbar_expect_tx(bfull_bar, (uint32_t)(present * p.kchunks) * B_BYTES);
      for (int t = 0; t < 9; ++t)
        if (p.wt[t] >= 0)
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_3d(smem_b + (size_t)(t * p.kchunks + kc) * B_BYTES, &p.tmB, bfull_bar, kc * kChunk, 0, p.wt[t]);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m = tile;
      const int tw = m % p.tilesW; m /= p.tilesW;
      const int th = m % p.tilesH;
      const int n0 = m / p.tilesH;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], kHaloBytes);
          tma_load_4d(smem_a + (size_t)stage * kHaloBytes, &p.tmA, &full_bar[stage], kc * kChunk, tw * 8 - 1, th * 16 - 1, n0);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    mbar_wait(bfull_bar, 0);
    tc_fence_after();
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t sb0 = smem_u32(smem_b);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem_a + (size_t)stage * kHaloBytes);
        const int ks = (kc == p.kchunks - 1) ? p.last_ksteps : 4;
        if (elect_one()) {
          uint32_t accumulate = kc == 0 ? 0u : 1u;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (p.wt[t] < 0) continue;      // tap subset (2x2 neighbourhood of a transposed-convolution phase)
            const int kh = t / 3, kw = t % 3;
            // base_offset stays 0: the 128B swizzle is a function of the absolute shared-memory address bits (measured on B200:
            // only base_offset = 0 reproduces the unshifted kernel), so a start that is not 1024-byte aligned needs no correction
            const uint64_t ad0 = make_desc(sa + (uint32_t)(kh * kHaloW + kw) * 128, 16, 2048);
            const uint64_t bd0 = make_desc(sb0 + (uint32_t)(t * p.kchunks + kc) * B_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < ks) { umma_f16(d_tmem, ad0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), p.idesc, accumulate); accumulate = 1u; }
          }
          umma_commit(&empty_bar[stage]);
          if (kc == p.kchunks - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ---- epilogue: tile row r = 8 * h_local + w_local
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    int acc = 0; uint32_t acc_phase = 0;
    const int wi = row & 7, hi = row >> 3;
    const int esz = p.out.dtype == DN_F32 ? 4 : 2;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m = tile;
      const int tw = m % p.tilesW; m /= p.tilesW;
      const int th = m % p.tilesH;
      const int n = m / p.tilesH;
      const int w = tw * 8 + wi, h = th * 16 + hi;
      const bool valid = (w < p.out.W) && (h < p.out.H);
      float* bs = bias_s + acc * BN;
      for (int c = et; c < BN; c += 128) bs[c] = (p.bias && c < p.out.C) ? p.bias[c] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      uint8_t* optr = (uint8_t*)p.out.ptr + (size_t)dn_off(p.out, n, h, w) * esz;
#pragma unroll 1
      for (int c0 = 0; c0 < p.n_mma; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (valid && c0 < p.c_eff) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = dn_act(__uint_as_float(r[j]) * p.out_scale + bs[c0 + j], p.act);
          if (p.out2) store_out2(p.out2, p.out2_dtype, (size_t)(dn_off(p.out, n, h, w) + c0), v, c0, p.c_eff);
          if (p.out.dtype == DN_F32) {
            float* o = (float*)optr + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.out.C) o[j] = p.accumulate ? o[j] + v[j] : v[j];
          } else if (p.out.dtype == DN_F16) {
            __half* o = (__half*)optr + c0;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
              if (c0 + 8 * hh < p.c_eff) {
                if (p.accumulate) {
                  float a[8];
                  Vec8<__half>::load(o + 8 * hh, a);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[8 * hh + j] += a[j];
                }
                Vec8<__half>::store(o + 8 * hh, v + 8 * hh);
              }
          } else {
            __nv_bfloat16* o = (__nv_bfloat16*)optr + c0;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
              if (c0 + 8 * hh < p.c_eff) {
                if (p.accumulate) {
                  float a[8];
                  Vec8<__nv_bfloat16>::load(o + 8 * hh, a);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[8 * hh + j] += a[j];
                }
                Vec8<__nv_bfloat16>::store(o + 8 * hh, v + 8 * hh);
              }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
struct WgradTcParams {
  CUtensorMap tmP[DN_MAX_SRC];
  CUtensorMap tmQ;
  TcTap taps[kMaxTcTaps];
  int ntaps;
  int tpc, ngroups;        // taps per work item (they share the dy tile), number of tap groups
  int wb, hb, nb;          // pixel box, wb*hb*nb == KPX
  int tilesW, tilesH, tilesN, num_ptiles;
  int cp_tiles, cq_tiles;  // output tiles: 128 x BNQ
  int cp_blocks;           // 64-channel blocks of P actually present in a cp tile (1 or 2)
  int splits, ptiles_per_split;
  int num_items;           // ngroups * cp_tiles * cq_tiles * splits
  int stages;
  int n_mma;               // MMA N per tap (multiple of 16, <= BNQ)
  int m64;                 // P has at most 64 channels: issue M = 64 MMAs (half the A-operand reads); 1 / 2 = TMEM row mapping variant
  int nacc;                // TMEM accumulator stages (2: the atomics of item i overlap the main loop of item i+1)
  int tmem_cols;
  int halo;                // 1: x is fetched once per pixel tile as a (8+2) x 16-pixel halo box, taps are shifted views
  uint32_t idesc;
  float* dw;
  int cp, cq, cp_pad, cq_pad;
  float scale;
};

constexpr int KPX = 64;   // pixels (= GEMM K) per pipeline stage

struct WgItem { int grp, cqt, cpt, split, t0, nt, src, pt_beg, n_iters; };

__device__ __forceinline__ WgItem wg_decode(const WgradTcParams& p, int item) {
  WgItem w;
  // tap group fastest so that CTAs running together share the same pixel range in L2
  w.grp = item % p.ngroups; item /= p.ngroups;
  w.cqt = item % p.cq_tiles; item /= p.cq_tiles;
  w.cpt = item % p.cp_tiles; item /= p.cp_tiles;
  w.split = item;
  w.t0 = w.grp * p.tpc;
  w.nt = (p.ntaps - w.t0) < p.tpc ? (p.ntaps - w.t0) : p.tpc;
  w.src = p.taps[w.t0].src;     // all taps of one launch share the dy view
  w.pt_beg = w.split * p.ptiles_per_split;
  int pt_end = w.pt_beg + p.ptiles_per_split;
  if (pt_end > p.num_ptiles) pt_end = p.num_ptiles;
  w.n_iters = pt_end - w.pt_beg;
  return w;
}

template <int BNQ>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
  dn_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t BLK_BYTES = KPX * 128;              // one [KPX px][64 ch] swizzled block
  constexpr uint32_t A_BYTES = 2 * BLK_BYTES;            // M = 128 channels of P
  constexpr uint32_t B_BYTES = (BNQ / 64) * BLK_BYTES;   // per tap (plain mode)
  constexpr uint32_t HALO_BLK = 16 * 10 * 128;           // one [10 rows][16 px][64 ch] halo block (halo mode)
  const uint32_t STAGE_BYTES = A_BYTES + (p.halo ? (uint32_t)(BNQ / 64) * HALO_BLK : (uint32_t)p.tpc * B_BYTES);
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmP[0]);
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  dn_pdl_wait();      // set-up above overlaps the previous kernel's tail; global memory only from here on
  const uint32_t acc_cols = (uint32_t)(p.tpc * p.n_mma);

  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem w = wg_decode(p, item);
      // bytes that really arrive per stage: only the P blocks that exist and the taps of this group are loaded
      const uint32_t tx_bytes = (uint32_t)p.cp_blocks * BLK_BYTES + (p.halo ? (uint32_t)(BNQ / 64) * HALO_BLK : (uint32_t)w.nt * B_BYTES);
      for (int it = 0; it < w.n_iters; ++it) {
        int m = w.pt_beg + it;
        const int tw = m % p.tilesW; m /= p.tilesW;
        const int th = m % p.tilesH;
        const int tn = m / p.tilesH;
        const int w0 = tw * p.wb, h0 = th * p.hb, n0 = tn * p.nb;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          for (int j = 0; j < p.cp_blocks; ++j)
            tma_load_4d(sa + j * BLK_BYTES, &p.tmP[w.src], &full_bar[stage], w.cpt * 128 + j * 64, w0, h0, n0);
          if (p.halo) {
#pragma unroll
            for (int j = 0; j < BNQ / 64; ++j)
              tma_load_4d(sa + A_BYTES + j * HALO_BLK, &p.tmQ, &full_bar[stage], w.cqt * BNQ + j * 64, w0 - 1, h0 - 1, n0);
          } else {
            for (int t = 0; t < w.nt; ++t) {
              const TcTap tap = p.taps[w.t0 + t];
              uint8_t* sb = sa + A_BYTES + (size_t)t * B_BYTES;
#pragma unroll
              for (int j = 0; j < BNQ / 64; ++j)
                tma_load_4d(sb + j * BLK_BYTES, &p.tmQ, &full_bar[stage], w.cqt * BNQ + j * 64, w0 + tap.dw, h0 + tap.dh, n0);
            }
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem w = wg_decode(p, item);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)acc * acc_cols;
      for (int it = 0; it < w.n_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        // MN-major, 128B swizzle: LBO = distance between 64-channel blocks, SBO = 8 pixel rows = 1024 B;
        // 16 pixels (one UMMA_K step) further down = +2048 bytes = +128 in the (addr >> 4) field
        const uint64_t ad0 = make_desc(sa, BLK_BYTES, 1024);
        if (elect_one()) {
          for (int t = 0; t < w.nt; ++t) {
            if (p.halo) {
              // pixel tile = 8 rows x 8 columns; tap (dh, dw) reads halo pixel (row + dh + 1, col + dw + 1): a view that starts
              // ((dh+1)*16 + (dw+1)) pixels into the 16-pixel-pitch halo block.  One UMMA_K step = 16 pixels = 2 image rows.
              const TcTap tap = p.taps[w.t0 + t];
              const uint32_t sb = sa + A_BYTES + (uint32_t)((tap.dh + 1) * 16 + (tap.dw + 1)) * 128;
              const uint64_t bd0 = make_desc(sb, HALO_BLK, 2048);
#pragma unroll
              for (int k = 0; k < KPX / 16; ++k)
                umma_f16(d_base + (uint32_t)(t * p.n_mma), ad0 + (uint64_t)(128 * k), bd0 + (uint64_t)(256 * k), p.idesc,
                         (it > 0 || k > 0) ? 1u : 0u);
            } else {
              const uint64_t bd0 = make_desc(sa + A_BYTES + (uint32_t)t * B_BYTES, BLK_BYTES, 1024);
#pragma unroll
              for (int k = 0; k < KPX / 16; ++k)
                umma_f16(d_base + (uint32_t)(t * p.n_mma), ad0 + (uint64_t)(128 * k), bd0 + (uint64_t)(128 * k), p.idesc,
                         (it > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (it == w.n_iters - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (p.nacc == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
      else acc_phase ^= 1;
    }
  } else {
    // ================= epilogue: TMEM -> scaled fp32 partial sums -> red.global.add =================
    const int q = warp & 3;
    const int row = q * 32 + lane;        // cp within the tile
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem w = wg_decode(p, item);
      // M = 64 accumulators: variant 1 = 16 rows per 32-lane TMEM quarter (row r in lane 32 * (r / 16) + r % 16),
      // variant 2 = rows in lanes 0..63
      int cp = w.cpt * 128 + row;
      bool row_ok = true;
      if (p.m64 == 1) { cp = w.cpt * 128 + q * 16 + lane; row_ok = lane < 16; }
      else if (p.m64 == 2) { row_ok = row < 64; }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      for (int t = 0; t < w.nt; ++t) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * p.n_mma);
        float* drow = p.dw + ((size_t)p.taps[w.t0 + t].wt * p.cp_pad + cp) * p.cq_pad + w.cqt * BNQ;
#pragma unroll 1
        for (int c0 = 0; c0 < p.n_mma; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          if (row_ok && cp < p.cp && w.cqt * BNQ + c0 < p.cq_pad) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add_v4(drow + c0 + j, __uint_as_float(r[j]) * p.scale, __uint_as_float(r[j + 1]) * p.scale,
                         __uint_as_float(r[j + 2]) * p.scale, __uint_as_float(r[j + 3]) * p.scale);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (p.nacc == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
      else acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)f;
  }
  return fn;
}

bool view_tma_ok(const dn_view& v) {
  if (v.dtype != DN_F16 && v.dtype != DN_BF16) return false;
  if (((uintptr_t)v.ptr % 16) != 0) return false;
  if ((v.sW % 8) || (v.sH % 8) || (v.sN % 8)) return false;
  if (v.sW <= 0 || v.sH <= 0 || v.sN <= 0) return false;
  return true;
}

// 4-D map (C, W, H, N) of an NHWC view, box {64, wb, hb, nb}, 128B swizzle, zero OOB fill
int make_view_map(CUtensorMap* tm, const dn_view& v, int wb, int hb, int nb) {
  auto enc = get_encode();
  if (!enc) return DN_E_UNSUPPORTED;
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)kChunk, (cuuint32_t)wb, (cuuint32_t)hb, (cuuint32_t)nb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapDataType dt = v.dtype == DN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(tm, dt, 4, v.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : DN_E_ARG;
}

// choose a pixel box with wb*hb*nb == rows that minimises the number of tiles
void choose_box(int N, int H, int W, int rows, int& wb, int& hb, int& nb) {
  long long best = -1;
  wb = rows; hb = 1; nb = 1;
  for (int a = 1; a <= rows; a <<= 1)
    for (int b = 1; a * b <= rows; b <<= 1) {
      int c = rows / (a * b);
      long long tiles = (long long)((W + a - 1) / a) * ((H + b - 1) / b) * ((N + c - 1) / c);
      if (best < 0 || tiles < best || (tiles == best && a > wb)) { best = tiles; wb = a; hb = b; nb = c; }
    }
}

uint32_t make_idesc(int a_dtype, int b_dtype, int a_mn_major, int b_mn_major, int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                                         // D format f32
  d |= (uint32_t)(a_dtype == DN_BF16 ? 1 : 0) << 7;     // A format
  d |= (uint32_t)(b_dtype == DN_BF16 ? 1 : 0) << 10;    // B format
  d |= (uint32_t)(a_mn_major ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn_major ? 1 : 0) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

int pick_bn(int cout_pad) {
  if (cout_pad > 256) {   // several N tiles: the largest tile width that divides the (16-aligned) channel count
    for (int bn = 256; bn >= 16; bn >>= 1)
      if (cout_pad % bn == 0) return bn;
  }
  if (cout_pad > 128) return 256;
  if (cout_pad > 64) return 128;
  if (cout_pad > 32) return 64;
  if (cout_pad > 16) return 32;
  return 16;
}

template <int BN>
int launch_igemm(const IgemmTcParams& P, cudaStream_t st) {
  constexpr uint32_t STAGE_BYTES = kRows * 128 + BN * 128;
  int stages = P.stages;
  size_t smem = (size_t)stages * STAGE_BYTES + 1024 + (2 * stages + 4) * 8 + 16 + 2 * BN * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int grid = P.num_tiles < dn_num_sms() ? P.num_tiles : dn_num_sms();
  dn_launch(igemm_tc_kernel<BN>, dim3(grid), dim3(192), smem, st, P);
  DN_CHECK_LAUNCH();
  return 0;
}

template <int BNQ>
int launch_wgrad(const WgradTcParams& P, int items, cudaStream_t st) {
  const uint32_t stage_bytes = 2 * KPX * 128 + (P.halo ? (uint32_t)(BNQ / 64) * 16 * 10 * 128 : (uint32_t)P.tpc * (BNQ / 64) * KPX * 128);
  size_t smem = (size_t)P.stages * stage_bytes + 1024 + (2 * P.stages + 4) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BNQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int grid = items < dn_num_sms() ? items : dn_num_sms();
  if (const char* e = getenv("DN_WGRAD_NONPERSISTENT")) { if (atoi(e) == 1) grid = items; }
  dn_launch(wgrad_tc_kernel<BNQ>, dim3(grid), dim3(192), smem, st, P);
  DN_CHECK_LAUNCH();
  return 0;
}


// ---- halo variant: eligibility + launch -------------------------------------------------------------------------
bool g_halo_enabled = true;
const bool g_halo_ring = []() { const char* e = getenv("DN_HALO_RING"); return !(e && e[0] == '0'); }();   // A/B knob

// fills wt[kh*3+kw]; true when the taps are exactly the 3x3 neighbourhood of one source
// taps form a subset of the 3x3 neighbourhood (all nine for a 3x3 convolution, a 2x2 corner for one output phase of a
// 4x4 / stride-2 transposed convolution): wt[kh * 3 + kw] = packed-weight index, -1 where the tap is absent
bool halo_taps(const dn_igemm* p, int* wt) {
  if (p->ntaps > 9 || p->ntaps < 3 || p->nsrc != 1 || p->stride != 1) return false;
  bool seen[9] = {false};
  for (int i = 0; i < 9; ++i) wt[i] = -1;
  for (int t = 0; t < p->ntaps; ++t) {
    int dh = p->taps[t].dh, dw = p->taps[t].dw;
    if (dh < -1 || dh > 1 || dw < -1 || dw > 1 || p->taps[t].src != 0) return false;
    int i = (dh + 1) * 3 + (dw + 1);
    if (seen[i]) return false;
    seen[i] = true;
    wt[i] = p->taps[t].wt;
  }
  return true;
}

bool halo_eligible(const dn_igemm* p, int* wt) {
  if (!g_halo_enabled || !halo_taps(p, wt)) return false;
  if (p->cout_pad > 256) return false;
  const int BN = pick_bn(p->cout_pad);
  const int kchunks = (p->in[0].C + kChunk - 1) / kChunk;
  const size_t b_bytes = (size_t)9 * kchunks * BN * 128;
  // weights stay resident next to >= 2 halo stages when they fit; else they stream through a ring, which measured a gain only
  // for one-chunk problems (features.7 forward 0.099 -> 0.093 ms): N <= 128 tiles with more K are bound by the 128 B/clk
  // shared-memory operand reads of cta_group::1 either way (A 4 KB + B 4 KB per 64-cycle 128x128x16 MMA)
  if (b_bytes + 2 * kHaloBytes > 200 * 1024 && (BN > 128 || kchunks > 1 || !g_halo_ring || p->ntaps != 9)) return false;
  const int H = p->out.H, W = p->out.W;
  if (H != p->in[0].H || W != p->in[0].W) return false;
  // tile = 16 rows x 8 columns: only worth it when little of the tile grid is padding
  const double eff = (double)H * W / ((double)((H + 15) / 16 * 16) * ((W + 7) / 8 * 8));
  return eff >= 0.85;
}

template <int BN>
int launch_halo(const HaloParams& P, cudaStream_t st) {
  const size_t b_bytes = (size_t)(P.ring ? P.bstages : 9 * P.kchunks) * BN * 128;
  size_t smem = b_bytes + (size_t)P.stages * kHaloBytes + 1024 + (2 * P.stages + 20) * 8 + 16 + 2 * BN * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int grid = P.num_tiles < dn_num_sms() ? P.num_tiles : dn_num_sms();
  dn_launch(igemm_halo_kernel<BN>, dim3(grid), dim3(192), smem, st, P);
  DN_CHECK_LAUNCH();
  return 0;
}

int dn_igemm_halo(const dn_igemm* p, const int* wt, cudaStream_t st) {
  HaloParams P;
  memset(&P, 0, sizeof(P));
  const int BN = pick_bn(p->cout_pad);
  auto enc = get_encode();
  if (!enc) return DN_E_UNSUPPORTED;
  {
    const dn_view& v = p->in[0];
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
    cuuint32_t box[4] = {(cuuint32_t)kChunk, (cuuint32_t)kHaloW, (cuuint32_t)kHaloH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMapDataType dt = v.dtype == DN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (enc(&P.tmA, dt, 4, v.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DN_E_ARG;
  }
  {
    int nw = 0;
    for (int t = 0; t < 9; ++t) nw = wt[t] + 1 > nw ? wt[t] + 1 : nw;
    cuuint64_t dims[3] = {(cuuint64_t)p->cin_pad, (cuuint64_t)p->cout_pad, (cuuint64_t)nw};
    cuuint64_t strides[2] = {(cuuint64_t)p->cin_pad * 2, (cuuint64_t)p->cin_pad * p->cout_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)kChunk, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUtensorMapDataType dt = p->w_dtype == DN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (enc(&P.tmB, dt, 3, (void*)p->w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DN_E_ARG;
  }
  for (int t = 0; t < 9; ++t) P.wt[t] = wt[t];
  const int Cin = p->in[0].C;
  P.kchunks = (Cin + kChunk - 1) / kChunk;
  P.last_ksteps = (Cin - (P.kchunks - 1) * kChunk + 15) / 16;
  P.tilesW = (p->out.W + 7) / 8;
  P.tilesH = (p->out.H + 15) / 16;
  P.num_tiles = P.tilesW * P.tilesH * p->out.N;
  const size_t b_bytes = (size_t)9 * P.kchunks * BN * 128;
  if (b_bytes + 2 * kHaloBytes > 200 * 1024) {
    P.ring = 1;
    P.stages = 2;
    P.bstages = (int)((200 * 1024 - 2 * kHaloBytes) / ((size_t)BN * 128));
    if (P.bstages > 8) P.bstages = 8;
  } else {
    P.stages = (int)((200 * 1024 - b_bytes) / kHaloBytes);
    if (P.stages > 4) P.stages = 4;
  }
  P.n_mma = p->cout_pad < BN ? p->cout_pad : BN;
  P.c_eff = (p->out.dtype != DN_F32 && p->out_pad_ok) ? (p->out.C + 7) / 8 * 8 : p->out.C;
  P.idesc = make_idesc(p->in[0].dtype, p->w_dtype, 0, 0, 128, P.n_mma);
  P.out = p->out;
  P.out2 = p->out2;
  P.out2_dtype = p->out2_dtype;
  P.bias = p->bias;
  P.act = p->act;
  P.accumulate = p->accumulate;
  P.out_scale = p->out_scale;
  switch (BN) {
    case 256: return launch_halo<256>(P, st);
    case 128: return launch_halo<128>(P, st);
    case 64: return launch_halo<64>(P, st);
    case 32: return launch_halo<32>(P, st);
    default: return launch_halo<16>(P, st);
  }
}

unsigned long long* g_tc_dbg = nullptr;

}  // namespace

// debug: device buffer of 8 counters (cycles summed over CTAs): [0] producer wait-empty, [1] producer total,
// [2] MMA wait-full, [3] MMA wait-tmem-empty, [4] MMA total, [5] epilogue wait-tmem-full, [6] epilogue drain+store
DN_EXPORT int dn_tc_set_debug(void* device_counters) {
  g_tc_dbg = (unsigned long long*)device_counters;
  return 0;
}

// 0 switches the shared-memory halo variant off (A/B comparisons, debugging)
DN_EXPORT int dn_tc_set_halo(int enabled) {
  g_halo_enabled = enabled != 0;
  return 0;
}

DN_EXPORT int dn_tc_available(void) {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cached = (major == 10 && get_encode() != nullptr) ? 1 : 0;
  }
  return cached;
}

DN_EXPORT int dn_igemm_tc_supported(const dn_igemm* p) {
  if (!p || p->stride != 1 || p->ntaps > kMaxTcTaps || p->ntaps < 1) return 0;
  if (p->w_dtype != DN_F16 && p->w_dtype != DN_BF16) return 0;
  if (p->out.dtype != DN_F32) {   // 16-byte vector epilogue
    if (!view_tma_ok(p->out)) return 0;
    if ((p->out.C % 8) != 0 && !p->out_pad_ok) return 0;
  }
  if (p->out2 && (p->out.dtype == DN_F32 || p->out2_dtype == DN_F32 || ((uintptr_t)p->out2 % 16) != 0 || p->accumulate)) return 0;
  if ((p->cin_pad % 64) != 0 || (p->cout_pad % 16) != 0 || ((uintptr_t)p->w % 16) != 0) return 0;
  for (int s = 0; s < p->nsrc; ++s) {
    if (!view_tma_ok(p->in[s]) || p->in[s].dtype != p->w_dtype) return 0;
    if (p->in[s].C != p->in[0].C) return 0;
  }
  return 1;
}

int dn_igemm_tc(const dn_igemm* p, cudaStream_t st) {
  {
    int wt[9];
    if (halo_eligible(p, wt)) return dn_igemm_halo(p, wt, st);
  }
  IgemmTcParams P;
  memset(&P, 0, sizeof(P));
  const int BN = pick_bn(p->cout_pad);
  choose_box(p->out.N, p->out.H, p->out.W, kRows, P.wb, P.hb, P.nb);
  for (int s = 0; s < p->nsrc; ++s) {
    int e = make_view_map(&P.tmA[s], p->in[s], P.wb, P.hb, P.nb);
    if (e) return e;
  }
  {
    auto enc = get_encode();
    if (!enc) return DN_E_UNSUPPORTED;
    int nw = 0;
    for (int t = 0; t < p->ntaps; ++t) nw = p->taps[t].wt + 1 > nw ? p->taps[t].wt + 1 : nw;
    cuuint64_t dims[3] = {(cuuint64_t)p->cin_pad, (cuuint64_t)p->cout_pad, (cuuint64_t)nw};
    cuuint64_t strides[2] = {(cuuint64_t)p->cin_pad * 2, (cuuint64_t)p->cin_pad * p->cout_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)kChunk, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUtensorMapDataType dt = p->w_dtype == DN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = enc(&P.tmB, dt, 3, (void*)p->w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DN_E_ARG;
  }
  for (int t = 0; t < p->ntaps; ++t) {
    P.taps[t].src = (int16_t)p->taps[t].src; P.taps[t].dh = (int16_t)p->taps[t].dh;
    P.taps[t].dw = (int16_t)p->taps[t].dw; P.taps[t].wt = (int16_t)p->taps[t].wt;
  }
  P.ntaps = p->ntaps;
  P.nsrc = p->nsrc;
  const int Cin = p->in[0].C;
  P.kchunks = (Cin + kChunk - 1) / kChunk;
  int rem = Cin - (P.kchunks - 1) * kChunk;
  P.last_ksteps = (rem + 15) / 16;
  P.tilesW = (p->out.W + P.wb - 1) / P.wb;
  P.tilesH = (p->out.H + P.hb - 1) / P.hb;
  P.tilesN = (p->out.N + P.nb - 1) / P.nb;
  P.ntile_n = (p->cout_pad + BN - 1) / BN;
  P.num_tiles = P.tilesW * P.tilesH * P.tilesN * P.ntile_n;
  const uint32_t stage_bytes = kRows * 128 + BN * 128;
  P.stages = (int)((200 * 1024) / stage_bytes);
  if (P.stages > 8) P.stages = 8;
  P.n_mma = p->cout_pad < BN ? p->cout_pad : BN;
  P.c_eff = (p->out.dtype != DN_F32 && p->out_pad_ok) ? (p->out.C + 7) / 8 * 8 : p->out.C;
  P.idesc = make_idesc(p->in[0].dtype, p->w_dtype, 0, 0, 128, P.n_mma);
  P.out = p->out;
  P.out2 = p->out2;
  P.out2_dtype = p->out2_dtype;
  P.bias = p->bias;
  P.act = p->act;
  P.accumulate = p->accumulate;
  P.out_scale = p->out_scale;
  P.dbg = g_tc_dbg;
  switch (BN) {
    case 256: return launch_igemm<256>(P, st);
    case 128: return launch_igemm<128>(P, st);
    case 64: return launch_igemm<64>(P, st);
    case 32: return launch_igemm<32>(P, st);
    default: return launch_igemm<16>(P, st);
  }
}

DN_EXPORT int dn_wgrad_tc_supported(const dn_wgrad* p) {
  if (!p || p->stride != 1 || p->ntaps > kMaxTcTaps || p->ntaps < 1) return 0;
  if (!view_tma_ok(p->q)) return 0;
  for (int s = 0; s < p->nsrc; ++s)
    if (!view_tma_ok(p->p[s])) return 0;
  if ((p->cq_pad % 64) != 0 || (p->cp_pad % 8) != 0 || ((uintptr_t)p->dw % 16) != 0) return 0;
  return 1;
}

int dn_wgrad_tc(const dn_wgrad* p, cudaStream_t st) {
  WgradTcParams P;
  memset(&P, 0, sizeof(P));
  const dn_view& P0 = p->p[0];
  // halo mode: the nine taps of a 3x3 stride-1 convolution read one shared x halo box
  bool halo = g_halo_enabled && p->ntaps == 9 && p->nsrc == 1 && P0.H == p->q.H && P0.W == p->q.W;
  if (halo) {
    bool seen[9] = {false};
    for (int t = 0; t < 9 && halo; ++t) {
      int dh = p->taps[t].dh, dw = p->taps[t].dw;
      if (dh < -1 || dh > 1 || dw < -1 || dw > 1 || seen[(dh + 1) * 3 + dw + 1]) halo = false;
      else seen[(dh + 1) * 3 + dw + 1] = true;
    }
    const double eff = (double)P0.H * P0.W / ((double)((P0.H + 7) / 8 * 8) * ((P0.W + 7) / 8 * 8));
    if (eff < 0.85) halo = false;
    // the halo box (20 KB per 64 channels) only pays off when several taps of one work item share it: a wide output
    // tile leaves TMEM room for one or two taps per item, and then nine plain 8 KB boxes are cheaper than nine halo boxes
    const int n16 = (p->q.C + 15) / 16 * 16;
    const int n_mma_est = p->cq_pad >= 256 ? 256 : (n16 < 64 ? n16 : (p->cq_pad >= 128 ? 128 : 64));
    if (512 / n_mma_est < 3 || (n_mma_est >= 128)) halo = false;
  }
  P.halo = halo ? 1 : 0;
  if (halo) { P.wb = 8; P.hb = 8; P.nb = 1; }
  else choose_box(P0.N, P0.H, P0.W, KPX, P.wb, P.hb, P.nb);
  for (int s = 0; s < p->nsrc; ++s) {
    int e = make_view_map(&P.tmP[s], p->p[s], P.wb, P.hb, P.nb);
    if (e) return e;
  }
  {
    int e = halo ? make_view_map(&P.tmQ, p->q, 16, 10, 1) : make_view_map(&P.tmQ, p->q, P.wb, P.hb, P.nb);
    if (e) return e;
  }
  for (int t = 0; t < p->ntaps; ++t) {
    P.taps[t].src = (int16_t)p->taps[t].src; P.taps[t].dh = (int16_t)p->taps[t].dh;
    P.taps[t].dw = (int16_t)p->taps[t].dw; P.taps[t].wt = (int16_t)p->taps[t].wt;
  }
  P.ntaps = p->ntaps;
  P.tilesW = (P0.W + P.wb - 1) / P.wb;
  P.tilesH = (P0.H + P.hb - 1) / P.hb;
  P.tilesN = (P0.N + P.nb - 1) / P.nb;
  P.num_ptiles = P.tilesW * P.tilesH * P.tilesN;
  const int BNQ = p->cq_pad >= 256 ? 256 : (p->cq_pad >= 128 ? 128 : 64);
  P.cp_tiles = (P0.C + 127) / 128;
  P.cq_tiles = (p->cq_pad + BNQ - 1) / BNQ;
  P.cp_blocks = P0.C > 64 ? 2 : 1;
  P.n_mma = BNQ;
  if (P.cq_tiles == 1) {
    int n16 = (p->q.C + 15) / 16 * 16;
    if (n16 < BNQ) P.n_mma = n16;
  }
  // taps that share one dy tile in a CTA: bounded by TMEM columns (512) and by keeping >= 3 pipeline stages
  const uint32_t a_bytes = 2 * KPX * 128, b_bytes = (BNQ / 64) * KPX * 128;
  int tpc = 512 / P.n_mma;
  int by_smem = (int)((64 * 1024 - a_bytes) / b_bytes);
  if (by_smem < 1) by_smem = 1;
  if (!halo && tpc > by_smem) tpc = by_smem;
  if (tpc > P.ntaps) tpc = P.ntaps;
  // wide tiles: prefer two accumulator stages (<= 256 columns per stage) over sharing the dy tile between more taps
  if (P.n_mma >= 128 && tpc * P.n_mma > 256) tpc = 256 / P.n_mma;
  // all taps of a group (consecutive taps [g * tpc, (g + 1) * tpc)) must read the same dy view: the four output phases of a
  // transposed convolution arrive as one problem with the taps ordered phase by phase
  auto uniform = [&](int n) {
    for (int t = 0; t < P.ntaps; ++t)
      if (p->taps[t].src != p->taps[(t / n) * n].src) return false;
    return true;
  };
  for (;;) {
    P.ngroups = (P.ntaps + tpc - 1) / tpc;
    P.tpc = (P.ntaps + P.ngroups - 1) / P.ngroups;
    P.ngroups = (P.ntaps + P.tpc - 1) / P.tpc;
    if (P.tpc == 1 || uniform(P.tpc)) break;
    tpc = P.tpc - 1;
  }
  int cols = P.tpc * P.n_mma;
  P.nacc = cols <= 256 ? 2 : 1;
  if (const char* e = getenv("DN_WGRAD_NACC")) { if (atoi(e) == 1) P.nacc = 1; }
  cols *= P.nacc;
  P.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  const int out_tiles = P.ngroups * P.cp_tiles * P.cq_tiles;
  // split-K factor: about two waves of CTAs, chosen so that the last wave is as full as possible
  int splits = 1;
  {
    const int sms = dn_num_sms();
    int lo = (sms + out_tiles - 1) / out_tiles, hi = (2 * sms + out_tiles - 1) / out_tiles;
    if (hi > P.num_ptiles) hi = P.num_ptiles;
    if (lo > hi) lo = hi;
    if (lo < 1) lo = 1;
    double best = -1.0;
    for (int sp = lo; sp <= hi; ++sp) {
      const int items = out_tiles * sp;
      const double eff = (double)items / ((double)((items + sms - 1) / sms) * sms);
      const double score = eff - 0.02 * (double)sp / (double)(hi > 0 ? hi : 1);   // mild preference for fewer atomics
      if (score > best) { best = score; splits = sp; }
    }
  }
  P.ptiles_per_split = (P.num_ptiles + splits - 1) / splits;
  P.splits = (P.num_ptiles + P.ptiles_per_split - 1) / P.ptiles_per_split;
  const uint32_t stage_bytes = a_bytes + (halo ? (uint32_t)(BNQ / 64) * 16 * 10 * 128 : (uint32_t)P.tpc * b_bytes);
  P.stages = (int)((200 * 1024) / stage_bytes);
  if (P.stages > 8) P.stages = 8;
  P.m64 = 0;
  if (P.cp_blocks == 1) {
    P.m64 = 1;
    if (const char* e = getenv("DN_WGRAD_M64")) P.m64 = atoi(e);
  }
  P.idesc = make_idesc(P0.dtype, p->q.dtype, 1, 1, P.m64 ? 64 : 128, P.n_mma);
  P.dw = p->dw;
  P.cp = P0.C; P.cq = p->q.C; P.cp_pad = p->cp_pad; P.cq_pad = p->cq_pad;
  P.scale = p->scale;
  const int items = out_tiles * P.splits;
  P.num_items = items;
  switch (BNQ) {
    case 256: return launch_wgrad<256>(P, items, st);
    case 128: return launch_wgrad<128>(P, items, st);
    default: return launch_wgrad<64>(P, items, st);
  }
}
