// tcgen05 / TMA / TMEM implicit-GEMM convolution kernels for sm_100a (backend 1 of dn_igemm_run / dn_wgrad_run).
//
//   igemm_tc_kernel : out[pixel][co] = sum_tap sum_ci in[pixel + tap][ci] * w[tap][co][ci]
//       A (activations) is fetched by 4-D TMA boxes {cb ch, wb, hb, nb} straight from the NHWC view -- the tap shift is a
//       coordinate offset and zero padding is TMA out-of-bounds fill, so there is no im2col buffer; cb = 64 / 32 / 16 channels
//       per row (128B / 64B / 32B swizzle) so that thin tensors are fetched as whole rows.  B (packed weights) by 3-D TMA.
//       tcgen05.mma (M=128, N<=256, K=16) with fp32 accumulators in TMEM.  Persistent CTAs, ten warps: TMA producer, MMA
//       issuer (one asm block per pipeline stage, warp-uniform), eight epilogue warps (two per TMEM lane quarter); a multi-stage
//       smem ring and two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.  The output
//       phases of a transposed convolution run as one launch (dn_igemm.nphase).
//   igemm_halo_kernel : the same for 3x3 stride-1 layers with many pixels: one (16+2) x 16-pixel halo box per 64-channel chunk,
//       the nine taps are nine shifted UMMA descriptors over it, weights resident in shared memory; two CTAs per SM for thin
//       tiles; channel-stacked output phases of thin transposed convolutions with a TMA-store epilogue.
//   wgrad_tc_kernel : dw[tap][cp][cq] += sum_pixel dy[pixel][cp] * x[pixel + tap][cq]
//       both operands are "MN-major" (the reduction index = pixel is the row index of the TMA box), split-K over pixel
//       tiles across CTAs, fp32 partial sums reduced with vector red.global.add; halo mode (one x box for nine taps) and an
//       M-stacked mode for thin dy (three column-shifted dy copies fill the MMA M dimension).
//
// Forward nn.Conv2d (stride 1), the output phases of nn.ConvTranspose2d, and the data gradients of both run on the igemm
// kernels with different tap tables / views; weight gradients of both on wgrad_tc_kernel.  What bounds
// these kernels was measured with tools/ubench_tc.cu and is written up in DESIGN.md 4.1.
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
// polling wait with a back-off: for roles that are not on the critical path (their polling must not take issue slots from the
// MMA-issuing warp that shares their scheduler)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (;;) {
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(200);
  }
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
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// TMA store of a shared-memory box (bulk-group completion)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)m), "r"(smem_u32(smem)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- "one lane issues" forms ------------------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit / TMA loads are issued by ONE lane.  Wrapping them in `if (elect_one()) { ... }` makes the
// branch divergent for the compiler: every stage then pays BRA.DIV / BSYNC scaffolding and moves its operands between the
// vector and the uniform register files (~100 SASS instructions per pipeline stage, more than the ~300 tensor-pipe cycles of
// a 4-MMA N=128 stage can hide).  With the election INSIDE the asm block the surrounding code stays warp-uniform: ptxas keeps
// descriptors and barrier addresses in uniform registers and emits a bare UTCHMMA / UTCBAR / UTMALDG.
__device__ __forceinline__ void mbar_expect_tx_e(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n.reg .pred E;\nelect.sync _|E, 0xffffffff;\n@E mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_e(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n.reg .pred E;\nelect.sync _|E, 0xffffffff;\n"
      "@E cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n}\n" ::"r"(
          smem_u32(smem)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_e(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n.reg .pred E;\nelect.sync _|E, 0xffffffff;\n"
      "@E cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n}\n" ::"r"(
          smem_u32(smem)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, E;\n"
      "elect.sync _|E, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@E tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_acc_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p, E;\n"
      "elect.sync _|E, 0xffffffff;\n"
      "setp.eq.u32 p, 1, 1;\n"
      "@E tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
  asm volatile("{\n.reg .pred E;\nelect.sync _|E, 0xffffffff;\n@E tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(
                   smem_u32(bar))
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
template <int NV>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t* r) {
  if (NV == 8) tmem_ld8(taddr, r);
  else if (NV == 16) tmem_ld16(taddr, r);
  else tmem_ld32(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// shared-memory matrix descriptor, 128B swizzle (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1
// <<46 | layout SWIZZLE_128B (2) << 61
// layout: 2 = SWIZZLE_128B (64 16-bit channels per row), 4 = SWIZZLE_64B (32 channels), 6 = SWIZZLE_32B (16 channels)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}


// The MMAs of one (tile, chunk): 9 shifted views x KS K-steps plus the commit(s), as ONE asm block.  The issuing warp's
// instruction stream paces thin layers (an M=128, N<=32 MMA takes ~50 cycles on the tensor pipe -- tools/ubench_tc.cu), so
// the block is branch-free: the election happens once, every operand offset is a PTX constant expression of the template
// constants, absent taps (tap subsets of transposed-convolution phases / row-expanded first layers) are predicated off.
// operands: %0 d_tmem, %1 A descriptor of the stage, %2 B descriptor of the chunk, %3 idesc, %4 accumulate flag of the first
// MMA, %5 empty barrier, %6 tap mask, %7 A row bytes, %8 B tile bytes, %9 tmem-full barrier, %10 "last chunk" flag
#define DN_HTAP_HEAD(KH, KW, T)                                   \
  "and.b32 m, %6, (1 << " #T ");\n"                               \
  "setp.ne.u32 Tp, m, 0;\n"                                       \
  "and.pred Q, Tp, E;\n"                                          \
  "add.u64 a, %1, ((" #KH " * 16 + " #KW ") * %7) / 16;\n"        \
  "add.u64 b, %2, (" #T " * %8) / 16;\n"                          \
  "@Q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, A;\n"    \
  "or.pred A, A, Tp;\n"
#define DN_HTAP_STEP                                              \
  "add.u64 a, a, 2;\n"                                            \
  "add.u64 b, b, 2;\n"                                            \
  "@Q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, TR;\n"
#define DN_HTAP1(KH, KW, T) DN_HTAP_HEAD(KH, KW, T)
#define DN_HTAP2(KH, KW, T) DN_HTAP_HEAD(KH, KW, T) DN_HTAP_STEP
#define DN_HTAP3(KH, KW, T) DN_HTAP_HEAD(KH, KW, T) DN_HTAP_STEP DN_HTAP_STEP
#define DN_HTAP4(KH, KW, T) DN_HTAP_HEAD(KH, KW, T) DN_HTAP_STEP DN_HTAP_STEP DN_HTAP_STEP
#define DN_HALO_BLOCK(TAP)                                                                                                   \
  asm volatile(                                                                                                              \
      "{\n.reg .pred E, A, TR, Tp, Q, L;\n.reg .b64 a, b;\n.reg .b32 m;\n"                                                   \
      "elect.sync _|E, 0xffffffff;\n"                                                                                        \
      "setp.ne.u32 A, %4, 0;\nsetp.eq.u32 TR, 1, 1;\n"                                                                       \
      TAP(0, 0, 0) TAP(0, 1, 1) TAP(0, 2, 2) TAP(1, 0, 3) TAP(1, 1, 4) TAP(1, 2, 5) TAP(2, 0, 6) TAP(2, 1, 7) TAP(2, 2, 8)   \
      "@E tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n"                                     \
      "setp.ne.u32 L, %10, 0;\nand.pred L, L, E;\n"                                                                          \
      "@L tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n"                                     \
      "}\n" ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc0), "r"(empty_bar), "r"(mask), "n"(A_RB), "n"(B_BYTES),      \
      "r"(tfull_bar), "r"(last)                                                                                              \
      : "memory")
template <int KS, int A_RB, int B_BYTES>
__device__ __forceinline__ void halo_issue(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t mask, uint32_t acc0,
                                           uint32_t empty_bar, uint32_t tfull_bar, uint32_t last) {
  if (KS == 1) DN_HALO_BLOCK(DN_HTAP1);
  else if (KS == 2) DN_HALO_BLOCK(DN_HTAP2);
  else if (KS == 3) DN_HALO_BLOCK(DN_HTAP3);
  else DN_HALO_BLOCK(DN_HTAP4);
}

// One (tap, chunk) of the plain kernel: KS MMAs on its A / B tiles, release of the stage, optional "accumulator complete"
// commit.  %0 d_tmem, %1 / %2 descriptors, %3 idesc, %4 accumulate flag of the first MMA, %5 empty barrier, %6 tmem-full
// barrier, %7 flags: bit 0 = last MMAs of the tile (commit tmem-full), bit 1 = more taps of the same stage follow (no release)
#define DN_STEP2 "add.u64 a, a, 2;\nadd.u64 b, b, 2;\n@E tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, TR;\n"
#define DN_STAGE_BLOCK(STEPS)                                                                                       \
  asm volatile(                                                                                                     \
      "{\n.reg .pred E, A, TR, L;\n.reg .b64 a, b;\n.reg .b32 fl;\n"                                               \
      "elect.sync _|E, 0xffffffff;\n"                                                                               \
      "setp.ne.u32 A, %4, 0;\nsetp.eq.u32 TR, 1, 1;\n"                                                              \
      "mov.b64 a, %1;\nmov.b64 b, %2;\n"                                                                            \
      "@E tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, A;\n" STEPS                                            \
      "and.b32 fl, %7, 2;\nsetp.eq.u32 L, fl, 0;\nand.pred L, L, E;\n"                                             \
      "@L tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n"                            \
      "and.b32 fl, %7, 1;\nsetp.ne.u32 L, fl, 0;\nand.pred L, L, E;\n"                                             \
      "@L tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"                            \
      "}\n" ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc0), "r"(empty_bar), "r"(tfull_bar), "r"(last)       \
      : "memory")
__device__ __forceinline__ void stage_issue(int ks, uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc0, uint32_t empty_bar,
                                            uint32_t tfull_bar, uint32_t last) {
  if (ks == 4) DN_STAGE_BLOCK(DN_STEP2 DN_STEP2 DN_STEP2);
  else if (ks == 1) DN_STAGE_BLOCK("");
  else if (ks == 2) DN_STAGE_BLOCK(DN_STEP2);
  else DN_STAGE_BLOCK(DN_STEP2 DN_STEP2);
}

// Four K steps (64 pixels) of one tap of the weight gradient: A advances by `ka`, B by `kb` 16-byte units per step
__device__ __forceinline__ void wg_issue4(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc0, uint32_t kb, uint32_t ka = 128) {
  asm volatile(
      "{\n.reg .pred E, A, TR;\n.reg .b64 a, b, ka, kb;\n"
      "elect.sync _|E, 0xffffffff;\n"
      "setp.ne.u32 A, %4, 0;\nsetp.eq.u32 TR, 1, 1;\n"
      "mov.b64 a, %1;\nmov.b64 b, %2;\ncvt.u64.u32 kb, %5;\ncvt.u64.u32 ka, %6;\n"
      "@E tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, A;\n"
      "add.u64 a, a, ka;\nadd.u64 b, b, kb;\n@E tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, TR;\n"
      "add.u64 a, a, ka;\nadd.u64 b, b, kb;\n@E tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, TR;\n"
      "add.u64 a, a, ka;\nadd.u64 b, b, kb;\n@E tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, TR;\n"
      "}\n" ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc0), "r"(kb), "r"(ka)
      : "memory");
}

// previous contents of an accumulate target (16-bit output): all 16-byte loads of the chunk are issued together -- and, where
// the caller can, before it waits for the accumulator -- instead of one dependent load per eight channels
template <int NV>
__device__ __forceinline__ void epi_prefetch(uint4* old, bool valid, int c_base, int c_eff, const uint8_t* optr_c) {
#pragma unroll
  for (int g = 0; g < NV / 8; ++g) {
    old[g] = make_uint4(0, 0, 0, 0);
    if (valid && c_base + 8 * g < c_eff) old[g] = *reinterpret_cast<const uint4*>(optr_c + 16 * g);
  }
}
__device__ __forceinline__ void epi_add_old(float* v, const uint4& u, int out_dtype) {
  if (out_dtype == DN_F16) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); v[2 * i] += t.x; v[2 * i + 1] += t.y; }
  } else {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); v[2 * i] += t.x; v[2 * i + 1] += t.y; }
  }
}

// one epilogue chunk: NV (8 / 16 / 32) accumulator columns of one pixel -> bias / activation -> one 16-byte store per 8 channels.
// `old` = epi_prefetch() of the same chunk when accumulate is set and the output is 16-bit, else unused
template <int NV>
__device__ __forceinline__ void epi_store(const uint32_t* r, const float* bs, float out_scale, int act, bool valid, int c_base, int c_eff, int out_C,
                                          int out_dtype, uint8_t* optr_c, int accumulate, void* out2, int out2_dtype, size_t elem_off,
                                          const uint4* old) {
  if (!valid) return;
#pragma unroll
  for (int g = 0; g < NV / 8; ++g) {
    const int c0 = c_base + 8 * g;
    if (c0 >= c_eff) break;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = dn_act(__uint_as_float(r[8 * g + j]) * out_scale + bs[8 * g + j], act);
    if (out2) {
      if (out2_dtype == DN_F16) Vec8<__half>::store((__half*)out2 + elem_off + 8 * g, v);
      else Vec8<__nv_bfloat16>::store((__nv_bfloat16*)out2 + elem_off + 8 * g, v);
    }
    if (out_dtype == DN_F32) {
      float* o = (float*)optr_c + 8 * g;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < out_C) o[j] = accumulate ? o[j] + v[j] : v[j];
    } else if (out_dtype == DN_F16) {
      if (accumulate) epi_add_old(v, old[g], out_dtype);
      Vec8<__half>::store((__half*)optr_c + 8 * g, v);
    } else {
      if (accumulate) epi_add_old(v, old[g], out_dtype);
      Vec8<__nv_bfloat16>::store((__nv_bfloat16*)optr_c + 8 * g, v);
    }
  }
}

// Epilogue chunk for the TMA-store path: NV accumulator columns of tile row `row` -> bias / activation -> 16-byte chunks of the
// staging tile(s) in shared memory, laid out as the TMA store boxes expect them: blocks of `ob` channels, [128 rows][2 * ob bytes]
// each, 16-byte chunks XOR-swizzled like the matching TMA swizzle mode (ob = 64 / 32 / 16 -> 128B / 64B / 32B).
// Why: a row-per-thread epilogue that stores to global memory touches 32 different 128-byte lines per st.v4 instruction
// (32 LSU wavefronts each); through shared memory the same bytes cost 4 wavefronts and one bulk store per tile.
template <int NV>
__device__ __forceinline__ void epi_stage(const uint32_t* r, const float* bs, float out_scale, int act, int c_base, int ob_shift, int row,
                                          uint8_t* stg, int out_dtype, uint8_t* stg2, int out2_dtype) {
  const int ob = 1 << ob_shift;
  const uint32_t swz = ob == 64 ? (uint32_t)(row & 7) : ob == 32 ? (uint32_t)((row >> 1) & 3) : (uint32_t)((row >> 2) & 1);
  const uint32_t row_off = (uint32_t)row * (uint32_t)(2 * ob);
  const uint32_t blk_bytes = 128u * (uint32_t)(2 * ob);
#pragma unroll
  for (int g = 0; g < NV / 8; ++g) {
    const int c0 = c_base + 8 * g;
    const uint32_t j = (uint32_t)(c0 >> ob_shift), ci = (uint32_t)((c0 & (ob - 1)) >> 3);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = dn_act(__uint_as_float(r[8 * g + i]) * out_scale + bs[8 * g + i], act);
    const uint32_t off = j * blk_bytes + row_off + ((ci ^ swz) << 4);
    if (out_dtype == DN_F16) Vec8<__half>::store((__half*)(stg + off), v);
    else Vec8<__nv_bfloat16>::store((__nv_bfloat16*)(stg + off), v);
    if (stg2) {
      if (out2_dtype == DN_F16) Vec8<__half>::store((__half*)(stg2 + off), v);
      else Vec8<__nv_bfloat16>::store((__nv_bfloat16*)(stg2 + off), v);
    }
  }
}

// epilogue chunk of a channel-stacked phase problem: every 8-channel group belongs to one phase (cpp is a multiple of 8) and goes
// to that phase's pixel; no accumulate form (forward of transposed convolutions only)
template <int NV>
__device__ __forceinline__ void epi_store_phases(const uint32_t* r, const float* bs, float out_scale, int act, bool valid, int c_base, int cpp_shift,
                                                 int c_eff, int out_dtype, uint8_t* out_px, const long long* phase_off, void* out2,
                                                 int out2_dtype, size_t px_off) {
  if (!valid) return;
#pragma unroll
  for (int g = 0; g < NV / 8; ++g) {
    const int c0 = c_base + 8 * g;
    const int ph = c0 >> cpp_shift, cc = c0 - (ph << cpp_shift);      // (cpp is a power of two: no division in the epilogue)
    if (ph >= 4 || cc >= c_eff) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = dn_act(__uint_as_float(r[8 * g + j]) * out_scale + bs[8 * g + j], act);
    const size_t e = (size_t)(phase_off[ph] + cc);
    if (out2) {
      if (out2_dtype == DN_F16) Vec8<__half>::store((__half*)out2 + px_off + e, v);
      else Vec8<__nv_bfloat16>::store((__nv_bfloat16*)out2 + px_off + e, v);
    }
    if (out_dtype == DN_F16) Vec8<__half>::store((__half*)out_px + e, v);
    else Vec8<__nv_bfloat16>::store((__nv_bfloat16*)out_px + e, v);
  }
}

constexpr int kMaxTcTaps = DN_MAX_TAPS;
constexpr int kRows = 128;          // pixels per tile = UMMA M
constexpr int kChunk = 64;          // channels per K chunk = one 128-byte swizzled row
// Thin tensors (C <= 32) are fetched as 32- or 16-channel rows instead (SWIZZLE_64B / SWIZZLE_32B boxes): a 64-channel box
// over a 17-channel tensor makes every box row partially out of bounds, and TMA fills such rows at ~7 cycles each
// (tools/ubench_tc.cu: 1 940 cycles per 18x16-pixel halo box against 750-850 for the same box with 32-channel rows).
__host__ __device__ __forceinline__ int thin_layout(int cb) { return cb == 64 ? 2 : cb == 32 ? 4 : 6; }
static inline int thin_cb(int C) { return C <= 16 ? 16 : C <= 32 ? 32 : 64; }

struct TcTap { int16_t src, dh, dw, wt; };

struct IgemmTcParams {
  CUtensorMap tmA[DN_MAX_SRC];
  CUtensorMap tmB;
  TcTap taps[kMaxTcTaps];
  int ntaps, nsrc;
  int kchunks;        // ceil(Cin / cb)
  int last_ksteps;    // UMMA_K steps in the last chunk (1..4)
  int cb;             // channels per A row in shared memory: 64 (128-byte rows), 32 or 16 (thin tensors, one chunk)
  int wb, hb, nb;     // pixel box (wb*hb*nb == 128)
  int tilesW, tilesH, tilesN, ntile_n, num_tiles;
  int nph, tph;       // merged output phases: tile index = ((((ph * tilesN + tn) * tilesH + th) * tilesW + tw) * ntile_n + nt; phase ph uses
                      // taps [ph * tph, (ph + 1) * tph) and writes at out + phase_off[ph]
  long long phase_off[4];
  int step[5];        // gridDim.x as mixed-radix digits (nt, tw, th, tn, ph): tiles advance by carries, not by divisions
  int stages;
  int tps;            // taps per pipeline stage (thin one-chunk problems with many taps: a 7x7 / 5x5 kernel, the 16 taps of a
                      // transposed-convolution data gradient): one barrier round trip per tps taps instead of one per tap
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

// position of a persistent CTA's current tile in the (nt, tw, th, tn, ph) index space
struct TilePos { int nt, tw, th, tn, ph; };
__device__ __forceinline__ void tile_init(TilePos& t, const IgemmTcParams& p, int tile) {
  t.nt = tile % p.ntile_n; tile /= p.ntile_n;
  t.tw = tile % p.tilesW; tile /= p.tilesW;
  t.th = tile % p.tilesH; tile /= p.tilesH;
  t.tn = tile % p.tilesN;
  t.ph = tile / p.tilesN;
}
__device__ __forceinline__ void tile_next(TilePos& t, const IgemmTcParams& p) {
  t.nt += p.step[0];
  int c = t.nt >= p.ntile_n ? 1 : 0; t.nt -= c * p.ntile_n;
  t.tw += p.step[1] + c;
  c = t.tw >= p.tilesW ? 1 : 0; t.tw -= c * p.tilesW;
  t.th += p.step[2] + c;
  c = t.th >= p.tilesH ? 1 : 0; t.th -= c * p.tilesH;
  t.tn += p.step[3] + c;
  c = t.tn >= p.tilesN ? 1 : 0; t.tn -= c * p.tilesN;
  t.ph += p.step[4] + c;
}

template <int BN>
__global__ void __launch_bounds__(320, 1) igemm_tc_kernel(const __grid_constant__ IgemmTcParams p) {
  dn_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t A_BYTES = (uint32_t)(kRows * 2 * p.cb);
  constexpr uint32_t B_BYTES = BN * 128;
  const uint32_t SUB_BYTES = A_BYTES + B_BYTES;                 // one (tap, chunk): A tile + B tile
  const uint32_t STAGE_BYTES = (uint32_t)p.tps * SUB_BYTES;
  const uint32_t a_layout = (uint32_t)thin_layout(p.cb), a_sbo = (uint32_t)(16 * p.cb);     // 8 rows of 2 * cb bytes
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr int EW = 8;                               // epilogue warps: two per TMEM lane quarter, half of the columns each
  constexpr int CPT = BN / 2 < 8 ? 8 : BN / 2;        // accumulator columns per epilogue thread
  constexpr int CH = CPT < 32 ? CPT : 32;
  // carve: [stages x (A|B)] | barriers | tmem ptr | bias tile
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tempty_bar + 2);
  float* bias_s = (float*)(tmem_ptr + 2);   // [2][BN]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // (tells the compiler the role branches are warp-uniform)
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EW); }     // one arrival per epilogue warp
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

  if (warp == 0) {
    // ================= TMA producer (whole warp runs the loop, one elected lane issues) =================
    int stage = 0; uint32_t phase = 0;
    long long dbg_acc0 = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    TilePos tp;
    tile_init(tp, p, blockIdx.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, tile_next(tp, p)) {
      const int nt = tp.nt;
      const int w0 = tp.tw * p.wb, h0 = tp.th * p.hb, n0 = tp.tn * p.nb;
      const int tbeg = tp.ph * p.tph;
      if (p.tps > 1) {      // (one chunk) several taps share a stage
        for (int t = tbeg; t < tbeg + p.tph; t += p.tps) {
          const int nsub = tbeg + p.tph - t < p.tps ? tbeg + p.tph - t : p.tps;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], (uint32_t)nsub * SUB_BYTES);
            for (int j = 0; j < nsub; ++j) {
              const TcTap tap = p.taps[t + j];
              tma_load_4d(sa + (size_t)j * SUB_BYTES, &p.tmA[tap.src], &full_bar[stage], 0, w0 + tap.dw, h0 + tap.dh, n0);
              tma_load_3d(sa + (size_t)j * SUB_BYTES + A_BYTES, &p.tmB, &full_bar[stage], 0, nt * BN, tap.wt);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      } else
      for (int t = tbeg; t < tbeg + p.tph; ++t) {
        const TcTap tap = p.taps[t];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.dbg) dbg_acc0 += clock64() - t0;
          if (elect_one()) {
            uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_4d(sa, &p.tmA[tap.src], &full_bar[stage], kc * p.cb, w0 + tap.dw, h0 + tap.dh, n0);
            tma_load_3d(sb, &p.tmB, &full_bar[stage], kc * p.cb, nt * BN, tap.wt);
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
    // The issuing thread's instruction stream paces N <= 128 tiles (a 128 x 128 x 16 MMA occupies the tensor pipe for ~70
    // cycles): the descriptors are running values advanced by one add per stage, the K steps are unrolled per count.
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint64_t ad_base = make_desc(smem_u32(smem), 16, a_sbo, a_layout);
    const uint64_t bd_base = make_desc(smem_u32(smem) + A_BYTES, 16, 1024);
    const uint64_t stage_inc = (uint64_t)(STAGE_BYTES >> 4);
    uint64_t ad = ad_base, bd = bd_base;
    const uint32_t idesc = p.idesc;
    const int kchunks = p.kchunks, last_ks = p.last_ksteps, full_ks = p.cb / 16, k_iters_m1 = p.tph * p.kchunks - 1;
    long long dbg_full = 0, dbg_te = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (p.dbg) dbg_te += clock64() - t0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      int kc = 0;
      if (p.tps > 1) {      // (one chunk) several taps per stage: one wait and one release per stage
        const uint64_t sub_inc = (uint64_t)(SUB_BYTES >> 4);
        for (int t = 0; t < p.tph; t += p.tps) {
          const int nsub = p.tph - t < p.tps ? p.tph - t : p.tps;
          t0 = p.dbg ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (p.dbg) dbg_full += clock64() - t0;
          tc_fence_after();
          const uint32_t eb = smem_u32(&empty_bar[stage]), tb = smem_u32(&tfull_bar[acc]);
          for (int j = 0; j < nsub; ++j) {
            const uint32_t flags = (j + 1 < nsub ? 2u : 0u) | ((t + j == p.tph - 1) ? 1u : 0u);
            stage_issue(last_ks, d_tmem, ad + (uint64_t)j * sub_inc, bd + (uint64_t)j * sub_inc, idesc, (t + j) == 0 ? 0u : 1u, eb, tb, flags);
          }
          ad += stage_inc; bd += stage_inc;
          if (++stage == stages) { stage = 0; phase ^= 1; ad = ad_base; bd = bd_base; }
        }
      } else
      for (int it = 0; it <= k_iters_m1; ++it) {
        t0 = p.dbg ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        if (p.dbg) dbg_full += clock64() - t0;
        tc_fence_after();
        const int ks = (kc == kchunks - 1) ? last_ks : full_ks;
        stage_issue(ks, d_tmem, ad, bd, idesc, it == 0 ? 0u : 1u, smem_u32(&empty_bar[stage]), smem_u32(&tfull_bar[acc]),
                    it == k_iters_m1 ? 1u : 0u);
        if (++kc == kchunks) kc = 0;
        ad += stage_inc; bd += stage_inc;
        if (++stage == stages) { stage = 0; phase ^= 1; ad = ad_base; bd = bd_base; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      atomicAdd(p.dbg + 2, (unsigned long long)dbg_full); atomicAdd(p.dbg + 3, (unsigned long long)dbg_te);
      atomicAdd(p.dbg + 4, (unsigned long long)(clock64() - tstart));
    }
  } else {
    // ================= epilogue warps (2..9): TMEM -> registers -> bias/act -> HBM =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // tile row = TMEM lane = pixel within the box
    const int et = threadIdx.x - 64;        // 0..255
    const int cbeg = (warp >= 6 ? 1 : 0) * CPT;
    int acc = 0; uint32_t acc_phase = 0;
    const int wi = row % p.wb;
    const int hi = (row / p.wb) % p.hb;
    const int ni = row / (p.wb * p.hb);
    const int esz = p.out.dtype == DN_F32 ? 4 : 2;
    const int out_dtype = p.out.dtype, out_C = p.out.C, c_eff = p.c_eff, act = p.act, accumulate = p.accumulate, n_mma = p.n_mma;
    const float out_scale = p.out_scale;
    float breg[CH];
    int nt_loaded = -1;
    TilePos tp;
    tile_init(tp, p, blockIdx.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, tile_next(tp, p)) {
      const int nt = tp.nt;
      const int w = tp.tw * p.wb + wi, h = tp.th * p.hb + hi, n = tp.tn * p.nb + ni;
      const bool valid = (w < p.out.W) && (h < p.out.H) && (n < p.out.N);
      const int co0 = nt * BN;
      float* bs = bias_s + acc * BN;
      if (CPT <= 32) {       // bias in registers, reloaded only when the N tile changes (it never does for Cout <= 256)
        if (nt != nt_loaded) {
#pragma unroll
          for (int j = 0; j < CH; ++j) breg[j] = (p.bias && co0 + cbeg + j < p.out.C) ? p.bias[co0 + cbeg + j] : 0.f;
          nt_loaded = nt;
        }
      } else {
        for (int c = et; c < BN; c += EW * 32) bs[c] = (p.bias && co0 + c < p.out.C) ? p.bias[co0 + c] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      }
      const size_t eoff = (size_t)(dn_off(p.out, n, h, w) + co0 + p.phase_off[tp.ph]);
      uint8_t* optr = (uint8_t*)p.out.ptr + eoff * esz;
      uint4 old[CH / 8];
      if (CPT <= 32 && accumulate && esz == 2) epi_prefetch<CH>(old, valid && cbeg < n_mma, cbeg, c_eff - co0, optr + (size_t)cbeg * esz);
      const long long te0 = p.dbg ? clock64() : 0;
      mbar_wait(&tfull_bar[acc], acc_phase);
      const long long te1 = p.dbg ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const int ce = c_eff - co0, oc = out_C - co0;      // channel limits relative to this N tile
      if (CPT <= 32) {
        uint32_t ra[CH];
        if (cbeg < n_mma) {
          tmem_ldn<CH>(taddr + cbeg, ra);
          tmem_ld_wait();
          epi_store<CH>(ra, breg, out_scale, act, valid, cbeg, ce, oc, out_dtype, optr + (size_t)cbeg * esz, accumulate, p.out2, p.out2_dtype, eoff + cbeg, old);
        }
      } else {
        uint32_t ra[32], rb[32];
        uint4 old4[4];
        if (cbeg < n_mma) tmem_ld32(taddr + cbeg, ra);
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + CPT && c0 < n_mma; c0 += 64) {
          if (accumulate && esz == 2) epi_prefetch<32>(old4, valid, c0, ce, optr + (size_t)c0 * esz);
          tmem_ld_wait();
          if (c0 + 32 < cbeg + CPT && c0 + 32 < n_mma) tmem_ld32(taddr + c0 + 32, rb);
          epi_store<32>(ra, bs + c0, out_scale, act, valid, c0, ce, oc, out_dtype, optr + (size_t)c0 * esz, accumulate, p.out2, p.out2_dtype, eoff + c0, old4);
          if (!(c0 + 32 < cbeg + CPT && c0 + 32 < n_mma)) break;
          if (accumulate && esz == 2) epi_prefetch<32>(old4, valid, c0 + 32, ce, optr + (size_t)(c0 + 32) * esz);
          tmem_ld_wait();
          if (c0 + 64 < cbeg + CPT && c0 + 64 < n_mma) tmem_ld32(taddr + c0 + 64, ra);
          epi_store<32>(rb, bs + c0 + 32, out_scale, act, valid, c0 + 32, ce, oc, out_dtype, optr + (size_t)(c0 + 32) * esz, accumulate, p.out2, p.out2_dtype, eoff + c0 + 32, old4);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
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


struct HaloParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int kchunks, last_ksteps;
  int cb;               // channels per halo-box row: 64, or 32 / 16 for thin inputs (one chunk)
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
  unsigned long long* dbg;   // optional per-role cycle counters (dn_tc_set_debug), same slots as igemm_tc_kernel
  CUtensorMap tmO[4], tmO2[4];   // TMA-store epilogue: output maps (one per phase; [0] alone otherwise), box {ob, 8, 16, 1}
  int tstore, ob_shift, nblk;    // tstore = 1: staged epilogue + bulk stores; the tile is nblk blocks of (1 << ob_shift) channels
  int nph, cpp;              // channel-stacked output phases (dn_igemm.phase_cout): column block i of a pixel goes to out + phase_off[i]
  long long phase_off[4];
  int step[3];               // gridDim.x as mixed-radix digits (tw, th, n): tiles advance by carries, not by divisions
  int dbg_flags;             // experiments (DN_TC_FLAGS): 1 = epilogue neither reads TMEM nor stores, 2 = epilogue waits with nanosleep back-off
};

// BN = MMA N tile, CB = channels per halo-box row (64 / 32 / 16).  Eight epilogue warps: two per TMEM lane quarter, each draining
// half of the accumulator columns (a single warp per scheduler is latency-bound: ~1 700 cycles per 128 x 16 tile measured).
// Thin tiles (BN <= 32) run two CTAs per SM for the same reason -- every role of this kernel is a single latency-bound warp.
template <int BN, int CB>
__global__ void __launch_bounds__(320, (BN <= 32 || (BN == 64 && CB <= 32)) ? 2 : 1) igemm_halo_kernel(const __grid_constant__ HaloParams p) {
  dn_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr int EW = 8;
  constexpr int CPT = BN / 2 < 8 ? 8 : BN / 2;      // accumulator columns per epilogue thread
  constexpr int CH = CPT < 32 ? CPT : 32;           // ... drained in chunks of CH
  constexpr uint32_t halo_bytes = kHaloW * kHaloH * 2 * CB;      // 36 864 for 64-channel rows
  constexpr uint32_t a_rb = 2 * CB, a_sbo = kHaloW * 2 * CB;
  constexpr uint32_t a_layout = CB == 64 ? 2 : CB == 32 ? 4 : 6;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  const int nb_tiles = p.ring ? p.bstages : 9 * p.kchunks;   // weight tiles in shared memory (ring slots or all of them)
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + (size_t)nb_tiles * B_BYTES;
  uint8_t* stg = smem_a + (size_t)stages * halo_bytes;                          // TMA-store staging tile(s), 1024-byte aligned
  const uint32_t stg_bytes = p.tstore ? (uint32_t)p.nblk * (256u << p.ob_shift) : 0u;      // nblk x [128 rows][2 * ob bytes]
  uint8_t* stg2 = (p.tstore && p.out2) ? stg + stg_bytes : nullptr;
  uint64_t* full_bar = (uint64_t*)(stg + (size_t)stg_bytes * (stg2 ? 2 : 1));
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bfull_bar = tempty_bar + 2;                      // resident: [0]; ring: full[0..7], empty[8..15]
  uint64_t* bempty_bar = bfull_bar + 8;
  uint32_t* tmem_ptr = (uint32_t*)(bfull_bar + 16);
  float* bias_s = (float*)(tmem_ptr + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // (tells the compiler the role branches are warp-uniform)
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EW); }      // one arrival per epilogue warp
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
      mbar_expect_tx_e(&full_bar[stage], halo_bytes);
      tma_load_4d_e(smem_a + (size_t)stage * halo_bytes, &p.tmA, &full_bar[stage], a_kc * CB, tw * 8 - 1, th * 16 - 1, n0);
      if (++stage == stages) { stage = 0; phase ^= 1; }
      if (++a_kc == p.kchunks) { a_kc = 0; a_tile += gridDim.x; }
    };
    load_a();
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < p.kchunks; ++kc) {
        load_a();
        for (int t = 0; t < 9; ++t) {
          mbar_wait(&bempty_bar[bs], bphase ^ 1);
          mbar_expect_tx_e(&bfull_bar[bs], B_BYTES);
          tma_load_3d_e(smem_b + (size_t)bs * B_BYTES, &p.tmB, &bfull_bar[bs], kc * kChunk, 0, p.wt[t]);
          if (++bs == p.bstages) { bs = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && p.ring) {
    // ---- MMA issuer, streamed weights
    int stage = 0; uint32_t phase = 0;
    int bs = 0; uint32_t bphase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint64_t ad_base = make_desc(smem_u32(smem_a), 16, a_sbo, a_layout);
    const uint64_t bd_base = make_desc(smem_u32(smem_b), 16, 1024);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t ad = ad_base + (uint64_t)((uint32_t)stage * (halo_bytes >> 4));
        const int ks = (kc == p.kchunks - 1) ? p.last_ksteps : 4;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          mbar_wait(&bfull_bar[bs], bphase);
          tc_fence_after();
          {
            const uint64_t ad0 = ad + (uint64_t)((((t / 3) * kHaloW + t % 3) * a_rb) >> 4);
            const uint64_t bd0 = bd_base + (uint64_t)((uint32_t)bs * (B_BYTES >> 4));
            umma_e(d_tmem, ad0, bd0, p.idesc, (kc == 0 && t == 0) ? 0u : 1u);
            if (ks > 1) umma_acc_e(d_tmem, ad0 + 2, bd0 + 2, p.idesc);
            if (ks > 2) umma_acc_e(d_tmem, ad0 + 4, bd0 + 4, p.idesc);
            if (ks > 3) umma_acc_e(d_tmem, ad0 + 6, bd0 + 6, p.idesc);
            umma_commit_e(&bempty_bar[bs]);
            if (t == 8) {
              umma_commit_e(&empty_bar[stage]);
              if (kc == p.kchunks - 1) umma_commit_e(&tfull_bar[acc]);
            }
          }
          if (++bs == p.bstages) { bs = 0; bphase ^= 1; }
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp == 0) {
    // ---- producer: weights once (tile (kc, t) at slot kc * 9 + t), then one halo box per (tile, chunk)
    if (elect_one()) {
      int present = 0;
      for (int t = 0; t < 9; ++t) present += p.wt[t] >= 0 ? 1 : 0;
      mbar_expect_tx(bfull_bar, (uint32_t)(present * p.kchunks) * B_BYTES);
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int t = 0; t < 9; ++t)
          if (p.wt[t] >= 0)
            tma_load_3d(smem_b + (size_t)(kc * 9 + t) * B_BYTES, &p.tmB, bfull_bar, kc * kChunk, 0, p.wt[t]);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    long long dbg_acc0 = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    int tw, th, n0;
    { int m = blockIdx.x; tw = m % p.tilesW; m /= p.tilesW; th = m % p.tilesH; n0 = m / p.tilesH; }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tw_c = tw, th_c = th, n0_c = n0;
      {     // next tile of this CTA
        tw += p.step[0];
        int c = tw >= p.tilesW ? 1 : 0; tw -= c * p.tilesW;
        th += p.step[1] + c;
        c = th >= p.tilesH ? 1 : 0; th -= c * p.tilesH;
        n0 += p.step[2] + c;
      }
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.dbg) dbg_acc0 += clock64() - t0;
        mbar_expect_tx_e(&full_bar[stage], halo_bytes);
        tma_load_4d_e(smem_a + (size_t)stage * halo_bytes, &p.tmA, &full_bar[stage], kc * CB, tw_c * 8 - 1, th_c * 16 - 1, n0_c);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    if (p.dbg && lane == 0) {
      atomicAdd(p.dbg + 0, (unsigned long long)dbg_acc0);
      atomicAdd(p.dbg + 1, (unsigned long long)(clock64() - tstart));
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    mbar_wait(bfull_bar, 0);
    tc_fence_after();
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t mask = 0;
    for (int t = 0; t < 9; ++t) mask |= (p.wt[t] >= 0 ? 1u : 0u) << t;
    const uint64_t ad_base = make_desc(smem_u32(smem_a), 16, a_sbo, a_layout);
    const uint64_t bd_base = make_desc(smem_u32(smem_b), 16, 1024);
    const uint32_t idesc = p.idesc;
    const int kchunks = p.kchunks, last_ks = p.last_ksteps;
    long long dbg_full = 0, dbg_te = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (p.dbg) dbg_te += clock64() - t0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kc = 0; kc < kchunks; ++kc) {
        t0 = p.dbg ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        if (p.dbg) dbg_full += clock64() - t0;
        tc_fence_after();
        const uint64_t ad = ad_base + (uint64_t)((uint32_t)stage * (halo_bytes >> 4));
        const uint64_t bd = bd_base + (uint64_t)((uint32_t)kc * (9 * B_BYTES >> 4));
        const uint32_t last = kc == kchunks - 1 ? 1u : 0u;
        {
          const uint32_t acc0 = kc == 0 ? 0u : 1u;
          const int ks = last ? last_ks : CB / 16;
          const uint32_t eb = smem_u32(&empty_bar[stage]), tb = smem_u32(&tfull_bar[acc]);
          if (CB == 16 || ks == 1) halo_issue<1, (int)a_rb, (int)B_BYTES>(d_tmem, ad, bd, idesc, mask, acc0, eb, tb, last);
          else if (CB == 32 || ks == 2) halo_issue<2, (int)a_rb, (int)B_BYTES>(d_tmem, ad, bd, idesc, mask, acc0, eb, tb, last);
          else if (ks == 3) halo_issue<3, (int)a_rb, (int)B_BYTES>(d_tmem, ad, bd, idesc, mask, acc0, eb, tb, last);
          else halo_issue<4, (int)a_rb, (int)B_BYTES>(d_tmem, ad, bd, idesc, mask, acc0, eb, tb, last);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      atomicAdd(p.dbg + 2, (unsigned long long)dbg_full); atomicAdd(p.dbg + 3, (unsigned long long)dbg_te);
      atomicAdd(p.dbg + 4, (unsigned long long)(clock64() - tstart));
    }
  } else {
    // ---- epilogue: tile row r = 8 * h_local + w_local; warps 2..5 drain columns [0, CPT), warps 6..9 columns [CPT, 2 CPT)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int cbeg = (warp >= 6 ? 1 : 0) * CPT;
    int acc = 0; uint32_t acc_phase = 0;
    const int wi = row & 7, hi = row >> 3;
    const int esz = p.out.dtype == DN_F32 ? 4 : 2;
    const int out_dtype = p.out.dtype, out_C = p.out.C, c_eff = p.c_eff, act = p.act, accumulate = p.accumulate, n_mma = p.n_mma;
    const float out_scale = p.out_scale;
    // bias: registers when a thread owns at most 32 columns, shared memory otherwise (one N tile: the same for every tile)
    float breg[CH];
    const int cpp = p.nph > 1 ? p.cpp : BN;        // (channel-stacked phases: bias / channel limits repeat every cpp columns)
    const int cpp_shift = 31 - __clz(cpp);
    if (CPT <= 32) {
#pragma unroll
      for (int j = 0; j < CH; ++j) breg[j] = (p.bias && (cbeg + j) % cpp < p.out.C) ? p.bias[(cbeg + j) % cpp] : 0.f;
    } else {
      for (int c = et; c < BN; c += EW * 32) bias_s[c] = (p.bias && c % cpp < p.out.C) ? p.bias[c % cpp] : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    }
    int tw, th, n;
    { int m = blockIdx.x; tw = m % p.tilesW; m /= p.tilesW; th = m % p.tilesH; n = m / p.tilesH; }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int w = tw * 8 + wi, h = th * 16 + hi;
      const int n_c = n;
      {     // next tile of this CTA
        tw += p.step[0];
        int c = tw >= p.tilesW ? 1 : 0; tw -= c * p.tilesW;
        th += p.step[1] + c;
        c = th >= p.tilesH ? 1 : 0; th -= c * p.tilesH;
        n += p.step[2] + c;
      }
      const bool valid = (w < p.out.W) && (h < p.out.H);
      const size_t eoff = (size_t)dn_off(p.out, n_c, h, w);
      uint8_t* optr = (uint8_t*)p.out.ptr + eoff * esz;
      uint4 old[CH / 8];
      if (CPT <= 32 && accumulate && esz == 2) epi_prefetch<CH>(old, valid && cbeg < n_mma, cbeg, c_eff, optr + (size_t)cbeg * esz);
      const long long te0 = p.dbg ? clock64() : 0;
      if (p.dbg_flags & 2) mbar_wait_sleep(&tfull_bar[acc], acc_phase); else mbar_wait(&tfull_bar[acc], acc_phase);
      const long long te1 = p.dbg ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if (p.dbg_flags & 1) {
      } else if (p.tstore) {
        // staged epilogue: the previous tile's bulk stores must have read the staging tile before it is overwritten
        if (et == 0) tma_store_wait_read();
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        uint32_t ra[CH];
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + CPT && c0 < n_mma; c0 += CH) {
          tmem_ldn<CH>(taddr + c0, ra);
          tmem_ld_wait();
          epi_stage<CH>(ra, CPT <= 32 ? breg : bias_s + c0, out_scale, act, c0, p.ob_shift, row, stg, out_dtype, stg2, p.out2_dtype);
        }
        fence_proxy_async();
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (et == 0) {
          const int w0 = (w - wi), h0 = (h - hi);
          const uint32_t blk = 256u << p.ob_shift;
          for (int j = 0; j < p.nblk; ++j) {
            const CUtensorMap* mo = p.nph > 1 ? &p.tmO[j] : &p.tmO[0];
            const int cj = p.nph > 1 ? 0 : (j << p.ob_shift);
            tma_store_4d(mo, stg + (size_t)j * blk, cj, w0, h0, n_c);
            if (stg2) tma_store_4d(p.nph > 1 ? &p.tmO2[j] : &p.tmO2[0], stg2 + (size_t)j * blk, cj, w0, h0, n_c);
          }
          tma_store_commit();
        }
      } else if (p.nph > 1) {
        // channel-stacked phases: chunks of CH columns, each 8-channel group scattered to its phase's pixel
        uint32_t ra[CH];
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + CPT && c0 < n_mma; c0 += CH) {
          tmem_ldn<CH>(taddr + c0, ra);
          tmem_ld_wait();
          epi_store_phases<CH>(ra, CPT <= 32 ? breg : bias_s + c0, out_scale, act, valid, c0, cpp_shift, c_eff, out_dtype, optr, p.phase_off, p.out2,
                               p.out2_dtype, eoff);
        }
      } else if (CPT <= 32) {
        uint32_t ra[CH];
        if (cbeg < n_mma) {
          tmem_ldn<CH>(taddr + cbeg, ra);
          tmem_ld_wait();
          if (p.dbg && et == 0) atomicAdd(p.dbg + 7, (unsigned long long)(clock64() - te1));
          epi_store<CH>(ra, breg, out_scale, act, valid, cbeg, c_eff, out_C, out_dtype, optr + (size_t)cbeg * esz, accumulate, p.out2, p.out2_dtype, eoff + cbeg, old);
        }
      } else {
        // chunks of 32 columns; the TMEM load of chunk i + 1 is in flight while chunk i is converted and stored
        uint32_t ra[32], rb[32];
        uint4 old4[4];
        if (cbeg < n_mma) tmem_ld32(taddr + cbeg, ra);
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + CPT && c0 < n_mma; c0 += 64) {
          if (accumulate && esz == 2) epi_prefetch<32>(old4, valid, c0, c_eff, optr + (size_t)c0 * esz);
          tmem_ld_wait();
          if (c0 + 32 < cbeg + CPT && c0 + 32 < n_mma) tmem_ld32(taddr + c0 + 32, rb);
          epi_store<32>(ra, bias_s + c0, out_scale, act, valid, c0, c_eff, out_C, out_dtype, optr + (size_t)c0 * esz, accumulate, p.out2, p.out2_dtype, eoff + c0, old4);
          if (!(c0 + 32 < cbeg + CPT && c0 + 32 < n_mma)) break;
          if (accumulate && esz == 2) epi_prefetch<32>(old4, valid, c0 + 32, c_eff, optr + (size_t)(c0 + 32) * esz);
          tmem_ld_wait();
          if (c0 + 64 < cbeg + CPT && c0 + 64 < n_mma) tmem_ld32(taddr + c0 + 64, ra);
          epi_store<32>(rb, bias_s + c0 + 32, out_scale, act, valid, c0 + 32, c_eff, out_C, out_dtype, optr + (size_t)(c0 + 32) * esz, accumulate, p.out2, p.out2_dtype, eoff + c0 + 32, old4);
        }
      }
      const long long te2 = p.dbg ? clock64() : 0;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (p.dbg && et == 0) {
        atomicAdd(p.dbg + 5, (unsigned long long)(te1 - te0));
        atomicAdd(p.dbg + 6, (unsigned long long)(clock64() - te1));
        atomicAdd(p.dbg + 15, (unsigned long long)(clock64() - te2));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tstore && et == 0) tma_store_wait_read();      // the staging tile must outlive the last bulk store's reads
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
  int p5, q5;              // dy / x maps are 5-D (64 ch, W, H, N, channel block): all 64-channel blocks of a stage in ONE TMA
  int cbq;                 // channels per x row in shared memory: 64, or 32 / 16 for thin x (one 64-wide cq tile, n_mma <= cbq)
  int mstack, cbp;         // thin dy (<= 32 channels), full 3x3: the M dimension holds THREE column-shifted copies of the dy tile
                           // (cbp = 16 / 32 channels per copy, slot s = dy shifted by s - 1 pixels), x is fetched as an (8+2)-row
                           // strip and the three MMAs per K step (one per kernel row) produce all nine taps: 12 MMAs per 64 pixels
                           // instead of 36 (thin weight gradients are bound by the MMA count, ~40 cycles each)
  uint32_t idesc;
  float* dw;
  int cp, cq, cp_pad, cq_pad;
  float scale;
  unsigned long long* dbg;   // optional per-role cycle counters (dn_tc_set_debug), slots 8..14 (same roles as slots 0..6)
};

constexpr int KPX = 64;   // pixels (= GEMM K) per pipeline stage
constexpr int kWgMaxTpc = 32;   // taps per work item: 512 TMEM columns / 16

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
  const uint32_t q_rb = (uint32_t)(2 * p.cbq), q_layout = (uint32_t)thin_layout(p.cbq);   // bytes per pixel row of x
  const uint32_t Q_BLK = KPX * q_rb;                     // one [KPX px][cbq ch] block of x
  const uint32_t B_BYTES = (BNQ / 64) * Q_BLK;           // per tap (plain mode)
  const uint32_t HALO_BLK = (p.mstack ? 8 : 16) * 10 * q_rb;   // one [10 rows][16 px][cbq ch] halo block (halo mode; 8 px wide when mstack)
  const uint32_t STAGE_BYTES = A_BYTES + ((p.halo || p.mstack) ? (uint32_t)(BNQ / 64) * HALO_BLK : (uint32_t)p.tpc * B_BYTES);
  const uint32_t p_rb = (uint32_t)(2 * p.cbp), P_SLOT = KPX * p_rb;       // mstack: bytes per dy pixel row, per shifted copy
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // (tells the compiler the role branches are warp-uniform)
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
    long long dbg_acc0 = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem w = wg_decode(p, item);
      // bytes that really arrive per stage: only the P blocks that exist and the taps of this group are loaded
      const uint32_t tx_bytes = p.mstack ? 3 * P_SLOT + (uint32_t)(BNQ / 64) * HALO_BLK
                                         : (uint32_t)p.cp_blocks * BLK_BYTES + (p.halo ? (uint32_t)(BNQ / 64) * HALO_BLK : (uint32_t)w.nt * B_BYTES);
      // pixel-tile coordinates: one division per item, then incremental (the divisions cost more than the TMA issue itself)
      int tw, th, tn;
      {
        int m = w.pt_beg;
        tw = m % p.tilesW; m /= p.tilesW;
        th = m % p.tilesH;
        tn = m / p.tilesH;
      }
      for (int it = 0; it < w.n_iters; ++it) {
        const int w0 = tw * p.wb, h0 = th * p.hb, n0 = tn * p.nb;
        if (++tw == p.tilesW) { tw = 0; if (++th == p.tilesH) { th = 0; ++tn; } }
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.dbg) dbg_acc0 += clock64() - t0;
        if (elect_one()) {
          uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          if (p.mstack) {
            for (int sft = 0; sft < 3; ++sft)
              tma_load_4d(sa + sft * P_SLOT, &p.tmP[0], &full_bar[stage], 0, w0 + sft - 1, h0, n0);
#pragma unroll
            for (int j = 0; j < BNQ / 64; ++j)
              tma_load_4d(sa + A_BYTES + j * HALO_BLK, &p.tmQ, &full_bar[stage], w.cqt * BNQ + j * 64, w0, h0 - 1, n0);
          } else {
          if (p.p5) tma_load_5d(sa, &p.tmP[w.src], &full_bar[stage], 0, w0, h0, n0, w.cpt * 2);
          else
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
              if (p.q5) tma_load_5d(sb, &p.tmQ, &full_bar[stage], 0, w0 + tap.dw, h0 + tap.dh, n0, w.cqt * (BNQ / 64));
              else {
#pragma unroll
                for (int j = 0; j < BNQ / 64; ++j)
                  tma_load_4d(sb + j * Q_BLK, &p.tmQ, &full_bar[stage], w.cqt * BNQ + j * 64, w0 + tap.dw, h0 + tap.dh, n0);
              }
            }
          }
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    if (p.dbg && lane == 0) {
      atomicAdd(p.dbg + 8, (unsigned long long)dbg_acc0);
      atomicAdd(p.dbg + 9, (unsigned long long)(clock64() - tstart));
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    long long dbg_full = 0, dbg_te = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    static_assert(KPX == 64, "four K steps per stage are unrolled below");
    const uint64_t ad_base = p.mstack ? make_desc(smem_u32(smem), P_SLOT, 8 * p_rb, (uint32_t)thin_layout(p.cbp)) : make_desc(smem_u32(smem), BLK_BYTES, 1024);
    const uint64_t bd_base = make_desc(smem_u32(smem), (p.halo || p.mstack) ? HALO_BLK : Q_BLK, p.halo ? 16 * q_rb : 8 * q_rb, q_layout);
    const uint32_t ka = p.mstack ? p_rb : 128u;          // 16 pixels further down the dy tile, in 16-byte units
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem w = wg_decode(p, item);
      long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (p.dbg) dbg_te += clock64() - t0;
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)acc * acc_cols;
      // per-tap operand offsets of this item, computed once (the inner loop then spends two adds per MMA: the issuing thread's
      // instruction stream is what paces thin problems, tools/ubench_tc.cu)
      uint32_t boff[kWgMaxTpc];
#pragma unroll
      for (int t = 0; t < kWgMaxTpc; ++t) {
        if (t >= w.nt) break;
        {
          if (p.mstack) {
            boff[t] = (A_BYTES + (uint32_t)t * 8 * q_rb) >> 4;       // kernel row t: the x strip starts t image rows (8 pixels each) further
          } else if (p.halo) {
            // pixel tile = 8 rows x 8 columns; tap (dh, dw) reads halo pixel (row + dh + 1, col + dw + 1): a view that starts
            // ((dh+1)*16 + (dw+1)) pixels into the 16-pixel-pitch halo block.  One UMMA_K step = 16 pixels = 2 image rows.
            const TcTap tap = p.taps[w.t0 + t];
            boff[t] = (A_BYTES + (uint32_t)((tap.dh + 1) * 16 + (tap.dw + 1)) * q_rb) >> 4;
          } else {
            boff[t] = (A_BYTES + (uint32_t)t * B_BYTES) >> 4;
          }
        }
      }
      const uint32_t kadv = p.halo ? 2 * q_rb : q_rb;      // 16 pixels further down the x block, in 16-byte units
      const uint32_t n_mma = (uint32_t)p.n_mma, idesc = p.idesc;
      for (int it = 0; it < w.n_iters; ++it) {
        t0 = p.dbg ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        if (p.dbg) dbg_full += clock64() - t0;
        tc_fence_after();
        // MN-major, 128B swizzle: LBO = distance between 64-channel blocks, SBO = 8 pixel rows = 1024 B;
        // 16 pixels (one UMMA_K step) further down = +2048 bytes = +128 in the (addr >> 4) field
        const uint64_t ad0 = ad_base + (uint64_t)((uint32_t)stage * (STAGE_BYTES >> 4));
        const uint64_t bd0 = bd_base + (uint64_t)((uint32_t)stage * (STAGE_BYTES >> 4));
        {
          const uint32_t first = it > 0 ? 1u : 0u;
          uint32_t d = d_base;
#pragma unroll
          for (int t = 0; t < kWgMaxTpc; ++t) {
            if (t >= w.nt) break;
            wg_issue4(d, ad0, bd0 + (uint64_t)boff[t], idesc, first, kadv, ka);
            d += n_mma;
          }
          umma_commit_e(&empty_bar[stage]);
          if (it == w.n_iters - 1) umma_commit_e(&tfull_bar[acc]);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (p.nacc == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
      else acc_phase ^= 1;
    }
    if (p.dbg && lane == 0) {
      atomicAdd(p.dbg + 10, (unsigned long long)dbg_full); atomicAdd(p.dbg + 11, (unsigned long long)dbg_te);
      atomicAdd(p.dbg + 12, (unsigned long long)(clock64() - tstart));
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
      if (p.mstack) { cp = lane; row_ok = q < 3 && lane < p.cbp; }      // TMEM quarter q = column shift slot, lane = dy channel
      else if (p.m64 == 1) { cp = w.cpt * 128 + q * 16 + lane; row_ok = lane < 16; }
      else if (p.m64 == 2) { row_ok = row < 64; }
      const long long te0 = p.dbg ? clock64() : 0;
      mbar_wait(&tfull_bar[acc], acc_phase);
      const long long te1 = p.dbg ? clock64() : 0;
      tc_fence_after();
      for (int t = 0; t < w.nt; ++t) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * p.n_mma);
        // mstack: accumulator t = kernel row t, slot q = dy shifted by q - 1 columns = kernel column 2 - q; taps[] is ordered (kh, kw)
        const int tap_i = p.mstack ? t * 3 + (q < 3 ? 2 - q : 0) : w.t0 + t;
        float* drow = p.dw + ((size_t)p.taps[tap_i].wt * p.cp_pad + cp) * p.cq_pad + w.cqt * BNQ;
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
      if (p.dbg && threadIdx.x == 64) {
        atomicAdd(p.dbg + 13, (unsigned long long)(te1 - te0));
        atomicAdd(p.dbg + 14, (unsigned long long)(clock64() - te1));
      }
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

CUtensorMapSwizzle swizzle_of(int cb) {
  return cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// 4-D map (C, W, H, N) of an NHWC view, box {cb, wb, hb, nb} (cb = 64 / 32 / 16 channels -> 128B / 64B / 32B swizzle), zero OOB fill
int make_view_map(CUtensorMap* tm, const dn_view& v, int wb, int hb, int nb, int cb = kChunk, bool exact = false) {
  auto enc = get_encode();
  if (!enc) return DN_E_UNSUPPORTED;
  // channel extent: whole box rows where the view says the bytes behind its channels are readable padding (dn_view.c_ext) -- a box
  // row that is partially out of bounds takes the TMA unit ~7 cycles instead of ~2.6 (tools/ubench_tc.cu)
  int ext = (v.C + cb - 1) / cb * cb;
  const int readable = v.c_ext > v.C ? v.c_ext : v.C;
  if (ext > readable) ext = readable;
  if (exact) ext = v.C;      // store maps: nothing beyond the view's channels may be written
  cuuint64_t dims[4] = {(cuuint64_t)ext, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)wb, (cuuint32_t)hb, (cuuint32_t)nb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapDataType dt = v.dtype == DN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(tm, dt, 4, v.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(cb),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : DN_E_ARG;
}

// 5-D map (64 ch, W, H, N, channel block) of a view whose channel count is a multiple of 64: box {64, wb, hb, nb, nblk} lands as
// nblk consecutive [pixels][64 ch] blocks -- the layout the weight-gradient kernel wants -- with one TMA instruction
int make_view_map5(CUtensorMap* tm, const dn_view& v, int wb, int hb, int nb, int nblk) {
  auto enc = get_encode();
  if (!enc) return DN_E_UNSUPPORTED;
  cuuint64_t dims[5] = {64, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N, (cuuint64_t)(v.C / 64)};
  cuuint64_t strides[4] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2, 128};
  cuuint32_t box[5] = {64, (cuuint32_t)wb, (cuuint32_t)hb, (cuuint32_t)nb, (cuuint32_t)nblk};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUtensorMapDataType dt = v.dtype == DN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(tm, dt, 5, v.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : DN_E_ARG;
}

const bool g_thin_rows = []() { const char* e = getenv("DN_THIN_ROWS"); return !(e && e[0] == '0'); }();   // A/B knob
int pick_cb(int C) { return g_thin_rows ? thin_cb(C) : kChunk; }

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
  const uint32_t STAGE_BYTES = (uint32_t)P.tps * (kRows * 2 * P.cb + BN * 128);
  int stages = P.stages;
  size_t smem = (size_t)stages * STAGE_BYTES + 1024 + (2 * stages + 4) * 8 + 16 + 2 * BN * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int grid = P.num_tiles < dn_num_sms() ? P.num_tiles : dn_num_sms();
  dn_launch(igemm_tc_kernel<BN>, dim3(grid), dim3(320), smem, st, P);
  DN_CHECK_LAUNCH();
  return 0;
}

template <int BNQ>
int launch_wgrad(const WgradTcParams& P, int items, cudaStream_t st) {
  const uint32_t q_rb = 2 * P.cbq;
  const uint32_t stage_bytes = 2 * KPX * 128 + (P.mstack ? (uint32_t)(BNQ / 64) * 8 * 10 * q_rb
                                                          : P.halo ? (uint32_t)(BNQ / 64) * 16 * 10 * q_rb : (uint32_t)P.tpc * (BNQ / 64) * KPX * q_rb);
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


unsigned long long* g_tc_dbg = nullptr;

// shared memory the halo kernel may spend on resident weights + staging + halo stages: 217 KB leaves room for two 64-channel halo
// stages next to 144 KB of weights (N = 64 over 128 input channels, N = 128 over 64), i.e. those layers keep their weights resident
// instead of falling back to the plain kernel / the streamed-weight ring; the rest of the 227 KB holds barriers and the bias tile
const size_t kHaloBudget = []() { const char* e = getenv("DN_HALO_BUDGET_KB"); return (size_t)(e ? atoi(e) : 217) * 1024; }();

// ---- halo variant: eligibility + launch -------------------------------------------------------------------------
bool g_halo_enabled = true;
const bool g_halo_ring = []() { const char* e = getenv("DN_HALO_RING"); return !(e && e[0] == '0'); }();   // A/B knob

// fills wt[kh*3+kw]; true when the taps are exactly the 3x3 neighbourhood of one source
// taps form a subset of the 3x3 neighbourhood (all nine for a 3x3 convolution, a 2x2 corner for one output phase of a
// 4x4 / stride-2 transposed convolution): wt[kh * 3 + kw] = packed-weight index, -1 where the tap is absent
bool halo_taps(const dn_igemm* p, int* wt) {
  if (p->ntaps > 9 || p->ntaps < 3 || p->nsrc != 1 || p->stride != 1 || (p->nphase > 1 && p->phase_cout <= 0)) return false;
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
  const int cb = pick_cb(p->in[0].C);
  const int kchunks = (p->in[0].C + cb - 1) / cb;
  const size_t b_bytes = (size_t)9 * kchunks * BN * 128;
  const size_t kHaloBytes = (size_t)kHaloW * kHaloH * 2 * cb;
  // weights stay resident next to >= 2 halo stages when they fit; else they stream through a ring, which measured a gain only
  // for one-chunk problems (features.7 forward 0.099 -> 0.093 ms): N <= 128 tiles with more K are bound by the 128 B/clk
  // shared-memory operand reads of cta_group::1 either way (A 4 KB + B 4 KB per 64-cycle 128x128x16 MMA)
  if (b_bytes + 2 * kHaloBytes > kHaloBudget && (BN > 128 || kchunks > 1 || !g_halo_ring || p->ntaps != 9)) return false;
  const int H = p->out.H, W = p->out.W;
  if (H != p->in[0].H || W != p->in[0].W) return false;
  // tile = 16 rows x 8 columns: only worth it when little of the tile grid is padding
  const double eff = (double)H * W / ((double)((H + 15) / 16 * 16) * ((W + 7) / 8 * 8));
  return eff >= 0.85;
}

template <int BN, int CB>
int launch_halo_cb(const HaloParams& P, cudaStream_t st) {
  const size_t b_bytes = (size_t)(P.ring ? P.bstages : 9 * P.kchunks) * BN * 128;
  const size_t kHaloBytes = (size_t)kHaloW * kHaloH * 2 * P.cb;
  const size_t stg = P.tstore ? (size_t)P.nblk * (256u << P.ob_shift) * (P.out2 ? 2 : 1) : 0;
  size_t smem = b_bytes + stg + (size_t)P.stages * kHaloBytes + 1024 + (2 * P.stages + 20) * 8 + 16 + 2 * BN * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_halo_kernel<BN, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  static const bool one_cta = []() { const char* e = getenv("DN_HALO_ONE_CTA"); return e && e[0] == '1'; }();      // A/B knob
  const int ctas = (((BN <= 32 || (BN == 64 && CB <= 32)) && smem <= 110 * 1024 && !one_cta) ? 2 : 1) * dn_num_sms();
  int grid = P.num_tiles < ctas ? P.num_tiles : ctas;
  HaloParams Q = P;
  Q.step[0] = grid % P.tilesW;
  Q.step[1] = (grid / P.tilesW) % P.tilesH;
  Q.step[2] = grid / P.tilesW / P.tilesH;
  dn_launch(igemm_halo_kernel<BN, CB>, dim3(grid), dim3(320), smem, st, Q);
  DN_CHECK_LAUNCH();
  return 0;
}

template <int BN>
int launch_halo(const HaloParams& P, cudaStream_t st) {
  if (P.cb == 16) return launch_halo_cb<BN, 16>(P, st);
  if (P.cb == 32) return launch_halo_cb<BN, 32>(P, st);
  return launch_halo_cb<BN, 64>(P, st);
}

int dn_igemm_halo(const dn_igemm* p, const int* wt, cudaStream_t st) {
  HaloParams P;
  memset(&P, 0, sizeof(P));
  const int BN = pick_bn(p->cout_pad);
  auto enc = get_encode();
  if (!enc) return DN_E_UNSUPPORTED;
  P.cb = pick_cb(p->in[0].C);
  const size_t kHaloBytes = (size_t)kHaloW * kHaloH * 2 * P.cb;
  {
    int e = make_view_map(&P.tmA, p->in[0], kHaloW, kHaloH, 1, P.cb);
    if (e) return e;
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
  P.kchunks = (Cin + P.cb - 1) / P.cb;
  P.last_ksteps = (Cin - (P.kchunks - 1) * P.cb + 15) / 16;
  P.tilesW = (p->out.W + 7) / 8;
  P.tilesH = (p->out.H + 15) / 16;
  P.num_tiles = P.tilesW * P.tilesH * p->out.N;
  size_t b_bytes = (size_t)9 * P.kchunks * BN * 128;
  // TMA-store epilogue (16-bit output, no accumulate, N tile <= 128): staging tile(s) in shared memory
  // Measured (B200): it pays for the channel-stacked phases of thin transposed convolutions, whose direct stores scatter 16-byte
  // pieces over four output pixels (upconv0 forward 62 -> 49 us); ordinary tiles (a thread stores its own pixel's channels) are
  // as fast or faster with direct stores (features.3 148 vs 158 us), so those keep them.  DN_TMA_STORE=0 / 2: never / always.
  static const int g_tstore = []() { const char* e = getenv("DN_TMA_STORE"); return e ? atoi(e) : 1; }();
  const bool stacked = p->nphase > 1 && p->phase_cout > 0;
  size_t stg_total = 0;
  if ((g_tstore == 2 || (g_tstore == 1 && stacked)) && p->out.dtype != DN_F32 && !p->accumulate && BN <= 128 && view_tma_ok(p->out) && (!p->out2 || ((uintptr_t)p->out2 % 16) == 0)) {
    const int ob = stacked ? p->phase_cout : (BN >= 64 ? 64 : BN);
    if (ob == 16 || ob == 32 || ob == 64) {
      P.tstore = 1;
      P.ob_shift = ob == 64 ? 6 : ob == 32 ? 5 : 4;
      P.nblk = ((stacked ? p->cout_pad : (p->cout_pad < BN ? p->cout_pad : BN)) + ob - 1) / ob;
      stg_total = (size_t)P.nblk * 256 * ob * (p->out2 ? 2 : 1);
      for (int i = 0; i < (stacked ? p->nphase : 1); ++i) {
        dn_view v = p->out;
        v.ptr = (char*)p->out.ptr + (stacked ? p->phase_off[i] : 0) * 2;
        if (make_view_map(&P.tmO[i], v, 8, 16, 1, ob, true)) return DN_E_ARG;
        if (p->out2) {
          v.ptr = (char*)p->out2 + (stacked ? p->phase_off[i] : 0) * 2;
          v.dtype = p->out2_dtype;
          if (make_view_map(&P.tmO2[i], v, 8, 16, 1, ob, true)) return DN_E_ARG;
        }
      }
    }
  }
  b_bytes += stg_total;        // (the staging area competes with the pipeline stages for shared memory)
  if (b_bytes + 2 * kHaloBytes > kHaloBudget) {
    if (stg_total) { P.tstore = 0; b_bytes -= stg_total; stg_total = 0; }
  }
  if (b_bytes + 2 * kHaloBytes > kHaloBudget) {
    P.ring = 1;
    P.stages = 2;
    P.bstages = (int)((200 * 1024 - 2 * kHaloBytes) / ((size_t)BN * 128));
    if (P.bstages > 8) P.bstages = 8;
  } else {
    // thin output tiles (BN <= 32): two CTAs per SM when the weights leave room for >= 3 stages in half of the shared memory
    size_t budget = kHaloBudget;
    if ((BN <= 32 || (BN == 64 && P.cb <= 32)) && b_bytes + 3 * kHaloBytes <= 108 * 1024) budget = 108 * 1024;
    P.stages = (int)((budget - b_bytes) / kHaloBytes);
    if (P.stages > (P.cb == kChunk ? 4 : 6)) P.stages = P.cb == kChunk ? 4 : 6;     // thin boxes are latency-bound: more in flight
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
  P.nph = (p->nphase > 1 && p->phase_cout > 0) ? p->nphase : 1;
  P.cpp = p->phase_cout;
  for (int i = 0; i < 4; ++i) P.phase_off[i] = (P.nph > 1 && i < p->nphase) ? p->phase_off[i] : 0;
  if (P.nph > 1) {       // channel limits are per phase
    if (p->out.dtype == DN_F32 || p->accumulate || (p->phase_cout % 8) != 0 || (p->phase_cout & (p->phase_cout - 1)) != 0) return DN_E_UNSUPPORTED;
    P.n_mma = p->cout_pad;
  }
  P.dbg = g_tc_dbg;
  { const char* e = getenv("DN_TC_FLAGS"); P.dbg_flags = e ? atoi(e) : 0; }
  switch (BN) {
    case 256: return launch_halo<256>(P, st);
    case 128: return launch_halo<128>(P, st);
    case 64: return launch_halo<64>(P, st);
    case 32: return launch_halo<32>(P, st);
    default: return launch_halo<16>(P, st);
  }
}

}  // namespace

// debug: device buffer of 16 counters (slots 8..14: the same roles of the weight-gradient kernel) (cycles summed over CTAs): [0] producer wait-empty, [1] producer total,
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
  if (p->nphase < 0 || p->nphase > 4 || (p->nphase > 1 && p->phase_cout <= 0 && (p->ntaps % p->nphase) != 0)) return 0;
  if (p->phase_cout > 0) {
    int wt[9];
    if (p->nphase < 2 || p->out.dtype == DN_F32 || p->accumulate || (p->phase_cout % 8) != 0 || (p->phase_cout & (p->phase_cout - 1)) != 0 ||
        !halo_eligible(p, wt))
      return 0;
  }
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
    if (p->phase_cout > 0) return DN_E_UNSUPPORTED;       // channel-stacked phases exist in the halo kernel only
  }
  IgemmTcParams P;
  memset(&P, 0, sizeof(P));
  int BN = pick_bn(p->cout_pad);
  choose_box(p->out.N, p->out.H, p->out.W, kRows, P.wb, P.hb, P.nb);
  {
    // few pixel tiles (the 4x13 ... 8x26 layers): a narrower N tile can put more SMs to work.  cost = waves x tensor-pipe cycles of
    // one 128 x BN x 16 MMA (tools/ubench_tc.cu: 55 / 70 / 128 cycles for N = 64 / 128 / 256)
    const long long m_tiles = (long long)((p->out.W + P.wb - 1) / P.wb) * ((p->out.H + P.hb - 1) / P.hb) * ((p->out.N + P.nb - 1) / P.nb) *
                              (p->nphase > 1 ? p->nphase : 1);
    const int sms = dn_num_sms();
    auto cost = [&](int bn) {
      const long long tiles = m_tiles * ((p->cout_pad + bn - 1) / bn);
      return (double)((tiles + sms - 1) / sms) * (bn >= 256 ? 128.0 : bn >= 128 ? 70.0 : 55.0) * ((p->cout_pad + bn - 1) / bn) /
             (double)((p->cout_pad + bn - 1) / bn);
    };
    while (BN > 64 && (p->cout_pad % (BN / 2)) == 0 && cost(BN / 2) < cost(BN)) BN /= 2;
  }
  P.cb = pick_cb(p->in[0].C);
  for (int s = 0; s < p->nsrc; ++s) {
    int e = make_view_map(&P.tmA[s], p->in[s], P.wb, P.hb, P.nb, P.cb);
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
  P.kchunks = (Cin + P.cb - 1) / P.cb;
  int rem = Cin - (P.kchunks - 1) * P.cb;
  P.last_ksteps = (rem + 15) / 16;
  P.tilesW = (p->out.W + P.wb - 1) / P.wb;
  P.tilesH = (p->out.H + P.hb - 1) / P.hb;
  P.tilesN = (p->out.N + P.nb - 1) / P.nb;
  P.ntile_n = (p->cout_pad + BN - 1) / BN;
  P.nph = p->nphase > 1 ? p->nphase : 1;
  P.tph = p->ntaps / P.nph;
  for (int i = 0; i < 4; ++i) P.phase_off[i] = (p->nphase > 1 && i < p->nphase) ? p->phase_off[i] : 0;
  P.num_tiles = P.tilesW * P.tilesH * P.tilesN * P.ntile_n * P.nph;
  {
    int g = P.num_tiles < dn_num_sms() ? P.num_tiles : dn_num_sms();      // = the grid size of launch_igemm
    const int radix[4] = {P.ntile_n, P.tilesW, P.tilesH, P.tilesN};
    for (int i = 0; i < 4; ++i) { P.step[i] = g % radix[i]; g /= radix[i]; }
    P.step[4] = g;
  }
  // thin one-chunk problems with many taps: up to four taps per stage
  static const bool g_tps = []() { const char* e = getenv("DN_TAPS_PER_STAGE"); return !(e && e[0] == '0'); }();
  P.tps = 1;
  if (g_tps && P.kchunks == 1 && P.cb <= 32 && P.tph >= 4) P.tps = 4;
  while (P.tps > 1 && (200 * 1024) / ((uint32_t)P.tps * (kRows * 2 * P.cb + BN * 128)) < 3) --P.tps;      // keep >= 3 stages
  const uint32_t stage_bytes = (uint32_t)P.tps * (kRows * 2 * P.cb + BN * 128);
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
  // M-stacked mode (see WgradTcParams::mstack): thin dy, full 3x3, one source
  static const bool g_mstack = []() { const char* e = getenv("DN_WGRAD_MSTACK"); return !(e && e[0] == '0'); }();
  bool mstack = false;
  int wt9[9];
  if (g_mstack && g_thin_rows && p->ntaps == 9 && p->nsrc == 1 && P0.H == p->q.H && P0.W == p->q.W && P0.C <= 32) {
    bool seen[9] = {false};
    mstack = true;
    for (int t = 0; t < 9 && mstack; ++t) {
      int dh = p->taps[t].dh, dw = p->taps[t].dw;
      if (dh < -1 || dh > 1 || dw < -1 || dw > 1 || seen[(dh + 1) * 3 + dw + 1]) mstack = false;
      else { seen[(dh + 1) * 3 + dw + 1] = true; wt9[(dh + 1) * 3 + dw + 1] = p->taps[t].wt; }
    }
    const int bnq = p->cq_pad >= 256 ? 256 : (p->cq_pad >= 128 ? 128 : 64);
    const int cq_tiles = (p->cq_pad + bnq - 1) / bnq;
    const int n16 = (p->q.C + 15) / 16 * 16;
    const int n_mma = (cq_tiles == 1 && n16 < bnq) ? n16 : bnq;
    if (3 * n_mma > 512) mstack = false;
  }
  if (mstack) halo = false;
  P.mstack = mstack ? 1 : 0;
  P.cbp = mstack ? thin_cb(P0.C) : kChunk;
  P.halo = halo ? 1 : 0;
  if (halo || mstack) { P.wb = 8; P.hb = 8; P.nb = 1; }
  else choose_box(P0.N, P0.H, P0.W, KPX, P.wb, P.hb, P.nb);
  const bool wg5 = []() { const char* e = getenv("DN_WGRAD_5D"); return !(e && e[0] == '0'); }();
  P.p5 = (wg5 && P0.C > 64 && P0.C % 128 == 0) ? 1 : 0;
  for (int s = 0; s < p->nsrc; ++s) {
    int e = P.p5 ? make_view_map5(&P.tmP[s], p->p[s], P.wb, P.hb, P.nb, 2) : make_view_map(&P.tmP[s], p->p[s], P.wb, P.hb, P.nb, P.cbp);
    if (e) return e;
  }
  // thin x (<= 32 channels in a single 64-wide cq tile): 32- / 16-channel rows, see thin_cb()
  P.cbq = (p->cq_pad <= 64) ? pick_cb(p->q.C) : kChunk;
  {
    const int bnq = p->cq_pad >= 256 ? 256 : (p->cq_pad >= 128 ? 128 : 64);
    P.q5 = (wg5 && !halo && !mstack && bnq > 64 && p->q.C % bnq == 0) ? 1 : 0;
    int e = P.q5 ? make_view_map5(&P.tmQ, p->q, P.wb, P.hb, P.nb, bnq / 64)
                 : (mstack ? make_view_map(&P.tmQ, p->q, 8, 10, 1, P.cbq)
                           : (halo ? make_view_map(&P.tmQ, p->q, 16, 10, 1, P.cbq) : make_view_map(&P.tmQ, p->q, P.wb, P.hb, P.nb, P.cbq)));
    if (e) return e;
  }
  for (int t = 0; t < p->ntaps; ++t) {
    P.taps[t].src = (int16_t)p->taps[t].src; P.taps[t].dh = (int16_t)p->taps[t].dh;
    P.taps[t].dw = (int16_t)p->taps[t].dw; P.taps[t].wt = (int16_t)p->taps[t].wt;
  }
  P.ntaps = p->ntaps;
  if (mstack) {       // the kernel sees three "taps" (kernel rows); taps[kh * 3 + kw] keeps the packed-weight index of every real tap
    for (int i = 0; i < 9; ++i) { P.taps[i].src = 0; P.taps[i].dh = (int16_t)(i / 3 - 1); P.taps[i].dw = (int16_t)(i % 3 - 1); P.taps[i].wt = (int16_t)wt9[i]; }
    P.ntaps = 3;
  }
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
  const uint32_t a_bytes = 2 * KPX * 128, b_bytes = (BNQ / 64) * KPX * 2 * P.cbq;
  int tpc = 512 / P.n_mma;
  int by_smem = (int)((64 * 1024 - a_bytes) / b_bytes);
  if (by_smem < 1) by_smem = 1;
  if (!halo && !mstack && tpc > by_smem) tpc = by_smem;
  if (tpc > P.ntaps) tpc = P.ntaps;
  // wide tiles: prefer two accumulator stages (<= 256 columns per stage) over sharing the dy tile between more taps
  if (!mstack && P.n_mma >= 128 && tpc * P.n_mma > 256) tpc = 256 / P.n_mma;
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
  const uint32_t stage_bytes = a_bytes + (mstack ? (uint32_t)(BNQ / 64) * 8 * 10 * 2 * P.cbq
                                                 : halo ? (uint32_t)(BNQ / 64) * 16 * 10 * 2 * P.cbq : (uint32_t)P.tpc * b_bytes);
  P.stages = (int)((200 * 1024) / stage_bytes);
  if (P.stages > 8) P.stages = 8;
  P.m64 = 0;
  if (P.cp_blocks == 1) {
    P.m64 = 1;
    if (const char* e = getenv("DN_WGRAD_M64")) P.m64 = atoi(e);
  }
  if (mstack) P.m64 = 0;
  P.idesc = make_idesc(P0.dtype, p->q.dtype, 1, 1, mstack ? 4 * P.cbp : (P.m64 ? 64 : 128), P.n_mma);
  P.dw = p->dw;
  P.cp = P0.C; P.cq = p->q.C; P.cp_pad = p->cp_pad; P.cq_pad = p->cq_pad;
  P.scale = p->scale;
  P.dbg = g_tc_dbg;
  const int items = out_tiles * P.splits;
  P.num_items = items;
  switch (BNQ) {
    case 256: return launch_wgrad<256>(P, items, st);
    case 128: return launch_wgrad<128>(P, items, st);
    default: return launch_wgrad<64>(P, items, st);
  }
}
