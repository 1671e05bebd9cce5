// Shared device/host helpers for libdispnet_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/dispnet_b200.h"

#define DN_CHECK_LAUNCH()                        \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

#define DN_EXPORT extern "C" __attribute__((visibility("default")))

static inline cudaStream_t dn_stream(void* s) { return (cudaStream_t)s; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel launched through dn_launch() begins with dn_pdl_trigger() (lets the NEXT kernel of the stream / graph start
// being scheduled as soon as all blocks of this one have started) and runs dn_pdl_wait() after its private set-up (barrier
// init, TMEM allocation, constants) and before its first access to global memory that an earlier kernel may still be
// reading or writing: the wait returns when the preceding grid has completed and flushed.  The launch latency, block
// scheduling ramp and prologue of kernel i+1 thus overlap the tail of kernel i (~300 launches per training step).
// DN_PDL=0 launches the same kernels with plain stream serialisation.
__device__ __forceinline__ void dn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void dn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
static inline bool dn_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DN_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
template <typename... KArgs, typename... Args>
static inline void dn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = dn_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- dtype-erased scalar access (generic kernels; the hot kernels use vector paths) -------------
__device__ __forceinline__ float dn_ld(const void* p, int dt, long long i) {
  if (dt == DN_F32) return ((const float*)p)[i];
  if (dt == DN_F16) return __half2float(((const __half*)p)[i]);
  return __bfloat162float(((const __nv_bfloat16*)p)[i]);
}
__device__ __forceinline__ void dn_st(void* p, int dt, long long i, float v) {
  if (dt == DN_F32) ((float*)p)[i] = v;
  else if (dt == DN_F16) ((__half*)p)[i] = __float2half_rn(v);
  else if (dt == DN_BF16_LO) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v - __bfloat162float(__float2bfloat16_rn(v)));
  else ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
}
__host__ __device__ __forceinline__ int dn_esize(int dt) { return dt == DN_F32 ? 4 : 2; }

__device__ __forceinline__ long long dn_off(const dn_view& v, int n, int h, int w) {
  return (long long)n * v.sN + (long long)h * v.sH + (long long)w * v.sW;
}

__device__ __forceinline__ float dn_act(float x, int act) {
  if (act == DN_ACT_RELU) return x > 0.f ? x : 0.f;
  if (act == DN_ACT_LRELU) return x > 0.f ? x : 0.1f * x;
  return x;
}
// derivative expressed on the OUTPUT of the activation (sign is preserved by relu / lrelu)
__device__ __forceinline__ float dn_act_grad(float out, int act) {
  if (act == DN_ACT_RELU) return out > 0.f ? 1.f : 0.f;
  if (act == DN_ACT_LRELU) return out > 0.f ? 1.f : 0.1f;
  return 1.f;
}

// ---- 8 x 16-bit vector <-> float[8] --------------------------------------------------------------
template <typename T> struct Vec8;
template <> struct Vec8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float* f) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store(__half* p, const float* f) {
    uint4 u; __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// true when `v` can be walked with 16-byte channel vectors
static inline bool dn_vec8_ok(const dn_view* v) {
  return v->dtype != DN_F32 && (v->C % 8) == 0 && ((uintptr_t)v->ptr % 16) == 0 && (v->sN % 8) == 0 &&
         (v->sH % 8) == 0 && (v->sW % 8) == 0;
}

// ... or with two float4 per 8 fp32 channels (precisions 'fp32' / 'tc32'): the generic channel-group walkers (ldc<8> / stc<8>)
// serve both; the 16-bit-only fast paths keep testing dn_vec8_ok
static inline bool dn_vec8_any(const dn_view* v) {
  if (v->dtype != DN_F32) return dn_vec8_ok(v);
  return (v->C % 8) == 0 && ((uintptr_t)v->ptr % 16) == 0 && (v->sN % 4) == 0 && (v->sH % 4) == 0 && (v->sW % 4) == 0;
}

__device__ __forceinline__ float dn_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double dn_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int dn_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
