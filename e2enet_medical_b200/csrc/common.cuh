// Shared helpers for the E2ENet B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "e2enet_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "e2enet_b200 kernels are written for sm_100a (B200) only"
#endif

void e2e_set_error(const char* fmt, ...);
void e2e_count_launch(int n = 1);

#define E2E_ARG(cond, ...)                                   \
  do {                                                       \
    if (!(cond)) {                                           \
      e2e_set_error(__VA_ARGS__);                            \
      return E2E_ERR_ARG;                                    \
    }                                                        \
  } while (0)

#define E2E_LAUNCHED(name)                                                         \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      e2e_set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));        \
      return E2E_ERR_CUDA;                                                         \
    }                                                                              \
    e2e_count_launch();                                                            \
  } while (0)

#define E2E_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) {                                                       \
      e2e_set_error("%s failed: %s", #call, cudaGetErrorString(_e));               \
      return E2E_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

// 16-bit storage type of activations, gradients and packed weights.  The default build is bf16 (north_star); the
// same sources compiled with -DE2E_FP16 give libe2enet_b200_fp16.so, whose operands are IEEE fp16 -- the reference's
// shipped AMP arithmetic (torch.cuda.amp.autocast, nnUNetTrainer_simple.py:552-557) -- at the same tcgen05
// kind::f16 rate.  fp16 needs a loss scale (training.TrainStep keeps one on the device).
#ifdef E2E_FP16
#include <cuda_fp16.h>
typedef __half act16;
#define E2E_UMMA_FMT 0u                                      /* instruction-descriptor a/b format: 0 = F16, 1 = BF16 */
#define E2E_TMAP_ACT CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define E2E_PRECISION_NAME "fp16"
#else
typedef __nv_bfloat16 act16;
#define E2E_UMMA_FMT 1u
#define E2E_TMAP_ACT CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define E2E_PRECISION_NAME "bf16"
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global -> shared with zero fill when src_bytes == 0
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
// D(16x8, f32) += A(16x16, act16, row) * B(16x8, act16, col)
__device__ __forceinline__ void mma_act_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
#ifdef E2E_FP16
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
#else
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
#endif
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

#ifdef E2E_FP16
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float act_lo(uint32_t v) { return __low2float(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ float act_hi(uint32_t v) { return __high2float(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ float act_round(float x) { return __half2float(__float2half_rn(x)); }
#else
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float act_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float act_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float act_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
#endif

// streaming 16-byte loads / stores that do not pollute L1
__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// per-device caches: the SM count and "function attribute already set" flags are properties of the
// CURRENT device, not of the process (one process may drive several GPUs)
constexpr int E2E_MAX_DEVICES = 64;
static inline int e2e_cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < E2E_MAX_DEVICES) ? dev : 0;
}
static inline int e2e_num_sms() {
  static int sms[E2E_MAX_DEVICES] = {0};
  const int dev = e2e_cur_device();
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}
// true exactly once per (call site flag array, current device)
struct E2eDevOnce {
  bool done[E2E_MAX_DEVICES] = {false};
  bool first() {
    const int dev = e2e_cur_device();
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};
