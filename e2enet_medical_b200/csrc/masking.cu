// DSFF Masking kernels: multi-tensor apply_mask, kernel-granular L1 magnitudes, exact k-th
// smallest selection, threshold kill, ordered dead-list compaction, growth scatter, counters.
// Index sets must be bit-exact with the reference
// (e2enet/training/network_training/sparselearning/core_channel.py:427-434, 647-666, 721-739),
// so every floating-point sum keeps the reference's association and selection is exact
// (radix select on the fp32 bit pattern, no approximate top-k).
#include "common.cuh"

namespace {

// grid (chunks, n_tensors)
__global__ void __launch_bounds__(256) mask_apply_multi_kernel(float* const* __restrict__ w, float* const* __restrict__ mom,
                                                               const float* const* __restrict__ mask,
                                                               const int64_t* __restrict__ numel) {
  const int t = blockIdx.y;
  const long long n = numel[t];
  float* wp = w[t];
  float* mp = mom ? mom[t] : nullptr;
  const float* kp = mask[t];
  const long long n4 = ((((uintptr_t)wp | (uintptr_t)kp | (uintptr_t)mp) & 15) == 0) ? (n >> 2) : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 k = reinterpret_cast<const float4*>(kp)[i];
    float4 a = reinterpret_cast<float4*>(wp)[i];
    a.x *= k.x; a.y *= k.y; a.z *= k.z; a.w *= k.w;
    reinterpret_cast<float4*>(wp)[i] = a;
    if (mp) {
      float4 b = reinterpret_cast<float4*>(mp)[i];
      b.x *= k.x; b.y *= k.y; b.z *= k.z; b.w *= k.w;
      reinterpret_cast<float4*>(mp)[i] = b;
    }
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float k = kp[i];
    wp[i] *= k;
    if (mp) mp[i] *= k;
  }
}

// One level of the reference's nested `sum(dim=-1)` (core_channel.py:653-655) over n <= 4 fp32 values, with the
// association of the torch build the reference runs on:
//   assoc 0 (torch CPU): left to right, ((a0 + a1) + a2) + a3
//   assoc 1 (torch CUDA, where the reference's Masking lives -- it hard-codes .cuda()): the reduce kernel strides
//           last_pow2(n) threads over the reduced dimension and combines them with a shuffle tree, i.e.
//           n = 3: (a0 + a2) + a1,  n = 4: (a0 + a2) + (a1 + a3)   [measured on B200 / torch 2.11:
//           tests/test_gpu_oracle_fullsize.py::test_cuda_sum_association..., SURVEY H6]
__device__ __forceinline__ float fold_level(const float* a, int n, int assoc) {
  if (assoc == 1 && n == 3) return __fadd_rn(__fadd_rn(a[0], a[2]), a[1]);
  if (assoc == 1 && n == 4) return __fadd_rn(__fadd_rn(a[0], a[2]), __fadd_rn(a[1], a[3]));
  float r = a[0];
  for (int i = 1; i < n; ++i) r = __fadd_rn(r, a[i]);
  return r;
}

// l1 = sum_kd( sum_kh( sum_kw |w| ) ): three nested folds of at most 4 values each
__global__ void __launch_bounds__(256) kernel_l1_kernel(const float* __restrict__ w, int n_kernels, int kd, int kh, int kw,
                                                        int assoc, float* __restrict__ l1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_kernels) return;
  const float* p = w + (long long)i * kd * kh * kw;
  float planes[4];
  for (int a = 0; a < kd; ++a) {
    float rows[4];
    for (int b = 0; b < kh; ++b) {
      float v[4];
      for (int c = 0; c < kw; ++c) v[c] = fabsf(p[c]);
      p += kw;
      rows[b] = fold_level(v, kw, assoc);
    }
    planes[a] = fold_level(rows, kh, assoc);
  }
  l1[i] = fold_level(planes, kd, assoc);
}

// exact k-th smallest of non-negative floats (bit pattern order == value order); one CTA
__global__ void __launch_bounds__(1024) kth_select_kernel(const float* __restrict__ v, int n, int rank, float* __restrict__ out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_rank;
  unsigned int prefix = 0, mask = 0;
  unsigned int r = (unsigned int)rank;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    unsigned int zeros = 0;   // dead kernels have L1 == +0: count them locally instead of hammering one bin
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned int u = __float_as_uint(v[i]);
      if (u == 0u) { ++zeros; continue; }
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 0xffu], 1u);
    }
    if (zeros && prefix == 0u) atomicAdd(&hist[0], zeros);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int acc = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] > r) break;
        acc += hist[b];
      }
      if (b > 255) b = 255;
      s_prefix = prefix | ((unsigned int)b << shift);
      s_rank = r - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    r = s_rank;
    mask |= 0xffu << shift;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = __uint_as_float(prefix);
}

__global__ void __launch_bounds__(256) mask_kill_kernel(const float* __restrict__ l1, const float* __restrict__ thr,
                                                        float* __restrict__ mask, int n_kernels, int ksize,
                                                        int* __restrict__ counts) {
  const float t = *thr;
  int alive = 0, dead = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_kernels; i += gridDim.x * blockDim.x) {
    float* m = mask + (long long)i * ksize;
    if (l1[i] <= t) {
      for (int k = 0; k < ksize; ++k) m[k] = 0.f;
      ++dead;
    } else {
      float s = 0.f;
      for (int k = 0; k < ksize; ++k) s += m[k];
      if (s < 1.f) ++dead; else ++alive;
    }
  }
  alive = (int)warp_sum((float)alive);   // counts < 2^24: exact in fp32
  dead = (int)warp_sum((float)dead);
  if ((threadIdx.x & 31) == 0) {
    if (alive) atomicAdd(&counts[0], alive);
    if (dead) atomicAdd(&counts[1], dead);
  }
}

// ordered compaction, one CTA of 1024 threads, contiguous chunk per thread
__global__ void __launch_bounds__(1024) dead_list_kernel(const float* __restrict__ mask, int n_kernels, int ksize,
                                                         int* __restrict__ dead, int* __restrict__ n_dead) {
  __shared__ int wsum[32];
  __shared__ int total;
  const int per = (n_kernels + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(n_kernels, lo + per);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) {
    const float* m = mask + (long long)i * ksize;
    float s = 0.f;
    for (int k = 0; k < ksize; ++k) s += fabsf(m[k]);
    cnt += (s < 1.f);
  }
  // block exclusive scan of cnt
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int v = wsum[lane];
    int vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= o) vi += y;
    }
    wsum[lane] = vi - v;
    if (lane == 31) total = vi;
  }
  __syncthreads();
  int pos = wsum[warp] + inc - cnt;
  for (int i = lo; i < hi; ++i) {
    const float* m = mask + (long long)i * ksize;
    float s = 0.f;
    for (int k = 0; k < ksize; ++k) s += fabsf(m[k]);
    if (s < 1.f) dead[pos++] = i;
  }
  if (threadIdx.x == 0) *n_dead = total;
}

__global__ void mask_grow_kernel(float* __restrict__ mask, const int* __restrict__ dead, const int* __restrict__ pick,
                                 int n_pick, int ksize) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pick * ksize) return;
  const int k = i % ksize, q = i / ksize;
  mask[(long long)dead[pick[q]] * ksize + k] = 1.f;
}

__global__ void __launch_bounds__(256) mask_counts_kernel(const float* __restrict__ mask, uint8_t* __restrict__ fired,
                                                          long long numel, int* __restrict__ nnz) {
  int a = 0, f = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel;
       i += (long long)gridDim.x * blockDim.x) {
    const bool on = mask[i] != 0.f;
    a += on;
    if (fired) {
      const uint8_t nf = fired[i] | (uint8_t)on;
      fired[i] = nf;
      f += (nf != 0);
    }
  }
  a = (int)warp_sum((float)a);
  f = (int)warp_sum((float)f);
  if ((threadIdx.x & 31) == 0) {
    if (a) atomicAdd(&nnz[0], a);
    if (f) atomicAdd(&nnz[1], f);
  }
}

}  // namespace

extern "C" int e2e_mask_apply_multi(float* const* w, float* const* mom, const float* const* mask, const int64_t* numel,
                                    int32_t n_tensors, int64_t max_numel, void* stream) {
  E2E_ARG(w && mask && numel && n_tensors > 0, "mask_apply_multi: bad arguments");
  long long chunks = (max_numel / 4 + 255) / 256;
  const long long cap = (e2e_num_sms() * 8 + n_tensors - 1) / n_tensors;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  mask_apply_multi_kernel<<<dim3((unsigned)chunks, n_tensors), 256, 0, (cudaStream_t)stream>>>(w, mom, mask, numel);
  E2E_LAUNCHED("mask_apply_multi");
  return E2E_OK;
}

extern "C" int e2e_mask_kernel_l1(const float* w, int32_t n_kernels, int32_t kd, int32_t kh, int32_t kw, int32_t assoc, float* l1,
                                  void* stream) {
  E2E_ARG(w && l1 && n_kernels > 0 && kd > 0 && kh > 0 && kw > 0, "mask_kernel_l1: bad arguments");
  E2E_ARG(kd <= 4 && kh <= 4 && kw <= 4, "mask_kernel_l1: kernel extents above 4 are not on the E2ENet path (got %d,%d,%d)", kd, kh, kw);
  E2E_ARG(assoc == 0 || assoc == 1, "mask_kernel_l1: assoc must be 0 (torch CPU order) or 1 (torch CUDA order)");
  kernel_l1_kernel<<<(n_kernels + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_kernels, kd, kh, kw, assoc, l1);
  E2E_LAUNCHED("mask_kernel_l1");
  return E2E_OK;
}

extern "C" int e2e_mask_kth(const float* l1, int32_t n, int32_t rank, float* thr, uint32_t* scratch, void* stream) {
  (void)scratch;
  E2E_ARG(l1 && thr && n > 0 && rank >= 0 && rank < n, "mask_kth: rank %d outside [0,%d)", rank, n);
  kth_select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(l1, n, rank, thr);
  E2E_LAUNCHED("mask_kth");
  return E2E_OK;
}

extern "C" int e2e_mask_kill(const float* l1, const float* thr, float* mask, int32_t n_kernels, int32_t ksize,
                             int32_t* counts, void* stream) {
  E2E_ARG(l1 && thr && mask && counts && n_kernels > 0 && ksize > 0, "mask_kill: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  E2E_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
  int blocks = (n_kernels + 255) / 256;
  if (blocks > e2e_num_sms() * 4) blocks = e2e_num_sms() * 4;
  mask_kill_kernel<<<blocks, 256, 0, st>>>(l1, thr, mask, n_kernels, ksize, counts);
  E2E_LAUNCHED("mask_kill");
  return E2E_OK;
}

extern "C" int e2e_mask_dead_list(const float* mask, int32_t n_kernels, int32_t ksize, int32_t* dead, int32_t* n_dead,
                                  int32_t* scratch, void* stream) {
  (void)scratch;
  E2E_ARG(mask && dead && n_dead && n_kernels > 0 && ksize > 0, "mask_dead_list: bad arguments");
  dead_list_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mask, n_kernels, ksize, dead, n_dead);
  E2E_LAUNCHED("mask_dead_list");
  return E2E_OK;
}

extern "C" int e2e_mask_grow(float* mask, const int32_t* dead, const int32_t* pick, int32_t n_pick, int32_t ksize,
                             void* stream) {
  E2E_ARG(mask && dead && (pick || n_pick == 0) && ksize > 0, "mask_grow: bad arguments");
  if (n_pick <= 0) return E2E_OK;
  const int total = n_pick * ksize;
  mask_grow_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mask, dead, pick, n_pick, ksize);
  E2E_LAUNCHED("mask_grow");
  return E2E_OK;
}

extern "C" int e2e_mask_counts(const float* mask, uint8_t* fired, int64_t numel, int32_t* nnz, void* stream) {
  E2E_ARG(mask && nnz && numel > 0, "mask_counts: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  E2E_CUDA(cudaMemsetAsync(nnz, 0, 2 * sizeof(int32_t), st));
  long long blocks = (numel + 255) / 256;
  if (blocks > e2e_num_sms() * 8) blocks = e2e_num_sms() * 8;
  mask_counts_kernel<<<(unsigned)blocks, 256, 0, st>>>(mask, fired, numel, nnz);
  E2E_LAUNCHED("mask_counts");
  return E2E_OK;
}
