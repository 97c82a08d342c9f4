// Sliding-window accumulate on the device (replaces the per-tile D2H copy + single-thread
// NumPy `agg[:, tile] += patch; nb[:, tile] += gaussian` of
// e2enet/network_architecture/neural_network.py:383-393 and the softmax / mirror / gaussian
// passes of :529-563, and the final `agg /= nb; argmax(0)` of :403-407).
// HBM-bound: per tile voxel we read ncls logits + 1 gaussian weight and RMW ncls+1 fp32
// accumulators; softmax lives in registers; accesses are coalesced along z.
#include "common.cuh"

namespace {

constexpr int MAXC = 32;

template <int NC>   // NC = compile-time upper bound on classes kept in registers
__global__ void __launch_bounds__(256) window_accumulate_kernel(const float* __restrict__ logits, const float* __restrict__ gauss,
                                                                float* __restrict__ agg, float* __restrict__ wsum, int ncls,
                                                                int px, int py, int pz, int X, int Y, int Z, int x0,
                                                                int y0, int z0, int flip, float scale, int add_weight, int apply_softmax) {
  const long long P = (long long)px * py * pz;
  const long long V = (long long)X * Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % pz);
    const long long t = i / pz;
    const int j = (int)(t % py), ii = (int)(t / py);
    // the network saw the tile flipped: prediction voxel of tile voxel (ii,j,k)
    const int si = (flip & 1) ? px - 1 - ii : ii;
    const int sj = (flip & 2) ? py - 1 - j : j;
    const int sk = (flip & 4) ? pz - 1 - k : k;
    const long long s = ((long long)si * py + sj) * pz + sk;
    float v[NC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      v[c] = (c < ncls) ? logits[c * P + s] : -INFINITY;
      mx = fmaxf(mx, v[c]);
    }
    float inv = 1.f;
    if (apply_softmax) {
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        v[c] = (c < ncls) ? expf(v[c] - mx) : 0.f;
        sum += v[c];
      }
      inv = 1.0f / sum;
    }
    const float g = gauss ? gauss[i] : 1.f;
    const long long dst = ((long long)(x0 + ii) * Y + (y0 + j)) * Z + (z0 + k);
    // all accumulator loads first, then all stores: `agg[c*V + dst] += ...` in one loop makes every
    // load wait for the previous store (the compiler cannot prove the class planes distinct)
    float a[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) a[c] = (c < ncls) ? agg[c * V + dst] : 0.f;
    const float w0 = add_weight ? wsum[dst] : 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < ncls) agg[c * V + dst] = a[c] + ((v[c] * inv) * scale) * g;
    if (add_weight) wsum[dst] = w0 + g;
  }
}

template <int NC>
__global__ void __launch_bounds__(256) window_finalize_kernel(float* __restrict__ agg, const float* __restrict__ wsum, int ncls,
                                                              long long V, long long* __restrict__ seg) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V;
       i += (long long)gridDim.x * blockDim.x) {
    const float w = wsum[i];
    float a[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) a[c] = (c < ncls) ? agg[c * V + i] : 0.f;     // independent loads in flight
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c >= ncls) break;
      const float p = a[c] / w;
      agg[c * V + i] = p;
      if (p > best) { best = p; bi = c; }
    }
    seg[i] = bi;
  }
}

}  // namespace

extern "C" int e2e_window_accumulate(const float* logits, const float* gauss, float* agg, float* wsum, int32_t ncls,
                                     int32_t px, int32_t py, int32_t pz, int32_t X, int32_t Y, int32_t Z, int32_t x0,
                                     int32_t y0, int32_t z0, int32_t flip, float scale, int32_t add_weight,
                                     int32_t apply_softmax, void* stream) {
  E2E_ARG(logits && agg && wsum, "window_accumulate: null pointer");
  E2E_ARG(ncls >= 1 && ncls <= MAXC, "window_accumulate: ncls %d outside [1,%d]", ncls, MAXC);
  E2E_ARG(x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + px <= X && y0 + py <= Y && z0 + pz <= Z,
          "window_accumulate: tile (%d,%d,%d)+(%d,%d,%d) outside volume (%d,%d,%d)", x0, y0, z0, px, py, pz, X, Y, Z);
  const long long P = (long long)px * py * pz;
  long long blocks = (P + 255) / 256;
  const long long cap = (long long)e2e_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (ncls <= 4)
    window_accumulate_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(logits, gauss, agg, wsum, ncls, px, py, pz, X, Y, Z, x0, y0, z0, flip, scale, add_weight, apply_softmax);
  else if (ncls <= 16)
    window_accumulate_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(logits, gauss, agg, wsum, ncls, px, py, pz, X, Y, Z, x0, y0, z0, flip, scale, add_weight, apply_softmax);
  else
    window_accumulate_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(logits, gauss, agg, wsum, ncls, px, py, pz, X, Y, Z, x0, y0, z0, flip, scale, add_weight, apply_softmax);
  E2E_LAUNCHED("window_accumulate");
  return E2E_OK;
}

extern "C" int e2e_window_finalize(float* agg, const float* wsum, int32_t ncls, int32_t X, int32_t Y, int32_t Z,
                                   int64_t* seg, void* stream) {
  E2E_ARG(agg && wsum && seg && ncls >= 1, "window_finalize: bad arguments");
  const long long V = (long long)X * Y * Z;
  long long blocks = (V + 255) / 256;
  const long long cap = (long long)e2e_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  E2E_ARG(ncls <= MAXC, "window_finalize: ncls %d outside [1,%d]", ncls, MAXC);
  cudaStream_t st = (cudaStream_t)stream;
  if (ncls <= 4) window_finalize_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(agg, wsum, ncls, V, (long long*)seg);
  else if (ncls <= 16) window_finalize_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(agg, wsum, ncls, V, (long long*)seg);
  else window_finalize_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(agg, wsum, ncls, V, (long long*)seg);
  E2E_LAUNCHED("window_finalize");
  return E2E_OK;
}
