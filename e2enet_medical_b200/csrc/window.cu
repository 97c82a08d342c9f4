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

// ------------------------------------------------------------------ fused 1x1x1 seg head + softmax + accumulate
// The last layer of the network at inference is Conv3d(C, ncls, 1) on the full-resolution feature map
// (unetpp_d.py:394-401,480-488), followed by softmax_helper, the un-mirroring, the Gaussian weighting and the
// `+=` into the accumulators (neural_network.py:529-563, 390-393).  With ncls <= 32 outputs the head is a
// 2*C*ncls-FLOP-per-voxel dot product: far too narrow for a tensor-core tile and cheap enough for the CUDA cores
// to hide under the memory traffic, so it is folded into the accumulate pass: the fp32 logits (ncls * 4 B per
// voxel written by a head kernel and read back here) never exist.  x: bf16 C8 [Cb][px][py][pz][8] of ONE tile;
// w: fp32 [ncls][C] (rounded to bf16 here, like the packed operand of the GEMM head); one thread per voxel.
template <int NC, int VPT>      // VPT voxels per thread: every weight fetched from shared memory is used VPT times
__global__ void __launch_bounds__(256) window_head_accumulate_kernel(const uint4* __restrict__ x, int Cb, const float* __restrict__ w,
                                                                     int C, const float* __restrict__ gauss, float* __restrict__ agg,
                                                                     float* __restrict__ wsum, int ncls, int px, int py, int pz,
                                                                     int X, int Y, int Z, int x0, int y0, int z0, int flip,
                                                                     float scale, int add_weight) {
  extern __shared__ float w_s[];                 // [Cb * 8][NC]: w_s[k * NC + c] = bf16(w[c][k]) (0 beyond C / ncls)
  for (int i = threadIdx.x; i < Cb * 8 * NC; i += blockDim.x) {
    const int k = i / NC, c = i - k * NC;
    w_s[i] = (k < C && c < ncls) ? act_round(w[c * C + k]) : 0.f;
  }
  __syncthreads();
  const long long P = (long long)px * py * pz;
  const long long V = (long long)X * Y * Z;
  const long long span = (long long)blockDim.x * VPT;
  for (long long i0 = blockIdx.x * span + threadIdx.x; i0 < P; i0 += (long long)gridDim.x * span) {
    long long src[VPT], dst[VPT], ii_[VPT];
    bool ok[VPT];
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      const long long i = i0 + (long long)q * blockDim.x;
      ok[q] = i < P;
      const long long ic = ok[q] ? i : 0;
      const int k = (int)(ic % pz);
      const long long t = ic / pz;
      const int j = (int)(t % py), ii = (int)(t / py);
      const int si = (flip & 1) ? px - 1 - ii : ii;
      const int sj = (flip & 2) ? py - 1 - j : j;
      const int sk = (flip & 4) ? pz - 1 - k : k;
      src[q] = ((long long)si * py + sj) * pz + sk;
      dst[q] = ((long long)(x0 + ii) * Y + (y0 + j)) * Z + (z0 + k);
      ii_[q] = ic;
    }
    float v[VPT][NC];
#pragma unroll
    for (int q = 0; q < VPT; ++q)
#pragma unroll
      for (int c = 0; c < NC; ++c) v[q][c] = 0.f;
    for (int cb = 0; cb < Cb; ++cb) {
      float f[VPT][8];
#pragma unroll
      for (int q = 0; q < VPT; ++q) {
        const uint4 u = ld_nc_16(x + (long long)cb * P + src[q]);
        f[q][0] = act_lo(u.x); f[q][1] = act_hi(u.x); f[q][2] = act_lo(u.y); f[q][3] = act_hi(u.y);
        f[q][4] = act_lo(u.z); f[q][5] = act_hi(u.z); f[q][6] = act_lo(u.w); f[q][7] = act_hi(u.w);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4* wr = reinterpret_cast<const float4*>(w_s + (cb * 8 + e) * NC);
#pragma unroll
        for (int c4 = 0; c4 < NC / 4; ++c4) {
          const float4 ww = wr[c4];
#pragma unroll
          for (int q = 0; q < VPT; ++q) {
            v[q][4 * c4 + 0] = __fmaf_rn(f[q][e], ww.x, v[q][4 * c4 + 0]);
            v[q][4 * c4 + 1] = __fmaf_rn(f[q][e], ww.y, v[q][4 * c4 + 1]);
            v[q][4 * c4 + 2] = __fmaf_rn(f[q][e], ww.z, v[q][4 * c4 + 2]);
            v[q][4 * c4 + 3] = __fmaf_rn(f[q][e], ww.w, v[q][4 * c4 + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      if (!ok[q]) continue;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < ncls) mx = fmaxf(mx, v[q][c]);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        v[q][c] = (c < ncls) ? expf(v[q][c] - mx) : 0.f;
        sum += v[q][c];
      }
      const float inv = 1.0f / sum;
      const float g = gauss ? gauss[ii_[q]] : 1.f;
      float a[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) a[c] = (c < ncls) ? agg[c * V + dst[q]] : 0.f;
      const float w0 = add_weight ? wsum[dst[q]] : 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < ncls) agg[c * V + dst[q]] = a[c] + ((v[q][c] * inv) * scale) * g;
      if (add_weight) wsum[dst[q]] = w0 + g;
    }
  }
}

// voxels [v0, v1) of the flattened volume (a range of x-planes: the tiled predictor finalises the planes no later tile
// touches while the remaining tiles are still being computed)
template <int NC>
__global__ void __launch_bounds__(256) window_finalize_kernel(float* __restrict__ agg, const float* __restrict__ wsum, int ncls,
                                                              long long V, long long v0, long long v1,
                                                              long long* __restrict__ seg) {
  for (long long i = v0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < v1;
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

extern "C" int e2e_window_head_accumulate(const void* x, int32_t Cb, const float* w, int32_t C, const float* gauss, float* agg,
                                          float* wsum, int32_t ncls, int32_t px, int32_t py, int32_t pz, int32_t X, int32_t Y,
                                          int32_t Z, int32_t x0, int32_t y0, int32_t z0, int32_t flip, float scale,
                                          int32_t add_weight, void* stream) {
  E2E_ARG(x && w && agg && wsum, "window_head_accumulate: null pointer");
  E2E_ARG(ncls >= 1 && ncls <= MAXC, "window_head_accumulate: ncls %d outside [1,%d]", ncls, MAXC);
  E2E_ARG(Cb >= 1 && C >= 1 && C <= 8 * Cb && Cb <= 40, "window_head_accumulate: bad channel counts (C %d, Cb %d)", C, Cb);
  E2E_ARG(x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + px <= X && y0 + py <= Y && z0 + pz <= Z,
          "window_head_accumulate: tile (%d,%d,%d)+(%d,%d,%d) outside volume (%d,%d,%d)", x0, y0, z0, px, py, pz, X, Y, Z);
  const long long P = (long long)px * py * pz;
  cudaStream_t st = (cudaStream_t)stream;
  const uint4* xp = (const uint4*)x;
#define E2E_WHA(NC, VPT)                                                                                                   \
  do {                                                                                                                     \
    long long blocks = (P + 256 * VPT - 1) / (256 * VPT);                                                                  \
    const long long cap = (long long)e2e_num_sms() * 8;                                                                    \
    if (blocks > cap) blocks = cap;                                                                                        \
    window_head_accumulate_kernel<NC, VPT><<<(unsigned)blocks, 256, (size_t)Cb * 8 * NC * 4, st>>>(                        \
        xp, Cb, w, C, gauss, agg, wsum, ncls, px, py, pz, X, Y, Z, x0, y0, z0, flip, scale, add_weight);                   \
  } while (0)
  if (ncls <= 4) E2E_WHA(4, 4);
  else if (ncls <= 16) E2E_WHA(16, 2);
  else E2E_WHA(32, 1);
#undef E2E_WHA
  E2E_LAUNCHED("window_head_accumulate");
  return E2E_OK;
}

static int window_finalize_impl(float* agg, const float* wsum, int32_t ncls, int32_t X, int32_t Y, int32_t Z, int32_t x0,
                                int32_t x1, int64_t* seg, void* stream) {
  const long long V = (long long)X * Y * Z, v0 = (long long)x0 * Y * Z, v1 = (long long)x1 * Y * Z;
  if (v1 <= v0) return E2E_OK;
  long long blocks = (v1 - v0 + 255) / 256;
  const long long cap = (long long)e2e_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (ncls <= 4) window_finalize_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(agg, wsum, ncls, V, v0, v1, (long long*)seg);
  else if (ncls <= 16) window_finalize_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(agg, wsum, ncls, V, v0, v1, (long long*)seg);
  else window_finalize_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(agg, wsum, ncls, V, v0, v1, (long long*)seg);
  E2E_LAUNCHED("window_finalize");
  return E2E_OK;
}

extern "C" int e2e_window_finalize(float* agg, const float* wsum, int32_t ncls, int32_t X, int32_t Y, int32_t Z,
                                   int64_t* seg, void* stream) {
  E2E_ARG(agg && wsum && seg && ncls >= 1, "window_finalize: bad arguments");
  E2E_ARG(ncls <= MAXC, "window_finalize: ncls %d outside [1,%d]", ncls, MAXC);
  return window_finalize_impl(agg, wsum, ncls, X, Y, Z, 0, X, seg, stream);
}

extern "C" int e2e_window_finalize_range(float* agg, const float* wsum, int32_t ncls, int32_t X, int32_t Y, int32_t Z,
                                         int32_t x0, int32_t x1, int64_t* seg, void* stream) {
  E2E_ARG(agg && wsum && seg && ncls >= 1 && x0 >= 0 && x0 <= x1 && x1 <= X, "window_finalize_range: bad arguments");
  E2E_ARG(ncls <= MAXC, "window_finalize_range: ncls %d outside [1,%d]", ncls, MAXC);
  return window_finalize_impl(agg, wsum, ncls, X, Y, Z, x0, x1, seg, stream);
}
