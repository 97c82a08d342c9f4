// Bandwidth-bound stages of the E2ENet hot path on the C8 layout (bf16 [B][C/8][V][8]):
// layout conversion, InstanceNorm statistics / apply (+LeakyReLU) / backward, MaxPool3d.
// All kernels move 16-byte vectors (one voxel x 8 channels), grid sized in multiples of the
// SM count, fp32 math, deterministic two-stage reductions (no float atomics).
#include "common.cuh"

#include <string.h>

namespace {

constexpr int EW_THREADS = 256;
constexpr int UNR = 4;            // independent 16-byte vectors in flight per thread and operand

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = act_lo(v.x); f[1] = act_hi(v.x); f[2] = act_lo(v.y); f[3] = act_hi(v.y);
  f[4] = act_lo(v.z); f[5] = act_hi(v.z); f[6] = act_lo(v.w); f[7] = act_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]), pack_act2(f[4], f[5]),
                    pack_act2(f[6], f[7]));
}

// ------------------------------------------------------------------ layout conversion
__global__ void nc_to_c8_kernel(const float* __restrict__ x, uint4* __restrict__ y, int B, int C, long long V) {
  const int Cb = (C + 7) / 8;
  const long long total = (long long)B * Cb * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % V;
    const long long t = i / V;
    const int cb = (int)(t % Cb), b = (int)(t / Cb);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      f[j] = (c < C) ? x[((long long)b * C + c) * V + v] : 0.f;
    }
    y[i] = pack8(f);
  }
}

__global__ void c8_to_nc_kernel(const uint4* __restrict__ x, float* __restrict__ y, int B, int C, long long V) {
  const int Cb = (C + 7) / 8;
  const long long total = (long long)B * Cb * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % V;
    const long long t = i / V;
    const int cb = (int)(t % Cb), b = (int)(t / Cb);
    float f[8];
    unpack8(x[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      if (c < C) y[((long long)b * C + c) * V + v] = f[j];
    }
  }
}

// ------------------------------------------------------------------ block reduction of 16 floats per thread
template <int NV>
__device__ __forceinline__ void block_reduce_store(float* acc, float* out /* NV floats */) {
  __shared__ float red[EW_THREADS / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < EW_THREADS / 32; ++w) s += red[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// grid (nchunk, B*Cb): partial[plane][chunk][0..7] = sum, [8..15] = sum of squares
__global__ void __launch_bounds__(EW_THREADS) in_stats_partial_kernel(const uint4* __restrict__ raw, long long V,
                                                                      int nchunk, float* __restrict__ partial) {
  const int plane = blockIdx.y, chunk = blockIdx.x;
  const long long per = (V + nchunk - 1) / nchunk;
  const long long lo = chunk * per, hi = min(V, lo + per);
  const uint4* base = raw + (long long)plane * V;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  // UNR independent 16-byte loads in flight per thread (the loop is latency-bound otherwise)
  for (long long v0 = lo + threadIdx.x; v0 < hi; v0 += UNR * EW_THREADS) {
    uint4 r[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * EW_THREADS;
      r[u] = v < hi ? ld_nc_16(base + v) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      float f[8];
      unpack8(r[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += f[j];
        acc[8 + j] += f[j] * f[j];
      }
    }
  }
  block_reduce_store<16>(acc, partial + ((long long)plane * nchunk + chunk) * 16);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one WARP per (plane, j): lanes stride over the chunk partials (fixed order -> deterministic)
__global__ void in_stats_final_kernel(const float* __restrict__ partial, int planes, int nchunk, long long V, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= planes * 8) return;
  const int plane = i >> 3, j = i & 7;
  double s = 0.0, q = 0.0;
  for (int c = lane; c < nchunk; c += 32) {
    s += partial[((long long)plane * nchunk + c) * 16 + j];
    q += partial[((long long)plane * nchunk + c) * 16 + 8 + j];
  }
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  if (lane) return;
  const double m = s / (double)V;
  double var = q / (double)V - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// statistics from the per-slot partials a fused conv epilogue wrote: stats[slot][b][{sum, sumsq}][C] (one slot per
// CTA of the conv launch).  grid (ceil(C / 32), B), 32 warps: lane = channel (coalesced rows), warp w takes slots
// w, w + 32, ...; the 32 partial sums are combined in a fixed order in fp64 (deterministic).
__global__ void __launch_bounds__(1024) in_stats_from_slots_kernel(const float* __restrict__ stats, int n_slots, int B, int C,
                                                                  long long V, float eps, float* __restrict__ mean,
                                                                  float* __restrict__ rstd) {
  __shared__ double red[32][2][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, c = blockIdx.x * 32 + lane;
  double s = 0.0, q = 0.0;
  if (c < C) {
#pragma unroll 4
    for (int k = w; k < n_slots; k += 32) {
      const float* row = stats + ((long long)(k * B + b) * 2) * C + c;
      s += (double)row[0];
      q += (double)row[C];
    }
  }
  red[w][0][lane] = s;
  red[w][1][lane] = q;
  __syncthreads();
  if (w != 0 || c >= C) return;
  s = 0.0; q = 0.0;
#pragma unroll
  for (int i = 0; i < 32; ++i) { s += red[i][0][lane]; q += red[i][1][lane]; }
  const double m = s / (double)V;
  double var = q / (double)V - m * m;
  if (var < 0.0) var = 0.0;
  mean[b * C + c] = (float)m;
  rstd[b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

// Block -> (plane, chunk) of the grid (nchunk, planes).  order 0: plane-major (blockIdx as is).  order 1 / 2: chunk-major
// (all planes of a voxel chunk before the next chunk: the order in which the GEMM kernels walk a tensor), descending /
// ascending.  A kernel that reads what its predecessor wrote LAST first finds that part still in the 126 MB L2: the forward
// apply and the first backward pass walk downwards (the conv / data-gradient kernel before them ended at the top), the
// second backward pass walks upwards again from where the first one ended.  E2E_EW_ORDER=0 restores plane-major order.
__device__ __forceinline__ void ew_block(int order, int& plane, int& chunk) {
  if (order == 0) { plane = blockIdx.y; chunk = blockIdx.x; return; }
  const int nch = gridDim.x, npl = gridDim.y;
  const int L = blockIdx.y * nch + blockIdx.x;
  const int inner = L % npl, outer = L / npl;
  plane = order == 1 ? npl - 1 - inner : inner;
  chunk = order == 1 ? nch - 1 - outer : outer;
}

static int ew_order(int which) {            // which: 1 = descending, 2 = ascending
  static int on = -1;
  if (on < 0) { const char* e = getenv("E2E_EW_ORDER"); on = e ? atoi(e) : 1; }
  return on ? which : 0;
}

// Plane statistics inside the consumer kernel (no separate tiny launch): every block of a (chunk, plane) grid forms
// the SAME fixed-order fp64 sums, so all blocks of a plane see bit-identical values; block (chunk 0) publishes them.
//   mean / rstd of plane (b, cb) from the per-CTA slots a fused conv epilogue wrote: stats[slot][b][{sum, sumsq}][C]
__device__ __forceinline__ void plane_mean_rstd_from_slots(const float* __restrict__ stats, int n_slots, int B, int C, int b,
                                                           int cb, long long V, float eps, float* s_mean /*[8]*/,
                                                           float* s_rstd /*[8]*/) {
  __shared__ double s_red[2][32][8];
  const int j = threadIdx.x & 7, g = threadIdx.x >> 3;             // 256 threads: 8 channels x 32 slot groups
  double s = 0.0, q = 0.0;
  const int c = cb * 8 + j;
  if (c < C) {
    for (int k = g; k < n_slots; k += 32) {
      const float* row = stats + ((long long)(k * B + b) * 2) * C + c;
      s += (double)__ldcg(row);
      q += (double)__ldcg(row + C);
    }
  }
  s_red[0][g][j] = s;
  s_red[1][g][j] = q;
  __syncthreads();
  if (threadIdx.x < 8) {
    double ss = 0.0, qq = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) { ss += s_red[0][i][threadIdx.x]; qq += s_red[1][i][threadIdx.x]; }
    const double m = ss / (double)V;
    double var = qq / (double)V - m * m;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)m;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
}
//   sum dz / sum dz*xhat of a plane from the chunk partials of the backward's first pass: partial[plane][chunk][16]
__device__ __forceinline__ void plane_sums_from_partials(const float* __restrict__ partial, int plane, int nchunk,
                                                         float* s_sums /*[16]*/) {
  __shared__ double s_red2[16][16];
  const int k = threadIdx.x & 15, g = threadIdx.x >> 4;            // 256 threads: 16 values x 16 chunk groups
  double s = 0.0;
  for (int c = g; c < nchunk; c += 16) s += (double)__ldcg(partial + ((long long)plane * nchunk + c) * 16 + k);
  s_red2[g][k] = s;
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += s_red2[i][threadIdx.x];
    s_sums[threadIdx.x] = (float)t;
  }
  __syncthreads();
}

// grid (nchunk, B*Cb)
__global__ void __launch_bounds__(EW_THREADS) in_apply_kernel(const uint4* __restrict__ raw, const float* __restrict__ mean,
                                                              const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float slope, int Cb,
                                                              long long V, int nchunk, uint4* __restrict__ out,
                                                              const float* __restrict__ stats, int n_slots, int B, float eps,
                                                              float* __restrict__ mean_out, float* __restrict__ rstd_out, int order) {
  int plane, chunk;
  ew_block(order, plane, chunk);
  const int cb = plane % Cb;
  __shared__ float s_mean[8], s_rstd[8];
  if (stats) {                      // statistics straight from the conv epilogue's slots (see plane_mean_rstd_from_slots)
    plane_mean_rstd_from_slots(stats, n_slots, B, Cb * 8, plane / Cb, cb, V, eps, s_mean, s_rstd);
    if (chunk == 0 && threadIdx.x < 8) {
      mean_out[plane * 8 + threadIdx.x] = s_mean[threadIdx.x];
      rstd_out[plane * 8 + threadIdx.x] = s_rstd[threadIdx.x];
    }
  }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float g = gamma[cb * 8 + j], r = stats ? s_rstd[j] : rstd[plane * 8 + j];
    const float m = stats ? s_mean[j] : mean[plane * 8 + j];
    sc[j] = g * r;
    sh[j] = beta[cb * 8 + j] - m * g * r;
  }
  const long long per = (V + nchunk - 1) / nchunk;
  const long long lo = chunk * per, hi = min(V, lo + per);
  const uint4* ib = raw + (long long)plane * V;
  uint4* ob = out + (long long)plane * V;
  for (long long v0 = lo + threadIdx.x; v0 < hi; v0 += UNR * EW_THREADS) {
    uint4 r[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * EW_THREADS;
      if (v < hi) r[u] = ld_nc_16(ib + v);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * EW_THREADS;
      if (v >= hi) break;
      float f[8];
      unpack8(r[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = f[j] * sc[j] + sh[j];
        f[j] = z > 0.f ? z : z * slope;
      }
      ob[v] = pack8(f);
    }
  }
}

// backward pass 1: partial[plane][chunk][0..7] = sum dz, [8..15] = sum dz*xhat
__global__ void __launch_bounds__(EW_THREADS) in_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ raw,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float slope, int Cb,
                                                                   long long V, int nchunk, float* __restrict__ partial, int order) {
  int plane, chunk;
  ew_block(order, plane, chunk);
  const int cb = plane % Cb;
  float mu[8], rs[8], ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = mean[plane * 8 + j]; rs[j] = rstd[plane * 8 + j];
    ga[j] = gamma[cb * 8 + j]; be[j] = beta[cb * 8 + j];
  }
  const long long per = (V + nchunk - 1) / nchunk;
  const long long lo = chunk * per, hi = min(V, lo + per);
  const uint4* xb = raw + (long long)plane * V;
  const uint4* gb = dy + (long long)plane * V;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (long long v0 = lo + threadIdx.x; v0 < hi; v0 += UNR * EW_THREADS) {
    uint4 xr[UNR], gr[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * EW_THREADS;
      if (v < hi) { xr[u] = ld_nc_16(xb + v); gr[u] = ld_nc_16(gb + v); }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (v0 + u * EW_THREADS >= hi) break;
      float x[8], g[8];
      unpack8(xr[u], x);
      unpack8(gr[u], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mu[j]) * rs[j];
        const float z = xh * ga[j] + be[j];
        const float dz = z > 0.f ? g[j] : g[j] * slope;
        acc[j] += dz;
        acc[8 + j] += dz * xh;
      }
    }
  }
  block_reduce_store<16>(acc, partial + ((long long)plane * nchunk + chunk) * 16);
}

// sums[plane*16 + k] (fp32) = total over chunks
__global__ void in_bwd_final_kernel(const float* __restrict__ partial, int planes, int nchunk, float* __restrict__ sums) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;     // one warp per value
  if (i >= planes * 16) return;
  const int plane = i >> 4, k = i & 15;
  double s = 0.0;
  for (int c = lane; c < nchunk; c += 32) s += partial[((long long)plane * nchunk + c) * 16 + k];
  s = warp_sum_d(s);
  if (lane == 0) sums[i] = (float)s;
}

// backward pass 2: draw = rstd*gamma*(dz - mean(dz) - xhat*mean(dz*xhat)); partial2[plane][chunk][0..7] = sum draw
__global__ void __launch_bounds__(EW_THREADS) in_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ raw,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const float* __restrict__ partial1,
                                                                  float* __restrict__ sums, float slope, int Cb,
                                                                  long long V, int nchunk, uint4* __restrict__ draw,
                                                                  float* __restrict__ partial2, int order) {
  int plane, chunk;
  ew_block(order, plane, chunk);
  const int cb = plane % Cb;
  const float invV = 1.0f / (float)V;
  __shared__ float s_sums[16];
  plane_sums_from_partials(partial1, plane, nchunk, s_sums);      // pass 1's totals, formed here (no tiny launch between)
  if (chunk == 0 && threadIdx.x < 16) sums[plane * 16 + threadIdx.x] = s_sums[threadIdx.x];
  float mu[8], rs[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = mean[plane * 8 + j]; rs[j] = rstd[plane * 8 + j];
    ga[j] = gamma[cb * 8 + j]; be[j] = beta[cb * 8 + j];
    m1[j] = s_sums[j] * invV;
    m2[j] = s_sums[8 + j] * invV;
  }
  const long long per = (V + nchunk - 1) / nchunk;
  const long long lo = chunk * per, hi = min(V, lo + per);
  const uint4* xb = raw + (long long)plane * V;
  const uint4* gb = dy + (long long)plane * V;
  uint4* ob = draw + (long long)plane * V;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (long long v0 = lo + threadIdx.x; v0 < hi; v0 += UNR * EW_THREADS) {
    uint4 xr[UNR], gr[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * EW_THREADS;
      if (v < hi) { xr[u] = ld_nc_16(xb + v); gr[u] = ld_nc_16(gb + v); }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * EW_THREADS;
      if (v >= hi) break;
      float x[8], g[8], o[8];
      unpack8(xr[u], x);
      unpack8(gr[u], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mu[j]) * rs[j];
        const float z = xh * ga[j] + be[j];
        const float dz = z > 0.f ? g[j] : g[j] * slope;
        o[j] = rs[j] * ga[j] * (dz - m1[j] - xh * m2[j]);
      }
      const uint4 pk = pack8(o);
      ob[v] = pk;
      float ro[8];
      unpack8(pk, ro);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += ro[j];
    }
  }
  block_reduce_store<8>(acc, partial2 + ((long long)plane * nchunk + chunk) * 8);
}

// dgamma[c] = sum_b sum dz*xhat ; dbeta[c] = sum_b sum dz ; dbias[c] = sum_b sum draw
__global__ void in_bwd_param_kernel(const float* __restrict__ sums, const float* __restrict__ partial2, int B, int Cb,
                                    int nchunk, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    float* __restrict__ dbias) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;     // one warp per channel
  if (c >= Cb * 8) return;
  const int cb = c >> 3, j = c & 7;
  double g = 0.0, bt = 0.0, bi = 0.0;
  for (int b = 0; b < B; ++b) {
    const int plane = b * Cb + cb;
    bt += sums[plane * 16 + j];
    g += sums[plane * 16 + 8 + j];
    for (int k = lane; k < nchunk; k += 32) bi += partial2[((long long)plane * nchunk + k) * 8 + j];
  }
  bi = warp_sum_d(bi);
  if (lane) return;
  dgamma[c] = (float)g;
  dbeta[c] = (float)bt;
  if (dbias) dbias[c] = (float)bi;
}

// ------------------------------------------------------------------ InstanceNorm + LeakyReLU fused with the
// MaxPool3d(kernel == stride) that consumes the same activation (the down* modules read x{i}_{j} right after
// its block produced it, unetpp_d.py:453-478): one thread per POOLED voxel normalises its window, writes the
// kd*kh*kw activations, the pooled maximum and its arg-max.  Saves the pool's re-read of the activation;
// backward folds the pooled gradient into the norm backward (no separate max-pool backward, no gradient add).
// Requires D % kd == H % kh == W % kw == 0 and kd*kh*kw <= 8.
struct PoolGeo { int D, H, W, kd, kh, kw; };

__device__ __forceinline__ void pool_decode(long long vo, const PoolGeo& g, int& od, int& oh, int& ow) {
  const int Wo = g.W / g.kw, Ho = g.H / g.kh;
  ow = (int)(vo % Wo);
  const long long t = vo / Wo;
  oh = (int)(t % Ho);
  od = (int)(t / Ho);
}

template <int KD, int KH, int KW>      // window shape as template: the per-window arrays must stay in registers
__global__ void __launch_bounds__(EW_THREADS) in_apply_pool_kernel(const uint4* __restrict__ raw, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float slope, int Cb, PoolGeo g,
                                                                   int nchunk, uint4* __restrict__ out, uint4* __restrict__ pooled,
                                                                   uint2* __restrict__ amax, const float* __restrict__ stats,
                                                                   int n_slots, int B, float eps, float* __restrict__ mean_out,
                                                                   float* __restrict__ rstd_out, int order) {
  int plane, chunk;
  ew_block(order, plane, chunk);
  const int cb = plane % Cb;
  const long long V = (long long)g.D * g.H * g.W;
  __shared__ float s_mean[8], s_rstd[8];
  if (stats) {
    plane_mean_rstd_from_slots(stats, n_slots, B, Cb * 8, plane / Cb, cb, V, eps, s_mean, s_rstd);
    if (chunk == 0 && threadIdx.x < 8) {
      mean_out[plane * 8 + threadIdx.x] = s_mean[threadIdx.x];
      rstd_out[plane * 8 + threadIdx.x] = s_rstd[threadIdx.x];
    }
  }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ga = gamma[cb * 8 + j], r = stats ? s_rstd[j] : rstd[plane * 8 + j];
    const float m = stats ? s_mean[j] : mean[plane * 8 + j];
    sc[j] = ga * r;
    sh[j] = beta[cb * 8 + j] - m * ga * r;
  }
  const long long Vo = V / (g.kd * g.kh * g.kw);
  const long long per = (Vo + nchunk - 1) / nchunk;
  const long long lo = chunk * per, hi = min(Vo, lo + per);
  const uint4* ib = raw + (long long)plane * V;
  uint4* ob = out + (long long)plane * V;
  for (long long vo = lo + threadIdx.x; vo < hi; vo += EW_THREADS) {
    int od, oh, ow;
    pool_decode(vo, g, od, oh, ow);
    constexpr int n = KD * KH * KW;
    uint4 r[n];
    long long src[n];
#pragma unroll
    for (int a = 0; a < KD; ++a)
#pragma unroll
      for (int b = 0; b < KH; ++b)
#pragma unroll
        for (int c = 0; c < KW; ++c) {
          const int i = (a * KH + b) * KW + c;
          src[i] = ((long long)(od * KD + a) * g.H + oh * KH + b) * g.W + ow * KW + c;
          r[i] = ld_nc_16(ib + src[i]);
        }
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float f[8];
      unpack8(r[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = f[j] * sc[j] + sh[j];
        f[j] = z > 0.f ? z : z * slope;
      }
      const uint4 pk = pack8(f);
      ob[src[i]] = pk;
      unpack8(pk, f);                    // the pool compares the bf16 activations, like the separate kernel
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (f[j] > best[j] || (f[j] != f[j] && best[j] == best[j])) { best[j] = f[j]; bi[j] = i; }
    }
    const long long po = (long long)plane * Vo + vo;
    pooled[po] = pack8(best);
    uint2 am;
    am.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
    am.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
    amax[po] = am;
  }
}

// backward with the pooled gradient folded in: dy_total[v] = dy[v] + (argmax(window(v)) == v ? dyp[window(v)] : 0)
// PASS 0: partial[plane][chunk][0..7] = sum dz, [8..15] = sum dz*xhat;  PASS 1: draw and partial2 = sum draw
template <int PASS, int KD, int KH, int KW>
__global__ void __launch_bounds__(EW_THREADS) in_bwd_pool_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ dyp,
                                                                 const uint2* __restrict__ amax, const uint4* __restrict__ raw,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ partial1, float* __restrict__ sums,
                                                                 float slope, int Cb, PoolGeo g, int nchunk,
                                                                 uint4* __restrict__ draw, float* __restrict__ partial, int order) {
  int plane, chunk;
  ew_block(order, plane, chunk);
  const int cb = plane % Cb;
  const long long V = (long long)g.D * g.H * g.W;
  const long long Vo = V / (g.kd * g.kh * g.kw);
  const float invV = 1.0f / (float)V;
  __shared__ float s_sums[16];
  if (PASS) {
    plane_sums_from_partials(partial1, plane, nchunk, s_sums);
    if (chunk == 0 && threadIdx.x < 16) sums[plane * 16 + threadIdx.x] = s_sums[threadIdx.x];
  }
  float mu[8], rs[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = mean[plane * 8 + j]; rs[j] = rstd[plane * 8 + j];
    ga[j] = gamma[cb * 8 + j]; be[j] = beta[cb * 8 + j];
    m1[j] = PASS ? s_sums[j] * invV : 0.f;
    m2[j] = PASS ? s_sums[8 + j] * invV : 0.f;
  }
  const long long per = (Vo + nchunk - 1) / nchunk;
  const long long lo = chunk * per, hi = min(Vo, lo + per);
  const uint4* xb = raw + (long long)plane * V;
  const uint4* gb = dy ? dy + (long long)plane * V : nullptr;
  uint4* ob = PASS ? draw + (long long)plane * V : nullptr;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (long long vo = lo + threadIdx.x; vo < hi; vo += EW_THREADS) {
    int od, oh, ow;
    pool_decode(vo, g, od, oh, ow);
    const long long po = (long long)plane * Vo + vo;
    const uint2 am = amax[po];
    float gp[8];
    unpack8(ld_nc_16(dyp + po), gp);
    constexpr int n = KD * KH * KW;
    uint4 xr[n], gr[n];
    long long src[n];
#pragma unroll
    for (int a = 0; a < KD; ++a)
#pragma unroll
      for (int b = 0; b < KH; ++b)
#pragma unroll
        for (int c = 0; c < KW; ++c) {
          const int i = (a * KH + b) * KW + c;
          src[i] = ((long long)(od * KD + a) * g.H + oh * KH + b) * g.W + ow * KW + c;
          xr[i] = ld_nc_16(xb + src[i]);
          gr[i] = gb ? ld_nc_16(gb + src[i]) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float x[8], gv[8], o[8];
      unpack8(xr[i], x);
      unpack8(gr[i], gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t a = ((j < 4 ? am.x : am.y) >> ((j & 3) * 8)) & 0xffu;
        const float gt = gv[j] + (a == (uint32_t)i ? gp[j] : 0.f);
        const float xh = (x[j] - mu[j]) * rs[j];
        const float z = xh * ga[j] + be[j];
        const float dz = z > 0.f ? gt : gt * slope;
        if (PASS == 0) {
          acc[j] += dz;
          acc[8 + j] += dz * xh;
        } else {
          o[j] = rs[j] * ga[j] * (dz - m1[j] - xh * m2[j]);
        }
      }
      if (PASS) {
        const uint4 pk = pack8(o);
        ob[src[i]] = pk;
        float ro[8];
        unpack8(pk, ro);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += ro[j];
      }
    }
  }
  if (PASS == 0)
    block_reduce_store<16>(acc, partial + ((long long)plane * nchunk + chunk) * 16);
  else
    block_reduce_store<8>(acc, partial + ((long long)plane * nchunk + chunk) * 8);
}

// ------------------------------------------------------------------ plane-resident InstanceNorm backward
// The two-kernel backward above reads dy and raw twice from HBM (5 tensor passes).  Here ONE kernel walks the (b, cb)
// planes; a group of Gs co-resident CTAs owns a plane at a time: phase A reduces sum dz / sum dz*xhat over the plane
// (dy and raw stream in from HBM and stay in the 126 MB L2: a full-resolution plane is 2 x 26 MB), a counter barrier
// among the group's CTAs, every CTA forms the totals from the group's partials in a fixed order (deterministic), and
// phase C re-reads ITS OWN slice -- now L2 hits -- to write draw: 3 HBM passes (dy, raw, draw) instead of 5.
// Groups are sized so that the planes in flight fit the L2; all CTAs are co-resident (grid = 2 per SM), so the
// spin-wait cannot deadlock.  POOL folds the pooled gradient in exactly like in_bwd_pool_kernel.
__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool POOL, int KD, int KH, int KW>
__global__ void __launch_bounds__(EW_THREADS, 2)
in_bwd_plane_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ dyp, const uint2* __restrict__ amax,
                    const uint4* __restrict__ raw, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float slope, int Cb, PoolGeo g, long long V,
                    int planes, int Gs, int ngroups, float* __restrict__ partial, float* __restrict__ partial2,
                    int* __restrict__ counters, float* __restrict__ sums_out, uint4* __restrict__ draw) {
  __shared__ float s_sums[16];
  const int group = blockIdx.x / Gs, c = blockIdx.x - group * Gs;
  if (group >= ngroups) return;
  constexpr int n = KD * KH * KW;                       // window voxels per pooled voxel (1 when !POOL)
  const long long Vw = POOL ? V / n : V;                // work items per plane: pooled voxels or voxels
  const long long per = (Vw + Gs - 1) / Gs;
  const long long lo = (long long)c * per, hi = min(Vw, lo + per);
  const float invV = 1.0f / (float)V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int plane = group; plane < planes; plane += ngroups) {
    const int cb = plane % Cb;
    float mu[8], rs[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = mean[plane * 8 + j]; rs[j] = rstd[plane * 8 + j];
      ga[j] = gamma[cb * 8 + j]; be[j] = beta[cb * 8 + j];
    }
    const uint4* xb = raw + (long long)plane * V;
    const uint4* gb = dy ? dy + (long long)plane * V : nullptr;
    uint4* ob = draw + (long long)plane * V;
    float m1[8], m2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m1[j] = 0.f; m2[j] = 0.f; }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.f;
      if (!POOL) {
        for (long long v0 = lo + threadIdx.x; v0 < hi; v0 += UNR * EW_THREADS) {
          uint4 xr[UNR], gr[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const long long v = v0 + u * EW_THREADS;
            if (v < hi) { xr[u] = ld_nc_16(xb + v); gr[u] = ld_nc_16(gb + v); }
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const long long v = v0 + u * EW_THREADS;
            if (v >= hi) break;
            float x[8], gv[8], o[8];
            unpack8(xr[u], x);
            unpack8(gr[u], gv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float xh = (x[j] - mu[j]) * rs[j];
              const float z = xh * ga[j] + be[j];
              const float dz = z > 0.f ? gv[j] : gv[j] * slope;
              if (pass == 0) { acc[j] += dz; acc[8 + j] += dz * xh; }
              else o[j] = rs[j] * ga[j] * (dz - m1[j] - xh * m2[j]);
            }
            if (pass) {
              const uint4 pk = pack8(o);
              ob[v] = pk;
              float ro[8];
              unpack8(pk, ro);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] += ro[j];
            }
          }
        }
      } else {
        for (long long vo = lo + threadIdx.x; vo < hi; vo += EW_THREADS) {
          int od, oh, ow;
          pool_decode(vo, g, od, oh, ow);
          const long long po = (long long)plane * Vw + vo;
          const uint2 am = amax[po];
          float gp[8];
          unpack8(ld_nc_16(dyp + po), gp);
          uint4 xr[n], gr[n];
          long long src[n];
#pragma unroll
          for (int a = 0; a < KD; ++a)
#pragma unroll
            for (int b = 0; b < KH; ++b)
#pragma unroll
              for (int cc = 0; cc < KW; ++cc) {
                const int i = (a * KH + b) * KW + cc;
                src[i] = ((long long)(od * KD + a) * g.H + oh * KH + b) * g.W + ow * KW + cc;
                xr[i] = ld_nc_16(xb + src[i]);
                gr[i] = gb ? ld_nc_16(gb + src[i]) : make_uint4(0, 0, 0, 0);
              }
#pragma unroll
          for (int i = 0; i < n; ++i) {
            float x[8], gv[8], o[8];
            unpack8(xr[i], x);
            unpack8(gr[i], gv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t a = ((j < 4 ? am.x : am.y) >> ((j & 3) * 8)) & 0xffu;
              const float gt = gv[j] + (a == (uint32_t)i ? gp[j] : 0.f);
              const float xh = (x[j] - mu[j]) * rs[j];
              const float z = xh * ga[j] + be[j];
              const float dz = z > 0.f ? gt : gt * slope;
              if (pass == 0) { acc[j] += dz; acc[8 + j] += dz * xh; }
              else o[j] = rs[j] * ga[j] * (dz - m1[j] - xh * m2[j]);
            }
            if (pass) {
              const uint4 pk = pack8(o);
              ob[src[i]] = pk;
              float ro[8];
              unpack8(pk, ro);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] += ro[j];
            }
          }
        }
      }
      if (pass == 0) {
        block_reduce_store<16>(acc, partial + ((long long)plane * Gs + c) * 16);
        // group barrier: publish this CTA's partial, wait until all Gs CTAs of the group have published theirs
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          atomicAdd(&counters[plane], 1);
          while (ld_acquire_i32(&counters[plane]) < Gs) __nanosleep(64);
        }
        __syncthreads();
        // totals of the plane, fixed order: warp w forms values w and w + 8 (fp64, lanes stride over the group's CTAs)
        for (int k = warp; k < 16; k += EW_THREADS / 32) {
          double sm = 0.0;
          for (int i = lane; i < Gs; i += 32) sm += (double)__ldcg(partial + ((long long)plane * Gs + i) * 16 + k);
          sm = warp_sum_d(sm);
          if (lane == 0) s_sums[k] = (float)sm;
        }
        __syncthreads();
        if (c == 0 && threadIdx.x < 16) sums_out[plane * 16 + threadIdx.x] = s_sums[threadIdx.x];
#pragma unroll
        for (int j = 0; j < 8; ++j) { m1[j] = s_sums[j] * invV; m2[j] = s_sums[8 + j] * invV; }
        __syncthreads();
      } else {
        block_reduce_store<8>(acc, partial2 + ((long long)plane * Gs + c) * 8);
        __syncthreads();
      }
    }
  }
}

// geometry of the plane-resident backward: CTAs per group, groups, grid
struct PlaneGeo { int grid, Gs, ngroups; };
// CTAs of in_bwd_plane_kernel that are guaranteed to be resident at the same time on the current device (the
// kernel's group barrier spins, so the grid must never exceed this)
static int plane_coresident_ctas() {
  static int cached[E2E_MAX_DEVICES] = {0};
  const int dev = e2e_cur_device();
  if (!cached[dev]) {
    int occ = 2, o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, in_bwd_plane_kernel<false, 1, 1, 1>, EW_THREADS, 0) == cudaSuccess && o < occ) occ = o;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, in_bwd_plane_kernel<true, 1, 2, 2>, EW_THREADS, 0) == cudaSuccess && o < occ) occ = o;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, in_bwd_plane_kernel<true, 2, 2, 2>, EW_THREADS, 0) == cudaSuccess && o < occ) occ = o;
    if (occ < 1) occ = 1;
    cached[dev] = occ * e2e_num_sms();
  }
  return cached[dev];
}
static PlaneGeo plane_geometry(int planes, long long V) {
  PlaneGeo pg;
  const int G = plane_coresident_ctas();                            // co-resident: __launch_bounds__(256, 2)
  const long long plane_bytes = V * 16 * 2;                         // dy + raw of one plane
  long long ng = (80ll << 20) / (plane_bytes > 0 ? plane_bytes : 1);   // planes in flight that fit the L2 comfortably
  if (ng < 1) ng = 1;
  if (ng > planes) ng = planes;
  long long gs = G / ng;
  const long long max_useful = (V + 511) / 512;                     // at least ~512 voxels per CTA
  if (gs > max_useful) gs = max_useful;
  if (gs < 1) gs = 1;
  ng = G / gs;
  if (ng > planes) ng = planes;
  pg.Gs = (int)gs; pg.ngroups = (int)ng; pg.grid = (int)(gs * ng);
  return pg;
}

// ------------------------------------------------------------------ MaxPool3d, kernel == stride
__global__ void __launch_bounds__(EW_THREADS) maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                                 uint2* __restrict__ amax, int BCb, int D, int H, int W,
                                                                 int kd, int kh, int kw) {
  const int Do = D / kd, Ho = H / kh, Wo = W / kw;
  const long long total = (long long)BCb * Do * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % Wo);
    long long t = i / Wo;
    const int oh = (int)(t % Ho);
    t /= Ho;
    const int od = (int)(t % Do);
    const long long plane = t / Do;
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
    int widx = 0;
    for (int a = 0; a < kd; ++a)
      for (int b = 0; b < kh; ++b)
        for (int c = 0; c < kw; ++c, ++widx) {
          const long long src = ((plane * D + od * kd + a) * H + oh * kh + b) * (long long)W + ow * kw + c;
          float f[8];
          unpack8(ld_nc_16(x + src), f);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (f[j] > best[j] || (f[j] != f[j] && best[j] == best[j])) { best[j] = f[j]; bi[j] = widx; }
        }
    y[i] = pack8(best);
    uint2 am;
    am.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
    am.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
    amax[i] = am;
  }
}

// one thread per POOLED voxel: reads dy + argmax once, writes the whole kd x kh x kw window of dx
// (zeros except at the arg-max position); trailing voxels of dx that no window covers are zeroed
// by the threads of the last window along each axis
__global__ void __launch_bounds__(EW_THREADS) maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ amax,
                                                                 uint4* __restrict__ dx, int BCb, int D, int H, int W,
                                                                 int kd, int kh, int kw) {
  const int Do = D / kd, Ho = H / kh, Wo = W / kw;
  const long long total = (long long)BCb * Do * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % Wo);
    long long t = i / Wo;
    const int oh = (int)(t % Ho);
    t /= Ho;
    const int od = (int)(t % Do);
    const long long plane = t / Do;
    const uint2 am = amax[i];
    const uint4 gv = ld_nc_16(dy + i);
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
    // extents of this thread's window (the last window along an axis also covers the remainder)
    const int d1 = (od == Do - 1) ? D : (od + 1) * kd;
    const int h1 = (oh == Ho - 1) ? H : (oh + 1) * kh;
    const int w1 = (ow == Wo - 1) ? W : (ow + 1) * kw;
    for (int d = od * kd; d < d1; ++d)
      for (int h = oh * kh; h < h1; ++h)
        for (int w = ow * kw; w < w1; ++w) {
          const int a = d - od * kd, b = h - oh * kh, c = w - ow * kw;
          const uint32_t widx = (a < kd && b < kh && c < kw) ? (uint32_t)((a * kh + b) * kw + c) : 0xffffu;
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t a0 = ((q < 2 ? am.x : am.y) >> ((2 * q & 3) * 8)) & 0xffu;
            const uint32_t a1 = ((q < 2 ? am.x : am.y) >> (((2 * q + 1) & 3) * 8)) & 0xffu;
            o[q] = (a0 == widx ? (gw[q] & 0xffffu) : 0u) | (a1 == widx ? (gw[q] & 0xffff0000u) : 0u);
          }
          dx[((plane * D + d) * H + h) * (long long)W + w] = make_uint4(o[0], o[1], o[2], o[3]);
        }
  }
}

// ------------------------------------------------------------------ stand-alone depth shift (torch_shift.forward,
// unetpp_d.py:45-59) on a plain NCDHW tensor of 2- or 4-byte elements: y[b,c,d] = x[b,c,d - s_c] with zero fill,
// s_c = sign * (c / ceil(C/shift_size) - shift_size/2).  Inside ConvDropoutNormNonlin the shift is folded into
// the conv's operand fetch; this kernel only serves direct calls of the exported module.  One thread moves one
// VEC-byte piece of one (b, c, d) row of H*W elements.
template <typename VEC>
__global__ void __launch_bounds__(EW_THREADS) shift_depth_kernel(const VEC* __restrict__ x, VEC* __restrict__ y, int C, int D,
                                                                 long long row_vecs, int group, int half, int sign,
                                                                 long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row_vecs, v = i - r * row_vecs;      // r = (b*C + c)*D + d
    const int d = (int)(r % D);
    const long long bc = r / D;
    const int c = (int)(bc % C);
    const int s = sign * (c / group - half);
    const int ds = d - s;
    VEC val;
    memset(&val, 0, sizeof(VEC));
    if (ds >= 0 && ds < D) val = x[(bc * D + ds) * row_vecs + v];
    y[i] = val;
  }
}

__global__ void __launch_bounds__(EW_THREADS) add_inplace_kernel(uint4* __restrict__ y, const uint4* __restrict__ x,
                                                                 long long n16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16;
       i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    unpack8(y[i], a);
    unpack8(ld_nc_16(x + i), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    y[i] = pack8(a);
  }
}

inline int ew_blocks(long long n) {
  long long b = (n + EW_THREADS - 1) / EW_THREADS;
  const long long cap = (long long)e2e_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int e2e_nc_to_c8(const float* x, void* y, int32_t B, int32_t C, int64_t V, void* stream) {
  E2E_ARG(x && y && B > 0 && C > 0 && V > 0, "nc_to_c8: bad arguments");
  const long long total = (long long)B * ((C + 7) / 8) * V;
  nc_to_c8_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(x, (uint4*)y, B, C, V);
  E2E_LAUNCHED("nc_to_c8");
  return E2E_OK;
}

extern "C" int e2e_c8_to_nc(const void* x, float* y, int32_t B, int32_t C, int64_t V, void* stream) {
  E2E_ARG(x && y && B > 0 && C > 0 && V > 0, "c8_to_nc: bad arguments");
  const long long total = (long long)B * ((C + 7) / 8) * V;
  c8_to_nc_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const uint4*)x, y, B, C, V);
  E2E_LAUNCHED("c8_to_nc");
  return E2E_OK;
}

extern "C" int e2e_in_stats(const void* raw, int32_t B, int32_t Cb, int64_t V, float eps, float* partial,
                            int32_t nchunk, float* mean, float* rstd, void* stream) {
  E2E_ARG(raw && partial && mean && rstd && B > 0 && Cb > 0 && V > 0 && nchunk > 0, "in_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  in_stats_partial_kernel<<<dim3(nchunk, B * Cb), EW_THREADS, 0, st>>>((const uint4*)raw, V, nchunk, partial);
  E2E_LAUNCHED("in_stats_partial");
  const int n = B * Cb * 8 * 32;
  in_stats_final_kernel<<<(n + 127) / 128, 128, 0, st>>>(partial, B * Cb, nchunk, V, eps, mean, rstd);
  E2E_LAUNCHED("in_stats_final");
  return E2E_OK;
}

extern "C" int e2e_in_stats_final(const float* stats, int32_t n_slots, int32_t B, int32_t C, int64_t V, float eps,
                                  float* mean, float* rstd, void* stream) {
  E2E_ARG(stats && mean && rstd && n_slots > 0 && B > 0 && C > 0 && V > 0, "in_stats_final: bad arguments");
  in_stats_from_slots_kernel<<<dim3((C + 31) / 32, B), 1024, 0, (cudaStream_t)stream>>>(stats, n_slots, B, C, V, eps, mean, rstd);
  E2E_LAUNCHED("in_stats_final");
  return E2E_OK;
}

extern "C" int e2e_in_apply(const void* raw, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, float slope, int32_t B, int32_t Cb, int64_t V, void* out,
                            void* stream) {
  E2E_ARG(raw && mean && rstd && gamma && beta && out && B > 0 && Cb > 0 && V > 0, "in_apply: bad arguments");
  int nchunk = (int)((V + 4095) / 4096);
  const int want = (e2e_num_sms() * 8 + B * Cb - 1) / (B * Cb);
  if (nchunk > want) nchunk = want;
  if (nchunk < 1) nchunk = 1;
  in_apply_kernel<<<dim3(nchunk, B * Cb), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const uint4*)raw, mean, rstd, gamma, beta, slope, Cb, V, nchunk, (uint4*)out, nullptr, 0, B, 0.f, nullptr, nullptr,
      ew_order(1));
  E2E_LAUNCHED("in_apply");
  return E2E_OK;
}

// e2e_in_stats_final + e2e_in_apply in one launch: mean / rstd come from the conv epilogue's statistic slots
// (stats[n_slots][B][2][Cb*8]); they are also written to mean_out / rstd_out (B*Cb*8 each) for the backward.
extern "C" int e2e_in_apply_from_slots(const void* raw, const float* stats, int32_t n_slots, float eps, const float* gamma,
                                       const float* beta, float slope, int32_t B, int32_t Cb, int64_t V, float* mean_out,
                                       float* rstd_out, void* out, void* stream) {
  E2E_ARG(raw && stats && gamma && beta && out && mean_out && rstd_out && n_slots > 0 && B > 0 && Cb > 0 && V > 0,
          "in_apply_from_slots: bad arguments");
  int nchunk = (int)((V + 4095) / 4096);
  const int want = (e2e_num_sms() * 8 + B * Cb - 1) / (B * Cb);
  if (nchunk > want) nchunk = want;
  if (nchunk < 1) nchunk = 1;
  in_apply_kernel<<<dim3(nchunk, B * Cb), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const uint4*)raw, nullptr, nullptr, gamma, beta, slope, Cb, V, nchunk, (uint4*)out, stats, n_slots, B, eps, mean_out,
      rstd_out, ew_order(1));
  E2E_LAUNCHED("in_apply_from_slots");
  return E2E_OK;
}

extern "C" int e2e_in_bwd(const void* dy, const void* raw, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, float slope, int32_t B, int32_t Cb, int64_t V, float* partial,
                          int32_t nchunk, float* sums, void* draw, float* dgamma, float* dbeta, float* dbias,
                          void* stream) {
  E2E_ARG(dy && raw && mean && rstd && gamma && beta && partial && sums && draw && dgamma && dbeta,
          "in_bwd: null pointer");
  E2E_ARG(B > 0 && Cb > 0 && V > 0 && nchunk > 0, "in_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(nchunk, B * Cb);
  in_bwd_reduce_kernel<<<grid, EW_THREADS, 0, st>>>((const uint4*)dy, (const uint4*)raw, mean, rstd, gamma, beta,
                                                    slope, Cb, V, nchunk, partial, ew_order(1));
  E2E_LAUNCHED("in_bwd_reduce");
  float* partial2 = partial + (size_t)B * Cb * nchunk * 16;        // pass 2 reads ALL of pass 1's partials: separate region
  in_bwd_apply_kernel<<<grid, EW_THREADS, 0, st>>>((const uint4*)dy, (const uint4*)raw, mean, rstd, gamma, beta, partial,
                                                   sums, slope, Cb, V, nchunk, (uint4*)draw, partial2, ew_order(2));
  E2E_LAUNCHED("in_bwd_apply");
  in_bwd_param_kernel<<<(Cb * 8 * 32 + 127) / 128, 128, 0, st>>>(sums, partial2, B, Cb, nchunk, dgamma, dbeta, dbias);
  E2E_LAUNCHED("in_bwd_param");
  return E2E_OK;
}

extern "C" int64_t e2e_in_bwd_scratch_floats(int32_t B, int32_t Cb, int64_t V) {
  if (B <= 0 || Cb <= 0 || V <= 0) return 0;
  const int planes = B * Cb;
  const PlaneGeo pg = plane_geometry(planes, V);
  return (int64_t)planes * pg.Gs * 24 + planes + 64;               // partial (16) + partial2 (8) per CTA, counters
}

// dyp / argmax null: plain backward; otherwise the pooled gradient is folded in (window kd,kh,kw must be (1,2,2) or
// (2,2,2) and divide the grid).  scratch: e2e_in_bwd_scratch_floats() floats.
extern "C" int e2e_in_bwd_fused(const void* dy, const void* dyp, const uint8_t* argmax, const void* raw, const float* mean,
                                const float* rstd, const float* gamma, const float* beta, float slope, int32_t B, int32_t Cb,
                                int32_t D, int32_t H, int32_t W, int32_t kd, int32_t kh, int32_t kw, float* scratch,
                                float* sums, void* draw, float* dgamma, float* dbeta, float* dbias, void* stream) {
  E2E_ARG(raw && mean && rstd && gamma && beta && scratch && sums && draw && dgamma && dbeta, "in_bwd_fused: null pointer");
  E2E_ARG(B > 0 && Cb > 0 && D > 0 && H > 0 && W > 0, "in_bwd_fused: bad sizes");
  const bool pool = dyp != nullptr;
  E2E_ARG(pool || dy != nullptr, "in_bwd_fused: dy is required without a pooled gradient");
  E2E_ARG(!pool || argmax != nullptr, "in_bwd_fused: pooled gradient without its arg-max");
  const bool k122 = kd == 1 && kh == 2 && kw == 2, k222 = kd == 2 && kh == 2 && kw == 2;
  if (pool && !((k122 || k222) && D % kd == 0 && H % kh == 0 && W % kw == 0)) {
    e2e_set_error("in_bwd_fused: pool window (%d,%d,%d) is not instantiated for grid (%d,%d,%d)", kd, kh, kw, D, H, W);
    return E2E_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = B * Cb;
  const long long V = (long long)D * H * W;
  const PlaneGeo pg = plane_geometry(planes, V);
  float* partial = scratch;
  float* partial2 = scratch + (size_t)planes * pg.Gs * 16;
  int* counters = reinterpret_cast<int*>(scratch + (size_t)planes * pg.Gs * 24);
  E2E_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * planes, st));
  const PoolGeo g{D, H, W, pool ? kd : 1, pool ? kh : 1, pool ? kw : 1};
#define E2E_PLANE(POOL, A, Bq, Cq)                                                                                         \
  in_bwd_plane_kernel<POOL, A, Bq, Cq><<<pg.grid, EW_THREADS, 0, st>>>(                                                    \
      (const uint4*)dy, (const uint4*)dyp, (const uint2*)argmax, (const uint4*)raw, mean, rstd, gamma, beta, slope, Cb, g, V, \
      planes, pg.Gs, pg.ngroups, partial, partial2, counters, sums, (uint4*)draw)
  if (!pool) E2E_PLANE(false, 1, 1, 1);
  else if (k122) E2E_PLANE(true, 1, 2, 2);
  else E2E_PLANE(true, 2, 2, 2);
#undef E2E_PLANE
  E2E_LAUNCHED("in_bwd_plane");
  in_bwd_param_kernel<<<(Cb * 8 * 32 + 127) / 128, 128, 0, st>>>(sums, partial2, B, Cb, pg.Gs, dgamma, dbeta, dbias);
  E2E_LAUNCHED("in_bwd_param");
  return E2E_OK;
}

static int in_apply_pool_impl(const void* raw, const float* mean, const float* rstd, const float* stats, int n_slots, float eps,
                              float* mean_out, float* rstd_out, const float* gamma, const float* beta, float slope, int32_t B,
                              int32_t Cb, int32_t D, int32_t H, int32_t W, int32_t kd, int32_t kh, int32_t kw, void* out,
                              void* pooled, uint8_t* argmax, void* stream) {
  E2E_ARG(kd >= 1 && kh >= 1 && kw >= 1 && kd * kh * kw <= 8 && D % kd == 0 && H % kh == 0 && W % kw == 0,
          "in_apply_pool: window (%d,%d,%d) must divide (%d,%d,%d) and hold at most 8 voxels", kd, kh, kw, D, H, W);
  const PoolGeo g{D, H, W, kd, kh, kw};
  const long long Vo = (long long)D * H * W / (kd * kh * kw);
  int nchunk = (int)((Vo + 1023) / 1024);
  const int want = (e2e_num_sms() * 8 + B * Cb - 1) / (B * Cb);
  if (nchunk > want) nchunk = want;
  if (nchunk < 1) nchunk = 1;
  const dim3 grid(nchunk, B * Cb);
  cudaStream_t st = (cudaStream_t)stream;
  if (kd == 1 && kh == 2 && kw == 2)
    in_apply_pool_kernel<1, 2, 2><<<grid, EW_THREADS, 0, st>>>((const uint4*)raw, mean, rstd, gamma, beta, slope, Cb, g, nchunk,
                                                                 (uint4*)out, (uint4*)pooled, (uint2*)argmax, stats, n_slots, B, eps,
                                                                 mean_out, rstd_out, ew_order(1));
  else if (kd == 2 && kh == 2 && kw == 2)
    in_apply_pool_kernel<2, 2, 2><<<grid, EW_THREADS, 0, st>>>((const uint4*)raw, mean, rstd, gamma, beta, slope, Cb, g, nchunk,
                                                                 (uint4*)out, (uint4*)pooled, (uint2*)argmax, stats, n_slots, B, eps,
                                                                 mean_out, rstd_out, ew_order(1));
  else {
    e2e_set_error("in_apply_pool: window (%d,%d,%d) is not instantiated (use the separate max-pool kernel)", kd, kh, kw);
    return E2E_ERR_UNSUPPORTED;
  }
  E2E_LAUNCHED("in_apply_pool");
  return E2E_OK;
}

extern "C" int e2e_in_apply_pool(const void* raw, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, float slope, int32_t B, int32_t Cb, int32_t D, int32_t H, int32_t W,
                                 int32_t kd, int32_t kh, int32_t kw, void* out, void* pooled, uint8_t* argmax, void* stream) {
  E2E_ARG(raw && mean && rstd && gamma && beta && out && pooled && argmax && B > 0 && Cb > 0, "in_apply_pool: bad arguments");
  return in_apply_pool_impl(raw, mean, rstd, nullptr, 0, 0.f, nullptr, nullptr, gamma, beta, slope, B, Cb, D, H, W, kd, kh, kw,
                            out, pooled, argmax, stream);
}

// e2e_in_stats_final + e2e_in_apply_pool in one launch (see e2e_in_apply_from_slots)
extern "C" int e2e_in_apply_pool_from_slots(const void* raw, const float* stats, int32_t n_slots, float eps, const float* gamma,
                                            const float* beta, float slope, int32_t B, int32_t Cb, int32_t D, int32_t H,
                                            int32_t W, int32_t kd, int32_t kh, int32_t kw, float* mean_out, float* rstd_out,
                                            void* out, void* pooled, uint8_t* argmax, void* stream) {
  E2E_ARG(raw && stats && n_slots > 0 && mean_out && rstd_out && gamma && beta && out && pooled && argmax && B > 0 && Cb > 0,
          "in_apply_pool_from_slots: bad arguments");
  return in_apply_pool_impl(raw, nullptr, nullptr, stats, n_slots, eps, mean_out, rstd_out, gamma, beta, slope, B, Cb, D, H, W,
                            kd, kh, kw, out, pooled, argmax, stream);
}

extern "C" int e2e_in_bwd_pool(const void* dy, const void* dyp, const uint8_t* argmax, const void* raw, const float* mean,
                               const float* rstd, const float* gamma, const float* beta, float slope, int32_t B, int32_t Cb,
                               int32_t D, int32_t H, int32_t W, int32_t kd, int32_t kh, int32_t kw, float* partial,
                               int32_t nchunk, float* sums, void* draw, float* dgamma, float* dbeta, float* dbias,
                               void* stream) {
  E2E_ARG(dyp && argmax && raw && mean && rstd && gamma && beta && partial && sums && draw && dgamma && dbeta,
          "in_bwd_pool: null pointer");
  E2E_ARG(B > 0 && Cb > 0 && nchunk > 0 && kd * kh * kw <= 8 && D % kd == 0 && H % kh == 0 && W % kw == 0,
          "in_bwd_pool: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const PoolGeo g{D, H, W, kd, kh, kw};
  const dim3 grid(nchunk, B * Cb);
  const bool k122 = kd == 1 && kh == 2 && kw == 2, k222 = kd == 2 && kh == 2 && kw == 2;
  if (!k122 && !k222) {
    e2e_set_error("in_bwd_pool: window (%d,%d,%d) is not instantiated", kd, kh, kw);
    return E2E_ERR_UNSUPPORTED;
  }
#define E2E_BWD_POOL(PASS, DRAW, SUMS, POUT)                                                                                   \
  do {                                                                                                                     \
    if (k122)                                                                                                              \
      in_bwd_pool_kernel<PASS, 1, 2, 2><<<grid, EW_THREADS, 0, st>>>((const uint4*)dy, (const uint4*)dyp, (const uint2*)argmax, \
                                                                     (const uint4*)raw, mean, rstd, gamma, beta, partial, SUMS, slope, \
                                                                     Cb, g, nchunk, DRAW, POUT, ew_order(PASS ? 2 : 1));  \
    else                                                                                                                   \
      in_bwd_pool_kernel<PASS, 2, 2, 2><<<grid, EW_THREADS, 0, st>>>((const uint4*)dy, (const uint4*)dyp, (const uint2*)argmax, \
                                                                     (const uint4*)raw, mean, rstd, gamma, beta, partial, SUMS, slope, \
                                                                     Cb, g, nchunk, DRAW, POUT, ew_order(PASS ? 2 : 1));  \
  } while (0)
  float* partial2 = partial + (size_t)B * Cb * nchunk * 16;
  E2E_BWD_POOL(0, nullptr, nullptr, partial);
  E2E_LAUNCHED("in_bwd_pool_reduce");
  E2E_BWD_POOL(1, (uint4*)draw, sums, partial2);
  E2E_LAUNCHED("in_bwd_pool_apply");
#undef E2E_BWD_POOL
  in_bwd_param_kernel<<<(Cb * 8 * 32 + 127) / 128, 128, 0, st>>>(sums, partial2, B, Cb, nchunk, dgamma, dbeta, dbias);
  E2E_LAUNCHED("in_bwd_param");
  return E2E_OK;
}

extern "C" int e2e_maxpool_fwd(const void* x, void* y, uint8_t* argmax, int32_t BCb, int32_t D, int32_t H, int32_t W,
                               int32_t kd, int32_t kh, int32_t kw, void* stream) {
  E2E_ARG(x && y && argmax, "maxpool_fwd: null pointer");
  E2E_ARG(kd >= 1 && kh >= 1 && kw >= 1 && kd * kh * kw <= 255, "maxpool_fwd: bad window");
  const long long total = (long long)BCb * (D / kd) * (H / kh) * (W / kw);
  if (total <= 0) return E2E_OK;
  maxpool_fwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y,
                                                                                (uint2*)argmax, BCb, D, H, W, kd, kh, kw);
  E2E_LAUNCHED("maxpool_fwd");
  return E2E_OK;
}

extern "C" int e2e_maxpool_bwd(const void* dy, const uint8_t* argmax, void* dx, int32_t BCb, int32_t D, int32_t H,
                               int32_t W, int32_t kd, int32_t kh, int32_t kw, void* stream) {
  E2E_ARG(dy && argmax && dx, "maxpool_bwd: null pointer");
  const long long total = (long long)BCb * (D / kd) * (H / kh) * (W / kw);
  if (total <= 0) return E2E_OK;
  maxpool_bwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint2*)argmax,
                                                                                (uint4*)dx, BCb, D, H, W, kd, kh, kw);
  E2E_LAUNCHED("maxpool_bwd");
  return E2E_OK;
}

extern "C" int e2e_add_inplace(void* y, const void* x, int64_t n_elems, void* stream) {
  E2E_ARG(y && x && n_elems % 8 == 0, "add_inplace: bad arguments");
  if (n_elems == 0) return E2E_OK;
  add_inplace_kernel<<<ew_blocks(n_elems / 8), EW_THREADS, 0, (cudaStream_t)stream>>>((uint4*)y, (const uint4*)x,
                                                                                      n_elems / 8);
  E2E_LAUNCHED("add_inplace");
  return E2E_OK;
}

extern "C" int e2e_shift_depth(const void* x, void* y, int32_t elem_bytes, int32_t B, int32_t C, int32_t D, int64_t HW,
                               int32_t shift_size, int32_t sign, void* stream) {
  E2E_ARG(x && y && B > 0 && C > 0 && D > 0 && HW > 0 && shift_size > 0, "shift_depth: bad arguments");
  E2E_ARG(elem_bytes == 2 || elem_bytes == 4, "shift_depth: element size must be 2 or 4 bytes");
  E2E_ARG(sign == 1 || sign == -1, "shift_depth: sign must be +1 (forward) or -1 (gradient)");
  const int group = (C + shift_size - 1) / shift_size, half = shift_size / 2;
  const long long row_bytes = HW * elem_bytes;
  const long long rows = (long long)B * C * D;
  cudaStream_t st = (cudaStream_t)stream;
  const bool a16 = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
  if (row_bytes % 16 == 0 && a16) {
    const long long total = rows * (row_bytes / 16);
    shift_depth_kernel<uint4><<<ew_blocks(total), EW_THREADS, 0, st>>>((const uint4*)x, (uint4*)y, C, D, row_bytes / 16, group,
                                                                         half, sign, total);
  } else if (elem_bytes == 4) {
    const long long total = rows * HW;
    shift_depth_kernel<uint32_t><<<ew_blocks(total), EW_THREADS, 0, st>>>((const uint32_t*)x, (uint32_t*)y, C, D, HW, group,
                                                                            half, sign, total);
  } else {
    const long long total = rows * HW;
    shift_depth_kernel<uint16_t><<<ew_blocks(total), EW_THREADS, 0, st>>>((const uint16_t*)x, (uint16_t*)y, C, D, HW, group,
                                                                            half, sign, total);
  }
  E2E_LAUNCHED("shift_depth");
  return E2E_OK;
}
