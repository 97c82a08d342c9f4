// Fused softmax statistics for the deep-supervision loss.
//
// The reference's DC_and_CE_loss (e2enet/training/loss_functions/dice_loss.py:302-359) runs, per
// deep-supervision output, softmax + one-hot scatter + tp/fp/fn reductions (SoftDiceLoss,
// dice_loss.py:155-190) and a second log-softmax + nll pass (crossentropy.py:4-11): ~10 passes
// over the fp32 logits.  Everything voxel-sized they compute is a function of three per-(b, c)
// sums and one scalar:
//     S_p[b,c] = sum_v p[b,c,v]      tp[b,c] = sum_v p[b,c,v] * [y[b,v] == c]
//     S_y[b,c] = #{v : y[b,v] == c}  ce_sum  = sum_{b,v} -log p[b, y[b,v], v]
// (fp = S_p - tp, fn = S_y - tp).  One pass produces them; the dice / CE formulas stay in torch on
// the tiny (B, C) tensors, and one more pass applies their gradients through the softmax:
//     dz[b,k,v] = p_k * (g_k - sum_j p_j g_j) + g_ce * (p_k - [y == k]),   g_j = gS_p[b,j] + gtp[b,j] * [y == j]
// HBM-bound: forward reads logits + target once, backward reads them once and writes dlogits once.
#include "common.cuh"

namespace {

// grid (chunks_per_b, B).  No atomics anywhere: every block writes its 3*C + 1 partial sums to
// partial[b][chunk][3*C + 1] (warps combined in a fixed order), a second kernel sums the chunks in a fixed order.
// A run-to-run difference of one ulp in these sums is amplified by the bf16 backward pass of the ~40 layers behind
// it into ~1 % differences of the deepest weight gradients, so the loss statistics must be bit-reproducible.
template <int NC>
__global__ void __launch_bounds__(256) softmax_stats_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                                int C, long long V, int chunks_per_b,
                                                                float* __restrict__ partial) {
  __shared__ float s_w[8][3 * NC + 1];
  const int b = blockIdx.y;
  const long long per = (V + chunks_per_b - 1) / chunks_per_b;
  const long long lo = (long long)blockIdx.x * per, hi = min(V, lo + per);
  const float* lb = logits + (long long)b * C * V;
  const float* tb = target + (long long)b * V;
  float sp[NC], tpv[NC], sy[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) { sp[c] = 0.f; tpv[c] = 0.f; sy[c] = 0.f; }
  float ce = 0.f;
  for (long long v = lo + threadIdx.x; v < hi; v += blockDim.x) {
    float z[NC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      z[c] = (c < C) ? lb[(long long)c * V + v] : -INFINITY;
      mx = fmaxf(mx, z[c]);
    }
    const int y = (int)tb[v];
    float sum = 0.f, zy = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float e = (c < C) ? expf(z[c] - mx) : 0.f;
      if (c == y) zy = z[c];
      z[c] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float p = z[c] * inv;
      sp[c] += p;
      if (c == y) { tpv[c] += p; sy[c] += 1.f; }
    }
    ce += logf(sum) + mx - zy;                     // -log softmax(z)[y]
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c >= C) break;
    const float a = warp_sum(sp[c]), t = warp_sum(tpv[c]), s = warp_sum(sy[c]);
    if (lane == 0) { s_w[warp][3 * c + 0] = a; s_w[warp][3 * c + 1] = t; s_w[warp][3 * c + 2] = s; }
  }
  ce = warp_sum(ce);
  if (lane == 0) s_w[warp][3 * NC] = ce;
  __syncthreads();
  float* out = partial + ((long long)b * chunks_per_b + blockIdx.x) * (3 * C + 1);
  for (int i = threadIdx.x; i < 3 * C + 1; i += blockDim.x) {
    const int k = i < 3 * C ? i : 3 * NC;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += s_w[w][k];
    out[i] = acc;
  }
}

// one warp per (b, value): lanes stride over the chunks, fp64, fixed order; stats / ce_sum are ADDED to (once,
// by one thread per value: the caller zeroes them, and different launches never overlap on a stream)
__global__ void softmax_stats_final_kernel(const float* __restrict__ partial, int B, int C, int chunks_per_b,
                                           float* __restrict__ stats, float* __restrict__ ce_sum) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nv = 3 * C + 1;
  if (i >= nv) return;
  double ce_all = 0.0;
  for (int b = 0; b < B; ++b) {
    double s = 0.0;
    for (int k = lane; k < chunks_per_b; k += 32) s += (double)partial[((long long)b * chunks_per_b + k) * nv + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (i < 3 * C) {
      if (lane == 0) stats[(long long)b * C * 3 + i] += (float)s;
    } else {
      ce_all += s;
    }
  }
  if (i == 3 * C && lane == 0) ce_sum[0] += (float)ce_all;
}

// DC_and_CE_loss from the statistics (dice_loss.py:155-190, 302-359; crossentropy.py:4-11), one block:
//   dc[b,c] = (2 tp + smooth) / (2 tp + fp + fn + smooth + 1e-8),  fp = S_p - tp, fn = S_y - tp  =>  den = S_p + S_y + smooth + 1e-8
//   loss = weight_ce * ce_sum / n_vox - weight_dice * mean_{b, c >= c0} dc        (batch_dice: tp, S_p, S_y summed over b first)
// and the coefficients the backward pass needs: gsp = d loss / d S_p, gtp = d loss / d tp, gce = d loss / d ce_sum.
// Replaces ~45 (B, C)-sized ATen launches per deep-supervision output (forward + backward) by this one.
__global__ void __launch_bounds__(256) dc_ce_from_stats_kernel(const float* __restrict__ stats, const float* __restrict__ ce_sum,
                                                               int B, int C, float n_vox, float smooth, int do_bg, int batch_dice,
                                                               float weight_ce, float weight_dice, float* __restrict__ loss,
                                                               float* __restrict__ gsp, float* __restrict__ gtp,
                                                               float* __restrict__ gce) {
  __shared__ float s_dc[256];
  const int c0 = do_bg ? 0 : 1;
  const int n = B * C;
  float acc = 0.f;                                   // sum of dc over this thread's (b, c) entries, fixed assignment
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = i / C, c = i - b * C;
    float tp, sp, sy;
    if (batch_dice) {
      tp = sp = sy = 0.f;
      for (int bb = 0; bb < B; ++bb) {
        const float* q = stats + ((long long)bb * C + c) * 3;
        sp += q[0]; tp += q[1]; sy += q[2];
      }
    } else {
      const float* q = stats + (long long)i * 3;
      sp = q[0]; tp = q[1]; sy = q[2];
    }
    const float num = 2.f * tp + smooth, den = sp + sy + smooth + 1e-8f;
    const float cnt = batch_dice ? (float)(C - c0) : (float)(B * (C - c0));
    const bool on = c >= c0;
    gtp[i] = on ? -weight_dice * (2.f / den) / cnt : 0.f;
    gsp[i] = on ? weight_dice * (num / (den * den)) / cnt : 0.f;
    if (on && (!batch_dice || b == 0)) acc += (num / den) / cnt;
  }
  s_dc[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float dc = 0.f;
    for (int i = 0; i < (int)blockDim.x; ++i) dc += s_dc[i];       // fixed order
    loss[0] = weight_ce * (ce_sum[0] / n_vox) - weight_dice * dc;
    gce[0] = weight_ce / n_vox;
  }
}

template <int NC>
__global__ void __launch_bounds__(256) softmax_stats_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                                const float* __restrict__ gsp, const float* __restrict__ gtp,
                                                                const float* __restrict__ gce, const float* __restrict__ gscale,
                                                                int C, long long V,
                                                                int chunks_per_b, float* __restrict__ dlogits) {
  const int b = blockIdx.y;
  const long long per = (V + chunks_per_b - 1) / chunks_per_b;
  const long long lo = (long long)blockIdx.x * per, hi = min(V, lo + per);
  const float* lb = logits + (long long)b * C * V;
  const float* tb = target + (long long)b * V;
  float* db = dlogits + (long long)b * C * V;
  float a[NC], t[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    a[c] = (c < C) ? gsp[b * C + c] : 0.f;
    t[c] = (c < C) ? gtp[b * C + c] : 0.f;
  }
  float gc = gce[0];
  if (gscale) {                                      // upstream gradient of the (scalar) loss, read on the device
    const float gs = gscale[0];
#pragma unroll
    for (int c = 0; c < NC; ++c) { a[c] *= gs; t[c] *= gs; }
    gc *= gs;
  }
  for (long long v = lo + threadIdx.x; v < hi; v += blockDim.x) {
    float z[NC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      z[c] = (c < C) ? lb[(long long)c * V + v] : -INFINITY;
      mx = fmaxf(mx, z[c]);
    }
    const int y = (int)tb[v];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      z[c] = (c < C) ? expf(z[c] - mx) : 0.f;
      sum += z[c];
    }
    const float inv = 1.0f / sum;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      z[c] *= inv;                                           // p_c
      dot += z[c] * (a[c] + (c == y ? t[c] : 0.f));
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c >= C) break;
      const float g = a[c] + (c == y ? t[c] : 0.f);
      db[(long long)c * V + v] = z[c] * (g - dot) + gc * (z[c] - (c == y ? 1.f : 0.f));
    }
  }
}

inline int chunks_for(long long V, int B) {
  long long want = ((long long)e2e_num_sms() * 8 + B - 1) / B;
  long long mx = (V + 1023) / 1024;
  if (want > mx) want = mx;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace

extern "C" int e2e_softmax_stats_partial_count(int32_t B, int32_t C, int64_t V) {
  if (B <= 0 || C <= 0 || V <= 0) return 0;
  return chunks_for(V, B) * B * (3 * C + 1);
}

extern "C" int e2e_softmax_stats_fwd(const float* logits, const float* target, int32_t B, int32_t C, int64_t V,
                                     float* partial, float* stats, float* ce_sum, void* stream) {
  E2E_ARG(logits && target && partial && stats && ce_sum && B > 0 && C > 0 && V > 0, "softmax_stats_fwd: bad arguments");
  E2E_ARG(C <= 32, "softmax_stats_fwd: at most 32 classes (got %d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(chunks_for(V, B), B);
  if (C <= 4) softmax_stats_fwd_kernel<4><<<grid, 256, 0, st>>>(logits, target, C, V, grid.x, partial);
  else if (C <= 16) softmax_stats_fwd_kernel<16><<<grid, 256, 0, st>>>(logits, target, C, V, grid.x, partial);
  else softmax_stats_fwd_kernel<32><<<grid, 256, 0, st>>>(logits, target, C, V, grid.x, partial);
  E2E_LAUNCHED("softmax_stats_fwd");
  const int nv = 3 * C + 1;
  softmax_stats_final_kernel<<<(nv * 32 + 127) / 128, 128, 0, st>>>(partial, B, C, grid.x, stats, ce_sum);
  E2E_LAUNCHED("softmax_stats_final");
  return E2E_OK;
}

extern "C" int e2e_dc_ce_from_stats(const float* stats, const float* ce_sum, int32_t B, int32_t C, int64_t n_vox, float smooth,
                                    int32_t do_bg, int32_t batch_dice, float weight_ce, float weight_dice, float* loss,
                                    float* gsp, float* gtp, float* gce, void* stream) {
  E2E_ARG(stats && ce_sum && loss && gsp && gtp && gce && B > 0 && C > 0 && n_vox > 0, "dc_ce_from_stats: bad arguments");
  E2E_ARG(C - (do_bg ? 0 : 1) > 0, "dc_ce_from_stats: no foreground class");
  dc_ce_from_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(stats, ce_sum, B, C, (float)n_vox, smooth, do_bg, batch_dice,
                                                               weight_ce, weight_dice, loss, gsp, gtp, gce);
  E2E_LAUNCHED("dc_ce_from_stats");
  return E2E_OK;
}

extern "C" int e2e_softmax_stats_bwd(const float* logits, const float* target, const float* gsp, const float* gtp,
                                     const float* gce, const float* gscale, int32_t B, int32_t C, int64_t V, float* dlogits,
                                     void* stream) {
  E2E_ARG(logits && target && gsp && gtp && gce && dlogits && B > 0 && C > 0 && V > 0, "softmax_stats_bwd: bad arguments");
  E2E_ARG(C <= 32, "softmax_stats_bwd: at most 32 classes (got %d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(chunks_for(V, B), B);
  if (C <= 4) softmax_stats_bwd_kernel<4><<<grid, 256, 0, st>>>(logits, target, gsp, gtp, gce, gscale, C, V, grid.x, dlogits);
  else if (C <= 16) softmax_stats_bwd_kernel<16><<<grid, 256, 0, st>>>(logits, target, gsp, gtp, gce, gscale, C, V, grid.x, dlogits);
  else softmax_stats_bwd_kernel<32><<<grid, 256, 0, st>>>(logits, target, gsp, gtp, gce, gscale, C, V, grid.x, dlogits);
  E2E_LAUNCHED("softmax_stats_bwd");
  return E2E_OK;
}
