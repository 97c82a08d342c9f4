// Fused optimizer step of the E2ENet training iteration: what the reference runs as
//   clip_grad_norm_(params, 12)            nnUNetTrainer_simple.py:560   (~150 ATen launches)
//   SGD(momentum .99, nesterov, wd 3e-5)   nnUNetTrainer_simple.py:367-371, optimizer.step() :561
//   Masking.apply_mask()                   core_channel.py:427-434       (w *= mask, momentum *= mask)
// becomes three multi-tensor launches over pointer tables: per-block sums of g^2 (no atomics), one block
// that turns them into the total norm and the clip coefficient ON THE DEVICE (no host read-back), and the
// update itself, which also applies the DSFF mask to the new weight and momentum.  Hyper-parameters live in a
// small device array so a captured CUDA graph follows learning-rate changes (poly-LR, :863-877).
// Optional dynamic loss scale (the reference's GradScaler, :553-562: scale -> unscale_ -> skip the step on inf / nan,
// halve the scale; double it after `growth_interval` clean steps) as device state scaler[5] = {scale, clean steps,
// growth interval, backoff factor, growth factor}: gradients are multiplied by 1 / scale on the fly and the state
// advances inside the coefficient kernel -- no host round trip, so it lives inside a captured graph.
#include "common.cuh"

namespace {

constexpr int OPT_THREADS = 256;

// grid (chunks, n_tensors): partial[t * chunks + c] = sum over this block's slice of (g * gscale)^2
__global__ void __launch_bounds__(OPT_THREADS) sgd_sumsq_kernel(const e2e_sgd_tensor_t* __restrict__ ts, const float* __restrict__ hp,
                                                                const float* __restrict__ scaler, float* __restrict__ partial) {
  const e2e_sgd_tensor_t t = ts[blockIdx.y];
  const float gs = scaler ? hp[4] / scaler[0] : hp[4];
  const long long n = t.numel;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  const long long n4 = (((uintptr_t)t.g & 15) == 0) ? (n >> 2) : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = reinterpret_cast<const float4*>(t.g)[i];
    const float a = g.x * gs, b = g.y * gs, c = g.z * gs, d = g.w * gs;
    acc += a * a + b * b + c * c + d * d;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = t.g[i] * gs;
    acc += a * a;
  }
  __shared__ float red[OPT_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
    partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

// one block: out[0] = total L2 norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (1 if max_norm <= 0),
// out[2] = 1 if the norm is inf / nan (the update is then skipped, like GradScaler.step does), out[3] = the gradient
// multiplier grad_scale / loss_scale this step used; then the loss-scale state advances (GradScaler.update)
__global__ void __launch_bounds__(1024) sgd_coef_kernel(const float* __restrict__ partial, int n, const float* __restrict__ hp,
                                                        float* __restrict__ scaler, float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];       // fixed order per thread
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float norm = (float)sqrt(tot);
    const float max_norm = hp[3];
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.f);
    const bool bad = !(norm == norm) || isinf(norm);
    out[0] = norm;
    out[1] = coef;
    out[2] = bad ? 1.f : 0.f;
    out[3] = scaler ? hp[4] / scaler[0] : hp[4];
    if (scaler) {
      float sc = scaler[0], clean = scaler[1];
      if (bad) {
        sc *= scaler[3];
        clean = 0.f;
      } else if (++clean >= scaler[2]) {
        sc *= scaler[4];
        clean = 0.f;
      }
      scaler[0] = sc;
      scaler[1] = clean;
    }
  }
}

// grid (chunks, n_tensors):  g = grad * gscale * coef + wd * p;  buf = momentum * buf + g;
//                            p -= lr * (nesterov ? g + momentum * buf : buf);  then p *= mask, buf *= mask
__global__ void __launch_bounds__(OPT_THREADS) sgd_update_kernel(const e2e_sgd_tensor_t* __restrict__ ts, const float* __restrict__ hp,
                                                                 const float* __restrict__ coef3, int nesterov) {
  if (coef3[2] != 0.f) return;                      // non-finite gradients: skip the step
  const e2e_sgd_tensor_t t = ts[blockIdx.y];
  const float lr = hp[0], mom = hp[1], wd = hp[2];
  const float gmul = coef3[3] * coef3[1];
  const long long n = t.numel;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.mom | (uintptr_t)t.mask) & 15) == 0);
  const long long n4 = vec ? (n >> 2) : 0;
  auto upd = [&](float& p, float g, float& b, float k) {
    g = __fmaf_rn(wd, p, g * gmul);
    b = __fmaf_rn(mom, b, g);
    const float d = nesterov ? __fmaf_rn(mom, b, g) : b;
    p = __fmaf_rn(-lr, d, p) * k;
    b *= k;
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(t.p)[i];
    const float4 g = reinterpret_cast<const float4*>(t.g)[i];
    float4 b = reinterpret_cast<float4*>(t.mom)[i];
    float4 k = make_float4(1.f, 1.f, 1.f, 1.f);
    if (t.mask) k = reinterpret_cast<const float4*>(t.mask)[i];
    upd(p.x, g.x, b.x, k.x); upd(p.y, g.y, b.y, k.y); upd(p.z, g.z, b.z, k.z); upd(p.w, g.w, b.w, k.w);
    reinterpret_cast<float4*>(t.p)[i] = p;
    reinterpret_cast<float4*>(t.mom)[i] = b;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float p = t.p[i], b = t.mom[i];
    upd(p, t.g[i], b, t.mask ? t.mask[i] : 1.f);
    t.p[i] = p;
    t.mom[i] = b;
  }
}

inline int opt_chunks(long long max_numel, int n_tensors) {
  long long c = (max_numel + OPT_THREADS * 16 - 1) / (OPT_THREADS * 16);
  const long long cap = ((long long)e2e_num_sms() * 16 + n_tensors - 1) / n_tensors;
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  return (int)c;
}

}  // namespace

extern "C" int e2e_sgd_partial_count(int32_t n_tensors, int64_t max_numel) {
  if (n_tensors <= 0 || max_numel <= 0) return 0;
  return opt_chunks(max_numel, n_tensors) * n_tensors;
}

extern "C" int e2e_sgd_clip_coef(const e2e_sgd_tensor_t* tensors, int32_t n_tensors, int64_t max_numel, const float* hyper,
                                 float* scaler, float* partial, float* norm_coef, void* stream) {
  E2E_ARG(tensors && hyper && partial && norm_coef && n_tensors > 0 && max_numel > 0, "sgd_clip_coef: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = opt_chunks(max_numel, n_tensors);
  sgd_sumsq_kernel<<<dim3(chunks, n_tensors), OPT_THREADS, 0, st>>>(tensors, hyper, scaler, partial);
  E2E_LAUNCHED("sgd_sumsq");
  sgd_coef_kernel<<<1, 1024, 0, st>>>(partial, chunks * n_tensors, hyper, scaler, norm_coef);
  E2E_LAUNCHED("sgd_coef");
  return E2E_OK;
}

extern "C" int e2e_sgd_update(const e2e_sgd_tensor_t* tensors, int32_t n_tensors, int64_t max_numel, const float* hyper,
                              const float* norm_coef, int32_t nesterov, void* stream) {
  E2E_ARG(tensors && hyper && norm_coef && n_tensors > 0 && max_numel > 0, "sgd_update: bad arguments");
  const int chunks = opt_chunks(max_numel, n_tensors);
  sgd_update_kernel<<<dim3(chunks, n_tensors), OPT_THREADS, 0, (cudaStream_t)stream>>>(tensors, hyper, norm_coef, nesterov);
  E2E_LAUNCHED("sgd_update");
  return E2E_OK;
}
