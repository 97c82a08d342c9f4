// Export path: resampling of the predicted class probabilities to the original voxel grid,
// fused with the arg-max that turns them into labels.
//
// The reference (e2enet/inference/segmentation_export.py:27-160 -> preprocessing.resample_data_or_seg,
// preprocessing.py:113-202) resizes every class volume on one CPU core with skimage.transform.resize(order,
// mode='edge', anti_aliasing=False) -- or, for anisotropic data, slice by slice in-plane followed by
// scipy.ndimage.map_coordinates(order_z, mode='nearest') along the coarse axis -- materialises the resampled
// (C, X', Y', Z') fp32 array and then takes argmax(0).  Both resamplers use the pixel-centre map
//     in = (out + 0.5) * (n_in / n_out) - 0.5            (preprocessing.py:167-174; skimage >= 0.19 / ndi.zoom grid_mode)
// with edge clamping, and for order <= 1 they are separable: per axis either linear (two taps) or nearest
// (floor(in + 0.5)).  One thread per OUTPUT voxel evaluates the <= 8 taps of every class, optionally stores the
// resampled probabilities, and keeps the first maximum (np.argmax tie rule): the resampled fp32 volume never has to
// exist when only labels are exported.  HBM-bound on the output side; the input taps hit L2.
#include "common.cuh"

namespace {

struct AxisMap { int n_in, n_out, mode; double scale; };      // mode 0: nearest (order 0), 1: linear (order 1)

__device__ __forceinline__ void axis_taps(const AxisMap& a, int o, int& i0, int& i1, float& f) {
  const double c = ((double)o + 0.5) * a.scale - 0.5;
  if (a.mode == 0) {
    int i = (int)floor(c + 0.5);
    i = i < 0 ? 0 : (i > a.n_in - 1 ? a.n_in - 1 : i);
    i0 = i1 = i;
    f = 0.f;
  } else {
    const double fl = floor(c);
    int lo = (int)fl, hi = lo + 1;
    f = (float)(c - fl);
    lo = lo < 0 ? 0 : (lo > a.n_in - 1 ? a.n_in - 1 : lo);
    hi = hi < 0 ? 0 : (hi > a.n_in - 1 ? a.n_in - 1 : hi);
    i0 = lo; i1 = hi;
  }
}

__global__ void __launch_bounds__(256) resample_argmax_kernel(const float* __restrict__ probs, int C, AxisMap ax, AxisMap ay, AxisMap az,
                                                              float* __restrict__ out_probs, uint8_t* __restrict__ labels) {
  const long long Vo = (long long)ax.n_out * ay.n_out * az.n_out;
  const long long Vi = (long long)ax.n_in * ay.n_in * az.n_in;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < Vo; i += (long long)gridDim.x * blockDim.x) {
    const int oz = (int)(i % az.n_out);
    const long long t = i / az.n_out;
    const int oy = (int)(t % ay.n_out), ox = (int)(t / ay.n_out);
    int x0, x1, y0, y1, z0, z1;
    float fx, fy, fz;
    axis_taps(ax, ox, x0, x1, fx);
    axis_taps(ay, oy, y0, y1, fy);
    axis_taps(az, oz, z0, z1, fz);
    const long long r00 = ((long long)x0 * ay.n_in + y0) * az.n_in, r01 = ((long long)x0 * ay.n_in + y1) * az.n_in;
    const long long r10 = ((long long)x1 * ay.n_in + y0) * az.n_in, r11 = ((long long)x1 * ay.n_in + y1) * az.n_in;
    float best = -INFINITY;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
      const float* p = probs + (long long)c * Vi;
      const float v000 = p[r00 + z0], v001 = p[r00 + z1], v010 = p[r01 + z0], v011 = p[r01 + z1];
      const float v100 = p[r10 + z0], v101 = p[r10 + z1], v110 = p[r11 + z0], v111 = p[r11 + z1];
      const float a00 = v000 + fz * (v001 - v000), a01 = v010 + fz * (v011 - v010);
      const float a10 = v100 + fz * (v101 - v100), a11 = v110 + fz * (v111 - v110);
      const float b0 = a00 + fy * (a01 - a00), b1 = a10 + fy * (a11 - a10);
      const float v = b0 + fx * (b1 - b0);
      if (out_probs) out_probs[(long long)c * Vo + i] = v;
      if (v > best) { best = v; bi = c; }
    }
    if (labels) labels[i] = (uint8_t)bi;
  }
}

}  // namespace

extern "C" int e2e_resample_argmax(const float* probs, int32_t C, int32_t X, int32_t Y, int32_t Z, int32_t Xo, int32_t Yo,
                                   int32_t Zo, int32_t mode_x, int32_t mode_y, int32_t mode_z, float* out_probs,
                                   uint8_t* labels, void* stream) {
  E2E_ARG(probs && (out_probs || labels), "resample_argmax: null pointer");
  E2E_ARG(C >= 1 && C <= 255 && X > 0 && Y > 0 && Z > 0 && Xo > 0 && Yo > 0 && Zo > 0, "resample_argmax: bad sizes");
  E2E_ARG((mode_x | mode_y | mode_z) >= 0 && mode_x <= 1 && mode_y <= 1 && mode_z <= 1,
          "resample_argmax: interpolation orders 0 (nearest) and 1 (linear) are implemented (got %d,%d,%d)", mode_x, mode_y, mode_z);
  const AxisMap ax{X, Xo, mode_x, (double)X / (double)Xo}, ay{Y, Yo, mode_y, (double)Y / (double)Yo},
      az{Z, Zo, mode_z, (double)Z / (double)Zo};
  const long long Vo = (long long)Xo * Yo * Zo;
  long long blocks = (Vo + 255) / 256;
  const long long cap = (long long)e2e_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  resample_argmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(probs, C, ax, ay, az, out_probs, labels);
  E2E_LAUNCHED("resample_argmax");
  return E2E_OK;
}
