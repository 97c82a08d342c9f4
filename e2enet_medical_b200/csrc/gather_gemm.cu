// Plan-driven gather GEMM on the C8 activation layout (mma.sync path).
//
// One kernel family covers every contraction of the E2ENet hot path that is not (yet)
// served by the tcgen05 kernel: depth-shifted (1,3,3) convs with any stride, their data
// gradients (per shift-group / per stride-parity variants built by the host plan),
// kernel==stride transposed convs and their data gradients, and the 1x1x1 seg heads.
// The reference's shift (unetpp_d.py:45-59) and torch.cat (unetpp_d.py:453-478) never
// materialise: a channel entry names (source tensor, 8-channel block, spatial offset).
//
// A-operand slabs are [128 voxels][8 ch] (16 B rows) gathered with zero-filling cp.async;
// B-operand slabs are [N][8 ch] rows of the pre-packed bf16 weights; both feed ldmatrix
// directly (8 rows x 16 B = one 8x8 fragment, conflict free).
#include "common.cuh"

namespace {

constexpr int BM = 128;      // voxels per CTA tile
constexpr int STAGES = 4;
constexpr int THREADS = 256;

struct RowInfo {             // per tile row: decoded iteration-grid coordinates
  int b, od, oh, ow;
};

template <int WN>            // n8 tiles per warp; CTA N tile = 16 * WN
__global__ void __launch_bounds__(THREADS, (WN <= 6) ? 2 : 1) gather_gemm_kernel(const __grid_constant__ e2e_gemm_t p) {
  constexpr int NT = 16 * WN;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                                  // STAGES * 4096
  uint8_t* sB = smem + STAGES * 4096;                  // STAGES * NT * 32
  __shared__ RowInfo rows[BM];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * NT;

  if (tid < BM) {
    long long m = m0 + tid;
    RowInfo ri;
    if (m < M) {
      int ow = (int)(m % p.Wo);
      long long t = m / p.Wo;
      int oh = (int)(t % p.Ho);
      t /= p.Ho;
      int od = (int)(t % p.Do);
      ri.b = (int)(t / p.Do);
      ri.od = od; ri.oh = oh; ri.ow = ow;
    } else {
      ri.b = -1; ri.od = 0; ri.oh = 0; ri.ow = 0;
    }
    rows[tid] = ri;
  }
  __syncthreads();

  const int r = tid & (BM - 1), half = tid >> 7;
  const RowInfo my = rows[r];
  const int ibd = my.od * p.isd + p.ivd, ibh = my.oh * p.ish + p.ivh, ibw = my.ow * p.isw + p.ivw;
  const int nk = (p.n_cent >> 1) * p.n_taps;
  const act16* wp = reinterpret_cast<const act16*>(p.wpacked);

  auto load_stage = [&](int ks, int st) {
    const int pr = ks / p.n_taps, t = ks - pr * p.n_taps;
    const e2e_centry_t ce = p.cents[2 * pr + half];
    const e2e_tap_t tp = p.taps[t];
    const int d = ibd + ce.dd + tp.dd, h = ibh + ce.dh + tp.dh, w = ibw + ce.dw + tp.dw;
    const bool ok = (my.b >= 0) && (unsigned)d < (unsigned)p.Di && (unsigned)h < (unsigned)p.Hi &&
                    (unsigned)w < (unsigned)p.Wi;
    const act16* sp = reinterpret_cast<const act16*>(p.src[ce.src]);
    size_t off = 0;
    if (ok) off = ((((size_t)my.b * p.src_cb[ce.src] + ce.blk) * p.Di + d) * p.Hi + h) * (size_t)p.Wi + w;
    cp_async_16(smem_u32(sA + st * 4096 + half * 2048 + r * 16), sp + off * 8, ok ? 16 : 0);
    if (tid < 2 * NT) {
      const int hb = tid / NT, nr = tid - hb * NT;
      const act16* ws = wp + (((size_t)ks * 2 + hb) * p.Npad + n0 + nr) * 8;
      cp_async_16(smem_u32(sB + st * (NT * 32) + tid * 16), ws, 16);
    }
  };

  float acc[2][WN][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int ks = 0; ks < nk; ++ks) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nx = ks + STAGES - 1;
      if (nx < nk) load_stage(nx, nx % STAGES);
      cp_async_commit();
    }
    const int st = ks % STAGES;
    const uint32_t a_base = smem_u32(sA + st * 4096);
    const uint32_t b_base = smem_u32(sB + st * (NT * 32));
    uint32_t a[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int row = wm * 32 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int slab = lane >> 4;
      ldmatrix_x4(a[mi][0], a[mi][1], a[mi][2], a[mi][3], a_base + slab * 2048 + row * 16);
    }
    uint32_t b[WN][2];
#pragma unroll
    for (int nj = 0; nj + 1 < WN; nj += 2) {
      const int mi = lane >> 3;
      const int nt = nj + (mi >> 1), slab = mi & 1;
      const int nrow = wn * (NT / 2) + nt * 8 + (lane & 7);
      ldmatrix_x4(b[nj][0], b[nj][1], b[nj + 1][0], b[nj + 1][1], b_base + slab * (NT * 16) + nrow * 16);
    }
    if (WN & 1) {
      const int slab = (lane >> 3) & 1;
      const int nrow = wn * (NT / 2) + (WN - 1) * 8 + (lane & 7);
      ldmatrix_x2(b[WN - 1][0], b[WN - 1][1], b_base + slab * (NT * 16) + nrow * 16);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nj = 0; nj < WN; ++nj) mma_act_16816(acc[mi][nj], a[mi], b[nj][0], b[nj][1]);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---------------------------------------------------------------- epilogue
  const int g = lane >> 2, t4 = lane & 3;
  if (p.out_mode == 0) {
    // stage as [colblk][row][8] bf16
    uint8_t* sO = smem;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nj = 0; nj < WN; ++nj) {
        const int cb = wn * WN + nj;
        const int row = wm * 32 + mi * 16 + g;
        *reinterpret_cast<uint32_t*>(sO + (cb * BM + row) * 16 + t4 * 4) = pack_act2(acc[mi][nj][0], acc[mi][nj][1]);
        *reinterpret_cast<uint32_t*>(sO + (cb * BM + row + 8) * 16 + t4 * 4) = pack_act2(acc[mi][nj][2], acc[mi][nj][3]);
      }
    __syncthreads();
    for (int idx = tid; idx < (NT / 8) * BM; idx += THREADS) {
      const int cb = idx / BM, row = idx - cb * BM;
      const RowInfo ri = rows[row];
      if (ri.b < 0) continue;
      const e2e_colblk_t col = p.cols[n0 / 8 + cb];
      if (col.dst < 0 || col.chmask == 0) continue;
      const int d = ri.od * p.osd + col.od, h = ri.oh * p.osh + col.oh, w = ri.ow * p.osw + col.ow;
      if ((unsigned)d >= (unsigned)p.Dd || (unsigned)h >= (unsigned)p.Hd || (unsigned)w >= (unsigned)p.Wd) continue;
      act16* dp = reinterpret_cast<act16*>(p.dst[col.dst]) +
                 (((((size_t)ri.b * p.dst_cb[col.dst] + col.blk) * p.Dd + d) * p.Hd + h) * (size_t)p.Wd + w) * 8;
      const uint4 v = *reinterpret_cast<const uint4*>(sO + (size_t)idx * 16);
      if (col.chmask == 0xff) {
        *reinterpret_cast<uint4*>(dp) = v;
      } else {
        const act16* vv = reinterpret_cast<const act16*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (col.chmask & (1 << j)) dp[j] = vv[j];
      }
    }
  } else {
    // fp32 NCDHW: stage as [col][row] fp32
    float* sO = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nj = 0; nj < WN; ++nj) {
        const int c = wn * (NT / 2) + nj * 8 + t4 * 2;
        const int row = wm * 32 + mi * 16 + g;
        sO[c * BM + row] = acc[mi][nj][0];
        sO[(c + 1) * BM + row] = acc[mi][nj][1];
        sO[c * BM + row + 8] = acc[mi][nj][2];
        sO[(c + 1) * BM + row + 8] = acc[mi][nj][3];
      }
    __syncthreads();
    const int C = p.dst_cb[0];
    float* out = reinterpret_cast<float*>(p.dst[0]);
    const size_t V = (size_t)p.Dd * p.Hd * p.Wd;
    for (int idx = tid; idx < NT * BM; idx += THREADS) {
      const int c = idx / BM, row = idx - c * BM;
      const RowInfo ri = rows[row];
      if (ri.b < 0 || n0 + c >= C) continue;
      out[((size_t)ri.b * C + n0 + c) * V + ((size_t)ri.od * p.Hd + ri.oh) * p.Wd + ri.ow] = sO[idx];
    }
  }
}

// ------------------------------------------------------------------------------------ wgrad
constexpr int KT = 64;        // voxels per K tile
constexpr int WSTAGES = 3;
constexpr int SG = 16;        // (entry, tap) slabs per CTA: 2 per warp

template <int WM>             // m16 tiles (rows of dW = output channels) per CTA; MT = 16 * WM
__global__ void __launch_bounds__(THREADS, 1) gather_wgrad_kernel(const __grid_constant__ e2e_wgrad_t p, int tiles_per_split) {
  constexpr int MT = 16 * WM;
  constexpr int G_BYTES = (MT / 8) * KT * 16;          // grad slabs  [nblk][v][8]
  constexpr int A_BYTES = SG * KT * 16;                // input slabs [slab][v][8]
  constexpr int STAGE_BYTES = G_BYTES + A_BYTES;
  extern __shared__ __align__(128) uint8_t smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.n_cent * p.n_taps;                   // slab s = (pair*n_taps + tap)*2 + half
  const int s0 = blockIdx.x * SG;
  const int nrow0 = blockIdx.y * MT;
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  const long long Vo = (long long)p.Do * p.Ho * p.Wo;
  const int ntile = (int)((M + KT - 1) / KT);
  const int tile_lo = blockIdx.z * tiles_per_split;
  const int tile_hi = min(ntile, tile_lo + tiles_per_split);
  if (tile_lo >= tile_hi) return;

  const int r = tid & (KT - 1), q = tid >> 6;          // row r, lane group q in 0..3
  const act16* gp = reinterpret_cast<const act16*>(p.grad);

  // my 4 slabs (q, q+4, q+8, q+12): decode (cent, tap) once
  int my_dd[4], my_dh[4], my_dw[4], my_src[4], my_blk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = s0 + q + 4 * i;
    if (s < S) {
      const int ks = s >> 1, hf = s & 1;
      const int pr = ks / p.n_taps, t = ks - pr * p.n_taps;
      const e2e_centry_t ce = p.cents[2 * pr + hf];
      const e2e_tap_t tp = p.taps[t];
      my_dd[i] = ce.dd + tp.dd; my_dh[i] = ce.dh + tp.dh; my_dw[i] = ce.dw + tp.dw;
      my_src[i] = ce.src; my_blk[i] = ce.blk;
    } else {
      my_src[i] = -1; my_dd[i] = my_dh[i] = my_dw[i] = my_blk[i] = 0;
    }
  }

  auto load_tile = [&](int tile, int st) {
    uint8_t* sG = smem + st * STAGE_BYTES;
    uint8_t* sA = sG + G_BYTES;
    const long long m = (long long)tile * KT + r;
    const bool rv = m < M;
    int b = 0, od = 0, oh = 0, ow = 0;
    long long v = 0;
    if (rv) {
      b = (int)(m / Vo);
      v = m - (long long)b * Vo;
      ow = (int)(v % p.Wo);
      long long t = v / p.Wo;
      oh = (int)(t % p.Ho);
      od = (int)(t / p.Ho);
    }
    // grad slabs
    for (int nb = q; nb < MT / 8; nb += 4) {
      const int gnb = nrow0 / 8 + nb;
      const bool ok = rv && gnb < p.grad_cb;
      const act16* sp = gp + (ok ? (((size_t)b * p.grad_cb + gnb) * Vo + v) * 8 : 0);
      cp_async_16(smem_u32(sG + (nb * KT + r) * 16), sp, ok ? 16 : 0);
    }
    const int ibd = od * p.isd + p.ivd, ibh = oh * p.ish + p.ivh, ibw = ow * p.isw + p.ivw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int sl = q + 4 * i;
      const int d = ibd + my_dd[i], h = ibh + my_dh[i], w = ibw + my_dw[i];
      const bool ok = rv && my_src[i] >= 0 && (unsigned)d < (unsigned)p.Di && (unsigned)h < (unsigned)p.Hi &&
                      (unsigned)w < (unsigned)p.Wi;
      const int si = ok ? my_src[i] : 0;
      const act16* sp = reinterpret_cast<const act16*>(p.src[si]);
      size_t off = 0;
      if (ok) off = ((((size_t)b * p.src_cb[si] + my_blk[i]) * p.Di + d) * p.Hi + h) * (size_t)p.Wi + w;
      cp_async_16(smem_u32(sA + (sl * KT + r) * 16), sp + off * 8, ok ? 16 : 0);
    }
  };

  float acc[WM][2][4];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  const int nt = tile_hi - tile_lo;
#pragma unroll
  for (int s = 0; s < WSTAGES - 1; ++s) {
    if (s < nt) load_tile(tile_lo + s, s);
    cp_async_commit();
  }
  for (int it = 0; it < nt; ++it) {
    cp_async_wait<WSTAGES - 2>();
    __syncthreads();
    {
      const int nx = it + WSTAGES - 1;
      if (nx < nt) load_tile(tile_lo + nx, nx % WSTAGES);
      cp_async_commit();
    }
    const int st = it % WSTAGES;
    const uint32_t g_base = smem_u32(smem + st * STAGE_BYTES);
    const uint32_t a_base = g_base + G_BYTES;
#pragma unroll
    for (int k16 = 0; k16 < KT / 16; ++k16) {
      // B operand: my two slabs, 16 voxels -> b[slab][0..1]
      uint32_t b[2][2];
      {
        const int mi = lane >> 3;
        const int sl = warp * 2 + (mi >> 1);
        const int vrow = k16 * 16 + (mi & 1) * 8 + (lane & 7);
        ldmatrix_x4_t(b[0][0], b[0][1], b[1][0], b[1][1], a_base + (sl * KT + vrow) * 16);
      }
#pragma unroll
      for (int mt = 0; mt < WM; ++mt) {
        // A operand = grad^T: matrices (nblk 2mt, v lo), (nblk 2mt+1, v lo), (nblk 2mt, v hi), (nblk 2mt+1, v hi)
        uint32_t a[4];
        const int mi = lane >> 3;
        const int nb = 2 * mt + (mi & 1);
        const int vrow = k16 * 16 + (mi >> 1) * 8 + (lane & 7);
        ldmatrix_x4_t(a[0], a[1], a[2], a[3], g_base + (nb * KT + vrow) * 16);
        mma_act_16816(acc[mt][0], a, b[0][0], b[0][1]);
        mma_act_16816(acc[mt][1], a, b[1][0], b[1][1]);
      }
    }
  }
  cp_async_wait<0>();

  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int sj = 0; sj < 2; ++sj) {
    const int s = s0 + warp * 2 + sj;
    if (s >= S) continue;
#pragma unroll
    for (int mt = 0; mt < WM; ++mt) {
      const int n = nrow0 + mt * 16 + g;
      float* base = p.dwp + ((size_t)s * p.Npad) * 8;
      if (n < p.Npad) {
        atomicAdd(base + (size_t)n * 8 + t4 * 2, acc[mt][sj][0]);
        atomicAdd(base + (size_t)n * 8 + t4 * 2 + 1, acc[mt][sj][1]);
      }
      if (n + 8 < p.Npad) {
        atomicAdd(base + (size_t)(n + 8) * 8 + t4 * 2, acc[mt][sj][2]);
        atomicAdd(base + (size_t)(n + 8) * 8 + t4 * 2 + 1, acc[mt][sj][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ pack / unpack
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ mask,
                                    const int32_t* __restrict__ rowoff, const int32_t* __restrict__ centoff,
                                    const int32_t* __restrict__ tapoff, const int32_t* __restrict__ emask,
                                    const int32_t* __restrict__ rclass, int n_cent, int n_taps, int Npad,
                                    act16* __restrict__ out) {
  // one thread per (ks, half, n): writes 8 bf16 (16 B)
  const long long total = (long long)n_cent * n_taps * Npad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Npad);
    const long long sl = i / Npad;                 // slab = ks*2 + half
    const int hf = (int)(sl & 1);
    const int ks = (int)(sl >> 1);
    const int pr = ks / n_taps, t = ks - pr * n_taps;
    const int e = 2 * pr + hf;
    int ro = rowoff[n];
    const int to = tapoff[t];
    if (emask && !((emask[e] >> (rclass ? rclass[n] : 0)) & 1)) ro = -1;     // entry does not feed this column class
    uint32_t o[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int co = centoff[e * 8 + jj * 2 + u];
        float x = 0.f;
        if (ro >= 0 && co >= 0) {
          const int idx = ro + co + to;
          x = w[idx];
          if (mask) x *= mask[idx];
        }
        v[u] = x;
      }
      o[jj] = pack_act2(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// all packed operands of a network in one launch: job j owns items [item_begin[j], item_begin[j+1])
__global__ void pack_weights_multi_kernel(const e2e_pack_job_t* __restrict__ jobs, int n_jobs, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_jobs - 1;                 // last job with item_begin <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].item_begin <= i) lo = mid; else hi = mid - 1;
    }
    const e2e_pack_job_t jb = jobs[lo];
    const long long li = i - jb.item_begin;
    const int n = (int)(li % jb.Npad);
    const long long sl = li / jb.Npad;             // slab = ks*2 + half
    const int hf = (int)(sl & 1);
    const int ks = (int)(sl >> 1);
    const int pr = ks / jb.n_taps, t = ks - pr * jb.n_taps;
    const int e = 2 * pr + hf;
    int ro = jb.rowoff[n];
    const int to = jb.tapoff[t];
    if (jb.emask && !((jb.emask[e] >> (jb.rclass ? jb.rclass[n] : 0)) & 1)) ro = -1;
    uint32_t o[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int co = jb.centoff[e * 8 + jj * 2 + u];
        float x = 0.f;
        if (ro >= 0 && co >= 0) {
          const int idx = ro + co + to;
          x = jb.w[idx];
          if (jb.mask) x *= jb.mask[idx];
        }
        v[u] = x;
      }
      o[jj] = pack_act2(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<act16*>(jb.out) + li * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, const int32_t* __restrict__ rowoff,
                                    const int32_t* __restrict__ centoff, const int32_t* __restrict__ tapoff,
                                    int n_cent, int n_taps, int Npad, float* __restrict__ grad) {
  const long long total = (long long)n_cent * n_taps * Npad * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    const long long i8 = i >> 3;
    const int n = (int)(i8 % Npad);
    const long long sl = i8 / Npad;
    const int hf = (int)(sl & 1);
    const int ks = (int)(sl >> 1);
    const int pr = ks / n_taps, t = ks - pr * n_taps;
    const int e = 2 * pr + hf;
    const int ro = rowoff[n], co = centoff[e * 8 + j];
    if (ro >= 0 && co >= 0) grad[ro + co + tapoff[t]] = dwp[i];
  }
}

template <int WN>
int launch_gemm(const e2e_gemm_t* p, cudaStream_t st) {
  constexpr int NT = 16 * WN;
  const long long M = (long long)p->B * p->Do * p->Ho * p->Wo;
  const int smem = STAGES * 4096 + STAGES * NT * 32;
  const int smem_epi = (p->out_mode == 0) ? NT * BM * 2 : NT * BM * 4;
  const int smem_bytes = smem > smem_epi ? smem : smem_epi;
  static E2eDevOnce attr_once;
  if (attr_once.first()) {
    E2E_CUDA(cudaFuncSetAttribute(gather_gemm_kernel<WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)(p->Npad / NT));
  gather_gemm_kernel<WN><<<grid, THREADS, smem_bytes, st>>>(*p);
  E2E_LAUNCHED("gather_gemm");
  return E2E_OK;
}

template <int WM>
int launch_wgrad(const e2e_wgrad_t* p, cudaStream_t st) {
  constexpr int MT = 16 * WM;
  constexpr int STAGE_BYTES = (MT / 8) * KT * 16 + SG * KT * 16;
  const int smem_bytes = WSTAGES * STAGE_BYTES;
  static E2eDevOnce attr_once;
  if (attr_once.first()) {
    E2E_CUDA(cudaFuncSetAttribute(gather_wgrad_kernel<WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  const long long M = (long long)p->B * p->Do * p->Ho * p->Wo;
  const int ntile = (int)((M + KT - 1) / KT);
  const int S = p->n_cent * p->n_taps;
  const int gx = (S + SG - 1) / SG, gy = (p->Npad + MT - 1) / MT;
  // enough voxel splits for ~2 waves of CTAs, at least 4 K tiles per CTA
  int want = (2 * e2e_num_sms() + gx * gy - 1) / (gx * gy);
  int maxsplit = (ntile + 3) / 4;
  int splits = want < 1 ? 1 : want;
  if (splits > maxsplit) splits = maxsplit;
  if (splits < 1) splits = 1;
  const int tps = (ntile + splits - 1) / splits;
  splits = (ntile + tps - 1) / tps;
  dim3 grid(gx, gy, splits);
  gather_wgrad_kernel<WM><<<grid, THREADS, smem_bytes, st>>>(*p, tps);
  E2E_LAUNCHED("gather_wgrad");
  return E2E_OK;
}

}  // namespace

int e2e_conv_tc_fwd(const e2e_gemm_t* p, int n, cudaStream_t st);   // conv_tc.cu
int e2e_conv_tc_supported(const e2e_gemm_t* p);
int e2e_conv_tc_stats_slots(const e2e_gemm_t* gs, int n);
int e2e_wgrad_tc(const e2e_wgrad_t* p, cudaStream_t st);
int e2e_wgrad_tc_supported(const e2e_wgrad_t* p);

static int gather_gemm_one(const e2e_gemm_t* p, void* stream, bool allow_tc);

extern "C" int e2e_gather_gemm(const e2e_gemm_t* p, void* stream) { return gather_gemm_one(p, stream, true); }

// n column chunks of one GEMM (identical except wpacked / cols / Npad): one launch on the tcgen05
// path, n launches otherwise
extern "C" int e2e_gather_gemm_multi(const e2e_gemm_t* p, int32_t n, void* stream) {
  E2E_ARG(p != nullptr && n >= 1, "gather_gemm_multi: bad arguments");
  if (n == 1) return gather_gemm_one(p, stream, true);
  bool tc = p[0].impl == 1 && n <= 12;
  for (int i = 0; i < n && tc; ++i) tc = p[i].impl == 1 && e2e_conv_tc_supported(p + i) == e2e_conv_tc_supported(p);
  if (tc && e2e_conv_tc_supported(p)) {
    const long long M = (long long)p->B * p->Do * p->Ho * p->Wo;
    if (M <= 0) return E2E_OK;
    return e2e_conv_tc_fwd(p, n, reinterpret_cast<cudaStream_t>(stream));
  }
  for (int i = 0; i < n; ++i) {
    const int rc = gather_gemm_one(p + i, stream, true);
    if (rc != E2E_OK) return rc;
  }
  return E2E_OK;
}

extern "C" int e2e_gather_gemm_on_tcgen05(const e2e_gemm_t* p, int32_t n) {
  if (p == nullptr || n < 1 || n > 12 || p->out_mode != 0) return 0;
  for (int i = 0; i < n; ++i)
    if (p[i].impl != 1 || !e2e_conv_tc_supported(p + i) || e2e_conv_tc_supported(p + i) != e2e_conv_tc_supported(p)) return 0;
  return 1;
}

extern "C" int e2e_gather_gemm_stats_slots(const e2e_gemm_t* p, int32_t n) {
  if (!e2e_gather_gemm_on_tcgen05(p, n)) return 0;
  if ((long long)p->B * p->Do * p->Ho * p->Wo <= 0) return 0;
  return e2e_conv_tc_stats_slots(p, n);
}

static int gather_gemm_one(const e2e_gemm_t* p, void* stream, bool allow_tc) {
  E2E_ARG(p != nullptr, "gather_gemm: null params");
  E2E_ARG(p->n_cent > 0 && (p->n_cent & 1) == 0, "gather_gemm: n_cent must be even and > 0 (got %d)", p->n_cent);
  E2E_ARG(p->n_taps > 0, "gather_gemm: n_taps must be > 0");
  E2E_ARG(p->Npad > 0 && p->Npad % 16 == 0, "gather_gemm: Npad must be a multiple of 16 (got %d)", p->Npad);
  E2E_ARG(p->n_src >= 1 && p->n_src <= E2E_MAX_SRC, "gather_gemm: bad n_src");
  E2E_ARG(p->out_mode == 0 || p->out_mode == 1, "gather_gemm: bad out_mode");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long M = (long long)p->B * p->Do * p->Ho * p->Wo;
  if (M <= 0) return E2E_OK;
  if (allow_tc && p->impl == 1 && e2e_conv_tc_supported(p)) return e2e_conv_tc_fwd(p, 1, st);
  E2E_ARG(p->stats == nullptr, "gather_gemm: fused statistics need the tcgen05 path (query e2e_gather_gemm_stats_slots first)");
  E2E_ARG(p->accumulate == 0, "gather_gemm: accumulate needs the tcgen05 path (query e2e_gather_gemm_on_tcgen05 first)");
  if (p->n_taps == 3) {
    e2e_set_error("gather_gemm: a kw-stacked plan (3 taps, N = 3 x Cout) is only executable by the tcgen05 kernel");
    return E2E_ERR_UNSUPPORTED;
  }
  const int N = p->Npad;
  if (N % 128 == 0) return launch_gemm<8>(p, st);
  if (N % 96 == 0) return launch_gemm<6>(p, st);
  if (N % 64 == 0) return launch_gemm<4>(p, st);
  if (N % 48 == 0) return launch_gemm<3>(p, st);
  if (N % 32 == 0) return launch_gemm<2>(p, st);
  return launch_gemm<1>(p, st);
}

extern "C" int e2e_gather_wgrad_direct_ok(const e2e_wgrad_t* p) {
  return (p != nullptr && p->impl == 1 && e2e_wgrad_tc_supported(p)) ? 1 : 0;
}

extern "C" int e2e_gather_wgrad(const e2e_wgrad_t* p, void* stream) {
  E2E_ARG(p != nullptr, "gather_wgrad: null params");
  E2E_ARG(p->n_cent > 0 && (p->n_cent & 1) == 0, "gather_wgrad: n_cent must be even and > 0");
  E2E_ARG(p->n_taps > 0, "gather_wgrad: n_taps must be > 0");
  E2E_ARG(p->Npad > 0 && p->Npad % 16 == 0, "gather_wgrad: Npad must be a multiple of 16");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long M = (long long)p->B * p->Do * p->Ho * p->Wo;
  if (M <= 0) return E2E_OK;
  if (p->grad_out != nullptr)
    E2E_ARG(p->rowoff && p->centoff && p->tapoff, "gather_wgrad: direct mode needs the rowoff / centoff / tapoff tables");
  else
    E2E_ARG(p->dwp != nullptr, "gather_wgrad: null dwp");
  if (p->impl == 1 && e2e_wgrad_tc_supported(p)) return e2e_wgrad_tc(p, st);
  E2E_ARG(p->grad_out == nullptr, "gather_wgrad: direct mode needs the tcgen05 path (query e2e_gather_wgrad_direct_ok first)");
  const int N = p->Npad;
  if (N <= 16) return launch_wgrad<1>(p, st);
  if (N <= 32) return launch_wgrad<2>(p, st);
  if (N <= 48) return launch_wgrad<3>(p, st);
  if (N <= 64 || N % 64 == 0) return launch_wgrad<4>(p, st);
  if (N <= 96 || N % 96 == 0) return launch_wgrad<6>(p, st);
  return launch_wgrad<8>(p, st);
}

extern "C" int e2e_pack_weights(const float* w, const float* mask, const int32_t* rowoff, const int32_t* centoff,
                                const int32_t* tapoff, const int32_t* emask, const int32_t* rclass, int32_t n_cent,
                                int32_t n_taps, int32_t Npad, void* wpacked, void* stream) {
  E2E_ARG(w && rowoff && centoff && tapoff && wpacked, "pack_weights: null pointer");
  const long long total = (long long)n_cent * n_taps * Npad;
  if (total <= 0) return E2E_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > e2e_num_sms() * 16) blocks = e2e_num_sms() * 16;
  pack_weights_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, mask, rowoff, centoff, tapoff, emask, rclass, n_cent, n_taps, Npad, reinterpret_cast<act16*>(wpacked));
  E2E_LAUNCHED("pack_weights");
  return E2E_OK;
}

extern "C" int e2e_pack_weights_multi(const e2e_pack_job_t* jobs, int32_t n_jobs, int64_t total_items, void* stream) {
  E2E_ARG(jobs && n_jobs > 0 && total_items >= 0, "pack_weights_multi: bad arguments");
  if (total_items == 0) return E2E_OK;
  long long blocks = (total_items + 255) / 256;
  if (blocks > e2e_num_sms() * 16) blocks = e2e_num_sms() * 16;
  pack_weights_multi_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(jobs, n_jobs, total_items);
  E2E_LAUNCHED("pack_weights_multi");
  return E2E_OK;
}

extern "C" int e2e_unpack_wgrad(const float* dwp, const int32_t* rowoff, const int32_t* centoff,
                                const int32_t* tapoff, int32_t n_cent, int32_t n_taps, int32_t Npad, float* grad,
                                void* stream) {
  E2E_ARG(dwp && rowoff && centoff && tapoff && grad, "unpack_wgrad: null pointer");
  const long long total = (long long)n_cent * n_taps * Npad * 8;
  if (total <= 0) return E2E_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > e2e_num_sms() * 16) blocks = e2e_num_sms() * 16;
  unpack_wgrad_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dwp, rowoff, centoff, tapoff, n_cent,
                                                                                  n_taps, Npad, grad);
  E2E_LAUNCHED("unpack_wgrad");
  return E2E_OK;
}
