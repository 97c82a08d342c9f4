// tcgen05 / TMA implicit-GEMM kernel for the stride-1 depth-shifted (1,3,3) convolutions of
// E2ENet (forward of every stride-1 ConvDropoutNormNonlin, and the data gradients of those
// layers, which are stride-1 correlations too).  Reference math: unetpp_d.py:45-59 (shift),
// :102-108 (conv), with the torch.cat of :453-478 folded in.
//
// Data flow per CTA (persistent, one CTA per SM, static round-robin over tiles):
//   tile  = 16 (H) x 8m (W) output voxels of one depth slice d of one sample b
//   K loop over channel-entry PAIRS (16 channels): per pair
//     A: 2 TMA tensor loads (one per 8-channel block) of the HALOED input window
//        [18][8m+2][8ch] at depth d - shift(block): TMA's out-of-bounds zero fill provides
//        both the conv padding and the shift's zero fill; all 9 filter taps and all m
//        sub-tiles read the same window through different UMMA descriptor start addresses
//        (no-swizzle K-major "interleaved" layout: a core matrix is 8 W-consecutive voxels
//        x 8 channels = 128 contiguous bytes, SBO = one window row).
//     B: 1 bulk copy of the pre-packed bf16 weights [9 taps][2][Npad][8] of this pair.
//     9*m tcgen05.mma (M=128, N=Npad, K=16) accumulate into m TMEM accumulators.
//   epilogue warps: tcgen05.ld -> bf16 -> 16-byte coalesced stores into the C8 destination(s);
//   TMEM is double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
//
// The same kernel has a second, "point" form (template HALO = false) for the 1-tap GEMMs of the
// path -- ConvTranspose3d(kernel == stride) forward (unetpp_d.py:521-522: every coarse voxel
// produces kd*kh*kw fine voxels; column block q is stored at fine voxel o*k + cols[q].off) and its
// data gradient (each K entry is a (tap, channel block) of dy fetched at o*k + tap: a TMA box with
// elementStrides = k and a per-entry start offset).  No halo, one MMA per (K step, sub-tile).
#include "common.cuh"

#include <type_traits>

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

namespace {

constexpr int TC_THREADS = 192;
constexpr int TH = 16;                 // tile rows (H)
constexpr int MAX_CENT = 384;
constexpr int TC_MAX_CHUNKS = 12;
constexpr int SMEM_BUDGET = 196 * 1024;    // dynamic pipeline stages; static tables take up to ~15 KB more

struct TcParams {
  int B, D, H, W;        // iteration grid (D, H, W); sources have Dsrc slices, destinations (Ddst, Hd, Wd)
  int Dsrc, Ddst, Hd, Wd;
  int isd, ish, isw, ivh, ivw;     // point form: source voxel = o * is + iv + entry offset
  int osd, osh, osw;               // destination voxel = o * os + column-block offset
  int rows;                        // window rows per slab: 18 (halo) or 16 (point)
  int need_bounds;                 // destination offsets / strides can leave the destination grid
  int merged;                      // source maps are 4-D with the merged (W, channel) inner dimension
  int serial_prod;                 // halo form: one thread issues all TMA ops of a stage (A/B test)
  int pps;                         // K pairs (16 channels each) per pipeline stage
  int poll;                        // E2E_TC_POLL: 0 all lanes poll, 1 one lane, 2 one lane + backoff
  int dbg;                         // E2E_TC_DBG (timing experiments, results wrong): 1 = no epilogue work, 2 = no MMAs
  int b_res;                       // packed weights of the CTA's (fixed) column chunk stay resident in smem
  int b_region_bytes;              // size of that region (then the A stages follow)
  int n_cent, Npad, m, stages, acc_stages;
  int ivd;
  int tiles_h, tiles_w, n_tiles;
  int a_slab_bytes;      // padded to 128
  int a_stage_bytes, b_stage_bytes, stage_bytes;
  int src_cb[E2E_MAX_SRC];
  const e2e_centry_t* cents;
  const e2e_tap_t* taps;
  // column chunks of one GEMM batched into one launch: work item = (tile, chunk); chunks differ in
  // packed weights, column table and width only (Npad above is the widest: it fixes the geometry)
  int n_chunks;
  const act16* wpacked[TC_MAX_CHUNKS];
  const e2e_colblk_t* cols[TC_MAX_CHUNKS];
  int npad[TC_MAX_CHUNKS];
  void* dst[E2E_MAX_SRC];
  int dst_cb[E2E_MAX_SRC];
  // InstanceNorm statistics of the (bf16-rounded) result, reduced in the epilogue: every CTA owns the slot
  // stats[(blockIdx.x * B + b) * 2 + {0: sum, 1: sum of squares}][stats_ctot]
  float* stats;
  int stats_ctot;                  // channels per row of `stats` (destination channel = col.blk * 8 + e)
  int stats_smem_off;              // byte offset of the per-warp accumulators [epi_warps][2][Npad] in dynamic smem
  int accum;                       // bit i: the result is ADDED to destination i (gradient fan-in of an activation)
};

struct alignas(64) TcMaps {
  CUtensorMap m[E2E_MAX_SRC];
};
// conv_tc3: m2 = the same tensors with a box of TWO consecutive 8-channel blocks (two adjacent K entries of one source
// with the same depth offset are fetched by one TMA instruction; the producer thread is an instruction stream)
struct alignas(64) S3Maps {
  CUtensorMap m[E2E_MAX_SRC];
  CUtensorMap m2[E2E_MAX_SRC];
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug becomes a trap (launch error), never a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// Warp-level waits: ONE lane polls the mbarrier (a 32-lane try_wait on one address is a 32-way
// serialised shared-memory access, and ncu showed the polling taking 5-20 % of the shared-memory
// pipe that the tensor core needs for its operands), the rest of the warp joins at __syncwarp().
// `backoff` > 0 adds a nanosleep between polls for waits that are expected to be long.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int backoff = 0, int mode = 1) {
  if (mode == 0) { mbar_wait(bar, parity); return; }
  if (mode == 1) backoff = 0;
  if ((threadIdx.x & 31) == 0) {
    if (!mbar_try_wait(bar, parity)) {
      const long long t0 = clock64();
      while (!mbar_try_wait(bar, parity)) {
        if (backoff) __nanosleep(backoff);
        if (clock64() - t0 > 8000000000LL) __trap();
      }
    }
  }
  __syncwarp();
}

// 4-D form for the stride-1 maps: (W, channel) is merged into one contiguous inner dimension of
// 32-bit elements (one voxel = 16 B = 4 elements), so a window row is ONE contiguous TMA row
// instead of 8m+2 separate 16-byte rows
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one lane of a converged warp; the compiler knows the guarded region is single-threaded and keeps
// descriptor arithmetic / UTCHMMA operands in uniform registers without a divergence waterfall
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// tcgen05.mma with the two 64-bit shared-memory descriptors given as (lo, hi) halves: all descriptor
// arithmetic of the issue loop stays 32-bit
__device__ __forceinline__ void tc_mma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

// Column sums over the 32 lanes of a warp for 32 columns at once: a halving butterfly (16 + 8 + 4 + 2 + 1 = 31
// shuffles instead of 32 x 5); on return lane L holds the sum over all lanes of v[L].  v is destroyed.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
// the same for 8 columns: lane L ends with the sum over all 32 lanes of v[L & 7]
__device__ __forceinline__ float warp_colsum8(float (&v)[8], int lane) {
#pragma unroll
  for (int o = 4; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float r = v[0];
  r += __shfl_xor_sync(0xffffffffu, r, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 16);
  return r;
}

struct ColInfo {            // one 8-column block of the result, decoded once per CTA
  int32_t off;              // voxel offset of (blk, od, oh, ow) inside one sample of the destination
  int32_t bstride;          // voxels per sample of that destination (dst_cb * Dd * Hd * Wd)
  int16_t blk;              // destination channel block (statistics are indexed by destination channel)
  int8_t od, oh, ow;        // small offsets: |od| <= 4 (shift-folded dgrad), oh / ow < 8 (transposed-conv taps)
  int8_t dst;               // -1: dead block
  uint8_t chmask;
  uint8_t pad_;
};
static_assert(sizeof(ColInfo) == 16, "ColInfo is read as one 16-byte shared-memory vector");

constexpr int EPI_WARPS = 16;            // upper bound; the launch picks 8 or 16 (blockDim.x = 64 + 32 * warps)
constexpr int TC_THREADS2 = 64 + 32 * EPI_WARPS;

// MS: 8-voxel-wide sub-tiles (accumulators) per tile.  MODE selects the epilogue: 0 plain store, 1 store +
// InstanceNorm statistics (always 8 epilogue warps: the smaller block leaves 204 registers per thread for the
// column-sum butterflies), 2 accumulate into the destination (gradient fan-in)
template <bool HALO, int MS, int MODE>
__global__ void __launch_bounds__(MODE == 1 ? 64 + 32 * 8 : TC_THREADS2, 1)
conv_tc_kernel(const __grid_constant__ TcParams p, const __grid_constant__ TcMaps maps) {
  constexpr int NT = HALO ? 9 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[32];
  __shared__ uint32_t tmem_base_s;
  __shared__ e2e_centry_t s_cents[MAX_CENT];
  __shared__ ColInfo s_cols[TC_MAX_CHUNKS * 32];
  __shared__ int s_tapoff[9];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Npad = p.Npad, S = p.stages, AS = p.acc_stages;
  constexpr int rowpitch = (8 * MS + (HALO ? 2 : 0)) * 16;   // bytes per window row
  const int npairs = p.n_cent >> 1;
  // barrier indices
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto tfull_bar = [&](int a) { return smem_u32(&bars[16 + a]); };
  auto tempty_bar = [&](int a) { return smem_u32(&bars[20 + a]); };
  const uint32_t bfull_bar = smem_u32(&bars[24]);

  const int epi_warps = ((int)blockDim.x >> 5) - 2;
  for (int i = threadIdx.x; i < p.n_cent; i += blockDim.x) s_cents[i] = p.cents[i];
  for (int i = threadIdx.x; i < p.n_chunks * 32; i += blockDim.x) {
    const int ch = i >> 5, q = i & 31;
    if (q >= (p.npad[ch] >> 3)) continue;
    const e2e_colblk_t c = p.cols[ch][q];
    ColInfo ci;
    const int dst = (c.dst >= 0 && c.chmask != 0) ? c.dst : -1;
    const int plane = p.Ddst * p.Hd * p.Wd;
    ci.dst = (int8_t)dst;
    ci.chmask = (uint8_t)c.chmask;
    ci.od = (int8_t)c.od; ci.oh = (int8_t)c.oh; ci.ow = (int8_t)c.ow;
    ci.blk = (int16_t)c.blk;
    ci.pad_ = 0;
    ci.off = dst >= 0 ? c.blk * plane + (c.od * p.Hd + c.oh) * p.Wd + c.ow : 0;
    ci.bstride = dst >= 0 ? p.dst_cb[dst] * plane : 0;
    s_cols[i] = ci;
  }
  if (threadIdx.x < 9) {
    int off = 0;
    if (HALO) {
      const e2e_tap_t t = p.taps[threadIdx.x];
      off = (1 + t.dh) * rowpitch + (1 + t.dw) * 16;
    }
    s_tapoff[threadIdx.x] = off;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < AS; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), epi_warps); }
    mbar_init(bfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;     // TMA destinations need 128 B; keep 1 KB

  if (warp == 0) {
    // ================================================= TMA producer WARP: lane 0 waits for the stage and
    // arms its barrier; then lanes 0 .. 2*np-1 issue the A slabs of the stage's np K pairs and lanes
    // 16 .. 16+np-1 their packed-weight copies, in parallel (a single issuing thread cannot keep up
    // with the short MMA bursts of the 1-tap form)
    int stage = 0, phase = 0;
    const uint32_t txa = 2u * (uint32_t)p.rows * (uint32_t)rowpitch;
    const int nwork = p.n_tiles * p.n_chunks;
    const int pps = p.pps;
    if (lane == 0 && p.b_res && (int)blockIdx.x < nwork) {
      // resident mode: gridDim.x is a multiple of n_chunks, so this CTA only ever sees one chunk;
      // its whole packed operand [pair][tap][2][npad][8] is fetched once
      const int ch = blockIdx.x % p.n_chunks;
      const uint32_t bbytes = (uint32_t)(NT * 2 * 16) * (uint32_t)p.npad[ch];
      mbar_expect_tx(bfull_bar, bbytes * (uint32_t)npairs);
      for (int pr = 0; pr < npairs; ++pr)
        bulk_copy_g2s(smem_base + pr * bbytes, p.wpacked[ch] + (size_t)pr * (bbytes / 2), bbytes, bfull_bar);
    }
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
      int t = work / p.n_chunks;
      const int ch = work - t * p.n_chunks;
      const uint32_t bbytes = (uint32_t)(NT * 2 * 16) * (uint32_t)p.npad[ch];
      const act16* wsrc = p.wpacked[ch];
      const int wt = t % p.tiles_w; t /= p.tiles_w;
      const int ht = t % p.tiles_h; t /= p.tiles_h;
      const int d = t % p.D;
      const int b = t / p.D;
      const int h0 = ht * TH, w0 = wt * 8 * MS;
      for (int pr0 = 0; pr0 < npairs; pr0 += pps) {
        const int np = min(pps, npairs - pr0);
        const uint32_t sa = smem_base + p.b_region_bytes + stage * p.stage_bytes;
        if (HALO && p.serial_prod) {
          if (lane == 0) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), p.b_res ? txa : txa + bbytes);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const e2e_centry_t ce = s_cents[2 * pr0 + hf];
              tma_load_4d(sa + hf * p.a_slab_bytes, &maps.m[ce.src], full_bar(stage), (w0 - 1) * 4, h0 - 1,
                          d + p.ivd + ce.dd, b * p.src_cb[ce.src] + ce.blk);
            }
            if (!p.b_res)
              bulk_copy_g2s(sa + p.a_stage_bytes, wsrc + (size_t)pr0 * (bbytes / 2), bbytes, full_bar(stage));
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
          continue;
        }
        if (lane == 0) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), (uint32_t)np * (p.b_res ? txa : txa + bbytes));
        }
        __syncwarp();
        if (lane < 2 * np) {
          const int i = lane >> 1, hf = lane & 1;
          const e2e_centry_t ce = s_cents[2 * (pr0 + i) + hf];
          const uint32_t dsta = sa + i * p.a_stage_bytes + hf * p.a_slab_bytes;
          if (HALO)
            tma_load_4d(dsta, &maps.m[ce.src], full_bar(stage), (w0 - 1) * 4, h0 - 1, d + p.ivd + ce.dd,
                        b * p.src_cb[ce.src] + ce.blk);
          else if (p.merged)
            tma_load_4d(dsta, &maps.m[ce.src], full_bar(stage), (w0 + p.ivw + ce.dw) * 4, h0 * p.ish + p.ivh + ce.dh,
                        d * p.isd + p.ivd + ce.dd, b * p.src_cb[ce.src] + ce.blk);
          else
            tma_load_5d(dsta, &maps.m[ce.src], full_bar(stage), 0, w0 * p.isw + p.ivw + ce.dw,
                        h0 * p.ish + p.ivh + ce.dh, d * p.isd + p.ivd + ce.dd, b * p.src_cb[ce.src] + ce.blk);
        } else if (!p.b_res && lane >= 16 && lane < 16 + np) {
          const int i = lane - 16;
          bulk_copy_g2s(sa + pps * p.a_stage_bytes + i * p.b_stage_bytes, wsrc + (size_t)(pr0 + i) * (bbytes / 2), bbytes,
                        full_bar(stage));
        }
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer (warp-uniform loop, one elected lane issues)
    int stage = 0, phase = 0, as = 0, aphase = 0;
    // descriptors differ only in the 14-bit start-address field (units of 16 B) of the low word
    const uint64_t adesc = make_desc(0, p.a_slab_bytes, rowpitch);
    const uint32_t a_hi = (uint32_t)(adesc >> 32), a_lo0 = (uint32_t)adesc;
    uint32_t tapu[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) tapu[t] = (uint32_t)s_tapoff[t] >> 4;
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
    const uint32_t a_stage_units = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t b_stage_units = (uint32_t)p.b_stage_bytes >> 4;
    const uint32_t sa0 = (smem_base + (uint32_t)p.b_region_bytes) >> 4;
    const uint32_t sb0 = smem_base >> 4;
    const int nwork = p.n_tiles * p.n_chunks;
    if (p.b_res && (int)blockIdx.x < nwork) {
      mbar_wait_warp(bfull_bar, 0);
      tc_fence_after();
    }
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
      const int npc = p.npad[work % p.n_chunks];           // width of this column chunk
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=npc, M=128
      const uint32_t idesc = (1u << 4) | (E2E_UMMA_FMT << 7) | (E2E_UMMA_FMT << 10) | ((uint32_t)(npc >> 3) << 17) | (8u << 24);
      const uint64_t bdesc = make_desc(0, npc * 16, 128);
      const uint32_t b_hi = (uint32_t)(bdesc >> 32), b_lo0 = (uint32_t)bdesc;
      const uint32_t b_tap_units = (uint32_t)(2 * npc * 16) >> 4;
      mbar_wait_warp(tempty_bar(as), aphase ^ 1, 0, p.poll);
      tc_fence_after();
      const uint32_t acc0 = tmem_base + (uint32_t)(as * MS * Npad);
      for (int pr0 = 0; pr0 < npairs; pr0 += p.pps) {
        const int np = min(p.pps, npairs - pr0);
        mbar_wait_warp(full_bar(stage), phase, 0, p.poll);
        tc_fence_after();
        if (elect_one_sync()) {
          for (int i = 0; i < np; ++i) {
            const int pr = pr0 + i;
            const uint32_t a_lo = a_lo0 + sa0 + (uint32_t)stage * stage_units + (uint32_t)i * a_stage_units;
            const uint32_t b_lo = p.b_res ? b_lo0 + sb0 + (uint32_t)pr * (uint32_t)(NT * 2) * (uint32_t)npc
                                          : b_lo0 + sa0 + (uint32_t)stage * stage_units + (uint32_t)p.pps * a_stage_units +
                                                (uint32_t)i * b_stage_units;
            const uint32_t first = pr ? 1u : 0u;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              if (p.dbg & 2) break;
#pragma unroll
              for (int j = 0; j < MS; ++j)
                tc_mma_f16_lh(acc0 + (uint32_t)(j * Npad), a_lo + tapu[t] + (uint32_t)(j * 8), a_hi,
                              b_lo + (uint32_t)t * b_tap_units, b_hi, idesc, t ? 1u : first);
            }
          }
          tc_commit(empty_bar(stage));            // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) tc_commit(tfull_bar(as));   // accumulators of this tile are complete
      __syncwarp();
      if (++as == AS) { as = 0; aphase ^= 1; }
    }
  } else {
    // ================================================= epilogue: EPI_WARPS / 4 warps per TMEM lane
    // quadrant; the warps of a quadrant take the 32-column chunks round-robin
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;                  // accumulator row = voxel (r / 8, r % 8) of a sub-tile
    const int need_bounds = p.need_bounds;      // bit 0: depth, 1: H, 2: W may leave the destination grid
    int as = 0, aphase = 0;
    const int nwork = p.n_tiles * p.n_chunks;
    // InstanceNorm statistics: this warp's accumulators [2][Npad] in shared memory, flushed to its own global
    // slot whenever the (sample, column chunk) of the work items changes (both are monotonic per CTA or alternate
    // with a short period) -- no atomics anywhere, so the statistics are bit-reproducible
    constexpr bool do_stats = MODE == 1;
    constexpr bool do_accum = MODE == 2;
    float* wstat_all = reinterpret_cast<float*>(smem + ((smem_base - smem_u32(smem)) + p.stats_smem_off));
    float* wstat = wstat_all + (warp - 2) * 2 * Npad;
    float* gslot = do_stats ? p.stats + (size_t)blockIdx.x * p.B * 2 * p.stats_ctot : nullptr;
    const int et = (int)threadIdx.x - 64, ethreads = epi_warps * 32;      // index among the epilogue threads
    int sb = -1, sch = -1;
    // register path: with one column chunk per launch and (warps per quadrant) % (32-column chunks) == 0 a warp
    // meets the same 32 columns in every work item, so each thread keeps running sums of its own rows and the
    // lanes are combined only at a flush (no shuffles per tile)
    const int nchunk0 = (p.npad[0] + 31) >> 5;
    const bool reg_stats = do_stats && p.n_chunks == 1 && ((epi_warps >> 2) % nchunk0) == 0;
    const int reg_c0 = (grp % nchunk0) << 5;
    float rs[do_stats ? 32 : 1], rq[do_stats ? 32 : 1];
#pragma unroll
    for (int i = 0; i < (do_stats ? 32 : 1); ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    // flush: executed by ALL epilogue warps at the same points of the (uniform) work sequence.  The per-warp
    // accumulators are summed in a fixed order and added to this CTA's global slot; named barrier 1 orders it.
    auto flush_stats = [&](int fb, int fch) {
      const int fn = p.npad[fch];
      const ColInfo* fc = s_cols + fch * 32;
      if constexpr (do_stats) {
        if (reg_stats) {
          const float s1 = warp_colsum32(rs, lane);       // destroys rs / rq: re-zeroed below
          const float s2 = warp_colsum32(rq, lane);
          if (reg_c0 + lane < fn) {
            wstat[reg_c0 + lane] += s1;
            wstat[Npad + reg_c0 + lane] += s2;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(ethreads) : "memory");
      for (int idx = et; idx < 2 * Npad; idx += ethreads) {
        const int k = idx >= Npad ? 1 : 0, c = idx - k * Npad;
        float sum = 0.f;
        for (int w = 0; w < epi_warps; ++w) {
          sum += wstat_all[w * 2 * Npad + idx];
          wstat_all[w * 2 * Npad + idx] = 0.f;
        }
        if (c < fn) {
          const ColInfo col = fc[c >> 3];
          if (col.dst >= 0) gslot[((size_t)fb * 2 + k) * p.stats_ctot + col.blk * 8 + (c & 7)] += sum;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(ethreads) : "memory");
    };
    if (do_stats) {
      for (int i = et; i < p.B * 2 * p.stats_ctot; i += ethreads) gslot[i] = 0.f;
      for (int c = lane; c < 2 * Npad; c += 32) wstat[c] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"r"(ethreads) : "memory");
    }
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
      int t = work / p.n_chunks;
      const int ch = work - t * p.n_chunks;
      const int npc = p.npad[ch];
      const int nchunk = (npc + 31) >> 5;
      const ColInfo* ccols = s_cols + ch * 32;
      const int wt = t % p.tiles_w; t /= p.tiles_w;
      const int ht = t % p.tiles_h; t /= p.tiles_h;
      const int d = t % p.D;
      const int b = t / p.D;
      if (do_stats && (b != sb || ch != sch)) {
        if (sb >= 0) flush_stats(sb, sch);
        sb = b; sch = ch;
      }
      const int h = ht * TH + (r >> 3);
      const int hs = h * p.osh, ds = d * p.osd;
      mbar_wait_warp(tfull_bar(as), aphase, 64, p.poll);
      tc_fence_after();
      const uint32_t acc0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * MS * Npad);
      for (int ci = (p.dbg & 1) ? MS * nchunk : grp; ci < MS * nchunk; ci += (epi_warps >> 2)) {
        const int j = ci / nchunk, c0 = (ci - j * nchunk) << 5;
        const int w = (wt * MS + j) * 8 + (r & 7);
        const int ws = w * p.osw;
        const bool inb = (h < p.H) && (w < p.W);
        const int tv = (ds * p.Hd + hs) * p.Wd + ws;
        uint32_t v[32];
        const bool two = c0 + 16 < npc;
        tc_ld16(acc0 + j * Npad + c0, v);
        if (two) tc_ld16(acc0 + j * Npad + c0 + 16, v + 16);
        // destination of column block u of this chunk (nullptr: nothing to store)
        auto dst_of = [&](int u) -> act16* {
          const ColInfo col = ccols[(c0 >> 3) + u];
          if (!inb || col.dst < 0) return nullptr;
          if (need_bounds) {
            if ((need_bounds & 1) && (unsigned)(ds + col.od) >= (unsigned)p.Ddst) return nullptr;
            if ((need_bounds & 2) && (unsigned)(hs + col.oh) >= (unsigned)p.Hd) return nullptr;
            if ((need_bounds & 4) && (unsigned)(ws + col.ow) >= (unsigned)p.Wd) return nullptr;
          }
          return reinterpret_cast<act16*>(p.dst[col.dst]) + (size_t)(uint32_t)(tv + col.off + b * col.bstride) * 8;
        };
        // accumulate mode: the values already stored at the destinations are fetched BEFORE the TMEM wait so that
        // their latency overlaps it (issued back to back: no store in between that they could alias)
        act16* dps[4] = {nullptr, nullptr, nullptr, nullptr};
        uint4 olds[4];
        if constexpr (do_accum) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            olds[u] = make_uint4(0, 0, 0, 0);
            if (u >= 2 && !two) continue;
            dps[u] = dst_of(u);
            if (dps[u] && ((p.accum >> ccols[(c0 >> 3) + u].dst) & 1)) olds[u] = *reinterpret_cast<const uint4*>(dps[u]);
          }
        }
        tc_wait_ld();
        uint4 pk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[u * 8 + e]);
          if constexpr (do_accum) {
            // gradient fan-in: another consumer of this activation already stored its contribution here
            const uint4 old = olds[u];
            f[0] += act_lo(old.x); f[1] += act_hi(old.x); f[2] += act_lo(old.y); f[3] += act_hi(old.y);
            f[4] += act_lo(old.z); f[5] += act_hi(old.z); f[6] += act_lo(old.w); f[7] += act_hi(old.w);
          }
          pk[u] = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]), pack_act2(f[4], f[5]),
                             pack_act2(f[6], f[7]));
        }
        if constexpr (do_stats) {
          // sums of the values exactly as stored (the packed bf16), rows outside the grid excluded
          float t1[32], t2[32];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool on = inb && (u < 2 || two);
            const uint32_t wv[4] = {pk[u].x, pk[u].y, pk[u].z, pk[u].w};
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              const float lo = on ? act_lo(wv[hh]) : 0.f, hi = on ? act_hi(wv[hh]) : 0.f;
              t1[u * 8 + 2 * hh] = lo; t1[u * 8 + 2 * hh + 1] = hi;
              t2[u * 8 + 2 * hh] = lo * lo; t2[u * 8 + 2 * hh + 1] = hi * hi;
            }
          }
          if (reg_stats) {
            // this warp always sees the same 32 columns: running sums per thread, lanes combined at the flush
#pragma unroll
            for (int i = 0; i < 32; ++i) { rs[i] += t1[i]; rq[i] += t2[i]; }
          } else {
            const float s1 = warp_colsum32(t1, lane);
            const float s2 = warp_colsum32(t2, lane);
            if (c0 + lane < npc) {
              wstat[c0 + lane] += s1;
              wstat[Npad + c0 + lane] += s2;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (u >= 2 && !two) break;
          act16* dp;
          if constexpr (do_accum) dp = dps[u]; else dp = dst_of(u);
          if (dp == nullptr) continue;
          const uint4 o = pk[u];
          const int cm = ccols[(c0 >> 3) + u].chmask;
          if (cm == 0xff) {
            *reinterpret_cast<uint4*>(dp) = o;
          } else {
            // shift groups are contiguous channel ranges: store the selected channels in the widest
            // aligned pieces (8 / 4 / 2 bytes)
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int m4 = (cm >> (4 * h2)) & 0xf;
              if (m4 == 0xf) {
                *reinterpret_cast<uint2*>(dp + 4 * h2) = make_uint2(ow[2 * h2], ow[2 * h2 + 1]);
              } else {
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                  const int m2 = (m4 >> (2 * q2)) & 3;
                  const int e = 4 * h2 + 2 * q2;
                  if (m2 == 3) {
                    *reinterpret_cast<uint32_t*>(dp + e) = ow[2 * h2 + q2];
                  } else if (m2) {
                    const act16* ov = reinterpret_cast<const act16*>(&ow[2 * h2 + q2]);
                    if (m2 & 1) dp[e] = ov[0];
                    if (m2 & 2) dp[e + 1] = ov[1];
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == AS) { as = 0; aphase ^= 1; }
    }
    if (do_stats && sb >= 0) flush_stats(sb, sch);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}


// =====================================================================================
// kw-stacked forward kernel for NARROW stride-1 3x3 layers (Cout <= 80).
//
// An SS-mode tcgen05.mma with M = 128, N = 48, K = 16 reads 4 KB of A and 1.5 KB of B from shared
// memory for 24 cycles of math: ncu shows conv_tc_kernel on these layers at 71 % of the tensor core's
// shared-memory read pipe and 27 % of its math pipe.  Here the three kw taps are stacked along N:
//     D'[v][kw * Np + n] = sum_{kh, c}  X[v + (kh - 1) rows][c] * W[kh][kw][c][n]      (N = 3 * Np)
//     out[h][w][n]       = D'[h][w - 1][0 * Np + n] + D'[h][w][1 * Np + n] + D'[h][w + 1][2 * Np + n]
// so A is read once per THREE taps (3 MMAs of N = 144 per 16 channels instead of 9 of N = 48) and the
// W shift happens in the epilogue as two lane shuffles: an M tile is 4 (H) x 32 (W) positions, TMEM
// lane = (row, w), one row per warp quadrant.  Lanes 0 and 31 only feed their neighbours, so tiles
// advance by 30 in W.  A work item is TWO vertically adjacent M tiles that share one A window
// [10 rows][32 voxels][8 ch] per 8-channel block (half the TMA boxes per voxel -- the TMA unit, not
// the tensor pipe, bounds this kernel -- and 1.25x instead of 1.5x H-halo traffic; no W halo: row
// pitch 512 B makes the 16 core matrices of an M tile contiguous, SBO = 128).  Their accumulators
// rotate through three 3*Np-column TMEM slots, so the epilogue of one M tile overlaps the MMAs of the
// next work item.  Packed weights [pair][kh][2][3*Np][8] stay resident in shared memory.
// =====================================================================================
// (b, d, ht, wt) of the tiles blockIdx.x, blockIdx.x + G, ... without a division per tile: the stride G is decomposed
// once and added digit by digit with carries.  The role warps of these kernels are single dependent instruction streams
// (one elected lane does the work), so the ~6 integer divisions of a tile decode cost several hundred cycles per work
// item -- measured on conv_tc3 with every load, MMA and epilogue switched off: 1.1 us per item of pure loop overhead.
struct TileWalk {
  int wt, ht, d, b, gw, gh, gd, gb, tw, th, D;
  __device__ __forceinline__ void init(int tile0, int G, int tiles_w, int tiles_h, int D_) {
    tw = tiles_w; th = tiles_h; D = D_;
    int t = tile0;
    wt = t % tw; t /= tw;
    ht = t % th; t /= th;
    d = t % D; b = t / D;
    t = G;
    gw = t % tw; t /= tw;
    gh = t % th; t /= th;
    gd = t % D; gb = t / D;
  }
  __device__ __forceinline__ void next() {
    wt += gw;
    int c = wt >= tw ? 1 : 0;
    wt -= c ? tw : 0;
    ht += gh + c;
    c = ht >= th ? 1 : 0;
    ht -= c ? th : 0;
    d += gd + c;
    c = d >= D ? 1 : 0;
    d -= c ? D : 0;
    b += gb + c;
  }
};

constexpr int S3_MT = 2;                            // M tiles (4 rows each) per work item: they share one window
constexpr int S3_TH = 4 * S3_MT, S3_TW = 32, S3_ADV = 30, S3_ROWS = S3_TH + 2;
constexpr int S3_SLAB = S3_ROWS * S3_TW * 16;      // 5120 B: [10 rows][32 voxels][8 ch]
#ifndef E2E_S3_PPS
#define E2E_S3_PPS 2
#endif
constexpr int S3_PPS = E2E_S3_PPS;                           // K pairs per pipeline stage (2 * 3 * S3_MT = 12 MMAs)
#ifndef E2E_S3_EPI_WARPS
#define E2E_S3_EPI_WARPS 12
#endif
constexpr int S3_EPI_WARPS = E2E_S3_EPI_WARPS;      // 3 epilogue warps per scheduler within the 128-register cap of 448 threads
                                                    // (measured against 8: loc4 forward 0.336 -> 0.333 ms, 48->48 0.269 -> 0.259 ms)
constexpr int S3_UB = (6 + S3_EPI_WARPS / 4 - 1) / (S3_EPI_WARPS / 4);    // 8-channel blocks per warp when statistics are fused (Np <= 48)
constexpr int S3_NPROD = 2;                         // TMA producer threads (lane 0 of warps 0 .. S3_NPROD-1), alternating stages
constexpr int S3_EPI0 = S3_NPROD + S3_MT;           // then S3_MT MMA issuer threads (one per M tile of a work item), then the epilogue
constexpr int S3_THREADS = 32 * (S3_EPI0 + S3_EPI_WARPS);

struct S3Params {
  int B, D, H, W;
  int n_cent, Np, N3, ivd;          // Np = padded Cout (multiple of 8), N3 = 3 * Np (multiple of 16)
  int no_runs;                      // E2E_TC3_RUNS=0: one TMA box per K entry (A/B)
  int seq_mt;                       // 1: one issuer thread per M tile of a work item (E2E_TC3_SEQ=0: one warp issues both, interleaved)
  int dbg;                          // E2E_TC3_DBG (timing experiments only, results are wrong): 1 = the epilogue only hands the
                                    // accumulator slots back, 2 = the issuer skips the MMAs, 4 = no TMA loads, 8 = 128-byte aligned boxes
  int stages, acc_stages;
  int tiles_h, tiles_w, n_tiles;
  int b_region_bytes, stage_bytes;
  int src_cb[E2E_MAX_SRC];
  int dst_cb;
  const e2e_centry_t* cents;
  const act16* wpacked;
  act16* dst;
  float* stats;                     // see TcParams::stats (slot = blockIdx.x)
  int stats_ctot;
  int stats_smem_off;               // per-warp accumulators [S3_EPI_WARPS][2][Np]
};

__device__ __forceinline__ void s3_commit(int dbg, uint32_t bar) {
  if (dbg & 16) mbar_arrive(bar);          // (timing experiment, only meaningful with dbg & 2: a plain arrive instead of tcgen05.commit)
  else tc_commit(bar);
}

// FULL: every epilogue warp owns exactly S3_UB 8-channel blocks (Np = 8 * S3_UB * S3_EPI_WARPS / 4 = 48, the layers this
// kernel mostly serves): one TMEM round trip per M tile and no per-block guards
template <bool FULL>
__global__ void __launch_bounds__(S3_THREADS, 1)
conv_tc3_kernel(const __grid_constant__ S3Params p, const __grid_constant__ S3Maps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[32];
  __shared__ uint32_t tmem_base_s;
  __shared__ int4 s_box[MAX_CENT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N3 = p.N3, Np = p.Np, S = p.stages, AS = p.acc_stages;
  const int npairs = p.n_cent >> 1;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto tfull_bar = [&](int a) { return smem_u32(&bars[16 + a]); };
  auto tempty_bar = [&](int a) { return smem_u32(&bars[20 + a]); };
  const uint32_t bfull_bar = smem_u32(&bars[24]);

  // box table, one record per K entry: {source | run << 8, depth offset, channel block, blocks per sample}.  Entries k, k+1
  // of one pipeline stage that are consecutive blocks of one source at the same depth form a run of 2 (one box through
  // maps.m2); the second entry of a run has run = 0 and is skipped by the producer.
  for (int g0 = threadIdx.x * 2 * S3_PPS; g0 < p.n_cent; g0 += blockDim.x * 2 * S3_PPS) {
    const int n = min(2 * S3_PPS, p.n_cent - g0);
    for (int k = 0; k < n;) {
      const e2e_centry_t ce = p.cents[g0 + k];
      int run = 1;
      if (k + 1 < n && !p.no_runs) {
        const e2e_centry_t c2 = p.cents[g0 + k + 1];
        if (c2.src == ce.src && c2.dd == ce.dd && c2.blk == ce.blk + 1) {
          run = 2;
          s_box[g0 + k + 1] = make_int4(c2.src, c2.dd, c2.blk, p.src_cb[c2.src]);
        }
      }
      s_box[g0 + k] = make_int4(ce.src | (run << 8), ce.dd, ce.blk, p.src_cb[ce.src]);
      k += run;
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), p.seq_mt ? S3_MT : 1); }
    for (int a = 0; a < AS; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), S3_EPI_WARPS); }
    mbar_init(bfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t bbytes = 3u * 2u * 16u * (uint32_t)N3;          // packed weights of one K pair

  if (warp < S3_NPROD) {
    // ================================================= TMA producers: ONE thread each (lane 0 of warps 0 .. S3_NPROD-1),
    // taking the pipeline stages alternately.  A producer is a single dependent instruction stream (wait, expect_tx, per
    // box: entry lookup, coordinates, UTMALDG from uniform registers); issuing the boxes of a stage "in parallel" from
    // several lanes compiled to a serial per-lane loop anyway (ELECT / R2UR.BROADCAST / UTMALDG / BRA.U.ANY), and with the
    // loads alone (E2E_TC3_DBG=3) one producer warp bounded the kernel at 0.14 - 0.18 ms.
    if (warp == 0 && lane == 0 && (int)blockIdx.x < p.n_tiles) {
      mbar_expect_tx(bfull_bar, bbytes * (uint32_t)npairs);
      for (int pr = 0; pr < npairs; ++pr)
        bulk_copy_g2s(smem_base + pr * bbytes, p.wpacked + (size_t)pr * (bbytes / 2), bbytes, bfull_bar);
    }
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t sa = smem_base + (uint32_t)p.b_region_bytes;
      TileWalk tw;
      tw.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h, p.D);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, tw.next()) {
        const int d = tw.d + p.ivd, b = tw.b;
        const int h0 = tw.ht * S3_TH - 1, c0 = (tw.wt * S3_ADV - 1) * 4;
        for (int pr0 = 0; pr0 < npairs; pr0 += S3_PPS) {
          if (stage % S3_NPROD == warp) {        // a stage index always has the same producer: mbarrier parity waits tell
                                                 // only one phase from the next, so nobody may wait two phases ahead
            const int nb = 2 * min(S3_PPS, npairs - pr0);
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (p.dbg & 4) {
              mbar_arrive(full_bar(stage));                      // (timing experiment: no loads at all)
            } else {
              mbar_expect_tx(full_bar(stage), (uint32_t)(nb * S3_SLAB));
              const uint32_t dst0 = sa + (uint32_t)(stage * p.stage_bytes);
              for (int k = 0; k < nb; ++k) {
                const int4 bx = s_box[2 * pr0 + k];
                const int run = bx.x >> 8, src = bx.x & 0xff;
                if (run == 0) continue;                          // covered by the previous entry's two-block box
                tma_load_4d(dst0 + (uint32_t)(k * S3_SLAB), run == 2 ? &maps.m2[src] : &maps.m[src], full_bar(stage), c0, h0,
                            d + bx.y, b * bx.w + bx.z);
              }
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp < S3_EPI0) {
    // ================================================= MMA issuers
    const uint32_t idesc = (1u << 4) | (E2E_UMMA_FMT << 7) | (E2E_UMMA_FMT << 10) | ((uint32_t)(N3 >> 3) << 17) | (8u << 24);
    const uint64_t adesc = make_desc(0, S3_SLAB, 128);             // LBO: next 8 channels, SBO: next 8 voxels
    const uint64_t bdesc = make_desc(0, N3 * 16, 128);
    const uint32_t a_hi = (uint32_t)(adesc >> 32), a_lo0 = (uint32_t)adesc;
    const uint32_t b_hi = (uint32_t)(bdesc >> 32), b_lo0 = (uint32_t)bdesc;
    const uint32_t sa0 = (smem_base + (uint32_t)p.b_region_bytes) >> 4, sb0 = smem_base >> 4;
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
    const uint32_t b_tap_units = (uint32_t)(2 * N3);               // one kh slice of a pair, in 16-byte units
    int stage = 0, phase = 0;
    uint32_t mcount = 0;                       // M tiles issued so far: slot = mcount % AS, phase = (mcount / AS) & 1
    if (p.seq_mt) {
      // One issuer THREAD per M tile of a work item (lane 0 of warps 1 and 2).  The issuer is a single dependent
      // instruction stream -- waits, descriptor arithmetic, MMAs, commits: ~6 cycles per instruction -- and with one warp
      // issuing both tiles (and electing a lane at every step) that stream, not the tensor pipe, bounded the kernel:
      // with every load, MMA and epilogue switched off (E2E_TC3_DBG=7) the loop alone took 1.1 us per work item.
      // The two issuers also decouple the tiles' accumulator slots: tile k waits only for the epilogue of tile k - 3
      // (3 slots), where the interleaved single issuer needed BOTH slots of the next item before its first MMA.
      // Both read the same K stages; a stage is released when both have committed (empty barrier count S3_MT).
      // (the loop is warp-uniform and one elected lane issues: descriptor arithmetic then stays in uniform registers, which
      // UTCHMMA reads directly -- inside an `if (lane == 0)` every operand needs an R2UR first)
      if ((int)blockIdx.x < p.n_tiles) {
        const int mt = warp - S3_NPROD;
        mbar_wait(bfull_bar, 0);
        tc_fence_after();
        int slot = mt % AS;
        uint32_t sphase = (uint32_t)(mt / AS) & 1u;
        const uint32_t a_mt = a_lo0 + sa0 + (uint32_t)(4 * mt * 32);
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
          mbar_wait(tempty_bar(slot), sphase ^ 1u);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(slot * N3);
          uint32_t b_st = b_lo0 + sb0;
          for (int pr0 = 0; pr0 < npairs; pr0 += S3_PPS) {
            const int np = min(S3_PPS, npairs - pr0);
            mbar_wait(full_bar(stage), (uint32_t)phase);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t a_st = a_mt + (uint32_t)stage * stage_units;
#pragma unroll
              for (int i = 0; i < S3_PPS; ++i) {
                if (i < np && !(p.dbg & 2)) {
                  const uint32_t a_lo = a_st + (uint32_t)(i * (2 * S3_SLAB >> 4));
                  const uint32_t b_lo = b_st + (uint32_t)i * 3u * b_tap_units;
#pragma unroll
                  for (int kh = 0; kh < 3; ++kh)
                    tc_mma_f16_lh(acc, a_lo + (uint32_t)(kh * 32), a_hi, b_lo + (uint32_t)kh * b_tap_units, b_hi, idesc,
                                  (i + kh) ? 1u : (pr0 ? 1u : 0u));
                }
              }
              s3_commit(p.dbg, empty_bar(stage));
            }
            __syncwarp();
            b_st += (uint32_t)S3_PPS * 3u * b_tap_units;
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          if (elect_one_sync()) s3_commit(p.dbg, tfull_bar(slot));
          __syncwarp();
          slot += S3_MT;
          if (slot >= AS) { slot -= AS; sphase ^= 1u; }
        }
      }
    } else if (warp == S3_NPROD) {
    if ((int)blockIdx.x < p.n_tiles) {
      mbar_wait(bfull_bar, 0);
      tc_fence_after();
    }
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      uint32_t acc[S3_MT];
      int slot[S3_MT];
#pragma unroll
      for (int mt = 0; mt < S3_MT; ++mt) {
        const uint32_t mc = mcount + mt;
        slot[mt] = (int)(mc % (uint32_t)AS);
        mbar_wait(tempty_bar(slot[mt]), ((mc / (uint32_t)AS) & 1u) ^ 1u);
        acc[mt] = tmem_base + (uint32_t)(slot[mt] * N3);
      }
      tc_fence_after();
      for (int pr0 = 0; pr0 < npairs; pr0 += S3_PPS) {
        const int np = min(S3_PPS, npairs - pr0);
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_st = a_lo0 + sa0 + (uint32_t)stage * stage_units;
          const uint32_t b_st = b_lo0 + sb0 + (uint32_t)pr0 * 3u * b_tap_units;
#pragma unroll
          for (int i = 0; i < S3_PPS; ++i) {      // fully unrolled issue sequence (np <= S3_PPS)
            if (i < np && !(p.dbg & 2)) {
              const uint32_t a_lo = a_st + (uint32_t)(i * (2 * S3_SLAB >> 4));
              const uint32_t b_lo = b_st + (uint32_t)i * 3u * b_tap_units;
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {    // M tile mt reads window rows 4*mt + kh .. +3: one row = 32 units
#pragma unroll
                for (int mt = 0; mt < S3_MT; ++mt)
                  tc_mma_f16_lh(acc[mt], a_lo + (uint32_t)((4 * mt + kh) * 32), a_hi, b_lo + (uint32_t)kh * b_tap_units,
                                b_hi, idesc, (i + kh) ? 1u : (pr0 ? 1u : 0u));
              }
            }
          }
          s3_commit(p.dbg, empty_bar(stage));
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) {
#pragma unroll
        for (int mt = 0; mt < S3_MT; ++mt) s3_commit(p.dbg, tfull_bar(slot[mt]));
      }
      __syncwarp();
      mcount += S3_MT;
    }
    }
  } else {
    // ================================================= epilogue: S3_EPI_WARPS / 4 warps per TMEM lane quadrant (= tile
    // row), the warps of a quadrant take the 8-channel blocks of the result round-robin
    const int q = warp & 3;                    // tile row
    const int grp = (warp - S3_EPI0) >> 2;
    const int nblk = min(Np >> 3, p.dst_cb);      // Np is padded to 16 channels: the destination may have one block less
    uint32_t mcount = 0;
    // InstanceNorm statistics: every thread keeps the running sums of ITS voxels' values for the (at most S3_UB) channel
    // blocks its warp handles in registers -- no shuffles per tile -- and the lanes / warps are combined only when
    // the sample changes and at the end (host guarantees nblk <= 6 when statistics are requested)
    const bool do_stats = p.stats != nullptr;
    float* wstat_all = reinterpret_cast<float*>(smem + ((smem_base - smem_u32(smem)) + p.stats_smem_off));
    float* wstat = wstat_all + (warp - S3_EPI0) * 2 * Np;
    float* gslot = do_stats ? p.stats + (size_t)blockIdx.x * p.B * 2 * p.stats_ctot : nullptr;
    const int et = (int)threadIdx.x - 32 * S3_EPI0;
    constexpr int ethreads = S3_EPI_WARPS * 32;
    float rs[S3_UB][8], rq[S3_UB][8];
#pragma unroll
    for (int u = 0; u < S3_UB; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) { rs[u][e] = 0.f; rq[u][e] = 0.f; }
    int sb = -1;
    auto flush_stats = [&](int fb) {
#pragma unroll
      for (int u = 0; u < S3_UB; ++u) {
        const int cb = grp + u * (S3_EPI_WARPS / 4);
        const float s1 = warp_colsum8(rs[u], lane);
        const float s2 = warp_colsum8(rq[u], lane);
        if (cb < nblk && lane < 8) {
          wstat[cb * 8 + lane] = s1;
          wstat[Np + cb * 8 + lane] = s2;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { rs[u][e] = 0.f; rq[u][e] = 0.f; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ethreads) : "memory");
      for (int idx = et; idx < 2 * Np; idx += ethreads) {
        const int k = idx >= Np ? 1 : 0, c = idx - k * Np;
        float sum = 0.f;
        for (int w = 0; w < S3_EPI_WARPS; ++w) {
          sum += wstat_all[w * 2 * Np + idx];
          wstat_all[w * 2 * Np + idx] = 0.f;
        }
        if (c < p.stats_ctot) gslot[((size_t)fb * 2 + k) * p.stats_ctot + c] += sum;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ethreads) : "memory");
    };
    if (do_stats) {
      for (int i = et; i < p.B * 2 * p.stats_ctot; i += ethreads) gslot[i] = 0.f;
      for (int c = lane; c < 2 * Np; c += 32) wstat[c] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(ethreads) : "memory");
    }
    TileWalk tw;
    tw.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h, p.D);
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t plane8 = (uint32_t)(p.D * p.H * p.W) * 8u;           // elements between consecutive channel blocks

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, tw.next()) {
      const int ht = tw.ht, d = tw.d, b = tw.b;
      const int w = tw.wt * S3_ADV - 1 + lane;
      if (do_stats && b != sb) {
        if (sb >= 0) flush_stats(sb);
        sb = b;
      }
     // voxel index of (b, channel block 0, d, first row of this quadrant, w); + plane per channel block, + 4 W per M tile
     const uint32_t vox0 = (uint32_t)((((b * p.dst_cb) * p.D + d) * p.H + (ht * S3_TH + q)) * p.W + w);   // (host: < 2^28 voxels)
     const bool wok = lane >= 1 && lane <= S3_ADV && w < p.W;
     for (int mt = 0; mt < S3_MT; ++mt) {
      const int h = ht * S3_TH + 4 * mt + q;
      const bool ok = wok && h < p.H;
      const uint32_t off0 = (vox0 + (uint32_t)(4 * mt * p.W)) * 8u;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * N3);
      constexpr int EB = S3_UB;                  // blocks per TMEM round trip (3 loads each in flight)
      constexpr int G = S3_EPI_WARPS / 4;
      if (!(p.dbg & 1)) {
        for (int cb0 = grp; cb0 < nblk; cb0 += EB * G) {
          uint32_t v[EB][3][8];
#pragma unroll
          for (int u = 0; u < EB; ++u) {
            const int cb = cb0 + u * G;
            if (FULL || cb < nblk) {
              tc_ld8(acc + cb * 8, v[u][0]);
              tc_ld8(acc + Np + cb * 8, v[u][1]);
              tc_ld8(acc + 2 * Np + cb * 8, v[u][2]);
            }
          }
          tc_wait_ld();
#pragma unroll
          for (int u = 0; u < EB; ++u) {
            const int cb = cb0 + u * G;
            if (!FULL && cb >= nblk) break;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v[u][0][e]), 1);     // D'[w - 1][kw = 0]
              const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v[u][2][e]), 1);  // D'[w + 1][kw = 2]
              o[e] = left + __uint_as_float(v[u][1][e]) + right;
            }
            if (ok) {
              const uint4 pk = make_uint4(pack_act2(o[0], o[1]), pack_act2(o[2], o[3]), pack_act2(o[4], o[5]),
                                          pack_act2(o[6], o[7]));
              *reinterpret_cast<uint4*>(p.dst + (off0 + (uint32_t)cb * plane8)) = pk;
              if (do_stats) {                       // sums of the values exactly as stored (the packed 16-bit values)
                const uint32_t wv[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                  const float lo = act_lo(wv[hh]), hi = act_hi(wv[hh]);
                  rs[u][2 * hh] += lo;
                  rs[u][2 * hh + 1] += hi;
                  rq[u][2 * hh] = __fmaf_rn(lo, lo, rq[u][2 * hh]);
                  rq[u][2 * hh + 1] = __fmaf_rn(hi, hi, rq[u][2 * hh + 1]);
                }
              }
            }
          }
          if (FULL) break;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == AS) { as = 0; aphase ^= 1u; }
     }
    }
    if (do_stats && sb >= 0) flush_stats(sb);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

}  // namespace

// 1: stride-1 3x3 halo form, 2: 1-tap point form, 0: not served by this kernel
int e2e_conv_tc_supported(const e2e_gemm_t* p) {
  if (p->out_mode != 0) return 0;
  if (p->Npad > 256 || p->Npad % 16 != 0 || p->n_cent > MAX_CENT) return 0;
  {
    // the epilogue indexes destination voxels with 32 bits
    long long mx = 0;
    for (int i = 0; i < p->n_dst && i < E2E_MAX_SRC; ++i)
      if (p->dst_cb[i] > mx) mx = p->dst_cb[i];
    if ((long long)p->B * mx * p->Dd * p->Hd * (long long)p->Wd >= (1ll << 31)) return 0;
  }
  if (p->n_taps == 9) {
    if (p->isd != 1 || p->ish != 1 || p->isw != 1 || p->osd != 1 || p->osh != 1 || p->osw != 1) return 0;
    if (p->ivh != 0 || p->ivw != 0) return 0;
    if (p->Ho != p->Hi || p->Wo != p->Wi) return 0;      // depth counts may differ (shift-folded dgrad)
    if (p->Hd != p->Hi || p->Wd != p->Wi) return 0;
    return 1;
  }
  if (p->n_taps == 1) {
    if (p->isd < 1 || p->ish < 1 || p->isw < 1 || p->isd > 8 || p->ish > 8 || p->isw > 8) return 0;
    if (16 * p->ish > 256 || 32 * p->isw > 256) return 0;
    return 2;
  }
  if (p->n_taps == 3) {
    // kw-stacked forward of a narrow stride-1 layer: Npad = 3 * (padded Cout), one destination, in place
    if (p->isd != 1 || p->ish != 1 || p->isw != 1 || p->osd != 1 || p->osh != 1 || p->osw != 1) return 0;
    if (p->ivh != 0 || p->ivw != 0 || p->Do != p->Di || p->Ho != p->Hi || p->Wo != p->Wi) return 0;
    if (p->Dd != p->Di || p->Hd != p->Hi || p->Wd != p->Wi || p->n_dst != 1) return 0;
    if (p->Npad % 48 != 0 || p->Npad > 240) return 0;
    if ((long long)p->B * p->dst_cb[0] * p->Dd * p->Hd * (long long)p->Wd >= (1ll << 28)) return 0;   // 32-bit element offsets in the epilogue
    const int region = ((p->n_cent / 2) * 3 * 2 * 16 * p->Npad + 1023) / 1024 * 1024;
    if (region + 3 * S3_PPS * 2 * S3_SLAB + S3_EPI_WARPS * 2 * (p->Npad / 3) * 4 + 128 > SMEM_BUDGET) return 0;   // resident weights + 3 stages + statistics
    return 3;
  }
  return 0;
}

static int conv_tc3_launch(const e2e_gemm_t* g, PFN_cuTensorMapEncodeTiled_v12000 encode, cudaStream_t st, int* slots_out) {
  S3Params p{};
  p.B = g->B; p.D = g->Di; p.H = g->Hi; p.W = g->Wi;
  p.n_cent = g->n_cent; p.N3 = g->Npad; p.Np = g->Npad / 3; p.ivd = g->ivd;
  p.acc_stages = 512 / p.N3;
  if (p.acc_stages > 4) p.acc_stages = 4;
  const int npairs = g->n_cent / 2;
  p.b_region_bytes = (npairs * 3 * 2 * 16 * p.N3 + 1023) / 1024 * 1024;
  p.stage_bytes = S3_PPS * 2 * S3_SLAB;
  const int stats_bytes = (S3_EPI_WARPS * 2 * p.Np * 4 + 127) / 128 * 128;      // always reserved: same geometry with / without
  int stages = (SMEM_BUDGET - p.b_region_bytes - stats_bytes) / p.stage_bytes;
  if (stages > 8) stages = 8;
  p.stages = stages;
  p.tiles_h = (p.H + S3_TH - 1) / S3_TH;
  p.tiles_w = (p.W + S3_ADV - 1) / S3_ADV;
  p.n_tiles = p.B * p.D * p.tiles_h * p.tiles_w;
  {
    int grid0 = e2e_num_sms();
    if (grid0 > p.n_tiles) grid0 = p.n_tiles;
    const bool can_stats = (p.Np >> 3) <= S3_UB * (S3_EPI_WARPS / 4);   // register accumulators: <= S3_UB blocks per warp
    if (slots_out) { *slots_out = can_stats ? grid0 : 0; return E2E_OK; }
    if (g->stats && !can_stats) {
      e2e_set_error("conv_tc3: fused statistics support at most %d output channels", 8 * S3_UB * (S3_EPI_WARPS / 4));
      return E2E_ERR_UNSUPPORTED;
    }
  }
  {
    static int seq = -1;
    if (seq < 0) { const char* e = getenv("E2E_TC3_SEQ"); seq = e ? atoi(e) : 1; }
    p.seq_mt = seq ? 1 : 0;
  }
  p.stats = g->stats;
  p.stats_ctot = g->stats_ctot;
  p.stats_smem_off = p.b_region_bytes + p.stages * p.stage_bytes;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("E2E_TC3_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
  }
  p.cents = g->cents;
  p.wpacked = reinterpret_cast<const act16*>(g->wpacked);
  p.dst = reinterpret_cast<act16*>(g->dst[0]);
  p.dst_cb = g->dst_cb[0];
  {
    static int runs = -1;
    if (runs < 0) { const char* e = getenv("E2E_TC3_RUNS"); runs = e ? atoi(e) : 1; }
    p.no_runs = runs ? 0 : 1;
  }
  S3Maps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < E2E_MAX_SRC; ++i) {
    const int si = i < g->n_src ? i : 0;
    p.src_cb[i] = g->src_cb[si];
    cuuint64_t gdim[4] = {(cuuint64_t)g->Wi * 4, (cuuint64_t)g->Hi, (cuuint64_t)g->Di, (cuuint64_t)p.B * g->src_cb[si]};
    cuuint64_t gstr[3] = {(cuuint64_t)g->Wi * 16, (cuuint64_t)g->Wi * g->Hi * 16, (cuuint64_t)g->Wi * g->Hi * g->Di * 16};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int nb = 1; nb <= 2; ++nb) {
      cuuint32_t box[4] = {(cuuint32_t)(S3_TW * 4), (cuuint32_t)S3_ROWS, 1, (cuuint32_t)nb};
      CUresult r = encode(nb == 1 ? &maps.m[i] : &maps.m2[i], CU_TENSOR_MAP_DATA_TYPE_INT32, 4, const_cast<void*>(g->src[si]), gdim,
                          gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        e2e_set_error("conv_tc3: cuTensorMapEncodeTiled failed with %d (src %d, %d blocks)", (int)r, i, nb);
        return E2E_ERR_CUDA;
      }
    }
  }
  const int smem_bytes = p.b_region_bytes + p.stages * p.stage_bytes + stats_bytes + 1024;
  static E2eDevOnce attr_once;
  if (attr_once.first()) {
    cudaFuncAttributes fa;
    E2E_CUDA(cudaFuncGetAttributes(&fa, conv_tc3_kernel<true>));
    E2E_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024 - (int)fa.sharedSizeBytes));
    E2E_CUDA(cudaFuncGetAttributes(&fa, conv_tc3_kernel<false>));
    E2E_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024 - (int)fa.sharedSizeBytes));
  }
  int grid = e2e_num_sms();
  if (grid > p.n_tiles) grid = p.n_tiles;
  if ((p.Np >> 3) == S3_UB * (S3_EPI_WARPS / 4) && p.dst_cb == (p.Np >> 3))
    conv_tc3_kernel<true><<<grid, S3_THREADS, smem_bytes, st>>>(p, maps);
  else
    conv_tc3_kernel<false><<<grid, S3_THREADS, smem_bytes, st>>>(p, maps);
  E2E_LAUNCHED("conv_tc3");
  return E2E_OK;
}

// gs[0..n): column chunks of ONE GEMM (same sources, grids, channel entries, taps and destinations;
// different packed weights / column tables / widths), executed by one launch
static int conv_tc_fwd_impl(const e2e_gemm_t* gs, int n, cudaStream_t st, int* slots_out);

int e2e_conv_tc_fwd(const e2e_gemm_t* gs, int n, cudaStream_t st) { return conv_tc_fwd_impl(gs, n, st, nullptr); }

// number of statistics slots (CTAs x epilogue warps) the launch for gs[0..n) would write; 0 if it is not served
// by a tcgen05 kernel
int e2e_conv_tc_stats_slots(const e2e_gemm_t* gs, int n) {
  int slots = 0;
  if (conv_tc_fwd_impl(gs, n, nullptr, &slots) != E2E_OK) return 0;
  return slots;
}

static int conv_tc_fwd_impl(const e2e_gemm_t* gs, int n, cudaStream_t st, int* slots_out) {
  const e2e_gemm_t* g = gs;
  int form = e2e_conv_tc_supported(g);
  if (n < 1 || n > TC_MAX_CHUNKS) form = 0;
  int npmax = 0;
  for (int i = 0; i < n && form; ++i) {
    if (e2e_conv_tc_supported(gs + i) != form) form = 0;
    if (gs[i].Npad > npmax) npmax = gs[i].Npad;
  }
  if (!form) {
    e2e_set_error("conv_tc_fwd: call is neither a stride-1 3x3 halo-form nor a 1-tap GEMM with Npad <= 256");
    return E2E_ERR_UNSUPPORTED;
  }
  const bool halo = form == 1;
  auto encode = get_encode_fn();
  if (!encode) {
    e2e_set_error("conv_tc_fwd: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return E2E_ERR_CUDA;
  }
  if (form == 3) {
    if (g->accumulate) {
      e2e_set_error("conv_tc_fwd: the kw-stacked forward form does not accumulate");
      return E2E_ERR_UNSUPPORTED;
    }
    if (n != 1) {
      e2e_set_error("conv_tc_fwd: the kw-stacked form takes a single column chunk");
      return E2E_ERR_UNSUPPORTED;
    }
    return conv_tc3_launch(g, encode, st, slots_out);
  }
  // host copies of the plan tables are not available here: in the halo form taps must have
  // |dh|,|dw| <= 1 and channel entries dh = dw = 0; the Python plan builder guarantees it.
  TcParams p{};
  p.B = g->B; p.D = g->Do; p.H = g->Ho; p.W = g->Wo;
  p.Dsrc = g->Di; p.Ddst = g->Dd; p.Hd = g->Hd; p.Wd = g->Wd;
  p.isd = g->isd; p.ish = g->ish; p.isw = g->isw; p.ivh = g->ivh; p.ivw = g->ivw;
  p.osd = g->osd; p.osh = g->osh; p.osw = g->osw;
  p.n_cent = g->n_cent; p.Npad = npmax; p.ivd = g->ivd;
  p.n_chunks = n;
  for (int i = 0; i < n; ++i) {
    p.wpacked[i] = reinterpret_cast<const act16*>(gs[i].wpacked);
    p.cols[i] = gs[i].cols;
    p.npad[i] = gs[i].Npad;
  }
  const int n_taps = halo ? 9 : 1;
  int m = 256 / npmax;
  if (m > 4) m = 4;
  if (m < 1) m = 1;
  while (m > 1 && 8 * (m - 1) >= g->Wo) --m;          // do not tile wider than the row
  p.m = m;
  p.acc_stages = (2 * m * npmax <= 512) ? 2 : 1;
  p.rows = halo ? 18 : 16;
  {
    static int poll = -1;
    if (poll < 0) { const char* e = getenv("E2E_TC_POLL"); poll = e ? atoi(e) : 0; }
    p.poll = poll;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("E2E_TC_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
  }
  p.merged = (halo || g->isw == 1) ? 1 : 0;
  p.need_bounds = g->col_bounds & 7;
  const int rowpitch = (8 * m + (halo ? 2 : 0)) * 16;
  p.a_slab_bytes = (p.rows * rowpitch + 127) / 128 * 128;
  p.a_stage_bytes = 2 * p.a_slab_bytes;
  p.b_stage_bytes = n_taps * 2 * npmax * 16;
  int grid = e2e_num_sms();
  const int npairs = g->n_cent / 2;
  // K pairs per pipeline stage: 1 for the halo form (9*m MMAs per pair), up to 4 for the 1-tap form
  // (m MMAs per pair: the barrier hand-shake would dominate)
  int pps = halo ? 1 : 4;
  {
    static int force = -1;
    if (force < 0) { const char* e = getenv("E2E_TC_PPS"); force = e ? atoi(e) : 0; }
    if (force >= 1 && force <= 8) pps = force;
  }
  if (pps > npairs) pps = npairs;
  {
    static int sp = -1;
    if (sp < 0) { const char* e = getenv("E2E_TC_SERIAL"); sp = e ? atoi(e) : 1; }
    p.serial_prod = (halo && pps == 1) ? sp : 0;
  }
  // statistics requested (or queried): 8 epilogue warps x [2][Npad] fp32 accumulators come out of the stage budget
  const bool want_stats = g->stats != nullptr || slots_out != nullptr;
  const int stats_bytes = want_stats ? (8 * 2 * npmax * 4 + 127) / 128 * 128 : 0;
  const int budget = SMEM_BUDGET - stats_bytes;
  // resident weights: the whole packed operand of a chunk (all K pairs) fits beside >= 3 A stages
  p.b_res = 0; p.b_region_bytes = 0;
  {
    const int region = (npairs * p.b_stage_bytes + 1023) / 1024 * 1024;
    const int gres = (grid / n) * n;
    static int allow = -1;
    if (allow < 0) { const char* e = getenv("E2E_TC_BRES"); allow = e ? atoi(e) : 1; }
    if (allow && region <= 112 * 1024 && gres >= n && region + 3 * pps * p.a_stage_bytes <= budget) {
      p.b_res = 1; p.b_region_bytes = region; grid = gres;
    }
  }
  int stages;
  for (;;) {
    p.stage_bytes = p.b_res ? pps * p.a_stage_bytes : (pps * (p.a_stage_bytes + p.b_stage_bytes) + 127) / 128 * 128;
    stages = (budget - p.b_region_bytes) / p.stage_bytes;
    if (stages >= 3 || pps == 1) break;
    --pps;
  }
  p.pps = pps;
  if (stages > 8) stages = 8;
  if (stages < 2) {
    e2e_set_error("conv_tc_fwd: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
    return E2E_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  p.tiles_h = (p.H + TH - 1) / TH;
  p.tiles_w = (p.W + 8 * m - 1) / (8 * m);
  p.n_tiles = p.B * p.D * p.tiles_h * p.tiles_w;
  p.cents = g->cents; p.taps = g->taps;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < E2E_MAX_SRC; ++i) {
    const int si = i < g->n_src ? i : 0;
    p.src_cb[i] = g->src_cb[si];
    CUresult r;
    if (p.merged) {
      cuuint64_t gdim[4] = {(cuuint64_t)g->Wi * 4, (cuuint64_t)g->Hi, (cuuint64_t)g->Di, (cuuint64_t)p.B * g->src_cb[si]};
      cuuint64_t gstr[3] = {(cuuint64_t)g->Wi * 16, (cuuint64_t)g->Wi * g->Hi * 16, (cuuint64_t)g->Wi * g->Hi * g->Di * 16};
      cuuint32_t box[4] = {(cuuint32_t)((8 * m + (halo ? 2 : 0)) * 4), (cuuint32_t)(halo ? 18 : 16 * g->ish), 1, 1};
      cuuint32_t estr[4] = {1, (cuuint32_t)(halo ? 1 : g->ish), 1, 1};
      r = encode(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_INT32, 4, const_cast<void*>(g->src[si]), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      // traversal strides: ceil(box / stride) elements are loaded per dimension
      cuuint64_t gdim[5] = {8, (cuuint64_t)g->Wi, (cuuint64_t)g->Hi, (cuuint64_t)g->Di, (cuuint64_t)p.B * g->src_cb[si]};
      cuuint64_t gstr[4] = {16, (cuuint64_t)g->Wi * 16, (cuuint64_t)g->Wi * g->Hi * 16,
                            (cuuint64_t)g->Wi * g->Hi * g->Di * 16};
      cuuint32_t box[5] = {8, (cuuint32_t)(8 * m * g->isw), (cuuint32_t)(16 * g->ish), 1, 1};
      cuuint32_t estr[5] = {1, (cuuint32_t)g->isw, (cuuint32_t)g->ish, 1, 1};
      r = encode(&maps.m[i], E2E_TMAP_ACT, 5, const_cast<void*>(g->src[si]), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      e2e_set_error("conv_tc_fwd: cuTensorMapEncodeTiled failed with %d (src %d)", (int)r, i);
      return E2E_ERR_CUDA;
    }
  }
  for (int i = 0; i < E2E_MAX_SRC; ++i) {
    p.dst[i] = i < g->n_dst ? g->dst[i] : nullptr;
    p.dst_cb[i] = i < g->n_dst ? g->dst_cb[i] : 0;
  }
  const int smem_bytes = p.b_region_bytes + p.stages * p.stage_bytes + stats_bytes + 1024;
  p.stats = g->stats;
  p.stats_ctot = g->stats_ctot;
  p.stats_smem_off = p.b_region_bytes + p.stages * p.stage_bytes;
  p.accum = g->accumulate & ((1 << E2E_MAX_SRC) - 1);
  typedef void (*kern_t)(const TcParams, const TcMaps);
#define E2E_TC_ROW(H, MODE) {conv_tc_kernel<H, 1, MODE>, conv_tc_kernel<H, 2, MODE>, conv_tc_kernel<H, 3, MODE>, conv_tc_kernel<H, 4, MODE>}
  static const kern_t kerns[3][2][4] = {{E2E_TC_ROW(false, 0), E2E_TC_ROW(true, 0)},
                                        {E2E_TC_ROW(false, 1), E2E_TC_ROW(true, 1)},
                                        {E2E_TC_ROW(false, 2), E2E_TC_ROW(true, 2)}};
#undef E2E_TC_ROW
  static E2eDevOnce attr_once;
  if (attr_once.first()) {
    for (int md = 0; md < 3; ++md)
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 4; ++b) {
          cudaFuncAttributes fa;
          E2E_CUDA(cudaFuncGetAttributes(&fa, kerns[md][a][b]));
          E2E_CUDA(cudaFuncSetAttribute(kerns[md][a][b], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024 - (int)fa.sharedSizeBytes));
        }
  }
  if (!p.b_res && grid > p.n_tiles * p.n_chunks) grid = p.n_tiles * p.n_chunks;
  // 16 epilogue warps when a tile has many accumulator columns per MMA (short K loops: data gradients
  // of narrow layers); 8 otherwise (measured: more warps slow the MMA-bound and the store-bound cases)
  int epi = 8;
  if (halo && m * npmax > 6 * npairs * n_taps) epi = 16;
  {
    static int force = -1;
    if (force < 0) { const char* e = getenv("E2E_TC_EPI"); force = e ? atoi(e) : 0; }
    if (force == 8 || force == 16) epi = force;
  }
  if (want_stats) epi = 8;                 // the statistics accumulators are sized for 8 epilogue warps
  if (slots_out) { *slots_out = grid; return E2E_OK; }
  if (g->stats && g->accumulate) {
    e2e_set_error("conv_tc_fwd: fused statistics and accumulate cannot be combined");
    return E2E_ERR_UNSUPPORTED;
  }
  const int mode = g->stats ? 1 : (g->accumulate ? 2 : 0);
  kerns[mode][halo ? 1 : 0][m - 1]<<<grid, 64 + 32 * epi, smem_bytes, st>>>(p, maps);
  E2E_LAUNCHED("conv_tc_fwd");
  return E2E_OK;
}

// =====================================================================================
// tcgen05 weight-gradient kernel: K = voxels.
//
//   dwp[(cent e, tap t)][n][j] += sum_v  x_e[v * is + off_e + tap_t][j] * g[v][n]
//
// GEMM view per (entry group, tap): D[M = 16 channel entries x 8 ch = 128][N = Nc] += A^T B with
// K = the 128 voxels of a 16(H) x 8(W) tile, 16 voxels (2 rows) per tcgen05.mma.  Both operands
// are read straight from C8 slabs as MN-major no-swizzle UMMA operands: a core matrix is 8
// W-consecutive voxels x 8 channels (128 contiguous bytes); LBO = one tile/window row (next 8
// voxels), SBO = one slab (next 8 channels).
//   HALO form (stride-1 3x3 convs): the 9 taps read the same haloed x window through different
//     descriptor start addresses; one CTA owns `nt` taps of one entry group.
//   POINT form (1 tap: transposed convs, strided convs as (tap, block) entries): every entry is its
//     own TMA box (elementStrides = input stride, per-entry start offset); one CTA owns `G`
//     consecutive entry groups that share the gradient tile.
// The G * nt accumulators (Nc columns each, G * nt * Nc <= 512) live in TMEM across ALL voxel
// tiles of the CTA; one fp32 atomic flush at the end.  Jobs: (entry-group set) x (tap group) x
// (column chunk of <= 256 gradient channels) x (voxel split); one CTA per job.
// =====================================================================================
namespace {

struct WgParams {
  int B, D, H, W;                  // iteration grid = grid of the gradient operand
  int n_cent, Npad, ivd, ivh, ivw;
  int isd, ish, isw;
  int tiles_h, tiles_w, n_tiles;
  int n_mgj, G, n_tg, taps_per_group, n_taps, n_chunks, Nc, splits, tiles_per_split;
  int x_slab_bytes, x_rowpitch, x_rows, g_slab_bytes, x_bytes, stage_bytes, stages;
  int merged;
  int src_cb[E2E_MAX_SRC];
  int grad_cb;
  const e2e_centry_t* cents;
  const e2e_tap_t* taps;
  float* dwp;
  // direct mode: the flush adds into the gradient in the PARAMETER layout (grad[rowoff[n] + centoff[e*8+j] + tapoff[t]],
  // the scatter e2e_unpack_wgrad would do) instead of the packed scratch
  float* gout;
  const int32_t* rowoff;
  const int32_t* centoff;
  const int32_t* tapoff;
};

struct alignas(64) WgMaps {
  CUtensorMap x[E2E_MAX_SRC];
  CUtensorMap g;
  CUtensorMap x2[E2E_MAX_SRC];     // wgrad_gshift: boxes of two consecutive 8-channel blocks (see S3Maps)
};

constexpr int WG_MAX_ENT = 64;             // entries per CTA (G <= 4 groups of 16)

#ifndef E2E_WG_NPROD
#define E2E_WG_NPROD 7
#endif
constexpr int WG_NPROD = E2E_WG_NPROD;       // 3 or 7 producer threads, then the issuer warp, then 4 epilogue warps
constexpr int WG_THREADS = 32 * (WG_NPROD + 1 + 4);

template <bool HALO>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ WgParams p, const __grid_constant__ WgMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[20];
  __shared__ uint32_t tmem_base_s;
  __shared__ e2e_centry_t s_cents[WG_MAX_ENT];
  __shared__ int s_tapoff[9];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Nc = p.Nc, S = p.stages;
  int job = blockIdx.x;
  const int split = job % p.splits; job /= p.splits;
  const int chunk = job % p.n_chunks; job /= p.n_chunks;
  const int tg = job % p.n_tg;
  const int mgj = job / p.n_tg;
  const int e0 = mgj * p.G * 16;
  const int ne = min(p.G * 16, p.n_cent - e0);              // channel entries of this job
  const int ng = (ne + 15) >> 4;                            // entry groups of this job
  const int t0 = tg * p.taps_per_group;
  const int nt = min(p.taps_per_group, p.n_taps - t0);      // taps of this job
  const int n0 = chunk * Nc;                                // first gradient channel of this job
  const int ncol = min(Nc, p.Npad - n0);                    // multiple of 16
  const int tile_lo = split * p.tiles_per_split;
  const int tile_hi = min(p.n_tiles, tile_lo + p.tiles_per_split);
  const int my_tiles = tile_hi - tile_lo;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  const uint32_t done_bar = smem_u32(&bars[16]);

  if (threadIdx.x < ne) s_cents[threadIdx.x] = p.cents[e0 + threadIdx.x];
  if (threadIdx.x < 9) {
    int off = 0;
    if (HALO) {
      const e2e_tap_t t = p.taps[threadIdx.x];
      off = (1 + t.dh) * p.x_rowpitch + (1 + t.dw) * 16;
    }
    s_tapoff[threadIdx.x] = off;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  if (my_tiles > 0) {
    if (warp < WG_NPROD) {
      // TMA producers: ONE thread each (lane 0 of warps 0 .. WG_NPROD-1); the x boxes of the job's entries and the
      // gradient tile are dealt round-robin (see wgrad_gshift_kernel: a box costs its issuing thread ~100 cycles, the
      // 1-tap jobs have up to 65 boxes per 128-voxel tile against 8 - 32 MMAs).  Producer 0 arms the barrier.
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = (uint32_t)ne * (uint32_t)(p.x_rows * p.x_rowpitch) + (uint32_t)ncol * 256u;
        int t = tile_lo;
        int wt = t % p.tiles_w; t /= p.tiles_w;
        int ht = t % p.tiles_h; t /= p.tiles_h;
        int d = t % p.D, b = t / p.D;
        const int ngb = ncol == Nc ? 1 : (ncol >> 3);      // gradient boxes: one sized for Nc, or block by block
        const int nb = ne + ngb;
        for (int tile = tile_lo; tile < tile_hi; ++tile) {
          const int h0 = ht * TH, w0 = wt * 8;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (warp == 0) mbar_expect_tx(full_bar(stage), tx);
          const uint32_t sx = smem_base + stage * p.stage_bytes;
          for (int k = warp; k < nb; k += WG_NPROD) {
            if (k < ne) {
              const e2e_centry_t ce = s_cents[k];
              const int cb = b * p.src_cb[ce.src] + ce.blk;
              if (HALO)
                tma_load_4d(sx + k * p.x_slab_bytes, &maps.x[ce.src], full_bar(stage), (w0 - 1) * 4, h0 - 1,
                            d + p.ivd + ce.dd, cb);
              else if (p.merged)
                tma_load_4d(sx + k * p.x_slab_bytes, &maps.x[ce.src], full_bar(stage), (w0 + p.ivw + ce.dw) * 4,
                            h0 * p.ish + p.ivh + ce.dh, d * p.isd + p.ivd + ce.dd, cb);
              else
                tma_load_5d(sx + k * p.x_slab_bytes, &maps.x[ce.src], full_bar(stage), 0, w0 * p.isw + p.ivw + ce.dw,
                            h0 * p.ish + p.ivh + ce.dh, d * p.isd + p.ivd + ce.dd, cb);
            } else if (ncol == Nc) {
              // gradient tile: ncol / 8 channel blocks starting at block n0 / 8 (the 4-D box is sized for Nc)
              tma_load_4d(sx + p.x_bytes, &maps.g, full_bar(stage), w0 * 4, h0, d, b * p.grad_cb + (n0 >> 3));
            } else {
              const int q = k - ne;                          // a narrower last chunk is loaded block by block
              tma_load_4d(sx + p.x_bytes + q * p.g_slab_bytes, &maps.x[E2E_MAX_SRC - 1], full_bar(stage), w0 * 4, h0, d,
                          b * p.grad_cb + (n0 >> 3) + q);
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
          if (++wt == p.tiles_w) {
            wt = 0;
            if (++ht == p.tiles_h) {
              ht = 0;
              if (++d == p.D) { d = 0; ++b; }
            }
          }
        }
      }
    } else if (warp == WG_NPROD) {
      // MMA issuer.  D=f32, A=B=the 16-bit type, A and B MN-major, N=ncol, M=128
      const uint32_t idesc = (1u << 4) | (E2E_UMMA_FMT << 7) | (E2E_UMMA_FMT << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(ncol >> 3) << 17) | (8u << 24);
      int stage = 0, phase = 0;
      const uint32_t krow_units = (uint32_t)(2 * p.x_rowpitch) >> 4;     // 16 voxels = 2 window rows
      const uint32_t grp_units = (uint32_t)(16 * p.x_slab_bytes) >> 4;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait_warp(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sx = smem_base + stage * p.stage_bytes;
          const uint32_t sg = sx + p.x_bytes;
          const uint64_t a0 = make_desc(sx, p.x_rowpitch, p.x_slab_bytes);
          const uint64_t b0 = make_desc(sg, 128, p.g_slab_bytes);
          const uint32_t a_hi = (uint32_t)(a0 >> 32), b_hi = (uint32_t)(b0 >> 32);
          const uint32_t a_lo0 = (uint32_t)a0, b_lo0 = (uint32_t)b0;
          uint32_t acc = tmem_base;
          for (int gi = 0; gi < ng; ++gi) {
            for (int tl = 0; tl < nt; ++tl) {
              const uint32_t a_t = a_lo0 + (uint32_t)gi * grp_units + ((uint32_t)s_tapoff[t0 + tl] >> 4);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                tc_mma_f16_lh(acc, a_t + (uint32_t)k * krow_units, a_hi, b_lo0 + (uint32_t)(k * 16), b_hi, idesc,
                              (it | k) ? 1u : 0u);
              acc += (uint32_t)Nc;
            }
          }
          tc_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) tc_commit(done_bar);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;             // accumulator row = (entry r/8, channel r%8)
      const int el = r >> 3, j = r & 7;
      mbar_wait_warp(done_bar, 0, 1000);      // the whole K loop runs before this fires: poll rarely
      tc_fence_after();
      for (int gi = 0; gi < ng; ++gi) {
        const int e = e0 + gi * 16 + el;
        const bool valid = gi * 16 + el < ne;
        const int co = (p.gout && valid) ? __ldg(p.centoff + e * 8 + j) : -1;
        for (int tl = 0; tl < nt; ++tl) {
          const int t = t0 + tl;
          // dwp index of (entry e, tap t): slab = ((e/2)*n_taps + t)*2 + e%2
          float* base = p.dwp + ((size_t)(((e >> 1) * p.n_taps + t) * 2 + (e & 1)) * p.Npad + n0) * 8 + j;
          const int to = p.gout ? __ldg(p.tapoff + t) : 0;
          const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((gi * nt + tl) * Nc);
          for (int c = 0; c < ncol; c += 16) {
            uint32_t v[16];
            tc_ld16(acc + c, v);
            tc_wait_ld();
            if (p.gout) {
              if (co >= 0) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                  const int ro = __ldg(p.rowoff + n0 + c + u);
                  if (ro >= 0) atomicAdd(p.gout + (size_t)(ro + co + to), __uint_as_float(v[u]));
                }
              }
            } else if (valid) {
#pragma unroll
              for (int u = 0; u < 16; ++u) atomicAdd(base + (size_t)(c + u) * 8, __uint_as_float(v[u]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}


// -------------------------------------------------------------------------------------
// "g-shift" weight-gradient kernel for narrow layers (Npad <= 48): the tap shift is applied to the
// GRADIENT operand instead of x,
//     dW[kh,kw][e][n] = sum_u  x[u][e] * g[u - (kh-1, kw-1)][n],
// so one unshifted x tile (A, M = 16 entries x 8 ch) is multiplied against the three kw-shifted copies
// of the haloed gradient window stacked along N (B, N = 3 * Npad, copies loaded by three TMA boxes
// into [kw][blk][rows][8 vox]) -- 3 MMAs (one per kh: a row offset of the B start address) of N = 144
// instead of 9 MMAs of N = 48 per 16 voxels.  The SS-mode MMA is bound by shared-memory operand
// reads (ncu: l1tex__data_pipe_tc_wavefronts_mem_shared at 79 % of peak in wgrad_tc_kernel); this
// form reads A once per 3 taps: 68 instead of 132 wavefronts per (16 voxels x 9 taps).
// 3 accumulators (kh) x 3*Npad columns live in TMEM across all voxel tiles of the CTA.
// Tile = 8 (H) x 8 (W) voxels; jobs = (16-entry group) x (voxel split).
// -------------------------------------------------------------------------------------
struct WgsParams {
  int B, D, H, W;
  int n_cent, Npad, ivd, no_runs;
  int tiles_h, tiles_w, n_tiles;
  int n_groups, splits, tiles_per_split;
  int x_slab_bytes, g_slab_bytes, x_bytes, stage_bytes, stages;
  int src_cb[E2E_MAX_SRC];
  int grad_cb;
  const e2e_centry_t* cents;
  const e2e_tap_t* taps;
  float* dwp;
  float* gout;                      // direct mode, see WgParams
  const int32_t* rowoff;
  const int32_t* centoff;
  const int32_t* tapoff;
};

constexpr int GS_TH = 8;                    // tile rows
constexpr int GS_WROWS = GS_TH + 2;         // gradient window rows (H halo)

#ifndef E2E_GS_NPROD
#define E2E_GS_NPROD 7
#endif
constexpr int GS_NPROD = E2E_GS_NPROD;      // TMA producer threads (3 or 7: the epilogue warps start at a multiple of 4); then
                                            // one MMA issuer thread, then 4 epilogue warps.  A box costs a producer ~100 cycles of
                                            // dependent instructions, a 64-voxel tile has up to 19 boxes and 860 cycles of MMAs
constexpr int GS_THREADS = 32 * (GS_NPROD + 1 + 4);

__global__ void __launch_bounds__(GS_THREADS, 1)
wgrad_gshift_kernel(const __grid_constant__ WgsParams p, const __grid_constant__ WgMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[20];
  __shared__ uint32_t tmem_base_s;
  __shared__ e2e_centry_t s_cents[16];
  __shared__ int4 s_xbox[16];               // x boxes of a tile: {source | blocks << 8 | first entry << 16, depth offset, block, blocks per sample}
  __shared__ int s_nrun;
  __shared__ int s_tap_of[9];               // [kh * 3 + kw] -> tap index of the plan

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Npad = p.Npad, N3 = 3 * p.Npad, S = p.stages;
  int job = blockIdx.x;
  const int split = job % p.splits;
  const int mg = job / p.splits;
  const int e0 = mg * 16;
  const int ne = min(16, p.n_cent - e0);
  const int tile_lo = split * p.tiles_per_split;
  const int tile_hi = min(p.n_tiles, tile_lo + p.tiles_per_split);
  const int my_tiles = tile_hi - tile_lo;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  const uint32_t done_bar = smem_u32(&bars[16]);

  if (threadIdx.x < ne) s_cents[threadIdx.x] = p.cents[e0 + threadIdx.x];
  if (threadIdx.x == 32) {
    // x boxes of a tile: runs of one or two K entries (consecutive blocks of one source at the same depth offset)
    int nr = 0;
    for (int k = 0; k < ne;) {
      const e2e_centry_t ce = p.cents[e0 + k];
      int run = 1;
      if (k + 1 < ne && !p.no_runs) {
        const e2e_centry_t c2 = p.cents[e0 + k + 1];
        if (c2.src == ce.src && c2.dd == ce.dd && c2.blk == ce.blk + 1) run = 2;
      }
      s_xbox[nr++] = make_int4(ce.src | (run << 8) | (k << 16), ce.dd, ce.blk, p.src_cb[ce.src]);
      k += run;
    }
    s_nrun = nr;
  }
  if (threadIdx.x < 9) {
    const e2e_tap_t t = p.taps[threadIdx.x];
    s_tap_of[(t.dh + 1) * 3 + (t.dw + 1)] = threadIdx.x;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  if (my_tiles > 0) {
    if (warp < GS_NPROD) {
      // TMA producers: ONE thread each (lane 0 of warps 0 .. GS_NPROD-1); the ne x boxes and 3 gradient boxes of a
      // tile are dealt round-robin to them.  (Issuing the 19 boxes "in parallel" from 19 lanes of one warp compiles to
      // a serial per-lane ELECT / R2UR / UTMALDG loop: ~1500 cycles per 64-voxel tile against 860 cycles of MMAs.)
      // Producer 0 arms the barrier with the bytes of ALL boxes; the others' bytes may land first (the transaction
      // count goes negative until the expect arrives, the phase cannot complete before that one pending arrival).
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = (uint32_t)ne * (uint32_t)(GS_TH * 128) + 3u * (uint32_t)(Npad >> 3) * (uint32_t)p.g_slab_bytes;
        int t = tile_lo;
        int wt = t % p.tiles_w; t /= p.tiles_w;
        int ht = t % p.tiles_h; t /= p.tiles_h;
        int d = t % p.D, b = t / p.D;
        const int nr = s_nrun, nb = nr + 3;
        for (int tile = tile_lo; tile < tile_hi; ++tile) {
          const int h0 = ht * GS_TH, w0 = wt * 8;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (warp == 0) mbar_expect_tx(full_bar(stage), tx);
          const uint32_t sx = smem_base + stage * p.stage_bytes;
          for (int k = warp; k < nb; k += GS_NPROD) {
            if (k < nr) {
              const int4 bx = s_xbox[k];
              const int src = bx.x & 0xff, e = bx.x >> 16;
              tma_load_4d(sx + e * p.x_slab_bytes, ((bx.x >> 8) & 0xff) == 2 ? &maps.x2[src] : &maps.x[src], full_bar(stage), w0 * 4,
                          h0, d + p.ivd + bx.y, b * bx.w + bx.z);
            } else {
              // three kw-shifted copies of the haloed gradient window: copy kw holds g[., w - (kw - 1)]
              const int kw = k - nr;
              tma_load_4d(sx + p.x_bytes + kw * (Npad >> 3) * p.g_slab_bytes, &maps.g, full_bar(stage),
                          (w0 - (kw - 1)) * 4, h0 - 1, d, b * p.grad_cb);
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
          if (++wt == p.tiles_w) {
            wt = 0;
            if (++ht == p.tiles_h) {
              ht = 0;
              if (++d == p.D) { d = 0; ++b; }
            }
          }
        }
      }
    } else if (warp == GS_NPROD) {
      // MMA issuer: one thread.  D=f32, A=B=the 16-bit type, A and B MN-major, N = 3*Npad, M=128
      if (lane == 0) {
        const uint32_t idesc = (1u << 4) | (E2E_UMMA_FMT << 7) | (E2E_UMMA_FMT << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(N3 >> 3) << 17) | (8u << 24);
        int stage = 0;
        uint32_t phase = 0;
        const uint64_t a00 = make_desc(smem_base, 128, p.x_slab_bytes);
        const uint64_t b00 = make_desc(smem_base + p.x_bytes, 128, p.g_slab_bytes);
        const uint32_t a_hi = (uint32_t)(a00 >> 32), b_hi = (uint32_t)(b00 >> 32);
        const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
        for (int it = 0; it < my_tiles; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_lo0 = (uint32_t)a00 + (uint32_t)stage * stage_units;
          const uint32_t b_lo0 = (uint32_t)b00 + (uint32_t)stage * stage_units;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int k = 0; k < GS_TH / 2; ++k)
              // x rows 2k, 2k+1 of the tile  x  gradient window rows 2k + 2 - kh, +1 (row = 128 B = 8 units)
              tc_mma_f16_lh(tmem_base + (uint32_t)(kh * N3), a_lo0 + (uint32_t)(2 * k * 8), a_hi,
                            b_lo0 + (uint32_t)((2 * k + 2 - kh) * 8), b_hi, idesc, (it | k) ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        tc_commit(done_bar);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;             // accumulator row = (entry r/8, channel r%8)
      const int el = r >> 3, j = r & 7;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int e = e0 + el;
      const bool valid = el < ne;
      const int co = (p.gout && valid) ? __ldg(p.centoff + e * 8 + j) : -1;
      for (int kh = 0; kh < 3; ++kh) {
        for (int kw = 0; kw < 3; ++kw) {
          const int t = s_tap_of[kh * 3 + kw];
          float* base = p.dwp + ((size_t)(((e >> 1) * 9 + t) * 2 + (e & 1)) * Npad) * 8 + j;
          const int to = p.gout ? __ldg(p.tapoff + t) : 0;
          const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kh * N3 + kw * Npad);
          for (int c = 0; c < Npad; c += 16) {
            uint32_t v[16];
            tc_ld16(acc + c, v);
            tc_wait_ld();
            if (p.gout) {
              if (co >= 0) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                  const int ro = __ldg(p.rowoff + c + u);
                  if (ro >= 0) atomicAdd(p.gout + (size_t)(ro + co + to), __uint_as_float(v[u]));
                }
              }
            } else if (valid) {
#pragma unroll
              for (int u = 0; u < 16; ++u) atomicAdd(base + (size_t)(c + u) * 8, __uint_as_float(v[u]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

}  // namespace

// 1: stride-1 3x3 halo form, 2: 1-tap point form, 0: not served
int e2e_wgrad_tc_supported(const e2e_wgrad_t* p) {
  if (p->Npad % 16 != 0) return 0;
  if (p->n_taps == 9) {
    if (p->isd != 1 || p->ish != 1 || p->isw != 1 || p->ivh != 0 || p->ivw != 0) return 0;
    if (p->Do != p->Di || p->Ho != p->Hi || p->Wo != p->Wi) return 0;
    return 1;
  }
  if (p->n_taps == 1) {
    if (p->isd < 1 || p->ish < 1 || p->isw < 1 || p->isd > 8 || p->ish > 8 || p->isw > 8) return 0;
    return 2;
  }
  return 0;
}

static int wgrad_gshift_launch(const e2e_wgrad_t* g, PFN_cuTensorMapEncodeTiled_v12000 encode, cudaStream_t st) {
  WgsParams p{};
  p.B = g->B; p.D = g->Do; p.H = g->Ho; p.W = g->Wo;
  p.n_cent = g->n_cent; p.Npad = g->Npad; p.ivd = g->ivd;
  p.tiles_h = (p.H + GS_TH - 1) / GS_TH;
  p.tiles_w = (p.W + 7) / 8;
  p.n_tiles = p.B * p.D * p.tiles_h * p.tiles_w;
  p.n_groups = (g->n_cent + 15) / 16;
  p.x_slab_bytes = GS_TH * 128;
  p.g_slab_bytes = GS_WROWS * 128;
  p.x_bytes = 16 * p.x_slab_bytes;
  p.stage_bytes = (p.x_bytes + 3 * (g->Npad / 8) * p.g_slab_bytes + 127) / 128 * 128;
  int stages = SMEM_BUDGET / p.stage_bytes;
  if (stages > 6) stages = 6;
  p.stages = stages;
  int splits = e2e_num_sms() / p.n_groups;
  if (splits > p.n_tiles) splits = p.n_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.splits = splits;
  p.cents = g->cents; p.taps = g->taps; p.dwp = g->dwp; p.grad_cb = g->grad_cb;
  p.gout = g->grad_out; p.rowoff = g->rowoff; p.centoff = g->centoff; p.tapoff = g->tapoff;
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < E2E_MAX_SRC; ++i) {
    const int si = i < g->n_src ? i : 0;
    p.src_cb[i] = g->src_cb[si];
    cuuint64_t gdim[4] = {(cuuint64_t)g->Wi * 4, (cuuint64_t)g->Hi, (cuuint64_t)g->Di, (cuuint64_t)p.B * g->src_cb[si]};
    cuuint64_t gstr[3] = {(cuuint64_t)g->Wi * 16, (cuuint64_t)g->Wi * g->Hi * 16, (cuuint64_t)g->Wi * g->Hi * g->Di * 16};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int nb = 1; nb <= 2; ++nb) {
      cuuint32_t box[4] = {32, (cuuint32_t)GS_TH, 1, (cuuint32_t)nb};
      CUresult r = encode(nb == 1 ? &maps.x[i] : &maps.x2[i], CU_TENSOR_MAP_DATA_TYPE_INT32, 4, const_cast<void*>(g->src[si]), gdim,
                          gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        e2e_set_error("wgrad_gshift: cuTensorMapEncodeTiled failed with %d (src %d, %d blocks)", (int)r, i, nb);
        return E2E_ERR_CUDA;
      }
    }
  }
  {
    static int runs = -1;
    if (runs < 0) { const char* e = getenv("E2E_TC3_RUNS"); runs = e ? atoi(e) : 1; }
    p.no_runs = runs ? 0 : 1;
  }
  {
    cuuint64_t gdim[4] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.B * g->grad_cb};
    cuuint64_t gstr[3] = {(cuuint64_t)p.W * 16, (cuuint64_t)p.W * p.H * 16, (cuuint64_t)p.W * p.H * p.D * 16};
    cuuint32_t box[4] = {32, (cuuint32_t)GS_WROWS, 1, (cuuint32_t)(g->Npad / 8)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&maps.g, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, const_cast<void*>(g->grad), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      e2e_set_error("wgrad_gshift: cuTensorMapEncodeTiled failed with %d (grad)", (int)r);
      return E2E_ERR_CUDA;
    }
  }
  const int smem_bytes = p.stages * p.stage_bytes + 1024;
  static E2eDevOnce attr_once;
  if (attr_once.first()) {
    E2E_CUDA(cudaFuncSetAttribute(wgrad_gshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 6 * 1024));
  }
  wgrad_gshift_kernel<<<p.n_groups * p.splits, GS_THREADS, smem_bytes, st>>>(p, maps);
  E2E_LAUNCHED("wgrad_gshift");
  return E2E_OK;
}

int e2e_wgrad_tc(const e2e_wgrad_t* g, cudaStream_t st) {
  const int form = e2e_wgrad_tc_supported(g);
  if (!form) {
    e2e_set_error("wgrad_tc: call is neither a stride-1 3x3 halo-form nor a 1-tap weight gradient");
    return E2E_ERR_UNSUPPORTED;
  }
  const bool halo = form == 1;
  auto encode = get_encode_fn();
  if (!encode) {
    e2e_set_error("wgrad_tc: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return E2E_ERR_CUDA;
  }
  {
    static int allow = -1;
    if (allow < 0) { const char* e = getenv("E2E_TC_GSHIFT"); allow = e ? atoi(e) : 1; }
    if (halo && allow && g->Npad <= 48 && g->n_src <= E2E_MAX_SRC) return wgrad_gshift_launch(g, encode, st);
  }
  WgParams p{};
  p.B = g->B; p.D = g->Do; p.H = g->Ho; p.W = g->Wo;
  p.n_cent = g->n_cent; p.Npad = g->Npad; p.ivd = g->ivd; p.ivh = g->ivh; p.ivw = g->ivw;
  p.isd = g->isd; p.ish = g->ish; p.isw = g->isw;
  p.n_taps = g->n_taps;
  p.merged = (halo || g->isw == 1) ? 1 : 0;
  p.tiles_h = (p.H + TH - 1) / TH;
  p.tiles_w = (p.W + 7) / 8;
  p.n_tiles = p.B * p.D * p.tiles_h * p.tiles_w;
  // column chunks: balanced, <= 256 (halo: <= 96 when more than one accumulator tap has to share
  // the 512 TMEM columns -- x windows are re-read per tap group, gradient tiles per entry group)
  int maxc = 256;
  if (halo && g->Npad > 96) maxc = (g->Npad % 96 == 0) ? 96 : ((g->Npad % 80 == 0) ? 160 : 128);
  p.n_chunks = (g->Npad + maxc - 1) / maxc;
  p.Nc = ((g->Npad / 16 + p.n_chunks - 1) / p.n_chunks) * 16;
  p.n_chunks = (g->Npad + p.Nc - 1) / p.Nc;
  const int n_groups = (g->n_cent + 15) / 16;
  p.x_rowpitch = halo ? 10 * 16 : 8 * 16;
  p.x_rows = halo ? 18 : 16;
  p.x_slab_bytes = (p.x_rows * p.x_rowpitch + 127) / 128 * 128;
  p.g_slab_bytes = 128 * 16;
  if (halo) {
    p.G = 1;
    p.taps_per_group = 512 / p.Nc;
    if (p.taps_per_group > 9) p.taps_per_group = 9;
    p.n_tg = (9 + p.taps_per_group - 1) / p.taps_per_group;
    p.taps_per_group = (9 + p.n_tg - 1) / p.n_tg;            // balance the groups
  } else {
    p.taps_per_group = 1; p.n_tg = 1;
    int G = 512 / p.Nc;
    if (G > WG_MAX_ENT / 16) G = WG_MAX_ENT / 16;
    if (G > n_groups) G = n_groups;
    // two stages of (G * 16 windows + gradient tile) must fit
    while (G > 1 && 2 * (G * 16 * p.x_slab_bytes + (p.Nc / 8) * p.g_slab_bytes) > SMEM_BUDGET) --G;
    p.G = G;
  }
  p.n_mgj = (n_groups + p.G - 1) / p.G;
  p.x_bytes = p.G * 16 * p.x_slab_bytes;
  p.stage_bytes = (p.x_bytes + (p.Nc / 8) * p.g_slab_bytes + 127) / 128 * 128;
  int stages = SMEM_BUDGET / p.stage_bytes;
  if (stages > 4) stages = 4;
  if (stages < 2) {
    e2e_set_error("wgrad_tc: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
    return E2E_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const int jobs = p.n_mgj * p.n_tg * p.n_chunks;
  int splits = e2e_num_sms() / jobs;            // one wave: jobs * splits <= SMs (one CTA per SM)
  if (splits > p.n_tiles) splits = p.n_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.splits = splits;
  p.cents = g->cents; p.taps = g->taps; p.dwp = g->dwp; p.grad_cb = g->grad_cb;
  p.gout = g->grad_out; p.rowoff = g->rowoff; p.centoff = g->centoff; p.tapoff = g->tapoff;
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int n_xmaps = g->n_src < E2E_MAX_SRC - 1 ? E2E_MAX_SRC - 1 : E2E_MAX_SRC;
  if (g->n_src > E2E_MAX_SRC - 1 && p.Npad % p.Nc != 0) {
    e2e_set_error("wgrad_tc: ragged column chunks need a free tensor-map slot (n_src <= %d)", E2E_MAX_SRC - 1);
    return E2E_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < n_xmaps; ++i) {
    const int si = i < g->n_src ? i : 0;
    p.src_cb[i] = g->src_cb[si];
    CUresult r;
    if (p.merged) {
      cuuint64_t gdim[4] = {(cuuint64_t)g->Wi * 4, (cuuint64_t)g->Hi, (cuuint64_t)g->Di, (cuuint64_t)p.B * g->src_cb[si]};
      cuuint64_t gstr[3] = {(cuuint64_t)g->Wi * 16, (cuuint64_t)g->Wi * g->Hi * 16, (cuuint64_t)g->Wi * g->Hi * g->Di * 16};
      cuuint32_t box[4] = {(cuuint32_t)(halo ? 40 : 32), (cuuint32_t)(halo ? 18 : 16 * g->ish), 1, 1};
      cuuint32_t estr[4] = {1, (cuuint32_t)(halo ? 1 : g->ish), 1, 1};
      r = encode(&maps.x[i], CU_TENSOR_MAP_DATA_TYPE_INT32, 4, const_cast<void*>(g->src[si]), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gdim[5] = {8, (cuuint64_t)g->Wi, (cuuint64_t)g->Hi, (cuuint64_t)g->Di, (cuuint64_t)p.B * g->src_cb[si]};
      cuuint64_t gstr[4] = {16, (cuuint64_t)g->Wi * 16, (cuuint64_t)g->Wi * g->Hi * 16,
                            (cuuint64_t)g->Wi * g->Hi * g->Di * 16};
      cuuint32_t box[5] = {8, (cuuint32_t)(8 * g->isw), (cuuint32_t)(16 * g->ish), 1, 1};
      cuuint32_t estr[5] = {1, (cuuint32_t)g->isw, (cuuint32_t)g->ish, 1, 1};
      r = encode(&maps.x[i], E2E_TMAP_ACT, 5, const_cast<void*>(g->src[si]), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      e2e_set_error("wgrad_tc: cuTensorMapEncodeTiled failed with %d (src %d)", (int)r, i);
      return E2E_ERR_CUDA;
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: box of Nc / 8 channel blocks; pass 1: single-block box for a ragged last chunk
    if (pass == 1 && p.Npad % p.Nc == 0) break;
    cuuint64_t gdim[4] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.B * g->grad_cb};
    cuuint64_t gstr[3] = {(cuuint64_t)p.W * 16, (cuuint64_t)p.W * p.H * 16, (cuuint64_t)p.W * p.H * p.D * 16};
    cuuint32_t box[4] = {32, 16, 1, (cuuint32_t)(pass == 0 ? p.Nc / 8 : 1)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(pass == 0 ? &maps.g : &maps.x[E2E_MAX_SRC - 1], CU_TENSOR_MAP_DATA_TYPE_INT32, 4,
                        const_cast<void*>(g->grad), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      e2e_set_error("wgrad_tc: cuTensorMapEncodeTiled failed with %d (grad)", (int)r);
      return E2E_ERR_CUDA;
    }
  }
  const int smem_bytes = p.stages * p.stage_bytes + 1024;
  static E2eDevOnce attr_once;
  if (attr_once.first()) {
    E2E_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 6 * 1024));
    E2E_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 6 * 1024));
  }
  if (halo)
    wgrad_tc_kernel<true><<<jobs * splits, WG_THREADS, smem_bytes, st>>>(p, maps);
  else
    wgrad_tc_kernel<false><<<jobs * splits, WG_THREADS, smem_bytes, st>>>(p, maps);
  E2E_LAUNCHED("wgrad_tc");
  return E2E_OK;
}
