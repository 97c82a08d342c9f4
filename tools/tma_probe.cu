// TMA throughput probe (sm_100a): streams tensor boxes of the shapes the conv kernels use from a C8
// activation tensor into shared memory (no MMA, no stores) and reports the aggregate GB/s, next to
// contiguous cp.async.bulk copies of the same size.  Answers "is the TMA box path the limiter?".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_probe tools/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b), "r"(ph) : "memory");
  }
}

struct alignas(64) Map { CUtensorMap m; };

// mode 0: 4-D tensor boxes; mode 1: contiguous bulk copies of box_bytes
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ Map map, const uint8_t* base, int mode, int box_bytes, int slot_bytes,
                                               int boxes_per_stage, int iters, int stages, int W, int H, int D, int NB,
                                               int box_w, int box_h, int lanes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(s32(&bars[s]), 1); mbar_init(s32(&bars[16 + s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const uint32_t sb = (s32(smem) + 1023u) & ~1023u;
  const int stage_bytes = boxes_per_stage * slot_bytes;     // slots are 128-byte aligned (TMA destination)
  const int stage_tx = boxes_per_stage * box_bytes;
  if (warp == 0) {
    int st = 0, ph = 0;
    uint32_t rng = blockIdx.x * 7919u + 13u;
    for (int it = 0; it < iters; ++it) {
      if (lane == 0) {
        mbar_wait(s32(&bars[16 + st]), ph ^ 1);
        mbar_expect(s32(&bars[st]), (uint32_t)stage_tx);
      }
      __syncwarp();
      // a pseudo-random tile position per iteration (same for the whole stage, different channel blocks)
      rng = rng * 1664525u + 1013904223u;
      const uint32_t r = __shfl_sync(0xffffffffu, rng, 0);
      const int w0 = (int)((r >> 4) % (uint32_t)(W - box_w + 1)), h0 = (int)((r >> 12) % (uint32_t)(H - box_h + 1));
      const int d = (int)((r >> 20) % (uint32_t)D);
      for (int i = (lanes > 1 ? lane : 0); i < boxes_per_stage; i += (lanes > 1 ? lanes : 1)) {
        if (lanes == 1 && lane != 0) break;
        const uint32_t dst = sb + st * stage_bytes + i * slot_bytes;
        const int blk = (int)((r + 17u * i) % (uint32_t)NB);
        if (mode == 0) {
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"((uint64_t)&map.m), "r"(s32(&bars[st])), "r"(w0 * 4), "r"(h0), "r"(d), "r"(blk) : "memory");
        } else {
          const size_t off = ((((size_t)blk * D + d) * H + h0) * (size_t)W + w0) * 16;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(base + (off & ~(size_t)15)), "r"(box_bytes), "r"(s32(&bars[st])) : "memory");
        }
      }
      if (++st == stages) { st = 0; ph ^= 1; }
    }
  } else {
    int st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (lane == 0) { mbar_wait(s32(&bars[st]), ph); mbar_arrive(s32(&bars[16 + st])); }
      __syncwarp();
      if (++st == stages) { st = 0; ph ^= 1; }
    }
  }
}

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fnp;
  struct Case { const char* name; int NB, D, H, W; };
  const Case tensors[2] = {{"HBM-sized (1.26 GB)", 48, 64, 160, 160}, {"L2-sized (79 MB)", 12, 16, 160, 160}};
  struct Box { const char* name; int bw, bh, per_stage, lanes; };
  const Box boxes[] = {{"halo m=4   [18][34] 9.8 KB x2", 34, 18, 2, 1},   {"halo m=4   [18][34] x2, 2 lanes", 34, 18, 2, 2},
                       {"halo m=2   [18][18] 5.2 KB x2", 18, 18, 2, 1},   {"stacked    [10][32] 5.1 KB x4, 4 lanes", 32, 10, 4, 4},
                       {"stacked    [6][32]  3.1 KB x8, 8 lanes", 32, 6, 8, 8}, {"wgrad win  [18][10] 2.9 KB x16, 16 lanes", 10, 18, 16, 16},
                       {"point      [16][8]  2.0 KB x8, 8 lanes", 8, 16, 8, 8},  {"big        [32][64] 32 KB x1", 64, 32, 1, 1}};
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (const Case& t : tensors) {
    const size_t bytes = (size_t)t.NB * t.D * t.H * t.W * 16;
    uint8_t* buf;
    CK(cudaMalloc(&buf, bytes + (1 << 20)));
    CK(cudaMemset(buf, 1, bytes + (1 << 20)));
    printf("== source tensor %s: [%d blk][%d][%d][%d][8ch bf16]\n", t.name, t.NB, t.D, t.H, t.W);
    for (const Box& b : boxes) {
      Map map;
      memset(&map, 0, sizeof(map));
      cuuint64_t gdim[4] = {(cuuint64_t)t.W * 4, (cuuint64_t)t.H, (cuuint64_t)t.D, (cuuint64_t)t.NB};
      cuuint64_t gstr[3] = {(cuuint64_t)t.W * 16, (cuuint64_t)t.W * t.H * 16, (cuuint64_t)t.W * t.H * t.D * 16};
      cuuint32_t box[4] = {(cuuint32_t)(b.bw * 4), (cuuint32_t)b.bh, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      if (encode(&map.m, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return 1;
      }
      const int box_bytes = b.bw * b.bh * 16;
      const int slot_bytes = (box_bytes + 127) / 128 * 128;
      const int stage_bytes = b.per_stage * slot_bytes;
      int stages = (180 * 1024) / stage_bytes;
      if (stages > 8) stages = 8;
      const int iters = 4000;
      for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<<<148, 64, stages * stage_bytes + 1024>>>(map, buf, mode, box_bytes, slot_bytes, b.per_stage, 200, stages, t.W, t.H, t.D, t.NB, b.bw, b.bh, b.lanes);
        cudaEventRecord(e0);
        probe<<<148, 64, stages * stage_bytes + 1024>>>(map, buf, mode, box_bytes, slot_bytes, b.per_stage, iters, stages, t.W, t.H, t.D, t.NB, b.bw, b.bh, b.lanes);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gb = 148.0 * iters * (double)(b.per_stage * box_bytes) / 1e9;
        printf("  %-44s %s  %8.0f GB/s  (%5.1f B/clk/SM @1.92 GHz, %d stages)\n", b.name, mode ? "bulk  " : "tensor",
               gb / (ms / 1e3), gb * 1e9 / (ms / 1e3) / 148 / 1.92e9, stages);
      }
    }
    cudaFree(buf);
  }
  return 0;
}
