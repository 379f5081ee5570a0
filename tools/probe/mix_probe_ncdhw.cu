// Developer microbenchmark: do TMA boxes and 16-byte cp.async gathers of an NCDHW level ([C][D][H][W] fp32, runs of 64
// bytes) add up when one CTA per SM runs both at once?  Warp 0 lane 0 streams TMA boxes {16 x, 8 rows, 32 channels}
// through a 2-slot ring; warps 1..4 gather the same kind of tile with cp.async (lane = 16-byte piece, loop over 32
// channels), 4 groups in flight per warp.  mode 1 = TMA only, 2 = cp.async only, 3 = both.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/mix_probe_ncdhw tools/probe/mix_probe_ncdhw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned s_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned m, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned m, unsigned b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(b) : "memory"); }
__device__ int g_timeout;
__device__ __forceinline__ void mbar_wait(unsigned m, unsigned par) {
  unsigned ok = 0;
  for (int i = 0; i < (1 << 20); ++i) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(m), "r"(par) : "memory");
    if (ok) return;
  }
  g_timeout = 1;
}
__device__ __forceinline__ void tma5(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, int c4, unsigned m) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(m) : "memory");
}

constexpr int BX = 16, BY = 8, BC = 32;
constexpr int BOX_BYTES = BX * BY * BC * 4;   // 16 KB
constexpr int NSLOT = 4;                       // TMA ring
constexpr int CPW = 4;                         // cp.async warps
constexpr int CPG = 4;                         // groups in flight per warp
constexpr int SMEM = NSLOT * BOX_BYTES + CPW * CPG * BOX_BYTES / 4 + 64;   // cp.async tiles: 8 channels each (4 KB)

__global__ void __launch_bounds__(160) probe(const __grid_constant__ CUtensorMap map, const float *feats, int iters, int mode, int C, int W, int H, int D,
                                             unsigned long long *cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const unsigned bar0 = s_u32(smem + SMEM - 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) mbar_init(bar0 + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  unsigned rng = (blockIdx.x * 160 + threadIdx.x) * 2654435761u + 12345u;
  if (warp == 0) {
    if (lane == 0 && (mode & 1)) {
      int issued = 0, done = 0;
      while (done < iters) {
        while (issued < iters && issued - done < NSLOT) {
          const int s = issued & (NSLOT - 1);
          mbar_expect(bar0 + s * 8, BOX_BYTES);
          rng = rng * 1664525u + 1013904223u;
          const int x = ((rng >> 8) & 15) * 4, y = (rng >> 16) & 63, z = (rng >> 4) & 31, c = ((rng >> 24) & (C / BC - 1)) * BC;
          tma5(s_u32(smem + s * BOX_BYTES), &map, x, y, z, c, 0, bar0 + s * 8);
          ++issued;
        }
        mbar_wait(bar0 + (done & (NSLOT - 1)) * 8, (done / NSLOT) & 1);
        ++done;
      }
    }
  } else if (mode & 2) {
    // each group: a tile of 8 rows x 16 floats x 8 channels = 4 KB: lane = (row, quarter): 32 pieces per channel
    unsigned char *mine = smem + NSLOT * BOX_BYTES + (warp - 1) * CPG * (BOX_BYTES / 4);
    const int row = lane >> 2, q = lane & 3;
    // the cp.async warps together move `iters` 16 KB tiles' worth: iters * 4 groups over CPW warps
    const int groups = iters * 4 / CPW;
    for (int g = 0; g < groups; ++g) {
      rng = rng * 1664525u + 1013904223u;
      const unsigned r = __shfl_sync(0xffffffffu, rng, 0);
      const int x = ((r >> 8) & 15) * 4, y = (r >> 16) & 63, z = (r >> 4) & 31, c = ((r >> 24) & (C / 8 - 1)) * 8;
      const float *src = feats + (((long long)c * D + z) * H + (y + row)) * W + x + q * 4;
      const unsigned dst = s_u32(mine + (g % CPG) * (BOX_BYTES / 4)) + lane * 16;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ch * 512), "l"(src + (long long)ch * D * H * W) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(CPG - 1) : "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int C = 256, W = 128, H = 128, D = 40;
  float *feats;
  const size_t n = (size_t)C * W * H * D;
  CK(cudaMalloc(&feats, n * 4));
  CK(cudaMemset(feats, 0, n * 4));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long *cyc;
  CK(cudaMalloc(&cyc, sms * 8));
  CUtensorMap map;
  const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, 1};
  const cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)D * H * W * 4, (cuuint64_t)C * D * H * W * 4};
  const cuuint32_t box[5] = {BX, BY, 1, BC, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, feats, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const int iters = 400;
  for (int mode = 1; mode <= 3; ++mode) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    probe<<<sms, 160, SMEM>>>(map, feats, 50, mode, C, W, H, D, cyc);
    cudaEventRecord(e0);
    probe<<<sms, 160, SMEM>>>(map, feats, iters, mode, C, W, H, D, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    int to = 0;
    CK(cudaMemcpyFromSymbol(&to, g_timeout, 4));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)sms * iters * BOX_BYTES * ((mode & 1) + ((mode >> 1) & 1));
    printf("mode %d (1 = TMA, 2 = cp.async 16 B, 3 = both)%s: %7.1f us  %6.2f TB/s  %6.1f B/clk/SM @1.9GHz\n", mode, to ? " TIMEOUT" : "", ms * 1e3,
           bytes / (ms * 1e-3) / 1e12, bytes / sms / (ms * 1e-3 * 1.9e9));
  }
  return 0;
}
