// NCDHW twin of tma_probe.cu (levels [B][C][D][H][W], boxes = bc channels x by rows x bx voxels: runs of bx*4 bytes).
// Developer microbenchmark: how fast does one SM's TMA unit deliver 5-D boxes of a channels-last level
// [B][D][H][W][C] fp32 when the contiguous run per voxel is 256 B / 512 B / 1 KB, and how does that depend on the
// number of boxes in flight?  One issuing thread per CTA, one CTA per SM, ring of NS slots.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_probe_ncdhw tools/probe/tma_probe_ncdhw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned s_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned m, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned m, unsigned b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(b) : "memory"); }
__device__ int g_mode;
__device__ int g_timeout;
__device__ int g_mis;  // 0 = try_wait, 1 = test_wait spin, 2 = try_wait with a 64 ns suspend hint
__device__ __forceinline__ void mbar_wait(unsigned m, unsigned par, int mode) {
  unsigned ok = 0;
  for (int i = 0; i < (1 << 18); ++i) {
    if (mode == 0)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(m), "r"(par) : "memory");
    else if (mode == 1)
      asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(m), "r"(par) : "memory");
    else
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(m), "r"(par), "r"(64u) : "memory");
    if (ok) return;
  }
  g_timeout = 1;
}
__device__ __forceinline__ void tma5(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, int c4, unsigned m) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(m) : "memory");
}

// every CTA walks `iters` boxes at pseudo-random positions; ops_per_box TMA ops fill one slot
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, int ns, int slot_bytes, int box_bytes, int ops_per_slot,
                                             int iters, int bc, int bx, int by, int C, int W, int H, int D, unsigned long long *cycles, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned bar0 = s_u32(smem + ns * slot_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) mbar_init(bar0 + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  unsigned rng = blockIdx.x * 2654435761u + 12345u;
  const int lg = 31 - __clz(ns);
  long long t0 = clock64();
  long long t_issue = 0, t_wait = 0;
  int issued = 0, done = 0;
  while (done < iters) {
    long long ta = clock64();
    while (issued < iters && issued - done < ns) {
      const int s = issued & (ns - 1);
      mbar_expect(bar0 + s * 8, (unsigned)(box_bytes * ops_per_slot));
      for (int o = 0; o < ops_per_slot; ++o) {
        rng = rng * 1664525u + 1013904223u;
        const int x = (rng >> 8) & 63, y = (rng >> 16) & 63, z = (rng >> 4) & 31, c = ((rng >> 24) & (C / bc - 1)) * bc;
        tma5(s_u32(smem + s * slot_bytes + o * box_bytes), &map, (g_mis < 0 ? (x & ~3) : x + (rng & g_mis)), y, z, c, 0, bar0 + s * 8);
      }
      ++issued;
    }
    long long tb = clock64();
    const int s = done & (ns - 1);
    mbar_wait(bar0 + s * 8, (done >> lg) & 1, mode);
    ++done;
    long long tc = clock64();
    t_issue += tb - ta, t_wait += tc - tb;
  }
  cycles[blockIdx.x] = clock64() - t0;
  if (blockIdx.x == 0) cycles[gridDim.x] = t_issue, cycles[gridDim.x + 1] = t_wait;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int mis = argc > 1 ? atoi(argv[1]) : 3;
  CK(cudaMemcpyToSymbol(g_mis, &mis, 4));
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
  CK(cudaMalloc(&cyc, (sms + 2) * 8));
  // (channels per box, voxels in x, rows): same bytes per op except the last ones
  const int shapes[][3] = {{64, 12, 1}, {64, 12, 2}, {64, 12, 3}, {64, 12, 8}, {64, 20, 1}, {64, 20, 4}, {64, 8, 8}, {64, 16, 8}, {64, 32, 4}, {32, 12, 8}};
  for (int mode = 0; mode < 1; ++mode)
  for (auto &sh : shapes) {
    const int bc = sh[0], bx = sh[1], by = sh[2];
    CUtensorMap map;
    const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, 1};
    const cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)D * H * W * 4, (cuuint64_t)C * D * H * W * 4};
    const cuuint32_t box[5] = {(cuuint32_t)bx, (cuuint32_t)by, 1, (cuuint32_t)bc, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, feats, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = bc * bx * by * 4;
    for (int ops = 1; ops <= 16; ops *= 4) {
      for (int ns = 2; ns <= 16; ns *= 2) {
        const int slot_bytes = box_bytes * ops;
        if ((size_t)ns * slot_bytes + 64 > 200 * 1024) continue;
        const int iters = 400;
        const size_t smem = (size_t)ns * slot_bytes + 64;
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        probe<<<sms, 128, smem>>>(map, ns, slot_bytes, box_bytes, ops, 50, bc, bx, by, C, W, H, D, cyc, mode);
        cudaEventRecord(e0);
        probe<<<sms, 128, smem>>>(map, ns, slot_bytes, box_bytes, ops, iters, bc, bx, by, C, W, H, D, cyc, mode);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        { int to = 0; CK(cudaMemcpyFromSymbol(&to, g_timeout, 4)); if (to) { printf("TIMEOUT box %d %d %d ops %d ns %d\n", bc, bx, by, ops, ns); to = 0; CK(cudaMemcpyToSymbol(g_timeout, &to, 4)); continue; } }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)sms * iters * slot_bytes;
        unsigned long long hc[2];
        CK(cudaMemcpy(hc, cyc + sms, 16, cudaMemcpyDeviceToHost));
        printf("mode %d issue %5.0f wait %5.0f cyc/iter | ", mode, (double)hc[0] / iters, (double)hc[1] / iters);
        printf("box %3dch x %2dvox x %drows (%6d B/op, run %4d B) ops/slot %d slots %d : %7.1f us  %6.2f TB/s  %6.1f B/cyc/SM @1.9GHz  %6.0f cyc per slot\n",
               bc, bx, by, box_bytes, bx * 4, ops, ns, ms * 1e3, bytes / (ms * 1e-3) / 1e12, bytes / sms / (ms * 1e-3 * 1.9e9),
               ms * 1e-3 * 1.9e9 / iters);
      }
    }
  }
  return 0;
}
