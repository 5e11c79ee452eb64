// l2_gather_bench.cu -- what is the L2 -> SM peak for the train path's access pattern?
//
// The row kernel of the train path (csrc/kge_train_split.cuh) gathers 8 KB entity rows with cp.async.bulk (TMA, 1-D)
// from a 120 MB table that mostly sits in the 126 MB L2.  This microbenchmark does ONLY that: a persistent grid, W warps
// per CTA, each warp keeps `depth` bulk copies of random rows in flight into its own shared-memory slots and (optionally)
// reads the row back with LDS.128 like the kernel's score sweep.  It prints GB/s for a table that fits L2 and one that
// does not, so that `roofline` can be quoted against the measured L2 -> SM gather ceiling next to the HBM copy peak.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/l2_gather_bench tools/l2_gather_bench.cu
//   tools/bin/l2_gather_bench            (prints one JSON line per configuration)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// rows_per_warp copies per warp; ids[warp_global * rows_per_warp + i] = row to fetch
template <int DEPTH, bool READ, bool TWO = false>
__global__ void gather_kernel(const float *__restrict__ table, const int *__restrict__ ids, int row_floats,
                              int rows_per_warp, float *__restrict__ sink) {
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float *slots = smem + (size_t)warp * DEPTH * row_floats;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nwarps * DEPTH * row_floats) + warp * DEPTH;
  if (lane == 0)
    for (int s = 0; s < DEPTH; ++s) mbar_init(bars + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int *my = ids + ((size_t)blockIdx.x * nwarps + warp) * rows_per_warp;
  const uint32_t bytes = (uint32_t)row_floats * 4u;
  auto issue = [&](int i) {
    if (lane == 0) {
      const int s = i % DEPTH;
      mbar_expect_tx(bars + s, bytes);
      if (TWO) {                                           // the row as two half-row copies (real | imaginary halves)
        bulk_g2s(slots + (size_t)s * row_floats, table + (size_t)my[i] * row_floats, bytes / 2, bars + s);
        bulk_g2s(slots + (size_t)s * row_floats + row_floats / 2, table + (size_t)my[i] * row_floats + row_floats / 2,
                 bytes / 2, bars + s);
      } else {
        bulk_g2s(slots + (size_t)s * row_floats, table + (size_t)my[i] * row_floats, bytes, bars + s);
      }
    }
  };
  for (int i = 0; i < DEPTH && i < rows_per_warp; ++i) issue(i);
  float acc = 0.f;
  for (int i = 0; i < rows_per_warp; ++i) {
    const int s = i % DEPTH;
    mbar_wait(bars + s, (i / DEPTH) & 1);
    if (READ) {
      const float4 *p = reinterpret_cast<const float4 *>(slots + (size_t)s * row_floats) + lane;
      for (int k = 0; k < row_floats / 128; ++k) {
        const float4 v = p[k * 32];
        acc += v.x + v.y + v.z + v.w;
      }
    }
    __syncwarp();
    if (i + DEPTH < rows_per_warp) issue(i + DEPTH);
  }
  if (READ && acc == 123.456f) sink[0] = acc;
}

template <int DEPTH, bool READ, bool TWO = false>
static double run(const float *table, const int *ids, int row_floats, int warps, int rows_per_warp, float *sink, int sms) {
  const size_t smem = (size_t)warps * DEPTH * row_floats * 4 + (size_t)warps * DEPTH * 8 + 16;
  auto k = gather_kernel<DEPTH, READ, TWO>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k<<<sms, warps * 32, smem, 0>>>(table, ids, row_floats, rows_per_warp, sink);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k<<<sms, warps * 32, smem, 0>>>(table, ids, row_floats, rows_per_warp, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  const double bytes = (double)sms * warps * rows_per_warp * row_floats * 4.0 * reps;
  return bytes / (ms * 1e-3) / 1e9;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int row_floats = 2000;                            // RotatE FB15k: D_e = 2000 fp32 = 8000 B per row
  float *sink;
  cudaMalloc(&sink, 4);
  struct Cfg { const char *name; int64_t rows; } tables[] = {{"fb15k_120MB_fits_L2", 14951}, {"4x_480MB_exceeds_L2", 59804}};
  for (auto &t : tables) {
    float *table;
    cudaMalloc(&table, (size_t)t.rows * row_floats * 4);
    cudaMemset(table, 0, (size_t)t.rows * row_floats * 4);
    for (int warps : {8, 12, 13}) {
      const int rows_per_warp = 262144 / (sms * warps) + 1;      // one train step's worth of pairs (1024 x 256)
      std::vector<int> h((size_t)sms * warps * rows_per_warp);
      uint64_t x = 88172645463325252ull;
      for (auto &v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (int)(x % (uint64_t)t.rows); }
      int *ids;
      cudaMalloc(&ids, h.size() * 4);
      cudaMemcpy(ids, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
      const double g2 = run<2, false>(table, ids, row_floats, warps, rows_per_warp, sink, sms);
      const double g2r = run<2, true>(table, ids, row_floats, warps, rows_per_warp, sink, sms);
      const double g3 = warps <= 8 ? run<3, false>(table, ids, row_floats, warps, rows_per_warp, sink, sms) : -1.0;
      const double g2two = run<2, false, true>(table, ids, row_floats, warps, rows_per_warp, sink, sms);
      printf("{\"table\": \"%s\", \"row_bytes\": %d, \"warps_per_sm\": %d, \"gather_gbs_depth2\": %.1f, "
             "\"gather_plus_lds_read_gbs_depth2\": %.1f, \"gather_gbs_depth3\": %.1f, "
             "\"gather_gbs_depth2_two_half_row_copies\": %.1f}\n",
             t.name, row_floats * 4, warps, g2, g2r, g3, g2two);
      cudaFree(ids);
    }
    cudaFree(table);
  }
  return 0;
}
