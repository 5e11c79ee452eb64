// kge_sampler.cu -- negative sampling on the device (reference: TrainDataset.__getitem__, dataloader.py:28-67).
//
// The reference draws, per positive triple, uniform entity ids with numpy, drops those that would form a true
// training triple ((h', r, t) for head-batch, (h, r, t') for tail-batch; np.in1d against true_head / true_tail) and
// keeps the first `negative_sample_size` survivors: i.i.d. uniform over the complement of the true set.  At
// 0.14-0.19 ms per sample per worker that loop is >100x slower than the fused train step, so it moves here:
// one thread per (row, negative), a counter-based Philox4x32-10 stream (restated bit-for-bit by the numpy oracle),
// id = mulhi(u32, nentity), rejection by binary search in the row's sorted true list.
#include "kge_common.cuh"

namespace kge {

struct Philox {
  uint32_t c[4];
  uint32_t k[2];
};

__host__ __device__ inline void philox_round(Philox &s) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * s.c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * s.c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ s.c[1] ^ s.k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ s.c[3] ^ s.k[1];
  const uint32_t n3 = (uint32_t)p0;
  s.c[0] = n0; s.c[1] = n1; s.c[2] = n2; s.c[3] = n3;
  s.k[0] += 0x9E3779B9u;
  s.k[1] += 0xBB67AE85u;
}

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
  Philox s{{c0, c1, c2, c3}, {k0, k1}};
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(s);
  out[0] = s.c[0]; out[1] = s.c[1]; out[2] = s.c[2]; out[3] = s.c[3];
}

__global__ void sample_negatives_kernel(const int64_t *__restrict__ triple_index, const int32_t *__restrict__ key_start,
                                        const int32_t *__restrict__ key_len, const int32_t *__restrict__ true_entities,
                                        int64_t B, int N, uint32_t nentity, uint64_t seed, uint64_t step,
                                        int64_t *__restrict__ negative) {
  const int64_t total = B * N;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / N;
    const int64_t t = triple_index[b];
    const int32_t *lst = true_entities + key_start[t];
    const int len = key_len[t];
    uint32_t id = 0;
    bool found = false;
    for (uint32_t attempt = 0; !found; ++attempt) {
      uint32_t r[4];
      // counter = (pair index lo, pair index hi, attempt block, step lo); key = seed (step hi folded into key 1)
      philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), attempt, (uint32_t)step, (uint32_t)seed,
                    (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32), r);
#pragma unroll
      for (int i = 0; i < 4 && !found; ++i) {
        const uint32_t cand = __umulhi(r[i], nentity);
        int lo = 0, hi = len;                              // binary search in the sorted true list
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((uint32_t)lst[mid] < cand) lo = mid + 1; else hi = mid;
        }
        if (!(lo < len && (uint32_t)lst[lo] == cand)) { id = cand; found = true; }
      }
      if (attempt > (1u << 20)) { id = 0; found = true; }   // every entity is "true": cannot happen for real data
    }
    negative[p] = (int64_t)id;
  }
}

}  // namespace kge

using namespace kge;

extern "C" int kge_sample_negatives(const int64_t *triple_index, const int32_t *key_start, const int32_t *key_len,
                                    const int32_t *true_entities, int64_t B, int64_t N, int64_t nentity, uint64_t seed,
                                    uint64_t step, int64_t *negative, void *stream) {
  KGE_REQUIRE(triple_index && key_start && key_len && true_entities && negative, "null pointer");
  KGE_REQUIRE(B >= 0 && N > 0 && nentity > 0 && nentity < (1ll << 32), "bad sizes");
  if (B == 0) return KGE_OK;
  const int64_t total = B * N;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  sample_negatives_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(triple_index, key_start, key_len, true_entities, B, (int)N,
                                                                 (uint32_t)nentity, seed, step, negative);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}
