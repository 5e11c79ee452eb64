// kge_train.cu -- the fused gather + score + loss + backward kernel ("row kernel") and its C ABI.
//
// One CTA owns one positive row b of the batch (model.py:261) and walks its N candidate rows twice:
//   phase 0  fold the fixed side (h,r) or (r,t) into the query vector q in shared memory
//   phase 1  one warp per candidate: 128-bit coalesced gather of the row, element op against q,
//            warp-shuffle reduction over the hidden dimension -> score s[n] in shared memory
//   phase 2  (train) self-adversarial softmax / logsigmoid loss of the row in the same CTA -> g[n]
//   phase 3  one warp per candidate again (rows now come from L2): recompute the element terms, emit
//            dL/dx with 16-byte vector atomics (red.global.add.v4.f32) into the dense gradient table,
//            keep dL/dq in registers, summed over the warp's candidates
//   phase 4  warps fold their dL/dq into shared memory in a fixed order (deterministic)
//   phase 5  chain rule q -> fixed entity row and relation row, atomics into the gradient tables
// The same kernel serves KGEModel.forward (phases 0-1), its autograd backward (0,3,4,5 with a given
// dL/ds) and train_step (all phases); 'single' mode is tail-batch with the positive tail as the only
// candidate (model.py:83-102 uses the non-head-batch association of every score function).
#include <stdlib.h>

#include "kge_rows.cuh"

namespace kge {

struct RowArgs {
  const float *E, *R, *modulus;
  const int64_t *positive;     // [B_total, 3]
  const int64_t *cand;         // candidate (b, n) = cand[b * cand_stride + n]
  int64_t cand_stride;
  int64_t row_begin;
  int row_count, N;
  int64_t nentity, nrelation;
  int d;                       // k-extent: hidden_dim for complex ops, entity_dim for real ops
  int De, Dr;
  float gamma, scale;
  int do_loss, loss_kind;
  float alpha;
  const float *weight, *wsum;
  float uniform_u;
  float *row_loss;
  float *pos_row_loss;         // split path only: also do the positive triple of each row (fused 'single' pass)
  float *score_out;
  const float *dscore;
  float *gE, *gR, *gM;
  int32_t *err;
  int *fused_positive;         // host-side out flag: the launched variant handled pos_row_loss itself
  int defer_entity;            // host: do not launch the entity-major pass (caller slices it, kge_train_entity_pass)
  int *entity_deferred;        // host out flag: the split path ran and its entity pass is still due
};

constexpr int kChunks = 8;     // units per lane per k-tile: 8 x float4 x (re,im) = 64 accumulator registers

template <int V>
__device__ __forceinline__ void load_global(float (&o)[V], const float *p) {
  if constexpr (V == 4) {
    float4 t = ldg_stream4(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  } else {
    o[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ void load_shared(float (&o)[V], const float *p) {
  if constexpr (V == 4) {
    float4 t = *reinterpret_cast<const float4 *>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  } else {
    o[0] = *p;
  }
}
template <int V>
__device__ __forceinline__ void red_global(float *p, const float (&v)[V]) {
  if constexpr (V == 4) red_add4(p, v[0], v[1], v[2], v[3]);
  else red_add1(p, v[0]);
}

__device__ __forceinline__ float block_reduce(float v, float *scratch, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();                       // protect scratch from the previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int w = 0; w < nw; ++w) r = is_max ? fmaxf(r, scratch[w]) : r + scratch[w];   // fixed order
  return r;
}

template <int MODEL, bool HEAD, int V>
__global__ void __launch_bounds__(512, 1) row_kernel(const RowArgs a) {
  constexpr int OP = op_of(MODEL, HEAD);
  constexpr bool CPLX = op_is_complex(OP);
  constexpr int H = CPLX ? 2 : 1;              // halves per unit
  extern __shared__ __align__(16) float smem[];
  const int Dq = CPLX ? 2 * a.d : a.d;
  float *q = smem;                             // [Dq]
  float *dq = q + ((Dq + 3) & ~3);             // [Dq]
  float *sc = dq + ((Dq + 3) & ~3);            // [N] scores (only when do_loss)
  float *gg = sc + (a.do_loss ? a.N : 0);      // [N] dL/ds   (only when do_loss)
  float *scratch = gg + (a.do_loss ? a.N : 0); // [32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int nunits = a.d / V;                  // host guarantees divisibility
  const float modulus = MODEL == KGE_PROTATE ? __ldg(a.modulus) : 1.f;
  const bool do_bwd = a.gE != nullptr;

  for (int rl = blockIdx.x; rl < a.row_count; rl += gridDim.x) {
    const int64_t b = a.row_begin + rl;
    int64_t hid = a.positive[b * 3 + 0], rid = a.positive[b * 3 + 1], tidx = a.positive[b * 3 + 2];
    int64_t fid = HEAD ? tidx : hid;
    if ((uint64_t)fid >= (uint64_t)a.nentity || (uint64_t)rid >= (uint64_t)a.nrelation) {
      if (tid == 0 && a.err) *a.err = 1;
      fid = 0; rid = 0;
    }
    const float *F = a.E + fid * a.De;
    const float *Rr = a.R + rid * a.Dr;
    const int64_t *cand = a.cand + b * a.cand_stride;

    // ---- phase 0: query vector -------------------------------------------------------------------
    __syncthreads();
    for (int k = tid; k < a.d; k += blockDim.x) {
      build_q<MODEL, HEAD>(F, Rr, k, a.d, a.scale, q);
      dq[k] = 0.f;
      if (CPLX) dq[a.d + k] = 0.f;
    }
    __syncthreads();

    // ---- phase 1: scores -------------------------------------------------------------------------
    if (a.do_loss || a.score_out) {
      for (int n = warp; n < a.N; n += nwarps) {
        int64_t id = cand[n];
        if ((uint64_t)id >= (uint64_t)a.nentity) { if (lane == 0 && a.err) *a.err = 1; id = 0; }
        const float *x = a.E + id * a.De;
        float part = 0.f;
        for (int kt = 0; kt < nunits; kt += 32 * kChunks) {
#pragma unroll
          for (int i = 0; i < kChunks; ++i) {
            const int u = kt + lane + 32 * i;
            if (u < nunits) {
              float x0[V], x1[V], q0[V], q1[V];
              load_global<V>(x0, x + u * V);
              load_shared<V>(q0, q + u * V);
              if constexpr (CPLX) {
                load_global<V>(x1, x + a.d + u * V);
                load_shared<V>(q1, q + a.d + u * V);
              }
#pragma unroll
              for (int j = 0; j < V; ++j)
                part += op_forward<OP>(q0[j], CPLX ? q1[j] : 0.f, x0[j], CPLX ? x1[j] : 0.f, a.scale);
            }
          }
        }
        part = warp_sum(part);
        const float s = finish_score<MODEL>(part, a.gamma, modulus);
        if (lane == 0) {
          if (a.do_loss) sc[n] = s;
          if (a.score_out) a.score_out[(int64_t)rl * a.N + n] = s;
        }
      }
    }
    if (!a.do_loss && !do_bwd) continue;

    // ---- phase 2: loss of this row (model.py:270-288) and dL/ds ---------------------------------------
    if (a.do_loss) {
      __syncthreads();
      const float u = a.weight ? a.weight[b] / a.wsum[0] : a.uniform_u;
      float row_val;
      if (a.loss_kind == KGE_LOSS_POSITIVE) {              // model.py:277-279
        const float s = sc[0];
        row_val = log_sigmoid(s);
        if (tid == 0) gg[0] = -0.5f * u * sigmoid(-s);
      } else {
        float zmax = -INFINITY;
        if (a.loss_kind == KGE_LOSS_NEG_ADVERSARIAL) {     // softmax(alpha * s).detach(), model.py:272
          for (int n = tid; n < a.N; n += blockDim.x) zmax = fmaxf(zmax, sc[n] * a.alpha);
          zmax = block_reduce(zmax, scratch, true);
        }
        float zsum = 0.f;
        for (int n = tid; n < a.N; n += blockDim.x) {
          const float e = a.loss_kind == KGE_LOSS_NEG_ADVERSARIAL ? expf(sc[n] * a.alpha - zmax) : 1.f;
          gg[n] = e;
          zsum += e;
        }
        zsum = block_reduce(zsum, scratch, false);
        float acc = 0.f;
        for (int n = tid; n < a.N; n += blockDim.x) {
          const float w = gg[n] / zsum;                    // = 1/N for the uniform case (model.py:275)
          const float s = sc[n];
          acc += w * log_sigmoid(-s);
          gg[n] = 0.5f * u * w * sigmoid(s);
        }
        row_val = block_reduce(acc, scratch, false);
      }
      if (tid == 0) a.row_loss[b] = row_val;
      __syncthreads();
    }
    if (!do_bwd) continue;
    const float *gsrc = a.do_loss ? gg : a.dscore + (int64_t)rl * a.N;

    // ---- phase 3/4: backward over the candidates, k-tiled so dL/dq stays in registers -----------------
    float gmod = 0.f;
    for (int kt = 0; kt < nunits; kt += 32 * kChunks) {
      float acc[kChunks][H][V];
#pragma unroll
      for (int i = 0; i < kChunks; ++i)
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int j = 0; j < V; ++j) acc[i][h][j] = 0.f;

      for (int n = warp; n < a.N; n += nwarps) {
        int64_t id = cand[n];
        if ((uint64_t)id >= (uint64_t)a.nentity) id = 0;
        const float g = gsrc[n];
        const float go = dsum_of<MODEL>(g, modulus);
        const float *x = a.E + id * a.De;
        float *gx = a.gE + id * a.De;
        float vsum = 0.f;
#pragma unroll
        for (int i = 0; i < kChunks; ++i) {
          const int u = kt + lane + 32 * i;
          if (u < nunits) {
            float x0[V], x1[V], q0[V], q1[V], dx0[V], dx1[V];
            load_global<V>(x0, x + u * V);
            load_shared<V>(q0, q + u * V);
            if constexpr (CPLX) {
              load_global<V>(x1, x + a.d + u * V);
              load_shared<V>(q1, q + a.d + u * V);
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
              float dq0 = 0.f, dq1 = 0.f, ex0 = 0.f, ex1 = 0.f;
              vsum += op_backward<OP>(q0[j], CPLX ? q1[j] : 0.f, x0[j], CPLX ? x1[j] : 0.f, a.scale, go,
                                      dq0, dq1, ex0, ex1);
              acc[i][0][j] += dq0;
              dx0[j] = ex0;
              if constexpr (CPLX) { acc[i][1][j] += dq1; dx1[j] = ex1; }
            }
            red_global<V>(gx + u * V, dx0);
            if constexpr (CPLX) red_global<V>(gx + a.d + u * V, dx1);
          }
        }
        if constexpr (MODEL == KGE_PROTATE) {
          vsum = warp_sum(vsum);
          gmod += -g * vsum;                               // d/dmodulus of gamma - modulus * sum
        }
      }
      // phase 4: fixed-order fold of the per-warp partial dL/dq into shared memory
      for (int w = 0; w < nwarps; ++w) {
        if (warp == w) {
#pragma unroll
          for (int i = 0; i < kChunks; ++i) {
            const int u = kt + lane + 32 * i;
            if (u < nunits) {
#pragma unroll
              for (int j = 0; j < V; ++j) {
                dq[u * V + j] += acc[i][0][j];
                if constexpr (CPLX) dq[a.d + u * V + j] += acc[i][1][j];
              }
            }
          }
        }
        __syncthreads();
      }
    }

    // ---- phase 5: chain rule into the fixed entity row and the relation row -----------------------------
    float *gF = a.gE + fid * a.De;
    float *gRr = a.gR + rid * a.Dr;
    for (int k = tid; k < a.d; k += blockDim.x) {
      float dF0, dF1, dR0, dR1;
      chain_q<MODEL, HEAD>(F, Rr, dq, k, a.d, a.scale, dF0, dF1, dR0, dR1);
      red_add1(gF + k, dF0);
      red_add1(gRr + k, dR0);
      if constexpr (CPLX) red_add1(gF + a.d + k, dF1);
      if constexpr (MODEL == KGE_COMPLEX) red_add1(gRr + a.d + k, dR1);
    }
    if constexpr (MODEL == KGE_PROTATE) {
      if (lane == 0 && gmod != 0.f && a.gM) red_add1(a.gM, gmod);
    }
  }
}

// ======================================================================================================
// TMA variant: candidate rows are gathered by the bulk-copy engine (cp.async.bulk global -> shared, one
// 1-D copy of the whole D_e*4-byte row per candidate, completion on an mbarrier) into a per-warp double
// buffer.  The bytes in flight are then bounded by shared memory (W warps x 2 slots x row bytes, ~200 KB
// per SM) instead of by registers, which is what the direct-load kernel above is limited by (ncu r1a:
// 68% of issue slots stalled on long-scoreboard with 16 KB in flight per SM).  Phases are the same.
// ======================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MODEL, bool HEAD>
__global__ void __launch_bounds__(512, 1) row_kernel_tma(const RowArgs a) {
  constexpr int OP = op_of(MODEL, HEAD);
  constexpr bool CPLX = op_is_complex(OP);
  constexpr int H = CPLX ? 2 : 1;
  constexpr int V = 4;
  constexpr int CH = CPLX ? 8 : 16;            // units per lane: 64 accumulator registers either way
  extern __shared__ __align__(128) float smem[];
  const int Dq = a.De;
  const int Dq4 = (Dq + 3) & ~3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // layout: [slots: nwarps x 2 x De | q | dq | sc | gg | scratch(32) | mbarriers]; every pointer is derived from
  // `smem` by element offsets so that the compiler keeps the shared address space (LDS, not generic LD)
  float *slot0 = smem + (size_t)(2 * warp) * a.De, *slot1 = slot0 + a.De;
  float *q = smem + (size_t)(2 * nwarps) * a.De;
  float *dq = q + Dq4;
  float *sc = dq + Dq4;
  float *gg = sc + (a.do_loss ? a.N : 0);
  float *scratch = gg + (a.do_loss ? a.N : 0);
  uint64_t *bars = reinterpret_cast<uint64_t *>(scratch + 32);
  const uint32_t rowbytes = (uint32_t)a.De * 4u;
  uint64_t *bar0 = bars + 2 * warp, *bar1 = bar0 + 1;
  uint32_t par0 = 0, par1 = 0;
  int64_t id0 = 0, id1 = 0;

  const int nunits = a.d / V;                  // <= 32 * CH (host)
  const float modulus = MODEL == KGE_PROTATE ? __ldg(a.modulus) : 1.f;
  const bool do_bwd = a.gE != nullptr;

  if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  for (int rl = blockIdx.x; rl < a.row_count; rl += gridDim.x) {
    const int64_t b = a.row_begin + rl;
    int64_t hid = a.positive[b * 3 + 0], rid = a.positive[b * 3 + 1], tidx = a.positive[b * 3 + 2];
    int64_t fid = HEAD ? tidx : hid;
    if ((uint64_t)fid >= (uint64_t)a.nentity || (uint64_t)rid >= (uint64_t)a.nrelation) {
      if (tid == 0 && a.err) *a.err = 1;
      fid = 0; rid = 0;
    }
    const float *F = a.E + fid * a.De;
    const float *Rr = a.R + rid * a.Dr;
    const int64_t *cand = a.cand + b * a.cand_stride;

    // issue the bulk copy of candidate n into this warp's slot s (all lanes run it; lane 0 talks to the engine)
    auto issue = [&](int s, int n) {
      int64_t id = cand[n];
      if ((uint64_t)id >= (uint64_t)a.nentity) { if (lane == 0 && a.err) *a.err = 1; id = 0; }
      if (s) id1 = id; else id0 = id;
      if (lane == 0) {
        uint64_t *bar = s ? bar1 : bar0;
        mbar_expect_tx(bar, rowbytes);
        bulk_g2s(s ? slot1 : slot0, a.E + id * a.De, rowbytes, bar);
      }
    };
    // first two candidates of phase 1 (or of phase 3 for the backward-only call) overlap the q build
    if (warp < a.N) issue(0, warp);
    if (warp + nwarps < a.N) issue(1, warp + nwarps);

    // ---- phase 0: query vector -------------------------------------------------------------------
    for (int k = tid; k < a.d; k += blockDim.x) {
      build_q<MODEL, HEAD>(F, Rr, k, a.d, a.scale, q);
      dq[k] = 0.f;
      if (CPLX) dq[a.d + k] = 0.f;
    }
    __syncthreads();

    // ---- phase 1: scores -------------------------------------------------------------------------
    const bool do_fwd = a.do_loss || a.score_out;
    if (do_fwd) {
      int it = 0;
      for (int n = warp; n < a.N; n += nwarps, ++it) {
        const int s = it & 1;
        if (s) { mbar_wait(bar1, par1); par1 ^= 1; } else { mbar_wait(bar0, par0); par0 ^= 1; }
        const float *x = s ? slot1 : slot0;
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int u = lane + 32 * i;
          if (u < nunits) {
            float x0[V], x1[V], q0[V], q1[V];
            load_shared<V>(x0, x + u * V);
            load_shared<V>(q0, q + u * V);
            if constexpr (CPLX) {
              load_shared<V>(x1, x + a.d + u * V);
              load_shared<V>(q1, q + a.d + u * V);
            }
#pragma unroll
            for (int j = 0; j < V; ++j)
              part += op_forward<OP>(q0[j], CPLX ? q1[j] : 0.f, x0[j], CPLX ? x1[j] : 0.f, a.scale);
          }
        }
        __syncwarp();                                   // every lane is done reading the slot
        if (n + 2 * nwarps < a.N) issue(s, n + 2 * nwarps);
        part = warp_sum(part);
        const float sv = finish_score<MODEL>(part, a.gamma, modulus);
        if (lane == 0) {
          if (a.do_loss) sc[n] = sv;
          if (a.score_out) a.score_out[(int64_t)rl * a.N + n] = sv;
        }
      }
      if (do_bwd) {                                     // restart the ring for phase 3 while the loss is computed
        if (warp < a.N) issue(0, warp);
        if (warp + nwarps < a.N) issue(1, warp + nwarps);
      }
    }
    if (!a.do_loss && !do_bwd) { __syncthreads(); continue; }

    // ---- phase 2: loss of this row (model.py:270-288) and dL/ds ---------------------------------------
    if (a.do_loss) {
      __syncthreads();
      const float u = a.weight ? a.weight[b] / a.wsum[0] : a.uniform_u;
      float row_val;
      if (a.loss_kind == KGE_LOSS_POSITIVE) {
        const float sv = sc[0];
        row_val = log_sigmoid(sv);
        if (tid == 0) gg[0] = -0.5f * u * sigmoid(-sv);
      } else {
        float zmax = -INFINITY;
        if (a.loss_kind == KGE_LOSS_NEG_ADVERSARIAL) {
          for (int n = tid; n < a.N; n += blockDim.x) zmax = fmaxf(zmax, sc[n] * a.alpha);
          zmax = block_reduce(zmax, scratch, true);
        }
        float zsum = 0.f;
        for (int n = tid; n < a.N; n += blockDim.x) {
          const float e = a.loss_kind == KGE_LOSS_NEG_ADVERSARIAL ? expf(sc[n] * a.alpha - zmax) : 1.f;
          gg[n] = e;
          zsum += e;
        }
        zsum = block_reduce(zsum, scratch, false);
        float acc = 0.f;
        for (int n = tid; n < a.N; n += blockDim.x) {
          const float w = gg[n] / zsum;
          const float sv = sc[n];
          acc += w * log_sigmoid(-sv);
          gg[n] = 0.5f * u * w * sigmoid(sv);
        }
        row_val = block_reduce(acc, scratch, false);
      }
      if (tid == 0) a.row_loss[b] = row_val;
      __syncthreads();
    }
    if (!do_bwd) continue;
    const float *gsrc = a.do_loss ? gg : a.dscore + (int64_t)rl * a.N;

    // ---- phase 3: backward over the candidates; dL/dq stays in registers ------------------------------------
    float gmod = 0.f;
    float acc[CH][H][V];
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[i][h][j] = 0.f;
    {
      int it = 0;
      for (int n = warp; n < a.N; n += nwarps, ++it) {
        const int s = it & 1;
        if (s) { mbar_wait(bar1, par1); par1 ^= 1; } else { mbar_wait(bar0, par0); par0 ^= 1; }
        const float *x = s ? slot1 : slot0;
        const int64_t id = s ? id1 : id0;
        const float g = gsrc[n];
        const float go = dsum_of<MODEL>(g, modulus);
        float *gx = a.gE + id * a.De;
        float vsum = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int u = lane + 32 * i;
          if (u < nunits) {
            float x0[V], x1[V], q0[V], q1[V], dx0[V], dx1[V];
            load_shared<V>(x0, x + u * V);
            load_shared<V>(q0, q + u * V);
            if constexpr (CPLX) {
              load_shared<V>(x1, x + a.d + u * V);
              load_shared<V>(q1, q + a.d + u * V);
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
              float dq0 = 0.f, dq1 = 0.f, ex0 = 0.f, ex1 = 0.f;
              vsum += op_backward<OP>(q0[j], CPLX ? q1[j] : 0.f, x0[j], CPLX ? x1[j] : 0.f, a.scale, go,
                                      dq0, dq1, ex0, ex1);
              acc[i][0][j] += dq0;
              dx0[j] = ex0;
              if constexpr (CPLX) { acc[i][1][j] += dq1; dx1[j] = ex1; }
            }
            red_global<V>(gx + u * V, dx0);
            if constexpr (CPLX) red_global<V>(gx + a.d + u * V, dx1);
          }
        }
        __syncwarp();
        if (n + 2 * nwarps < a.N) issue(s, n + 2 * nwarps);
        if constexpr (MODEL == KGE_PROTATE) {
          vsum = warp_sum(vsum);
          gmod += -g * vsum;
        }
      }
    }
    // ---- phase 4: fixed-order fold of the per-warp partial dL/dq into shared memory -----------------------
    for (int w = 0; w < nwarps; ++w) {
      if (warp == w) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int u = lane + 32 * i;
          if (u < nunits) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              dq[u * V + j] += acc[i][0][j];
              if constexpr (CPLX) dq[a.d + u * V + j] += acc[i][1][j];
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- phase 5: chain rule into the fixed entity row and the relation row -----------------------------
    float *gF = a.gE + fid * a.De;
    float *gRr = a.gR + rid * a.Dr;
    for (int k = tid; k < a.d; k += blockDim.x) {
      float dF0, dF1, dR0, dR1;
      chain_q<MODEL, HEAD>(F, Rr, dq, k, a.d, a.scale, dF0, dF1, dR0, dR1);
      red_add1(gF + k, dF0);
      red_add1(gRr + k, dR0);
      if constexpr (CPLX) red_add1(gF + a.d + k, dF1);
      if constexpr (MODEL == KGE_COMPLEX) red_add1(gRr + a.d + k, dR1);
    }
    if constexpr (MODEL == KGE_PROTATE) {
      if (lane == 0 && gmod != 0.f && a.gM) red_add1(a.gM, gmod);
    }
    __syncthreads();                                    // q / dq are rebuilt by the next row
  }
}

}  // namespace kge

#include "kge_train_split.cuh"

namespace kge {

// ---- host side ---------------------------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t split_workspace_bytes(int64_t rows, int64_t N, int64_t De, int64_t nentity) {
  return align256(rows * N * 4) + align256(rows * De * 4) + align256((nentity + 1) * 4 + nentity * 4 + 64) +
         align256(((nentity + 1023) / 1024) * 4) + 2 * align256(rows * N * 4);
}

static SplitWs carve_split_ws(void *workspace, int64_t rows, int64_t N, int64_t De, int64_t nentity) {
  char *wp = (char *)workspace;
  SplitWs ws;
  ws.G = (float *)wp;      wp += align256((size_t)rows * N * 4);
  ws.Qtab = (float *)wp;   wp += align256((size_t)rows * De * 4);
  ws.cnt = (int *)wp;
  ws.cursor = ws.cnt + (nentity + 1);
  ws.queue = ws.cursor + nentity;                          // 16 queue counters (one per entity slice)
  wp += align256((size_t)(nentity + 1) * 4 + (size_t)nentity * 4 + 64);
  ws.tile_tot = (int *)wp; wp += align256((size_t)((nentity + 1023) / 1024) * 4);
  ws.perm = (int *)wp;     wp += align256((size_t)rows * N * 4);
  ws.gsorted = (float *)wp;
  return ws;
}

template <int MODEL, bool HEAD>
static int launch_entity_pass(const RowArgs &a, const SplitWs &ws, int64_t ent_begin, int64_t ent_end, int slot,
                              cudaStream_t st, int reserve_sms = 0) {
  constexpr bool CPLX = op_is_complex(op_of(MODEL, HEAD));
  if (ent_end <= ent_begin) return KGE_OK;
  const int nunits = a.d / 4;
  EntArgs e{};
  e.E = a.E; e.modulus = a.modulus; e.gE = a.gE; e.gM = a.gM; e.gsorted = ws.gsorted; e.Qtab = ws.Qtab;
  e.off = ws.cnt; e.perm = ws.perm; e.queue = ws.queue + slot; e.nentity = a.nentity;
  e.ent_begin = ent_begin; e.ent_count = ent_end - ent_begin;
  e.N = a.N; e.d = a.d; e.De = a.De; e.scale = a.scale;
  e.need_gmod = (MODEL == KGE_PROTATE && !a.do_loss) ? 1 : 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (reserve_sms > 0 && sms > 2 * reserve_sms) sms -= reserve_sms;   // leave SMs for the concurrent NCCL kernel
  const bool two = nunits >= 64;                     // enough work per lane to split the row in two parts
  e.upp = two ? (nunits + 1) / 2 : nunits;
  // slots have the kernel's compile-time half stride: CH chunks of 32 float4 units, CH = (complex ? 8 : 16) / parts
  const size_t slotbytes = (size_t)(CPLX ? 2 : 1) * ((CPLX ? 8 : 16) / (two ? 2 : 1)) * 32 * 16;
  int We = (int)((227 * 1024 - 16) / (2 * slotbytes + 16));
  const int wmax = two ? 20 : 12;
  if (We > wmax) We = wmax;
  const size_t esmem = 16 + (size_t)We * (2 * slotbytes + 16);
  if (two) {
    auto k = entity_kernel<MODEL, HEAD, 2>;
    KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
    k<<<sms, We * 32, esmem, st>>>(e);
  } else {
    auto k = entity_kernel<MODEL, HEAD, 1>;
    KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
    k<<<sms, We * 32, esmem, st>>>(e);
  }
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

template <int MODEL, bool HEAD>
static int launch_rows_v(const RowArgs &a, bool vec4, int threads, size_t smem, void *workspace, size_t workspace_bytes,
                         cudaStream_t st) {
  int grid = a.row_count;
  // TMA ring variants: rows are 16-byte multiples, one k-tile covers the row, and >= 4 warps get a double buffer
  constexpr bool CPLX = op_is_complex(op_of(MODEL, HEAD));
  const int nunits = a.d / 4;
  if (vec4 && nunits <= 32 * (CPLX ? 8 : 16) && !getenv("KGE_NO_TMA")) {
    const size_t rowbytes = (size_t)a.De * 4;
    const bool split = workspace && a.gE && a.N >= 8 && !(a.do_loss && a.loss_kind == KGE_LOSS_POSITIVE) &&
                       workspace_bytes >= split_workspace_bytes(a.row_count, a.N, a.De, a.nentity) &&
                       a.nentity < (1ll << 31) && (int64_t)a.row_count * a.N < (1ll << 31) && !getenv("KGE_NO_SPLIT") &&
                       // the entity-major pass pays a fixed cost per touched entity: it wins when an entity collects
                       // several pairs (17 at FB15k shapes: 0.73 vs 0.96 ms) and loses when pairs are sparse
                       // (3.3 at YAGO3-10 shapes: 1.59 vs 0.99 ms)
                       ((int64_t)a.row_count * a.N >= 6 * a.nentity || getenv("KGE_FORCE_SPLIT"));
    // ---- single-read path: row-major forward + dL/dq, counting sort, entity-major dL/dx -----------------
    if (split) {
      constexpr int Hs = CPLX ? 2 : 1;
      const int chunks = (nunits + 31) / 32;                 // 128-float chunks per half row
      const int nch = chunks <= 4 ? 4 : (chunks <= 8 ? 8 : 16);
      const size_t hs = (size_t)Hs * 128 * nch;              // padded slot (floats)
      const size_t fixed_s = sizeof(float) * (hs + (size_t)((a.De + 3) & ~3) + 2 * (size_t)a.N + 32) + 16;
      const size_t per_warp = 2 * hs * sizeof(float) + 16;
      int Ws = (int)((227 * 1024 - fixed_s) / per_warp);
      const int wcap = split_max_threads(CPLX, nch) / 32;
      if (Ws > wcap) Ws = wcap;
      if (a.N < 4 * Ws) Ws = a.N >= 16 ? (a.N + 3) / 4 : 4;              // short candidate lists: fewer, busier warps
      if (Ws > wcap) Ws = wcap;
      if (Ws >= 4 && fixed_s + Ws * per_warp <= 227 * 1024) {
        const size_t total = fixed_s + Ws * per_warp;
        SplitWs ws = carve_split_ws(workspace, a.row_count, a.N, a.De, a.nentity);
        KGE_CUDA_OK(cudaMemsetAsync(ws.cnt, 0, ((size_t)(2 * a.nentity + 1) + 16) * 4, st));
        // persistent CTAs (one per SM: the slots take the whole shared memory), rows are dealt round-robin
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int sgrid = grid < sms ? grid : sms;
#define KGE_SPLIT_LAUNCH(NCH)                                                                          \
  do {                                                                                                 \
    auto k = row_kernel_split<MODEL, HEAD, NCH>;                                                       \
    KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));     \
    k<<<sgrid, Ws * 32, total, st>>>(a, ws);                                                           \
  } while (0)
        if (nch == 4) KGE_SPLIT_LAUNCH(4);
        else if (nch == 8) KGE_SPLIT_LAUNCH(8);
        else {
          if constexpr (CPLX) { set_error("row too wide"); return KGE_ERR_INVALID; }      // unreachable: nunits <= 256
          else KGE_SPLIT_LAUNCH(16);
        }
#undef KGE_SPLIT_LAUNCH
        KGE_CUDA_OK(cudaGetLastError());
        if (a.fused_positive && a.pos_row_loss) *a.fused_positive = 1;
        {
          const int tiles = (int)((a.nentity + 1023) / 1024);
          scan_tiles_kernel<<<tiles, 1024, 0, st>>>(ws.cnt, ws.cursor, ws.tile_tot, a.nentity);
          KGE_CUDA_OK(cudaGetLastError());
          scan_apply_kernel<<<tiles, 1024, 0, st>>>(ws.cnt, ws.cursor, ws.tile_tot, a.nentity);
          KGE_CUDA_OK(cudaGetLastError());
        }
        {
          const int64_t pairs = (int64_t)a.row_count * a.N;
          int g2 = (int)((pairs + 255) / 256);
          if (g2 > 148 * 16) g2 = 148 * 16;
          scatter_pairs_kernel<<<g2, 256, 0, st>>>(a.cand, a.cand_stride, a.row_begin, a.row_count, a.N, a.nentity,
                                                   ws.G, ws.cursor, ws.perm, ws.gsorted);
          KGE_CUDA_OK(cudaGetLastError());
        }
        if (a.defer_entity) {
          if (a.entity_deferred) *a.entity_deferred = 1;
          return KGE_OK;
        }
        return launch_entity_pass<MODEL, HEAD>(a, ws, 0, a.nentity, 0, st);
      }
    }
    // ---- two-sweep TMA kernel ------------------------------------------------------------------------------
    const size_t base = sizeof(float) * (2 * (size_t)((a.De + 3) & ~3) + (a.do_loss ? 2 * (size_t)a.N : 0) + 32);
    const size_t fixed = base + 16;
    int W = (int)((227 * 1024 - fixed) / (2 * rowbytes + 16));
    if (W > 16) W = 16;
    if (a.N < 4 * W) W = a.N >= 16 ? (a.N + 3) / 4 : 4;                // short candidate lists: fewer, busier warps
    if (W >= 4 && fixed + W * (2 * rowbytes + 16) <= 227 * 1024) {
      const size_t total = fixed + W * (2 * rowbytes + 16);
      auto k = row_kernel_tma<MODEL, HEAD>;
      KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
      k<<<grid, W * 32, total, st>>>(a);
      KGE_CUDA_OK(cudaGetLastError());
      return KGE_OK;
    }
  }
  if (vec4) {
    auto k = row_kernel<MODEL, HEAD, 4>;
    if (smem > 48 * 1024) KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, threads, smem, st>>>(a);
  } else {
    auto k = row_kernel<MODEL, HEAD, 1>;
    if (smem > 48 * 1024) KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, threads, smem, st>>>(a);
  }
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

static int launch_rows(const kge_model_t *m, bool head, RowArgs &a, cudaStream_t st, void *workspace = nullptr,
                       size_t workspace_bytes = 0) {
  const bool cplx = m->model == KGE_COMPLEX || m->model == KGE_ROTATE;
  a.E = m->entity; a.R = m->relation; a.modulus = m->modulus;
  a.nentity = m->nentity; a.nrelation = m->nrelation;
  a.De = (int)m->entity_dim; a.Dr = (int)m->relation_dim;
  a.d = cplx ? (int)(m->entity_dim / 2) : (int)m->entity_dim;
  a.gamma = m->gamma;
  a.scale = phase_scale(m);
  if (a.row_count <= 0 || a.N <= 0) return KGE_OK;
  const bool aligned = (((uintptr_t)m->entity | (uintptr_t)a.gE) & 15) == 0;
  const bool vec4 = aligned && (a.d % 4 == 0) && (m->entity_dim % 4 == 0);
  const int Dq = (int)m->entity_dim;
  size_t smem = sizeof(float) * (2 * (size_t)((Dq + 3) & ~3) + (a.do_loss ? 2 * (size_t)a.N : 0) + 32);
  KGE_REQUIRE(smem <= 227 * 1024, "entity_dim=%d with %d candidates per row needs %zu B of shared memory (max 232448)",
              Dq, a.N, smem);
  // one warp per candidate in flight; a single-candidate pass ('single' mode) only needs a few warps for q
  int threads = a.N >= 16 ? 512 : (a.N >= 4 ? 256 : 128);
#define KGE_ROWS(MODEL)                                                            \
  case MODEL:                                                                      \
    return head ? launch_rows_v<MODEL, true>(a, vec4, threads, smem, workspace, workspace_bytes, st)           \
                : launch_rows_v<MODEL, false>(a, vec4, threads, smem, workspace, workspace_bytes, st);
  switch (m->model) {
    KGE_ROWS(KGE_TRANSE)
    KGE_ROWS(KGE_DISTMULT)
    KGE_ROWS(KGE_COMPLEX)
    KGE_ROWS(KGE_ROTATE)
    KGE_ROWS(KGE_PROTATE)
  }
#undef KGE_ROWS
  set_error("model %d not supported", m->model);
  return KGE_ERR_INVALID;
}

static int resolve_mode(int mode, const int64_t *positive, const int64_t *negative, int64_t N, RowArgs &a,
                        bool &head) {
  if (mode == KGE_SINGLE) {                 // model.py:83-102: candidate = the positive tail, N = 1
    KGE_REQUIRE(N == 1, "single mode scores one triple per row (N=%lld)", (long long)N);
    a.cand = positive + 2; a.cand_stride = 3; head = false;
  } else if (mode == KGE_HEAD_BATCH || mode == KGE_TAIL_BATCH) {
    KGE_REQUIRE(negative != nullptr, "mode needs the negative sample");
    a.cand = negative; a.cand_stride = N; head = mode == KGE_HEAD_BATCH;
  } else {
    set_error("mode %d not supported", mode);        // model.py:149
    return KGE_ERR_INVALID;
  }
  return KGE_OK;
}

}  // namespace kge

using namespace kge;

extern "C" int kge_score_forward(const kge_model_t *m, int mode, const int64_t *positive, const int64_t *negative,
                                 int64_t B, int64_t N, float *score, int32_t *err_flag, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(positive && score, "null pointer");
  RowArgs a{};
  bool head;
  if ((rc = resolve_mode(mode, positive, negative, N, a, head))) return rc;
  if ((rc = set_device(m))) return rc;
  a.positive = positive; a.row_begin = 0; a.row_count = (int)B; a.N = (int)N;
  a.score_out = score; a.err = err_flag;
  return launch_rows(m, head, a, (cudaStream_t)stream);
}

extern "C" int kge_score_backward(const kge_model_t *m, int mode, const int64_t *positive, const int64_t *negative,
                                  int64_t B, int64_t N, const float *dscore, float *grad_entity,
                                  float *grad_relation, float *grad_modulus, void *workspace, int64_t workspace_bytes,
                                  int32_t *err_flag, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(positive && dscore && grad_entity && grad_relation, "null pointer");
  KGE_REQUIRE(m->model != KGE_PROTATE || grad_modulus, "pRotatE needs grad_modulus");
  RowArgs a{};
  bool head;
  if ((rc = resolve_mode(mode, positive, negative, N, a, head))) return rc;
  if ((rc = set_device(m))) return rc;
  a.positive = positive; a.row_begin = 0; a.row_count = (int)B; a.N = (int)N;
  a.dscore = dscore; a.gE = grad_entity; a.gR = grad_relation; a.gM = grad_modulus; a.err = err_flag;
  return launch_rows(m, head, a, (cudaStream_t)stream, workspace, (size_t)workspace_bytes);
}

extern "C" int kge_train_rows(const kge_model_t *m, int mode, int loss_kind, float adversarial_temperature,
                              const int64_t *positive, const int64_t *negative, const float *weight,
                              const float *weight_sum, int64_t B_total, int64_t row_begin, int64_t row_count,
                              int64_t N, float *row_loss, float *pos_row_loss, float *grad_entity,
                              float *grad_relation, float *grad_modulus, float *score_out, void *workspace,
                              int64_t workspace_bytes, int32_t *err_flag, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(positive && row_loss && grad_entity && grad_relation, "null pointer");
  KGE_REQUIRE(m->model != KGE_PROTATE || grad_modulus, "pRotatE needs grad_modulus");
  KGE_REQUIRE(loss_kind >= KGE_LOSS_NEG_ADVERSARIAL && loss_kind <= KGE_LOSS_POSITIVE, "bad loss_kind %d", loss_kind);
  KGE_REQUIRE(!weight || weight_sum, "subsampling weights need their sum (kge_weight_sum)");
  KGE_REQUIRE(row_begin >= 0 && row_begin + row_count <= B_total, "row slice [%lld,+%lld) outside batch of %lld",
              (long long)row_begin, (long long)row_count, (long long)B_total);
  RowArgs a{};
  bool head;
  if (loss_kind == KGE_LOSS_POSITIVE) { mode = KGE_SINGLE; N = 1; }
  if ((rc = resolve_mode(mode, positive, negative, N, a, head))) return rc;
  if ((rc = set_device(m))) return rc;
  a.positive = positive; a.row_begin = row_begin; a.row_count = (int)row_count; a.N = (int)N;
  a.do_loss = 1; a.loss_kind = loss_kind; a.alpha = adversarial_temperature;
  a.weight = weight; a.wsum = weight_sum; a.uniform_u = 1.0f / (float)B_total;
  a.row_loss = row_loss; a.score_out = score_out;
  a.pos_row_loss = loss_kind == KGE_LOSS_POSITIVE ? nullptr : pos_row_loss;
  a.gE = grad_entity; a.gR = grad_relation; a.gM = grad_modulus; a.err = err_flag;
  int fused_positive = 0;
  a.fused_positive = &fused_positive;
  rc = launch_rows(m, head, a, (cudaStream_t)stream, workspace, (size_t)workspace_bytes);
  if (rc || !a.pos_row_loss || fused_positive) return rc;
  // the kernel variant that ran has no fused positive pass: run the 'single' pass as its own launch
  return kge_train_rows(m, KGE_SINGLE, KGE_LOSS_POSITIVE, 1.0f, positive, negative, weight, weight_sum, B_total, row_begin,
                        row_count, 1, pos_row_loss, nullptr, grad_entity, grad_relation, grad_modulus, nullptr, nullptr, 0,
                        err_flag, stream);
}

extern "C" int64_t kge_train_workspace_bytes(const kge_model_t *m, int64_t rows, int64_t N) {
  if (!m || rows <= 0 || N <= 0) return 0;
  return (int64_t)split_workspace_bytes(rows, N, m->entity_dim, m->nentity);
}

// ---- sliced variant for multi-GPU runs: the entity-major pass is launched per entity range so that the all-reduce of
// a finished slice of the gradient table overlaps the computation of the next one --------------------------------------
extern "C" int kge_train_rows_begin(const kge_model_t *m, int mode, int loss_kind, float adversarial_temperature,
                                    const int64_t *positive, const int64_t *negative, const float *weight,
                                    const float *weight_sum, int64_t B_total, int64_t row_begin, int64_t row_count,
                                    int64_t N, float *row_loss, float *pos_row_loss, float *grad_entity,
                                    float *grad_relation, float *grad_modulus, void *workspace, int64_t workspace_bytes,
                                    int32_t *err_flag, int32_t *host_entity_pass_pending, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(positive && negative && row_loss && pos_row_loss && grad_entity && grad_relation && host_entity_pass_pending,
              "null pointer");
  KGE_REQUIRE(m->model != KGE_PROTATE || grad_modulus, "pRotatE needs grad_modulus");
  KGE_REQUIRE(loss_kind == KGE_LOSS_NEG_ADVERSARIAL || loss_kind == KGE_LOSS_NEG_UNIFORM, "bad loss_kind %d", loss_kind);
  KGE_REQUIRE(!weight || weight_sum, "subsampling weights need their sum (kge_weight_sum)");
  KGE_REQUIRE(row_begin >= 0 && row_begin + row_count <= B_total, "row slice outside the batch");
  RowArgs a{};
  bool head;
  if ((rc = resolve_mode(mode, positive, negative, N, a, head))) return rc;
  if ((rc = set_device(m))) return rc;
  a.positive = positive; a.row_begin = row_begin; a.row_count = (int)row_count; a.N = (int)N;
  a.do_loss = 1; a.loss_kind = loss_kind; a.alpha = adversarial_temperature;
  a.weight = weight; a.wsum = weight_sum; a.uniform_u = 1.0f / (float)B_total;
  a.row_loss = row_loss; a.pos_row_loss = pos_row_loss;
  a.gE = grad_entity; a.gR = grad_relation; a.gM = grad_modulus; a.err = err_flag;
  int fused_positive = 0, deferred = 0;
  a.fused_positive = &fused_positive;
  a.defer_entity = 1;
  a.entity_deferred = &deferred;
  rc = launch_rows(m, head, a, (cudaStream_t)stream, workspace, (size_t)workspace_bytes);
  *host_entity_pass_pending = deferred;
  if (rc || fused_positive) return rc;
  return kge_train_rows(m, KGE_SINGLE, KGE_LOSS_POSITIVE, 1.0f, positive, negative, weight, weight_sum, B_total, row_begin,
                        row_count, 1, pos_row_loss, nullptr, grad_entity, grad_relation, grad_modulus, nullptr, nullptr, 0,
                        err_flag, stream);
}

extern "C" int kge_train_entity_pass(const kge_model_t *m, int mode, void *workspace, int64_t row_count, int64_t N,
                                     int64_t ent_begin, int64_t ent_end, int slice_index, float *grad_entity,
                                     float *grad_modulus, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(workspace && grad_entity, "null pointer");
  KGE_REQUIRE(mode == KGE_HEAD_BATCH || mode == KGE_TAIL_BATCH, "mode %d not supported", mode);
  KGE_REQUIRE(ent_begin >= 0 && ent_begin <= ent_end && ent_end <= m->nentity && slice_index >= 0 && slice_index < 16,
              "bad entity slice");
  if ((rc = set_device(m))) return rc;
  const bool cplx = m->model == KGE_COMPLEX || m->model == KGE_ROTATE;
  RowArgs a{};
  a.E = m->entity; a.modulus = m->modulus; a.nentity = m->nentity;
  a.De = (int)m->entity_dim; a.d = cplx ? (int)(m->entity_dim / 2) : (int)m->entity_dim;
  a.scale = phase_scale(m); a.N = (int)N; a.row_count = (int)row_count;
  a.gE = grad_entity; a.gM = grad_modulus; a.do_loss = 1;
  const SplitWs ws = carve_split_ws(workspace, row_count, N, m->entity_dim, m->nentity);
  const bool head = mode == KGE_HEAD_BATCH;
  cudaStream_t st = (cudaStream_t)stream;
  // the persistent entity kernel would otherwise hold every SM and the all-reduce of the previous slice could not start
  const char *rs = getenv("KGE_ENTITY_SMS_RESERVE");
  const int reserve = rs ? atoi(rs) : 0;       // measured at 2 GPUs: 0, 8, 24 equal, 48 slower
#define KGE_ENT(MODEL)                                                                            \
  case MODEL:                                                                                     \
    return head ? launch_entity_pass<MODEL, true>(a, ws, ent_begin, ent_end, slice_index, st, reserve)     \
                : launch_entity_pass<MODEL, false>(a, ws, ent_begin, ent_end, slice_index, st, reserve);
  switch (m->model) {
    KGE_ENT(KGE_TRANSE)
    KGE_ENT(KGE_DISTMULT)
    KGE_ENT(KGE_COMPLEX)
    KGE_ENT(KGE_ROTATE)
    KGE_ENT(KGE_PROTATE)
  }
#undef KGE_ENT
  set_error("model %d not supported", m->model);
  return KGE_ERR_INVALID;
}
