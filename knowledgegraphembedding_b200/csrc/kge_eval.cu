// kge_eval.cu -- filtered all-entity ranking without a sort (KGEModel.test_step, model.py:346-427).
//
// The reference scores every entity as a candidate, adds filter_bias, argsorts each row and looks up the
// position of the positive.  The position only depends on how many unfiltered candidates beat the
// positive, so this file computes exactly that count:
//   query_vectors    fold the fixed side of every query into q[Q, D_e]           (model.py:214-223 ...)
//   positive_scores  s_pos[q], bit-identical to the tile kernel's value for that column
//   count_ranks      register-tiled [64 queries x 64 entities] CTA tiles streamed over k through a
//                    double-buffered cp.async pipeline; epilogue compares with s_pos, applies the filter
//                    bitmap and accumulates integer counts (bit-exact, order independent)
// All score arithmetic here is written with un-contractable IEEE primitives and accumulated over k in
// index order, so the CPU oracle (oracle/kge_oracle.c) reproduces every score bit-for-bit.
#include "kge_rows.cuh"

namespace kge {

// ---- exact element ops -----------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ float op_exact(float q0, float q1, float x0, float x1) {
  if constexpr (OP == OP_SUBABS) return fabsf(fsub(q0, x0));                       // model.py:170,172
  else if constexpr (OP == OP_ADDABS) return fabsf(fadd(x0, q0));                  // model.py:168,172
  else if constexpr (OP == OP_MUL) return fmul(q0, x0);                            // model.py:177-179
  else if constexpr (OP == OP_CMUL) return fadd(fmul(q0, x0), fmul(q1, x1));       // model.py:192,196
  else if constexpr (OP == OP_CDIST) {                                             // model.py:217-226
    const float a = fsub(q0, x0), b = fsub(q1, x1);
    return fsqrt(ffma(b, b, fmul(a, a)));
  } else if constexpr (OP == OP_SUBSIN) return fabsf(sin_rep(fsub(q0, x0)));       // model.py:243-246 (x = phase table)
  else return fabsf(sin_rep(fadd(x0, q0)));                                        // model.py:241
}

// Fast (non-canonical) element op of the two-stage RotatE evaluation: contraction allowed, one-instruction
// approximate sqrt.  Per-term relative deviation from op_exact is below 2^-20 (sqrt.approx <= 2^-21, the a^2+b^2
// association <= 2^-22, exact rounding <= 2^-24); see kFastBand.
// pRotatE: |sin(t)| has period pi, so t is reduced by the nearest multiple of pi (two-constant Cody-Waite: the product
// k * 3.140625 is exact for |k| < 2^15; each of the two FMAs rounds once, <= 1.2e-7 for |r| <= 2; the residual constant
// is good to 2^-24 relative, i.e. <= |k| * 6e-11) to r in [-pi/2, pi/2], where sin.approx (one MUFU instruction) is
// within 2^-20.9 = 5.1e-7 of sin (PTX ISA, quadrant 00): 8 instructions instead of the ~35 of sin_rep.  For |t| < 4e4
// (|k| < 1.3e4) the reduction error is <= 2.4e-7 + 8e-7; beyond that the result is NaN, which the epilogue turns into
// "undecidable" (exact re-score), never into a wrong count.
__device__ __forceinline__ float abs_sin_fast(float t) {
  const float k = rintf(t * 0.318309886183790672f);
  float r = fmaf(-k, 3.140625f, t);
  r = fmaf(-k, 9.67653589793e-4f, r);
  float s;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(r));
  return fabsf(t) < 4.0e4f ? fabsf(s) : __int_as_float(0x7fc00000);
}
template <int OP>
__device__ __forceinline__ float op_fast(float q0, float q1, float x0, float x1) {
  static_assert(OP == OP_CDIST || OP == OP_SUBSIN || OP == OP_ADDSIN, "ops with a fast variant");
  if constexpr (OP == OP_CDIST) {
    const float a = q0 - x0, b = q1 - x1;
    return sqrt_approx(fmaf(a, a, b * b));
  } else if constexpr (OP == OP_SUBSIN) return abs_sin_fast(q0 - x0);
  else return abs_sin_fast(x0 + q0);
}
// |s_fast - s_canonical| <= kFastBand * (gamma - s_fast): 2^-20 per term + two blocked fp32 sums of the same
// blocks (each within (32 + d/32) * 2^-24 <= 6e-6 of the exact sum for d <= 2048), rounded up.
constexpr float kFastBand = 1.6e-5f;
// pRotatE: per term |fast - canonical| <= 5.1e-7 + 1.04e-6 (above) + 1.8e-7 (sin_rep's own 1.5 ulp) < kFastSinTerm; sums as above
constexpr float kFastSinTerm = 2.0e-6f, kFastSinSum = 1.2e-5f;

template <int OP>
__device__ __forceinline__ float finish_exact(float acc, float gamma, float modulus) {
  if constexpr (OP == OP_MUL || OP == OP_CMUL) return acc;
  else if constexpr (OP == OP_SUBSIN || OP == OP_ADDSIN) return fsub(gamma, fmul(acc, modulus));
  else return fsub(gamma, acc);
}

// ---- query vectors ---------------------------------------------------------------------------------------
template <int MODEL, bool HEAD>
__global__ void query_vectors_kernel(const float *__restrict__ E, const float *__restrict__ R,
                                     const int64_t *__restrict__ queries, int64_t Q, int64_t nentity,
                                     int64_t nrelation, int d, int De, int Dr, float scale,
                                     float *__restrict__ qvec, int32_t *err) {
  for (int64_t qi = blockIdx.x; qi < Q; qi += gridDim.x) {
    int64_t fid = queries[qi * 3 + (HEAD ? 2 : 0)], rid = queries[qi * 3 + 1];
    const int64_t pid = queries[qi * 3 + (HEAD ? 0 : 2)];
    if ((uint64_t)pid >= (uint64_t)nentity && threadIdx.x == 0 && err) *err = 1;
    if ((uint64_t)fid >= (uint64_t)nentity || (uint64_t)rid >= (uint64_t)nrelation) {
      if (threadIdx.x == 0 && err) *err = 1;
      fid = 0; rid = 0;
    }
    const float *F = E + fid * De, *Rr = R + rid * Dr;
    float *q = qvec + qi * De;
    for (int k = threadIdx.x; k < d; k += blockDim.x) build_q<MODEL, HEAD>(F, Rr, k, d, scale, q);
  }
}

__global__ void phase_table_kernel(const float *__restrict__ E, int64_t n, float scale, float *__restrict__ P) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    P[i] = fdiv(E[i], scale);
}

// Accumulation order of every evaluation score (the "canonical order"): the k axis is cut
// into blocks of KC consecutive indices; each block is summed in index order from 0.0f, and the block sums are
// added in index order.  (Error grows with sqrt(KC) + sqrt(d/KC) instead of sqrt(d).)
constexpr int TQ = 64, TJ = 64, KC = 32, ST = KC + 4, EVAL_THREADS = 256;

// ---- canonical score of one (query vector, candidate row) pair, computed by a whole warp ------------------------
// Lane l sums block l (32 consecutive k, index order, from 0.0f); lane 0 then adds the block sums in index order.
// Bit-identical to the sequential definition, ~32x shorter dependency chain.  Result valid in every lane.
template <int OP>
__device__ __forceinline__ float canonical_pair_score(const float *__restrict__ q, const float *__restrict__ x, int d) {
  constexpr bool CPLX = op_is_complex(OP);
  const int lane = threadIdx.x & 31;
  const int nblocks = (d + KC - 1) / KC;
  const bool vec4 = (d % 4 == 0) && ((((uintptr_t)q | (uintptr_t)x) & 15) == 0);
  float acc = 0.f;
  for (int b0 = 0; b0 < nblocks; b0 += 32) {
    const int blk = b0 + lane;
    float part = 0.f;
    if (blk < nblocks) {
      const int k0 = blk * KC, k1 = k0 + KC < d ? k0 + KC : d;
      if (vec4) {                                          // 16-byte loads; same element order
        for (int k = k0; k < k1; k += 4) {
          const float4 qa = *reinterpret_cast<const float4 *>(q + k), xa = *reinterpret_cast<const float4 *>(x + k);
          float4 qb = make_float4(0.f, 0.f, 0.f, 0.f), xb = qb;
          if constexpr (CPLX) {
            qb = *reinterpret_cast<const float4 *>(q + d + k);
            xb = *reinterpret_cast<const float4 *>(x + d + k);
          }
          part = fadd(part, op_exact<OP>(qa.x, qb.x, xa.x, xb.x));
          part = fadd(part, op_exact<OP>(qa.y, qb.y, xa.y, xb.y));
          part = fadd(part, op_exact<OP>(qa.z, qb.z, xa.z, xb.z));
          part = fadd(part, op_exact<OP>(qa.w, qb.w, xa.w, xb.w));
        }
      } else {
        for (int k = k0; k < k1; ++k)
          part = fadd(part, op_exact<OP>(q[k], CPLX ? q[d + k] : 0.f, x[k], CPLX ? x[d + k] : 0.f));
      }
    }
    const int nb = nblocks - b0 < 32 ? nblocks - b0 : 32;
    for (int l = 0; l < nb; ++l) acc = fadd(acc, __shfl_sync(0xffffffffu, part, l));
  }
  return acc;
}

// ---- positive scores ----------------------------------------------------------------------------------------
template <int OP>
__global__ void positive_scores_kernel(const float *__restrict__ qvec, const float *__restrict__ X,
                                       const int64_t *__restrict__ queries, int pos_col, int64_t Q, int64_t nentity,
                                       int d, int De, float gamma, const float *__restrict__ modulus,
                                       float *__restrict__ pos_score) {
  const int64_t qi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // one warp per query
  if (qi >= Q) return;
  int64_t pid = queries[qi * 3 + pos_col];
  if ((uint64_t)pid >= (uint64_t)nentity) pid = 0;
  const float acc = canonical_pair_score<OP>(qvec + qi * De, X + pid * De, d);
  if ((threadIdx.x & 31) == 0) pos_score[qi] = finish_exact<OP>(acc, gamma, modulus ? modulus[0] : 1.f);
}

// ---- tiled score + count ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;                       // src-size 0 => 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct EvalArgs {
  const float *qvec;            // [Q, De]
  const float *X;               // candidate table [nentity, De] (entity table, or phase table for pRotatE)
  const int64_t *queries;       // [Q, 3]
  const float *pos_score;       // [Q]
  const uint32_t *filter_bits;  // [Q, words]
  int32_t *counts;              // [Q]
  float *scores_out;            // [Q, nentity] or null
  const float *modulus;
  int64_t Q, nentity, ent_begin, ent_end;
  int d, De, pos_col, words;
  float gamma;
  int2 *amb;                    // two-stage mode: ambiguous (q, j) pairs for the exact re-score
  int *amb_count;               // [0] appended, [1] overflow
  int amb_capacity;
};

// Stage one k-chunk of the query tile and the entity tile:  smem[row][half][ST]
template <bool CPLX, bool ALIGNED>
__device__ __forceinline__ void stage_tiles(const EvalArgs &a, float *sq, float *sx, int64_t q0, int64_t j0, int k0) {
  constexpr int H = CPLX ? 2 : 1;
  constexpr int VPR = KC / 4;                              // float4 per row-half per chunk
  constexpr int TOTAL = (TQ + TJ) * H * VPR;
  for (int idx = threadIdx.x; idx < TOTAL; idx += EVAL_THREADS) {
    const int v = idx % VPR;
    const int h = (idx / VPR) % H;
    const int row = idx / (VPR * H);
    const int k = k0 + v * 4;
    const bool is_q = row < TQ;
    int64_t g = is_q ? q0 + row : j0 + (row - TQ);
    const int64_t lim = is_q ? a.Q : a.ent_end;
    if (g >= lim) g = lim - 1;                             // clamp: the duplicate row's results are masked out
    const float *src = (is_q ? a.qvec : a.X) + g * a.De + h * a.d + k;
    float *dst = (is_q ? sq + row * (H * ST) : sx + (row - TQ) * (H * ST)) + h * ST + v * 4;
    if constexpr (ALIGNED) {
      cp_async16(dst, k < a.d ? src : (is_q ? a.qvec : a.X), k < a.d);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) dst[e] = (k + e < a.d) ? src[e] : 0.f;
    }
  }
}

template <int OP, bool ALIGNED, bool FAST = false>
__global__ void __launch_bounds__(EVAL_THREADS, 2) count_ranks_kernel(const EvalArgs a) {
  constexpr bool CPLX = op_is_complex(OP);
  constexpr int H = CPLX ? 2 : 1;
  extern __shared__ __align__(16) float smem[];
  constexpr int TILE_Q = TQ * H * ST, TILE_X = TJ * H * ST;
  constexpr int STAGE = TILE_Q + TILE_X;
  __shared__ int cnt_sh[TQ];

  const int tid = threadIdx.x;
  const int tq = tid % 16, tj = tid / 16;
  const int64_t j0 = a.ent_begin + (int64_t)blockIdx.x * TJ;
  const int64_t q0 = (int64_t)blockIdx.y * TQ;
  if (tid < TQ) cnt_sh[tid] = 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nchunks = (a.d + KC - 1) / KC;
  stage_tiles<CPLX, ALIGNED>(a, smem, smem + TILE_Q, q0, j0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int cur = c & 1;
    if (c + 1 < nchunks)
      stage_tiles<CPLX, ALIGNED>(a, smem + (cur ^ 1) * STAGE, smem + (cur ^ 1) * STAGE + TILE_Q, q0, j0, (c + 1) * KC);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float *Q_ = smem + cur * STAGE, *X_ = Q_ + TILE_Q;
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
    constexpr bool PACKED = FAST && OP == OP_CDIST;        // FADD2 / FMUL2 / FFMA2: 3.5 issue slots per term, not 6
    f2 part2[PACKED ? 4 : 1][PACKED ? 4 : 1];
    if constexpr (PACKED) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part2[i][j] = pack2(0.f, 0.f);
    }
#pragma unroll 2
    for (int kk = 0; kk < KC; kk += 4) {
      float4 qa[4], qb[4], xa[4], xb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qa[i] = *reinterpret_cast<const float4 *>(Q_ + (tq + 16 * i) * (H * ST) + kk);
        xa[i] = *reinterpret_cast<const float4 *>(X_ + (tj + 16 * i) * (H * ST) + kk);
        if constexpr (CPLX) {
          qb[i] = *reinterpret_cast<const float4 *>(Q_ + (tq + 16 * i) * (H * ST) + ST + kk);
          xb[i] = *reinterpret_cast<const float4 *>(X_ + (tj + 16 * i) * (H * ST) + ST + kk);
        } else {
          qb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          xb[i] = qb[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {       // k ascending inside the group => index-order accumulation
          if constexpr (PACKED) {
            // |q - x| for four k at once: differences, squares and the partial sums run on packed pairs; only the four
            // square roots (MUFU) are scalar.  The approximate scores stay inside the band of the two-stage scheme
            // (same per-term error; the block sum is formed as (even k) + (odd k), still at most KC terms deep).
            const f2 a01 = sub2(pack2(qa[i].x, qa[i].y), pack2(xa[j].x, xa[j].y));
            const f2 a23 = sub2(pack2(qa[i].z, qa[i].w), pack2(xa[j].z, xa[j].w));
            const f2 b01 = sub2(pack2(qb[i].x, qb[i].y), pack2(xb[j].x, xb[j].y));
            const f2 b23 = sub2(pack2(qb[i].z, qb[i].w), pack2(xb[j].z, xb[j].w));
            const f2 m01 = fma2(b01, b01, mul2(a01, a01));
            const f2 m23 = fma2(b23, b23, mul2(a23, a23));
            float s0, s1, s2, s3;
            unpack2(m01, s0, s1);
            unpack2(m23, s2, s3);
            const f2 r01 = pack2(sqrt_approx(s0), sqrt_approx(s1));
            const f2 r23 = pack2(sqrt_approx(s2), sqrt_approx(s3));
            part2[i][j] = add2(part2[i][j], add2(r01, r23));
          } else if constexpr (FAST) {
            part[i][j] += op_fast<OP>(qa[i].x, qb[i].x, xa[j].x, xb[j].x);
            part[i][j] += op_fast<OP>(qa[i].y, qb[i].y, xa[j].y, xb[j].y);
            part[i][j] += op_fast<OP>(qa[i].z, qb[i].z, xa[j].z, xb[j].z);
            part[i][j] += op_fast<OP>(qa[i].w, qb[i].w, xa[j].w, xb[j].w);
          } else {
            part[i][j] = fadd(part[i][j], op_exact<OP>(qa[i].x, qb[i].x, xa[j].x, xb[j].x));
            part[i][j] = fadd(part[i][j], op_exact<OP>(qa[i].y, qb[i].y, xa[j].y, xb[j].y));
            part[i][j] = fadd(part[i][j], op_exact<OP>(qa[i].z, qb[i].z, xa[j].z, xb[j].z));
            part[i][j] = fadd(part[i][j], op_exact<OP>(qa[i].w, qb[i].w, xa[j].w, xb[j].w));
          }
        }
    }
    if constexpr (PACKED) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo, hi;
          unpack2(part2[i][j], lo, hi);
          part[i][j] = lo + hi;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fadd(acc[i][j], part[i][j]);
    __syncthreads();
  }

  // ---- epilogue: compare with the positive, apply the filter, count -------------------------------------------
  const float modulus = a.modulus ? a.modulus[0] : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t qi = q0 + tq + 16 * i;
    if (qi >= a.Q) continue;
    const float sp = a.pos_score[qi];
    int64_t pid = a.queries[qi * 3 + a.pos_col];
    int c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t ej = j0 + tj + 16 * j;
      if (ej >= a.ent_end) continue;
      float s = finish_exact<OP>(acc[i][j], a.gamma, modulus);
      const bool filtered = ej != pid && ((a.filter_bits[qi * a.words + (ej >> 5)] >> (ej & 31)) & 1u);
      if (filtered) s = fadd(sp, -1.0f);                   // candidate replaced by the positive, bias -1
      else if (ej != pid) {
        if constexpr (FAST) {                              // s is only within eps of the canonical score
          float eps;
          if constexpr (OP == OP_CDIST) eps = kFastBand * (a.gamma - s) + 1e-12f;
          else eps = fabsf(modulus) * (float)a.d * kFastSinTerm + kFastSinSum * fabsf(a.gamma - s) + 1e-12f;
          if (s - eps > sp) ++c;
          else if (!(s + eps < sp)) {                      // undecidable here (or NaN): exact re-score later
            const int slot = atomicAdd(a.amb_count, 1);
            if (slot < a.amb_capacity) a.amb[slot] = make_int2((int)qi, (int)ej);
            else a.amb_count[1] = 1;
          }
        } else {
          c += (s > sp) || (s == sp && ej < pid);
        }
      }
      if (a.scores_out) a.scores_out[qi * a.nentity + ej] = s;
    }
    if (c) atomicAdd(&cnt_sh[tq + 16 * i], c);
  }
  __syncthreads();
  if (tid < TQ && cnt_sh[tid] && q0 + tid < a.Q) atomicAdd(&a.counts[q0 + tid], cnt_sh[tid]);
}

// exact canonical re-score of a list of ambiguous (query, entity) pairs (tcgen05 path, kge_eval_gemm.cu)
template <int OP>
__global__ void rescore_pairs_kernel(const int2 *__restrict__ amb, const int *__restrict__ amb_count, int capacity,
                                     const float *__restrict__ qvec, const float *__restrict__ E, int d, int De,
                                     const float *__restrict__ pos_score, const int64_t *__restrict__ queries,
                                     int pos_col, float gamma, int32_t *__restrict__ counts,
                                     const float *__restrict__ modulus = nullptr) {
  int n = amb_count[0];
  if (n > capacity) n = capacity;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = wid; i < n; i += nw) {
    const int2 p = amb[i];
    const float s = finish_exact<OP>(canonical_pair_score<OP>(qvec + (int64_t)p.x * De, E + (int64_t)p.y * De, d),
                                     gamma, modulus ? modulus[0] : 1.f);
    if (lane == 0) {
      const float sp = pos_score[p.x];
      const int64_t pid = queries[(int64_t)p.x * 3 + pos_col];
      if (s > sp || (s == sp && p.y < pid)) atomicAdd(counts + p.x, 1);
    }
  }
}

int launch_rescore_pairs(bool cplx, const void *amb, const int *amb_count, int capacity, const float *qvec,
                         const float *E, int d, int De, const float *pos_score, const int64_t *queries, int pos_col,
                         int32_t *counts, cudaStream_t st) {
  if (cplx)
    rescore_pairs_kernel<OP_CMUL><<<148 * 8, 256, 0, st>>>((const int2 *)amb, amb_count, capacity, qvec, E, d, De,
                                                           pos_score, queries, pos_col, 0.f, counts);
  else
    rescore_pairs_kernel<OP_MUL><<<148 * 8, 256, 0, st>>>((const int2 *)amb, amb_count, capacity, qvec, E, d, De,
                                                          pos_score, queries, pos_col, 0.f, counts);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

__global__ void filter_bits_kernel(const int64_t *__restrict__ offsets, const int32_t *__restrict__ ents, int64_t Q,
                                   int64_t nentity, int words, uint32_t *__restrict__ bits) {
  for (int64_t qi = blockIdx.x; qi < Q; qi += gridDim.x) {
    const int64_t b = offsets[qi], e = offsets[qi + 1];
    for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int32_t j = ents[i];
      if (j >= 0 && j < nentity) atomicOr(&bits[qi * words + (j >> 5)], 1u << (j & 31));
    }
  }
}

// Filter bitmap straight from the device-resident index of all true triples: one CTA per query looks its key
// ((r, t) for head-batch, (h, r) for tail-batch) up in the sorted key table and sets the bits of that key's run.
__global__ void filter_lookup_kernel(const int64_t *__restrict__ keys, const int64_t *__restrict__ key_offsets,
                                     const int32_t *__restrict__ values, int64_t nkeys,
                                     const int64_t *__restrict__ queries, int64_t Q, int head_batch, int64_t nentity,
                                     int64_t nrelation, int words, uint32_t *__restrict__ bits) {
  __shared__ int64_t run[2];
  for (int64_t qi = blockIdx.x; qi < Q; qi += gridDim.x) {
    if (threadIdx.x == 0) {
      const int64_t h = queries[qi * 3], r = queries[qi * 3 + 1], t = queries[qi * 3 + 2];
      const int64_t key = head_batch ? r * nentity + t : h * nrelation + r;
      int64_t lo = 0, hi = nkeys;                                      // lower bound
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
      }
      const bool found = lo < nkeys && keys[lo] == key;
      run[0] = found ? key_offsets[lo] : 0;
      run[1] = found ? key_offsets[lo + 1] : 0;
    }
    __syncthreads();
    const int64_t b = run[0], e = run[1];
    for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int32_t j = values[i];
      if (j >= 0 && j < nentity) atomicOr(&bits[qi * words + (j >> 5)], 1u << (j & 31));
    }
    __syncthreads();
  }
}

// ---- device-side build of the filter index (dataloader.py:122-162 keeps all true triples in a python set and probes it
// once per entity per query): a direct-address CSR over the key space nentity x nrelation, by counting sort --
// histogram of the keys, tiled exclusive scan (the train path's scan kernels), scatter.  offsets[key] .. offsets[key + 1]
// then delimit the true heads of (r, t) / true tails of (h, r); no sorted key table, no binary search per query.
__global__ void __launch_bounds__(1024) scan_tiles_kernel(const int *__restrict__ cnt, int *__restrict__ cursor,
                                                          int *__restrict__ tile_tot, int64_t n);      // kge_train.cu
__global__ void __launch_bounds__(1024) scan_apply_kernel(int *__restrict__ cnt, int *__restrict__ cursor,
                                                          const int *__restrict__ tile_tot, int64_t n);

template <bool PLACE>
__global__ void filter_index_kernel(const int64_t *__restrict__ tri, int64_t n, int head_batch, int64_t nentity,
                                    int64_t nrelation, int *__restrict__ cnt_or_cursor, int32_t *__restrict__ entities,
                                    int32_t *err) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t h = tri[3 * i], r = tri[3 * i + 1], t = tri[3 * i + 2];
    if ((uint64_t)h >= (uint64_t)nentity || (uint64_t)t >= (uint64_t)nentity || (uint64_t)r >= (uint64_t)nrelation) {
      if (err) *err = 1;                                   // (a triple outside the tables can never be a query's answer)
      continue;
    }
    const int64_t key = head_batch ? r * nentity + t : h * nrelation + r;
    const int pos = atomicAdd(cnt_or_cursor + key, 1);
    if (PLACE) entities[pos] = (int32_t)(head_batch ? h : t);
  }
}

__global__ void filter_lookup_dense_kernel(const int32_t *__restrict__ offsets, const int32_t *__restrict__ values,
                                           const int64_t *__restrict__ queries, int64_t Q, int head_batch,
                                           int64_t nentity, int64_t nrelation, int words, uint32_t *__restrict__ bits) {
  for (int64_t qi = blockIdx.x; qi < Q; qi += gridDim.x) {
    const int64_t h = queries[qi * 3], r = queries[qi * 3 + 1], t = queries[qi * 3 + 2];
    const int64_t fixed = head_batch ? t : h;
    if ((uint64_t)fixed >= (uint64_t)nentity || (uint64_t)r >= (uint64_t)nrelation) continue;   // flagged by the query kernels
    const int64_t key = head_batch ? r * nentity + t : h * nrelation + r;
    const int b = offsets[key], e = offsets[key + 1];
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int32_t j = values[i];
      atomicOr(&bits[qi * words + (j >> 5)], 1u << (j & 31));
    }
  }
}

template <int OP>
static int launch_count(const EvalArgs &a, bool aligned, cudaStream_t st) {
  constexpr int H = op_is_complex(OP) ? 2 : 1;
  const size_t smem = sizeof(float) * 2 * (TQ + TJ) * H * ST;
  dim3 grid((unsigned)((a.ent_end - a.ent_begin + TJ - 1) / TJ), (unsigned)((a.Q + TQ - 1) / TQ));
  if constexpr (OP == OP_CDIST || OP == OP_SUBSIN || OP == OP_ADDSIN) {
    if (a.amb && aligned && !a.scores_out && a.d <= 2048) {          // two-stage: fast tile pass + exact re-score
      KGE_CUDA_OK(cudaMemsetAsync(a.amb_count, 0, 2 * sizeof(int), st));
      auto k = count_ranks_kernel<OP, true, true>;
      KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, EVAL_THREADS, smem, st>>>(a);
      KGE_CUDA_OK(cudaGetLastError());
      rescore_pairs_kernel<OP><<<148 * 4, 256, 0, st>>>(a.amb, a.amb_count, a.amb_capacity, a.qvec, a.X, a.d, a.De,
                                                        a.pos_score, a.queries, a.pos_col, a.gamma, a.counts, a.modulus);
      KGE_CUDA_OK(cudaGetLastError());
      return KGE_OK;
    }
  }
  if (aligned) {
    auto k = count_ranks_kernel<OP, true>;
    KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, EVAL_THREADS, smem, st>>>(a);
  } else {
    auto k = count_ranks_kernel<OP, false>;
    KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, EVAL_THREADS, smem, st>>>(a);
  }
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

static int eval_mode(int mode, bool &head) {
  if (mode == KGE_HEAD_BATCH) head = true;
  else if (mode == KGE_TAIL_BATCH) head = false;
  else {
    set_error("negative batch mode %d not supported", mode);     // dataloader.py:147
    return KGE_ERR_INVALID;
  }
  return KGE_OK;
}

}  // namespace kge

using namespace kge;

extern "C" int kge_eval_query_vectors(const kge_model_t *m, int mode, const int64_t *queries, int64_t Q, float *qvec,
                                      int32_t *err_flag, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  bool head;
  if ((rc = eval_mode(mode, head))) return rc;
  KGE_REQUIRE(queries && qvec, "null pointer");
  if (Q <= 0) return KGE_OK;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  const bool cplx = m->model == KGE_COMPLEX || m->model == KGE_ROTATE;
  const int d = cplx ? (int)(m->entity_dim / 2) : (int)m->entity_dim;
  const int grid = (int)(Q < 148 * 16 ? Q : 148 * 16);
  const float scale = phase_scale(m);
  cudaStream_t st = (cudaStream_t)stream;
#define KGE_QV(MODEL)                                                                                              \
  case MODEL:                                                                                                      \
    if (head)                                                                                                      \
      query_vectors_kernel<MODEL, true><<<grid, 256, 0, st>>>(m->entity, m->relation, queries, Q, m->nentity,       \
                                                              m->nrelation, d, (int)m->entity_dim,                 \
                                                              (int)m->relation_dim, scale, qvec, err_flag);        \
    else                                                                                                           \
      query_vectors_kernel<MODEL, false><<<grid, 256, 0, st>>>(m->entity, m->relation, queries, Q, m->nentity,      \
                                                               m->nrelation, d, (int)m->entity_dim,                \
                                                               (int)m->relation_dim, scale, qvec, err_flag);       \
    break;
  switch (m->model) {
    KGE_QV(KGE_TRANSE)
    KGE_QV(KGE_DISTMULT)
    KGE_QV(KGE_COMPLEX)
    KGE_QV(KGE_ROTATE)
    KGE_QV(KGE_PROTATE)
  }
#undef KGE_QV
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

extern "C" int kge_eval_phase_table(const kge_model_t *m, float *phase_table, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(phase_table, "null pointer");
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  phase_table_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(m->entity, m->nentity * m->entity_dim, phase_scale(m),
                                                              phase_table);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

static int fill_eval_args(const kge_model_t *m, bool head, const float *qvec, const int64_t *queries, int64_t Q,
                          const float *phase_table, EvalArgs &a) {
  const bool cplx = m->model == KGE_COMPLEX || m->model == KGE_ROTATE;
  KGE_REQUIRE(m->model != KGE_PROTATE || phase_table, "pRotatE evaluation needs kge_eval_phase_table");
  a.qvec = qvec;
  a.X = m->model == KGE_PROTATE ? phase_table : m->entity;
  a.queries = queries;
  a.modulus = m->modulus;
  a.Q = Q;
  a.nentity = m->nentity;
  a.d = cplx ? (int)(m->entity_dim / 2) : (int)m->entity_dim;
  a.De = (int)m->entity_dim;
  a.pos_col = head ? 0 : 2;
  a.words = (int)((m->nentity + 31) / 32);
  a.gamma = m->gamma;
  return KGE_OK;
}

extern "C" int kge_eval_positive_scores(const kge_model_t *m, int mode, const float *qvec, const int64_t *queries,
                                        int64_t Q, const float *phase_table, float *pos_score, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  bool head;
  if ((rc = eval_mode(mode, head))) return rc;
  KGE_REQUIRE(qvec && queries && pos_score, "null pointer");
  if (Q <= 0) return KGE_OK;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  EvalArgs a{};
  if ((rc = fill_eval_args(m, head, qvec, queries, Q, phase_table, a))) return rc;
  const int grid = (int)((Q + 3) / 4);                   // 4 warps per CTA, one warp per query
  cudaStream_t st = (cudaStream_t)stream;
#define KGE_PS(OP)                                                                                              \
  case OP:                                                                                                      \
    positive_scores_kernel<OP><<<grid, 128, 0, st>>>(a.qvec, a.X, a.queries, a.pos_col, a.Q, a.nentity, a.d, a.De, \
                                                     a.gamma, a.modulus, pos_score);                            \
    break;
  switch (op_of(m->model, head)) {
    KGE_PS(OP_SUBABS) KGE_PS(OP_ADDABS) KGE_PS(OP_MUL) KGE_PS(OP_CMUL) KGE_PS(OP_CDIST) KGE_PS(OP_SUBSIN)
    KGE_PS(OP_ADDSIN)
  }
#undef KGE_PS
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

static int count_ranks_impl(const kge_model_t *m, int mode, const float *qvec, const int64_t *queries, int64_t Q,
                            const float *phase_table, const float *pos_score, const uint32_t *filter_bits,
                            int64_t ent_begin, int64_t ent_end, int32_t *counts, float *scores_out, void *amb_pairs,
                            int64_t amb_capacity, int32_t *amb_count, void *stream);

extern "C" int kge_eval_count_ranks(const kge_model_t *m, int mode, const float *qvec, const int64_t *queries, int64_t Q,
                                    const float *phase_table, const float *pos_score, const uint32_t *filter_bits,
                                    int64_t ent_begin, int64_t ent_end, int32_t *counts, float *scores_out,
                                    void *stream) {
  return count_ranks_impl(m, mode, qvec, queries, Q, phase_table, pos_score, filter_bits, ent_begin, ent_end, counts,
                          scores_out, nullptr, 0, nullptr, stream);
}

extern "C" int kge_eval_count_ranks_two_stage(const kge_model_t *m, int mode, const float *qvec, const int64_t *queries,
                                              int64_t Q, const float *phase_table, const float *pos_score,
                                              const uint32_t *filter_bits, int64_t ent_begin, int64_t ent_end,
                                              int32_t *counts, void *amb_pairs, int64_t amb_capacity,
                                              int32_t *amb_count, void *stream) {
  KGE_REQUIRE(amb_pairs && amb_count && amb_capacity > 0, "two-stage ranking needs the ambiguous-pair buffers");
  return count_ranks_impl(m, mode, qvec, queries, Q, phase_table, pos_score, filter_bits, ent_begin, ent_end, counts,
                          nullptr, amb_pairs, amb_capacity, amb_count, stream);
}

static int count_ranks_impl(const kge_model_t *m, int mode, const float *qvec, const int64_t *queries, int64_t Q,
                            const float *phase_table, const float *pos_score, const uint32_t *filter_bits,
                            int64_t ent_begin, int64_t ent_end, int32_t *counts, float *scores_out, void *amb_pairs,
                            int64_t amb_capacity, int32_t *amb_count, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  bool head;
  if ((rc = eval_mode(mode, head))) return rc;
  KGE_REQUIRE(qvec && queries && pos_score && filter_bits && counts, "null pointer");
  KGE_REQUIRE(ent_begin >= 0 && ent_begin <= ent_end && ent_end <= m->nentity, "entity slice [%lld,%lld) outside [0,%lld)",
              (long long)ent_begin, (long long)ent_end, (long long)m->nentity);
  if (Q <= 0 || ent_begin == ent_end) return KGE_OK;
  KGE_REQUIRE((Q + TQ - 1) / TQ <= 65535, "at most %d queries per call", 65535 * TQ);
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  EvalArgs a{};
  if ((rc = fill_eval_args(m, head, qvec, queries, Q, phase_table, a))) return rc;
  a.pos_score = pos_score; a.filter_bits = filter_bits; a.counts = counts; a.scores_out = scores_out;
  a.ent_begin = ent_begin; a.ent_end = ent_end;
  a.amb = (int2 *)amb_pairs; a.amb_count = amb_count; a.amb_capacity = (int)amb_capacity;
  const bool aligned = (a.d % 4 == 0) && (a.De % 4 == 0) && ((((uintptr_t)a.qvec | (uintptr_t)a.X) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
#define KGE_CR(OP) \
  case OP:         \
    return launch_count<OP>(a, aligned, st);
  switch (op_of(m->model, head)) {
    KGE_CR(OP_SUBABS) KGE_CR(OP_ADDABS) KGE_CR(OP_MUL) KGE_CR(OP_CMUL) KGE_CR(OP_CDIST) KGE_CR(OP_SUBSIN)
    KGE_CR(OP_ADDSIN)
  }
#undef KGE_CR
  set_error("unreachable");
  return KGE_ERR_INVALID;
}

extern "C" int kge_eval_filter_bits(const int64_t *csr_offsets, const int32_t *csr_entities, int64_t Q, int64_t nentity,
                                    uint32_t *filter_bits, void *stream) {
  KGE_REQUIRE(csr_offsets && filter_bits && Q >= 0 && nentity > 0, "bad arguments");
  if (Q == 0) return KGE_OK;
  const int words = (int)((nentity + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  KGE_CUDA_OK(cudaMemsetAsync(filter_bits, 0, sizeof(uint32_t) * (size_t)Q * words, st));
  const int grid = (int)(Q < 148 * 16 ? Q : 148 * 16);
  filter_bits_kernel<<<grid, 128, 0, st>>>(csr_offsets, csr_entities, Q, nentity, words, filter_bits);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

extern "C" int kge_eval_filter_bits_lookup(const int64_t *index_keys, const int64_t *index_offsets,
                                           const int32_t *index_entities, int64_t nkeys, const int64_t *queries,
                                           int64_t Q, int mode, int64_t nentity, int64_t nrelation,
                                           uint32_t *filter_bits, void *stream) {
  KGE_REQUIRE(filter_bits && queries && Q >= 0 && nentity > 0 && nrelation > 0 && nkeys >= 0, "bad arguments");
  KGE_REQUIRE(nkeys == 0 || (index_keys && index_offsets && index_entities), "null filter index");
  KGE_REQUIRE(mode == KGE_HEAD_BATCH || mode == KGE_TAIL_BATCH, "negative batch mode %d not supported", mode);
  if (Q == 0) return KGE_OK;
  const int words = (int)((nentity + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  KGE_CUDA_OK(cudaMemsetAsync(filter_bits, 0, sizeof(uint32_t) * (size_t)Q * words, st));
  if (nkeys == 0) return KGE_OK;
  const int grid = (int)(Q < 148 * 16 ? Q : 148 * 16);
  filter_lookup_kernel<<<grid, 128, 0, st>>>(index_keys, index_offsets, index_entities, nkeys, queries, Q,
                                             mode == KGE_HEAD_BATCH, nentity, nrelation, words, filter_bits);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

static size_t align256e(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int64_t kge_eval_filter_index_scratch_bytes(int64_t nentity, int64_t nrelation) {
  if (nentity <= 0 || nrelation <= 0) return 0;
  const int64_t nkeys = nentity * nrelation;
  return (int64_t)(align256e((size_t)nkeys * 4) + align256e((size_t)((nkeys + 1023) / 1024) * 4));
}

extern "C" int kge_eval_filter_index_build(const int64_t *triples, int64_t ntriples, int mode, int64_t nentity,
                                           int64_t nrelation, int32_t *offsets, int32_t *entities, void *scratch,
                                           int64_t scratch_bytes, int32_t *err_flag, void *stream) {
  KGE_REQUIRE(mode == KGE_HEAD_BATCH || mode == KGE_TAIL_BATCH, "negative batch mode %d not supported", mode);
  KGE_REQUIRE(nentity > 0 && nrelation > 0 && ntriples >= 0 && ntriples < (1ll << 31), "bad sizes");
  const int64_t nkeys = nentity * nrelation;
  KGE_REQUIRE(nkeys < (1ll << 31) - 1024, "key space of %lld entries is too large for the direct-address index",
              (long long)nkeys);
  KGE_REQUIRE(offsets && scratch && (entities || ntriples == 0) && (triples || ntriples == 0), "null pointer");
  KGE_REQUIRE(scratch_bytes >= kge_eval_filter_index_scratch_bytes(nentity, nrelation), "scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  int *cnt = offsets;
  int *cursor = (int *)scratch;
  int *tile_tot = (int *)((char *)scratch + align256e((size_t)nkeys * 4));
  KGE_CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)(nkeys + 1) * 4, st));
  int grid = (int)((ntriples + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  const int head = mode == KGE_HEAD_BATCH;
  filter_index_kernel<false><<<grid, 256, 0, st>>>(triples, ntriples, head, nentity, nrelation, cnt, nullptr, err_flag);
  KGE_CUDA_OK(cudaGetLastError());
  const int tiles = (int)((nkeys + 1023) / 1024);
  scan_tiles_kernel<<<tiles, 1024, 0, st>>>(cnt, cursor, tile_tot, nkeys);
  KGE_CUDA_OK(cudaGetLastError());
  scan_apply_kernel<<<tiles, 1024, 0, st>>>(cnt, cursor, tile_tot, nkeys);
  KGE_CUDA_OK(cudaGetLastError());
  filter_index_kernel<true><<<grid, 256, 0, st>>>(triples, ntriples, head, nentity, nrelation, cursor, entities, nullptr);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

extern "C" int kge_eval_filter_bits_lookup_dense(const int32_t *index_offsets, const int32_t *index_entities,
                                                 const int64_t *queries, int64_t Q, int mode, int64_t nentity,
                                                 int64_t nrelation, uint32_t *filter_bits, void *stream) {
  KGE_REQUIRE(mode == KGE_HEAD_BATCH || mode == KGE_TAIL_BATCH, "negative batch mode %d not supported", mode);
  KGE_REQUIRE(index_offsets && index_entities && filter_bits && Q >= 0 && nentity > 0 && nrelation > 0, "bad arguments");
  KGE_REQUIRE(queries || Q == 0, "null pointer");
  if (Q == 0) return KGE_OK;
  const int words = (int)((nentity + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  KGE_CUDA_OK(cudaMemsetAsync(filter_bits, 0, sizeof(uint32_t) * (size_t)Q * words, st));
  const int grid = (int)(Q < 148 * 16 ? Q : 148 * 16);
  filter_lookup_dense_kernel<<<grid, 128, 0, st>>>(index_offsets, index_entities, queries, Q, mode == KGE_HEAD_BATCH,
                                                   nentity, nrelation, words, filter_bits);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}
