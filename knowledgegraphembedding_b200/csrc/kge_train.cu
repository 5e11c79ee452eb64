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

#include "kge_train_args.cuh"

namespace kge {

// Exclusive scan of the histogram cnt[0..n) -> offsets (in place), cursor = copy, cnt[n] = total.  Two launches over
// 1024-entry tiles (a single-CTA scan cost 18 us at FB15k's 14,951 entities and 157 us at YAGO3-10's 123,182):
// scan_tiles_kernel scans each tile and leaves its total, scan_apply_kernel adds the totals of the preceding tiles.
__device__ __forceinline__ int block_exclusive_scan_1024(int v, int *warp_tot, int &total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int s = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += s;
    }
    warp_tot[lane] = t;                                    // inclusive totals of the warps
  }
  __syncthreads();
  total = warp_tot[31];
  return (warp ? warp_tot[warp - 1] : 0) + incl - v;
}

__global__ void __launch_bounds__(1024) scan_tiles_kernel(const int *__restrict__ cnt, int *__restrict__ cursor,
                                                          int *__restrict__ tile_tot, int64_t n) {
  __shared__ int warp_tot[32];
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  int total;
  const int excl = block_exclusive_scan_1024(i < n ? cnt[i] : 0, warp_tot, total);
  if (i < n) cursor[i] = excl;                             // offset inside the tile
  if (threadIdx.x == 0) tile_tot[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_apply_kernel(int *__restrict__ cnt, int *__restrict__ cursor,
                                                          const int *__restrict__ tile_tot, int64_t n) {
  __shared__ int warp_tot[32];
  __shared__ int prefix_sh;
  // sum of the totals of tiles [0, blockIdx.x): every thread adds a strided share, one block reduction
  int part = 0;
  for (int t = threadIdx.x; t < (int)blockIdx.x; t += 1024) part += tile_tot[t];
  int total;
  block_exclusive_scan_1024(part, warp_tot, total);
  if (threadIdx.x == 0) prefix_sh = total;
  __syncthreads();
  const int prefix = prefix_sh;
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  if (i < n) {
    const int o = cursor[i] + prefix;
    cnt[i] = o;
    cursor[i] = o;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) cnt[n] = prefix + tile_tot[blockIdx.x];
}

__global__ void scatter_pairs_kernel(const int64_t *__restrict__ cand, int64_t cand_stride, int64_t row_begin, int rows,
                                     int N, int64_t nentity, const int *__restrict__ ids32, const float *__restrict__ G,
                                     const int *__restrict__ dids, int *__restrict__ cursor, int *__restrict__ perm,
                                     float *__restrict__ gsorted) {
  const int64_t pairs = (int64_t)rows * N;
  const int64_t total = pairs + (dids ? 3 * (int64_t)rows : 0);
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    if (p < pairs) {
      const int rl = (int)(p / N), n = (int)(p % N);
      int64_t id = ids32 ? ids32[p] : cand[(row_begin + rl) * cand_stride + n];     // (ids32: clamped by the row kernel)
      if ((uint64_t)id >= (uint64_t)nentity) id = 0;
      const int pos = atomicAdd(cursor + id, 1);
      perm[pos] = rl;                                      // the entity pass only needs the q row and dL/ds
      gsorted[pos] = G[p];
    } else {                                               // direct gradient row i of the positive triples (fused optimizer)
      const int i = (int)(p - pairs);
      const int pos = atomicAdd(cursor + dids[i], 1);
      perm[pos] = -(1 + i);
      gsorted[pos] = 0.f;
    }
  }
}

// ---- entity-sharded multi-GPU step: the owner of an entity range sorts the pairs gathered from ALL ranks ----------
// Gather area of every rank's peer block (identical layout; R = row capacity per rank, section s is written by rank s):
//   ids [G][R][N] int32 candidate ids | Gs [G][R][N] dL/ds | Qtab [G][R][De] query vectors |
//   Dvec [G][3R][De] gradient rows of the positive triples (a row is present only in its owner's block) | dids [G][3R] |
//   hist [nentity] this rank's pairs per entity (all entities; local atomics of its row kernel): the owner of an entity
//   range sums the G histograms of that range with NVLink loads
struct GatherLayout { size_t ids, Gs, Qtab, Dvec, dids, hist, total; };
static size_t align256(size_t x);
static GatherLayout gather_layout(int world, int64_t R, int64_t N, int64_t De, int64_t nentity) {
  GatherLayout g;
  size_t o = 0;
  g.ids = o;  o += align256((size_t)world * R * N * 4);
  g.Gs = o;   o += align256((size_t)world * R * N * 4);
  g.Qtab = o; o += align256((size_t)world * R * De * 4);
  g.Dvec = o; o += align256((size_t)world * 3 * R * De * 4);
  g.dids = o; o += align256((size_t)world * 3 * R * 4);
  g.hist = o; o += align256((size_t)nentity * 4);
  g.total = o;
  return g;
}

struct GatherArgs {
  const int *ids;
  const float *Gs;
  const int *dids;
  int world, R, N;
  int rows_of[KGE_PEER_MAX_RANKS];
  int eb, ee;                                              // entity range this rank owns
};

// this rank's candidate ids, as int32, into section `rank` of every block (clamped like the row kernel's gathers)
__global__ void push_ids_kernel(const int64_t *__restrict__ neg, int64_t n, int64_t nentity, int *dst, const Mirror mir,
                                int32_t *err) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t id = neg[i];
    if ((uint64_t)id >= (uint64_t)nentity) { if (err) *err = 1; id = 0; }
    store_all(mir, dst + i, (int)id);
  }
}

// scan_tiles_kernel over the sum of all ranks' histograms, restricted to the owned entity range
__global__ void __launch_bounds__(1024) scan_tiles_gathered_kernel(const int *hist, const Mirror mir, int eb, int ee,
                                                                   int *__restrict__ cursor, int *__restrict__ tile_tot,
                                                                   int64_t n) {
  __shared__ int warp_tot[32];
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  int v = 0;
  if (i >= eb && i < ee) {
    int x[KGE_PEER_MAX_RANKS];                             // all peers' loads in flight together (one NVLink round trip)
#pragma unroll
    for (int r = 0; r < KGE_PEER_MAX_RANKS; ++r)
      if (r < mir.world) asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(x[r]) : "l"(at_rank(mir, hist + i, r)));
#pragma unroll
    for (int r = 0; r < KGE_PEER_MAX_RANKS; ++r)
      if (r < mir.world) v += x[r];
  }
  int total;
  const int excl = block_exclusive_scan_1024(v, warp_tot, total);
  if (i < n) cursor[i] = excl;
  if (threadIdx.x == 0) tile_tot[blockIdx.x] = total;
}

// Scatter of the gathered pairs and direct rows whose entity lies in the owned range [eb, ee): perm = gathered row index
// s * R + rl, or -(1 + gathered direct index).  blockIdx.y = source rank; VEC ids (one 16-byte load) per thread and trip.
template <int VEC>
__global__ void gathered_scatter_kernel(const GatherArgs g, int *__restrict__ cursor, int *__restrict__ perm,
                                        float *__restrict__ gsorted) {
  const int s = blockIdx.y, rows = g.rows_of[s];
  const int n = rows * g.N;                                // (world * R * (N + 3) < 2^31 is checked by the host)
  const size_t base = (size_t)s * g.R * g.N;
  const int stride = gridDim.x * blockDim.x * VEC;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < n; i += stride) {
    int id[VEC];
    float gv[VEC];
    if constexpr (VEC == 4) {
      const int4 a = *reinterpret_cast<const int4 *>(g.ids + base + i);
      const float4 b = *reinterpret_cast<const float4 *>(g.Gs + base + i);
      id[0] = a.x; id[1] = a.y; id[2] = a.z; id[3] = a.w;
      gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
    } else {
      id[0] = g.ids[base + i];
      gv[0] = g.Gs[base + i];
    }
    const int row = s * g.R + i / g.N;                     // VEC == 4: N % 4 == 0, the four pairs share their row
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (id[v] < g.eb || id[v] >= g.ee) continue;
      const int pos = atomicAdd(cursor + id[v], 1);
      perm[pos] = row;
      gsorted[pos] = gv[v];
    }
  }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < 3 * rows; j += gridDim.x * blockDim.x) {
    const int i = s * 3 * g.R + j;
    const int id = g.dids[i];
    if (id < g.eb || id >= g.ee) continue;
    const int pos = atomicAdd(cursor + id, 1);
    perm[pos] = -(1 + i);
    gsorted[pos] = 0.f;
  }
}

// debug instrumentation of row_kernel_split (KGE_ROW_PHASES=1): where do the cycles of a row go?
static unsigned long long *g_phase_dev = nullptr;
unsigned long long *row_phase_counters() {
  if (!getenv("KGE_ROW_PHASES")) return nullptr;
  if (!g_phase_dev) {
    if (cudaMalloc(&g_phase_dev, 8 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    cudaMemset(g_phase_dev, 0, 8 * sizeof(unsigned long long));
  }
  return g_phase_dev;
}

// ---- host side ---------------------------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t split_workspace_bytes(int64_t rows, int64_t N, int64_t De, int64_t nentity) {
  return align256(rows * N * 4) + align256(rows * De * 4) + align256((nentity + 1) * 4 + nentity * 4 + 64) +
         align256(((nentity + 1023) / 1024) * 4) + 2 * align256(rows * (N + 3) * 4) + align256(rows * 3 * De * 4) +
         align256(rows * 3 * 4) + align256(rows * N * 4);
}

SplitWs carve_split_ws(void *workspace, int64_t rows, int64_t N, int64_t De, int64_t nentity) {
  char *wp = (char *)workspace;
  SplitWs ws;
  ws.G = (float *)wp;      wp += align256((size_t)rows * N * 4);
  ws.Qtab = (float *)wp;   wp += align256((size_t)rows * De * 4);
  ws.cnt = (int *)wp;
  ws.cursor = ws.cnt + (nentity + 1);
  ws.queue = ws.cursor + nentity;                          // 16 queue counters (one per entity slice)
  wp += align256((size_t)(nentity + 1) * 4 + (size_t)nentity * 4 + 64);
  ws.tile_tot = (int *)wp; wp += align256((size_t)((nentity + 1023) / 1024) * 4);
  ws.perm = (int *)wp;     wp += align256((size_t)rows * (N + 3) * 4);
  ws.gsorted = (float *)wp; wp += align256((size_t)rows * (N + 3) * 4);
  ws.Dvec = (float *)wp;   wp += align256((size_t)rows * 3 * De * 4);
  ws.dids = (int *)wp;     wp += align256((size_t)rows * 3 * 4);
  ws.ids32 = (int *)wp;
  return ws;
}

// The single-read path needs 16-byte rows that fit one k-tile, >= 8 candidates and 32-bit pair counts.  With dense
// gradients it pays a fixed cost per touched entity, so it wins when an entity collects several pairs (17 at FB15k
// shapes: 0.73 vs 0.96 ms) and loses when pairs are sparse (3.3 at YAGO3-10 shapes: 1.59 vs 0.99 ms).  With the entity
// table's Adam update fused into the entity pass every entity is visited anyway (dense Adam) and the pass replaces the
// zero-fill, the atomic scatter and the optimizer's 7 table streams: it wins at every density (YAGO3-10 shapes: 1.06 vs
// 1.63 ms per step).  KGE_FORCE_SPLIT / KGE_NO_SPLIT override.
bool split_path_shape_ok(int64_t rows, int64_t N, int64_t De, int64_t d, bool cplx, int64_t nentity, bool fused_adam) {
  if (getenv("KGE_NO_SPLIT") || getenv("KGE_NO_TMA")) return false;
  if (d % 4 || De % 4 || d / 4 > 32 * (cplx ? 8 : 16)) return false;
  if (N < 8 || nentity >= (1ll << 31) || rows * (N + 3) >= (1ll << 31)) return false;
  return fused_adam || rows * N >= 6 * nentity || getenv("KGE_FORCE_SPLIT");
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
  const bool aligned = (((uintptr_t)m->entity | (uintptr_t)a.gE | (uintptr_t)m->relation) & 15) == 0 && m->relation_dim % 4 == 0;
  const bool vec4 = aligned && (a.d % 4 == 0) && (m->entity_dim % 4 == 0);
  const int Dq = (int)m->entity_dim;
  size_t smem = sizeof(float) * (2 * (size_t)((Dq + 3) & ~3) + (a.do_loss ? 2 * (size_t)a.N : 0) + 32);
  KGE_REQUIRE(smem <= 227 * 1024, "entity_dim=%d with %d candidates per row needs %zu B of shared memory (max 232448)",
              Dq, a.N, smem);
  // one warp per candidate in flight; a single-candidate pass ('single' mode) only needs a few warps for q
  int threads = a.N >= 16 ? 512 : (a.N >= 4 ? 256 : 128);
#define KGE_ROWS(MODEL) \
  case MODEL: return launch_rows_model<MODEL>(head, a, vec4, threads, smem, workspace, workspace_bytes, st);
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
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
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
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  a.positive = positive; a.row_begin = 0; a.row_count = (int)B; a.N = (int)N;
  a.dscore = dscore; a.gE = grad_entity; a.gR = grad_relation; a.gM = grad_modulus; a.err = err_flag;
  return launch_rows(m, head, a, (cudaStream_t)stream, workspace, (size_t)workspace_bytes);
}

extern "C" int64_t kge_train_workspace_bytes(const kge_model_t *m, int64_t rows, int64_t N);

static int train_rows_impl(const kge_model_t *m, int mode, int loss_kind, float adversarial_temperature,
                           const int64_t *positive, const int64_t *negative, const float *weight,
                           const float *weight_sum, int64_t B_total, int64_t row_begin, int64_t row_count,
                           int64_t N, float *row_loss, float *pos_row_loss, float *grad_entity,
                           float *grad_relation, float *grad_modulus, float *score_out, void *workspace,
                           int64_t workspace_bytes, int32_t *err_flag, const EntityAdam *entity_adam, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(positive && row_loss && grad_relation && (grad_entity || entity_adam), "null pointer");
  KGE_REQUIRE(m->model != KGE_PROTATE || grad_modulus, "pRotatE needs grad_modulus");
  KGE_REQUIRE(loss_kind >= KGE_LOSS_NEG_ADVERSARIAL && loss_kind <= KGE_LOSS_POSITIVE, "bad loss_kind %d", loss_kind);
  KGE_REQUIRE(!weight || weight_sum, "subsampling weights need their sum (kge_weight_sum)");
  KGE_REQUIRE(row_begin >= 0 && row_begin + row_count <= B_total, "row slice [%lld,+%lld) outside batch of %lld",
              (long long)row_begin, (long long)row_count, (long long)B_total);
  RowArgs a{};
  bool head;
  if (loss_kind == KGE_LOSS_POSITIVE) { mode = KGE_SINGLE; N = 1; }
  if ((rc = resolve_mode(mode, positive, negative, N, a, head))) return rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  a.positive = positive; a.row_begin = row_begin; a.row_count = (int)row_count; a.N = (int)N;
  a.do_loss = 1; a.loss_kind = loss_kind; a.alpha = adversarial_temperature;
  a.weight = weight; a.wsum = weight_sum; a.uniform_u = 1.0f / (float)B_total;
  a.row_loss = row_loss; a.score_out = score_out;
  a.pos_row_loss = loss_kind == KGE_LOSS_POSITIVE ? nullptr : pos_row_loss;
  a.gE = grad_entity; a.gR = grad_relation; a.gM = grad_modulus; a.err = err_flag;
  int fused_positive = 0, adam_applied = 0;
  a.fused_positive = &fused_positive;
  a.entity_adam = entity_adam;
  a.entity_adam_applied = &adam_applied;
  rc = launch_rows(m, head, a, (cudaStream_t)stream, workspace, (size_t)workspace_bytes);
  if (rc) return rc;
  KGE_REQUIRE(!entity_adam || adam_applied || row_count == 0,
              "internal: kge_train_plan promised the fused entity optimizer for this shape but the launcher declined");
  if (!a.pos_row_loss || fused_positive) return rc;
  // the kernel variant that ran has no fused positive pass: run the 'single' pass as its own launch
  return train_rows_impl(m, KGE_SINGLE, KGE_LOSS_POSITIVE, 1.0f, positive, negative, weight, weight_sum, B_total,
                         row_begin, row_count, 1, pos_row_loss, nullptr, grad_entity, grad_relation, grad_modulus,
                         nullptr, nullptr, 0, err_flag, nullptr, stream);
}

extern "C" int kge_train_rows(const kge_model_t *m, int mode, int loss_kind, float adversarial_temperature,
                              const int64_t *positive, const int64_t *negative, const float *weight,
                              const float *weight_sum, int64_t B_total, int64_t row_begin, int64_t row_count,
                              int64_t N, float *row_loss, float *pos_row_loss, float *grad_entity,
                              float *grad_relation, float *grad_modulus, float *score_out, void *workspace,
                              int64_t workspace_bytes, int32_t *err_flag, void *stream) {
  KGE_REQUIRE(grad_entity, "null pointer");
  return train_rows_impl(m, mode, loss_kind, adversarial_temperature, positive, negative, weight, weight_sum, B_total,
                         row_begin, row_count, N, row_loss, pos_row_loss, grad_entity, grad_relation, grad_modulus,
                         score_out, workspace, workspace_bytes, err_flag, nullptr, stream);
}

static bool plan_split(const kge_model_t *m, int64_t rows, int64_t N, bool fused_adam) {
  const bool cplx = m->model == KGE_COMPLEX || m->model == KGE_ROTATE;
  const int64_t d = cplx ? m->entity_dim / 2 : m->entity_dim;
  return (((uintptr_t)m->entity | (uintptr_t)m->relation) & 15) == 0 && m->relation_dim % 4 == 0 && N <= 8192 &&
         split_path_shape_ok(rows, N, m->entity_dim, d, cplx, m->nentity, fused_adam);
}

extern "C" int kge_train_plan(const kge_model_t *m, int64_t rows, int64_t N) {
  if (check_model(m) || rows <= 0 || N <= 0) return 0;
  // the fit of the kernel's shared-memory carve-up (>= 4 warps next to q, dq, 2N scores) holds for every N that passes
  // launch_rows' own limit when rows fit one k-tile, except absurdly long candidate lists
  return (plan_split(m, rows, N, false) ? KGE_PLAN_SINGLE_READ : 0) | (plan_split(m, rows, N, true) ? KGE_PLAN_ENTITY_ADAM : 0);
}

extern "C" int kge_train_rows_adam(const kge_model_t *m, int mode, int loss_kind, float adversarial_temperature,
                                   const int64_t *positive, const int64_t *negative, const float *weight,
                                   const float *weight_sum, int64_t B_total, int64_t row_count, int64_t N,
                                   float *row_loss, float *pos_row_loss, float *grad_relation, float *grad_modulus,
                                   void *workspace, int64_t workspace_bytes, int32_t *err_flag,
                                   const kge_entity_adam_t *host_entity_adam, void *stream) {
  KGE_REQUIRE(m && host_entity_adam && host_entity_adam->exp_avg && host_entity_adam->exp_avg_sq &&
              host_entity_adam->step >= 1, "bad entity optimizer state");
  KGE_REQUIRE(loss_kind == KGE_LOSS_NEG_ADVERSARIAL || loss_kind == KGE_LOSS_NEG_UNIFORM, "bad loss_kind %d", loss_kind);
  KGE_REQUIRE(pos_row_loss && workspace, "the fused optimizer needs pos_row_loss and the workspace");
  KGE_REQUIRE(row_count == B_total, "the fused entity optimizer is a single-device path (whole batch)");
  KGE_REQUIRE(kge_train_plan(m, row_count, N) & KGE_PLAN_ENTITY_ADAM,
              "kge_train_plan does not offer the fused entity optimizer for this shape");
  KGE_REQUIRE(workspace_bytes >= kge_train_workspace_bytes(m, row_count, N), "workspace too small");
  const kge_entity_adam_t &o = *host_entity_adam;
  EntityAdam ea{};
  ea.exp_avg = o.exp_avg; ea.exp_avg_sq = o.exp_avg_sq;
  ea.s = adam_scalars(o.lr, o.beta1, o.beta2, o.eps, o.l3_coefficient, o.step);
  ea.l3 = o.l3_coefficient != 0.0;
  ea.reg_partials = o.reg_partials; ea.n_reg_partials = o.n_reg_partials;
  return train_rows_impl(m, mode, loss_kind, adversarial_temperature, positive, negative, weight, weight_sum, B_total, 0,
                         row_count, N, row_loss, pos_row_loss, nullptr, grad_relation, grad_modulus, nullptr, workspace,
                         workspace_bytes, err_flag, &ea, stream);
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
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
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
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  const bool cplx = m->model == KGE_COMPLEX || m->model == KGE_ROTATE;
  RowArgs a{};
  a.E = m->entity; a.modulus = m->modulus; a.nentity = m->nentity;
  a.De = (int)m->entity_dim; a.d = cplx ? (int)(m->entity_dim / 2) : (int)m->entity_dim;
  a.scale = phase_scale(m); a.N = (int)N; a.row_count = (int)row_count;
  a.gE = grad_entity; a.gM = grad_modulus; a.do_loss = 1;
  SplitWs ws = carve_split_ws(workspace, row_count, N, m->entity_dim, m->nentity);
  ws.Dvec = nullptr;                                      // (the sliced multi-GPU flow keeps the dense gradient)
  const bool head = mode == KGE_HEAD_BATCH;
  cudaStream_t st = (cudaStream_t)stream;
  // the persistent entity kernel would otherwise hold every SM and the all-reduce of the previous slice could not start
  const char *rs = getenv("KGE_ENTITY_SMS_RESERVE");
  const int reserve = rs ? atoi(rs) : 0;       // measured at 2 GPUs: 0, 8, 24 equal, 48 slower
#define KGE_ENT(MODEL) \
  case MODEL: return launch_entity_model<MODEL>(head, a, ws, ent_begin, ent_end, slice_index, st, reserve);
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

// ---- entity-sharded multi-GPU step (owner computes): see include/kge_b200.h ------------------------------------------
static int shard_mirror(const kge_model_t *m, const kge_shard_t *sh, int64_t N, Mirror &mir, GatherLayout &L) {
  KGE_REQUIRE(sh && sh->world >= 2 && sh->world <= KGE_PEER_MAX_RANKS && sh->rank >= 0 && sh->rank < sh->world,
              "bad shard description");
  KGE_REQUIRE(sh->rows_max >= 1 && m->nentity >= sh->world && m->nentity < (1ll << 31), "bad shard shape");
  KGE_REQUIRE(sh->world * sh->rows_max * (N + 3) < (1ll << 31), "too many pairs for 32-bit sort positions");
  L = gather_layout(sh->world, sh->rows_max, N, m->entity_dim, m->nentity);
  KGE_REQUIRE(sh->gather_offset >= 0 && sh->gather_offset % 256 == 0 &&
                  (size_t)sh->gather_offset + L.total <= (size_t)sh->block_bytes,
              "gather area does not fit the peer block");
  const char *base = (const char *)sh->block[sh->rank];
  KGE_REQUIRE(base && (const char *)m->entity >= base &&
                  (const char *)(m->entity + m->nentity * m->entity_dim) <= base + sh->block_bytes,
              "the entity table must live inside the local peer block (the owners store updated rows into it)");
  mir = Mirror{};
  mir.world = sh->world; mir.rank = sh->rank;
  mir.ent_base = (int)(m->nentity / sh->world); mir.ent_rem = (int)(m->nentity % sh->world);
  for (int r = 0; r < sh->world; ++r) {
    KGE_REQUIRE(sh->block[r] && sh->rows_of[r] >= 0 && sh->rows_of[r] <= sh->rows_max, "bad shard entry %d", r);
    mir.delta[r] = (long long)((const char *)sh->block[r] - base);
  }
  if (sh->multicast) {
    KGE_REQUIRE(((uintptr_t)sh->multicast & 15) == 0 && ((uintptr_t)base & 15) == 0, "multicast mapping is not 16-byte aligned");
    mir.mc_delta = (long long)((const char *)sh->multicast - base);
  }
  return KGE_OK;
}

extern "C" int64_t kge_train_gather_bytes(const kge_model_t *m, int world, int64_t rows_max, int64_t N) {
  if (!m || world < 1 || rows_max < 1 || N < 1) return 0;
  return (int64_t)gather_layout(world, rows_max, N, m->entity_dim, m->nentity).total;
}

extern "C" int64_t kge_train_shard_workspace_bytes(const kge_model_t *m, int world, int64_t rows_max, int64_t N) {
  if (!m || world < 1 || rows_max < 1 || N < 1) return 0;
  return (int64_t)(align256((size_t)(2 * m->nentity + 1) * 4 + 64) + align256((size_t)((m->nentity + 1023) / 1024) * 4) +
                   2 * align256((size_t)world * rows_max * (N + 3) * 4));
}

extern "C" int kge_train_rows_sharded(const kge_model_t *m, int mode, int loss_kind, float adversarial_temperature,
                                      const int64_t *positive, const int64_t *negative, const float *weight,
                                      const float *weight_sum, int64_t B_total, int64_t row_count, int64_t N,
                                      float *row_loss, float *pos_row_loss, float *grad_relation, float *grad_modulus,
                                      const kge_shard_t *host_shard, int32_t *err_flag, void *stream, void *aux_stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(loss_kind == KGE_LOSS_NEG_ADVERSARIAL || loss_kind == KGE_LOSS_NEG_UNIFORM, "bad loss_kind %d", loss_kind);
  KGE_REQUIRE(mode == KGE_HEAD_BATCH || mode == KGE_TAIL_BATCH, "mode %d not supported", mode);
  KGE_REQUIRE(m->model != KGE_PROTATE || grad_modulus, "pRotatE needs grad_modulus");
  KGE_REQUIRE(!weight || weight_sum, "subsampling weights need their sum (kge_weight_sum)");
  Mirror mir;
  GatherLayout L;
  if ((rc = shard_mirror(m, host_shard, N, mir, L))) return rc;
  const kge_shard_t &sh = *host_shard;
  KGE_REQUIRE(row_count == sh.rows_of[sh.rank] && row_count <= B_total, "row_count does not match the shard description");
  KGE_REQUIRE(kge_train_plan(m, sh.rows_max, N) & KGE_PLAN_ENTITY_ADAM,
              "kge_train_plan does not offer the fused entity optimizer for this shape");
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char *g = (char *)sh.block[sh.rank] + sh.gather_offset;
  // this rank's histogram: the owners sum it whether or not the rank holds rows in this step (a short last batch can
  // leave a rank empty after steps in which it counted pairs); every peer's scan of the last step is done
  KGE_CUDA_OK(cudaMemsetAsync(g + L.hist, 0, (size_t)m->nentity * 4, st));
  if (row_count == 0) return KGE_OK;
  KGE_REQUIRE(positive && negative && row_loss && pos_row_loss && grad_relation, "null pointer");
  const int64_t R = sh.rows_max, De = m->entity_dim;
  // the id mirror is independent of the row kernel: with an auxiliary stream it runs next to it (joined at the end)
  cudaStream_t aux = aux_stream ? (cudaStream_t)aux_stream : st;
  static thread_local cudaEvent_t fork_ev[16] = {}, join_ev[16] = {};
  const int dv = m->device & 15;
  if (aux != st) {
    if (!fork_ev[dv]) {
      KGE_CUDA_OK(cudaEventCreateWithFlags(&fork_ev[dv], cudaEventDisableTiming));
      KGE_CUDA_OK(cudaEventCreateWithFlags(&join_ev[dv], cudaEventDisableTiming));
    }
    KGE_CUDA_OK(cudaEventRecord(fork_ev[dv], st));
    KGE_CUDA_OK(cudaStreamWaitEvent(aux, fork_ev[dv], 0));
  }
  {
    const int64_t n = row_count * N;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    push_ids_kernel<<<grid, 256, 0, aux>>>(negative, n, m->nentity, (int *)(g + L.ids) + (size_t)sh.rank * R * N, mir,
                                           err_flag);
    KGE_CUDA_OK(cudaGetLastError());
  }
  if (aux != st) KGE_CUDA_OK(cudaEventRecord(join_ev[dv], aux));
  SplitWs ws{};
  ws.G = (float *)(g + L.Gs) + (size_t)sh.rank * R * N;
  ws.Qtab = (float *)(g + L.Qtab) + (size_t)sh.rank * R * De;
  ws.Dvec = (float *)(g + L.Dvec) + (size_t)sh.rank * 3 * R * De;
  ws.dids = (int *)(g + L.dids) + (size_t)sh.rank * 3 * R;
  ws.cnt = (int *)(g + L.hist);                            // this rank's histogram; the owners read their ranges of it
  RowArgs a{};
  bool head;
  if ((rc = resolve_mode(mode, positive, negative, N, a, head))) return rc;
  a.positive = positive; a.row_begin = 0; a.row_count = (int)row_count; a.N = (int)N;
  a.do_loss = 1; a.loss_kind = loss_kind; a.alpha = adversarial_temperature;
  a.weight = weight; a.wsum = weight_sum; a.uniform_u = 1.0f / (float)B_total;
  a.row_loss = row_loss; a.pos_row_loss = pos_row_loss;
  a.gR = grad_relation; a.gM = grad_modulus; a.err = err_flag;
  int fused_positive = 0, deferred = 0;
  a.fused_positive = &fused_positive;
  a.defer_entity = 1;
  a.entity_deferred = &deferred;
  a.mir = mir;
  a.shard_ws = &ws;
  rc = launch_rows(m, head, a, st);
  if (aux != st) KGE_CUDA_OK(cudaStreamWaitEvent(st, join_ev[dv], 0));
  if (rc) return rc;
  KGE_REQUIRE(deferred && fused_positive, "internal: the single-read row kernel did not run for the entity-sharded step");
  return KGE_OK;
}

extern "C" int kge_train_entity_sharded(const kge_model_t *m, int mode, int64_t N, const kge_shard_t *host_shard,
                                        void *workspace, int64_t workspace_bytes,
                                        const kge_entity_adam_t *host_entity_adam, int32_t *err_flag, void *stream) {
  int rc = check_model(m);
  if (rc) return rc;
  KGE_REQUIRE(mode == KGE_HEAD_BATCH || mode == KGE_TAIL_BATCH, "mode %d not supported", mode);
  KGE_REQUIRE(workspace && host_entity_adam && host_entity_adam->exp_avg && host_entity_adam->exp_avg_sq &&
                  host_entity_adam->step >= 1, "bad entity optimizer state");
  Mirror mir;
  GatherLayout L;
  if ((rc = shard_mirror(m, host_shard, N, mir, L))) return rc;
  const kge_shard_t &sh = *host_shard;
  KGE_REQUIRE(workspace_bytes >= kge_train_shard_workspace_bytes(m, sh.world, sh.rows_max, N), "workspace too small");
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(m->device))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t R = sh.rows_max, nentity = m->nentity;
  const char *g = (const char *)sh.block[sh.rank] + sh.gather_offset;
  // local sort arrays
  SplitWs ws{};
  char *wp = (char *)workspace;
  ws.cnt = (int *)wp;
  ws.cursor = ws.cnt + (nentity + 1);
  ws.queue = ws.cursor + nentity;
  wp += align256((size_t)(2 * nentity + 1) * 4 + 64);
  ws.tile_tot = (int *)wp; wp += align256((size_t)((nentity + 1023) / 1024) * 4);
  const size_t cap = (size_t)sh.world * R * (N + 3);
  ws.perm = (int *)wp;      wp += align256(cap * 4);
  ws.gsorted = (float *)wp;
  ws.Qtab = (float *)(g + L.Qtab);
  ws.Dvec = (float *)(g + L.Dvec);
  KGE_CUDA_OK(cudaMemsetAsync(ws.queue, 0, 16 * 4, st));
  GatherArgs ga{};
  ga.ids = (const int *)(g + L.ids); ga.Gs = (const float *)(g + L.Gs); ga.dids = (const int *)(g + L.dids);
  ga.world = sh.world; ga.R = (int)R; ga.N = (int)N;
  for (int r = 0; r < sh.world; ++r) ga.rows_of[r] = sh.rows_of[r];
  ga.eb = sh.rank * mir.ent_base + (sh.rank < mir.ent_rem ? sh.rank : mir.ent_rem);
  ga.ee = ga.eb + mir.ent_base + (sh.rank < mir.ent_rem ? 1 : 0);
  {
    // every rank histogrammed its own pairs while its row kernel ran; the tile scan sums the G histograms of the owned
    // range (NVLink loads) and treats every other entity as empty
    const int *hist = (const int *)(g + L.hist);
    const int tiles = (int)((nentity + 1023) / 1024);
    scan_tiles_gathered_kernel<<<tiles, 1024, 0, st>>>(hist, mir, ga.eb, ga.ee, ws.cursor, ws.tile_tot, nentity);
    KGE_CUDA_OK(cudaGetLastError());
    scan_apply_kernel<<<tiles, 1024, 0, st>>>(ws.cnt, ws.cursor, ws.tile_tot, nentity);
    KGE_CUDA_OK(cudaGetLastError());
    const bool vec = N % 4 == 0;
    const int64_t per = R * N / (vec ? 4 : 1);
    int gx = (int)((per + 255) / 256);
    if (gx > 148 * 16 / sh.world) gx = 148 * 16 / sh.world;
    if (gx < 1) gx = 1;
    if (vec) gathered_scatter_kernel<4><<<dim3(gx, sh.world), 256, 0, st>>>(ga, ws.cursor, ws.perm, ws.gsorted);
    else gathered_scatter_kernel<1><<<dim3(gx, sh.world), 256, 0, st>>>(ga, ws.cursor, ws.perm, ws.gsorted);
    KGE_CUDA_OK(cudaGetLastError());
  }
  const kge_entity_adam_t &o = *host_entity_adam;
  EntityAdam ea{};
  ea.exp_avg = o.exp_avg; ea.exp_avg_sq = o.exp_avg_sq;
  ea.s = adam_scalars(o.lr, o.beta1, o.beta2, o.eps, o.l3_coefficient, o.step);
  ea.l3 = o.l3_coefficient != 0.0;
  ea.reg_partials = o.reg_partials; ea.n_reg_partials = o.n_reg_partials;
  const bool cplx = m->model == KGE_COMPLEX || m->model == KGE_ROTATE;
  RowArgs a{};
  a.E = m->entity; a.modulus = m->modulus; a.nentity = nentity;
  a.De = (int)m->entity_dim; a.d = cplx ? (int)(m->entity_dim / 2) : (int)m->entity_dim;
  a.scale = phase_scale(m); a.N = (int)N; a.row_count = (int)(sh.world * R);
  a.do_loss = 1; a.err = err_flag; a.entity_adam = &ea; a.mir = mir;
  const bool head = mode == KGE_HEAD_BATCH;
#define KGE_ENT(MODEL) \
  case MODEL: return launch_entity_model<MODEL>(head, a, ws, ga.eb, ga.ee, 0, st, 0);
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

/* debug: cycles of thread 0 of every row_kernel_split CTA, summed over rows and launches since the last reset, per
 * phase: [0] query vector, [1] candidate loop, [2] row loss, [3] fold, [4] chain rule, [5] positive triple,
 * [6] of [1]: waiting on the TMA mbarrier (warp 0).  Only counted when KGE_ROW_PHASES=1. */
extern "C" int kge_debug_row_phase_cycles(uint64_t *host_out8, int reset) {
  KGE_REQUIRE(host_out8, "null pointer");
  for (int i = 0; i < 8; ++i) host_out8[i] = 0;
  if (!g_phase_dev) return KGE_OK;
  KGE_CUDA_OK(cudaMemcpy(host_out8, g_phase_dev, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) KGE_CUDA_OK(cudaMemset(g_phase_dev, 0, 8 * sizeof(unsigned long long)));
  return KGE_OK;
}
