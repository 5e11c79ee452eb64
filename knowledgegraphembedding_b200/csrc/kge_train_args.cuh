// kge_train_args.cuh -- argument blocks of the train-path kernels and the host entry points that are instantiated
// per model in their own translation units (kge_train_inst.cu, built once per model: parallel compiles).
#pragma once
#include "kge_adam.cuh"
#include "kge_rows.cuh"

namespace kge {

// Adam update of the entity table fused into the entity-major pass (kge_train_rows_adam): host-side description
struct EntityAdam {
  float *exp_avg, *exp_avg_sq;   // [nentity, De]
  AdamScalars s;
  int l3;                        // 1: add 3*l3*x*|x| to the gradient and accumulate sum|x|^3 (model.py:290-297)
  double *reg_partials;          // [>= grid] doubles, zeroed by the launcher
  int64_t n_reg_partials;
};

// Multi-GPU run with the entity-sharded optimizer (kge_train_rows_sharded / kge_train_entity_sharded): every rank maps
// every rank's peer block (identical layout), rank g owns the entity rows [g*base + min(g, rem), ...) and is the only
// one to update them.  delta[r] = byte distance from an address inside the local block to the same address inside
// rank r's block, so a value is mirrored with one extra store per peer.  world == 0: single device, nothing mirrored.
struct Mirror {
  int world, rank;
  int ent_base, ent_rem;                   // balanced contiguous entity ranges: the first `rem` ranks own base + 1 rows
  long long delta[KGE_PEER_MAX_RANKS];
  long long mc_delta;                      // != 0: byte distance to the NVSwitch multicast mapping of the blocks -- one
                                           // multimem.st replaces the local store and the world - 1 peer stores
};

struct SplitWs;

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
  const EntityAdam *entity_adam;   // host: fuse the entity table's Adam update into the entity-major pass (or NULL)
  int *entity_adam_applied;    // host out flag: the update was applied (gE was not written; the caller skips E in Adam)
  int ring;                    // single-read path: slots per row group in the TMA ring (2..4)
  int l2_hints;                // single-read path: L2 residency hints on (KGE_L2_HINTS=1)
  unsigned long long *phase_cycles;   // debug (KGE_ROW_PHASES=1): [8] cycles of thread 0 per phase, summed over CTAs and rows
  Mirror mir;                  // entity-sharded multi-GPU step: where the row kernel's outputs are mirrored to
  const SplitWs *shard_ws;     // host: with mir.world > 1, the (peer-visible) arrays the row kernel writes
};

#if defined(__CUDACC__)
__device__ __forceinline__ int owner_of(const Mirror &m, int64_t id) {
  const int64_t cut = (int64_t)m.ent_rem * (m.ent_base + 1);
  return id < cut ? (int)(id / (m.ent_base + 1)) : m.ent_rem + (int)((id - cut) / m.ent_base);
}
template <typename T>
__device__ __forceinline__ T *at_rank(const Mirror &m, T *p, int r) {
  return reinterpret_cast<T *>(reinterpret_cast<uintptr_t>(p) + m.delta[r]);
}
// store to the local block and to the same place in every peer's block (plain stores: the cross-GPU barrier that
// follows the kernel fences them at system scope)
template <typename T>
__device__ __forceinline__ void store_all(const Mirror &m, T *p, T v) {
  static_assert(sizeof(T) == 4, "store_all moves 32-bit values");
  if (m.mc_delta) {
    uint32_t bits;
    memcpy(&bits, &v, 4);
    multimem_st_b32(reinterpret_cast<char *>(p) + m.mc_delta, bits);
    return;
  }
  *p = v;
  for (int r = 0; r < m.world; ++r)
    if (r != m.rank) *at_rank(m, p, r) = v;
}
#endif

struct SplitWs {             // carved from the caller's workspace
  float *G;                  // [rows, N]   dL/ds of every negative pair
  float *Qtab;               // [rows, De]  query vectors
  int *cnt;                  // [nentity + 1] histogram -> exclusive offsets (row kernel of the entity-sharded step: this
                             //             rank's peer-visible histogram, which the owners read over NVLink)
  int *cursor;               // [nentity]   scatter cursors
  int *queue;                // [16]        dynamic entity queues of entity_kernel (one per entity slice)
  int *tile_tot;             // [ceil(nentity / 1024)] totals of the scan tiles
  int *perm;                 // [rows * (N + 3)]  per entity: row index (b - row_begin) of each of its pairs, or
                             //             -(1 + i) for the direct gradient row i of Dvec
  float *gsorted;            // [rows * (N + 3)]  dL/ds of every pair in the same order
  float *Dvec;               // [rows * 3, De] direct gradient rows (fixed entity, positive head, positive tail), or NULL:
                             //             those rows are added to gE with atomics instead
  int *dids;                 // [rows * 3]  target entity of every direct row
  int *ids32;                // [rows, N]   the row kernel's clamped int32 copy of the candidate ids (what the counting sort
                             //             reads: the caller's int64 ids may sit in pinned host memory), or NULL
};

size_t split_workspace_bytes(int64_t rows, int64_t N, int64_t De, int64_t nentity);
SplitWs carve_split_ws(void *workspace, int64_t rows, int64_t N, int64_t De, int64_t nentity);
// does the launcher take the single-read path for this shape (same predicate in kge_train_plan and launch_rows_model)?
bool split_path_shape_ok(int64_t rows, int64_t N, int64_t De, int64_t d, bool cplx, int64_t nentity, bool fused_adam);

// per-model entry points (explicitly instantiated in kge_train_inst.cu)
template <int MODEL>
int launch_rows_model(bool head, const RowArgs &a, bool vec4, int threads, size_t smem, void *workspace,
                      size_t workspace_bytes, cudaStream_t st);
template <int MODEL>
int launch_entity_model(bool head, const RowArgs &a, const SplitWs &ws, int64_t ent_begin, int64_t ent_end, int slot,
                        cudaStream_t st, int reserve_sms);

}  // namespace kge
