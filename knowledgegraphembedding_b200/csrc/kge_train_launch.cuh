// kge_train_launch.cuh -- host-side launch logic of the train-path kernels for one model (template over MODEL; the
// explicit instantiations live in kge_train_inst.cu, one translation unit per model).
#pragma once
#include <stdlib.h>

#include "kge_train_split.cuh"

namespace kge {

template <int MODEL, bool HEAD>
static int launch_entity_pass(const RowArgs &a, const SplitWs &ws, int64_t ent_begin, int64_t ent_end, int slot,
                              cudaStream_t st, int reserve_sms = 0) {
  constexpr bool CPLX = op_is_complex(op_of(MODEL, HEAD));
  if (ent_end <= ent_begin) return KGE_OK;
  const int nunits = a.d / 4;
  EntArgs e{};
  e.E = const_cast<float *>(a.E); e.modulus = a.modulus; e.gE = a.gE; e.gM = a.gM; e.gsorted = ws.gsorted;
  e.Qtab = ws.Qtab; e.Dvec = ws.Dvec;
  e.off = ws.cnt; e.perm = ws.perm; e.queue = ws.queue + slot; e.nentity = a.nentity;
  e.ent_begin = ent_begin; e.ent_count = ent_end - ent_begin;
  e.N = a.N; e.d = a.d; e.De = a.De; e.scale = a.scale;
  e.need_gmod = (MODEL == KGE_PROTATE && !a.do_loss) ? 1 : 0;
  e.mir = a.mir;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (reserve_sms > 0 && sms > 2 * reserve_sms) sms -= reserve_sms;   // leave SMs for the concurrent NCCL kernel
  const bool fused = a.entity_adam != nullptr && ws.Dvec != nullptr;
  if (fused) {
    const EntityAdam &o = *a.entity_adam;
    e.exp_avg = o.exp_avg; e.exp_avg_sq = o.exp_avg_sq; e.adam = o.s; e.l3 = o.l3 && o.s.l3x3 != 0.f;
    e.err = a.err;
    if (e.l3) {
      KGE_REQUIRE(o.reg_partials && o.n_reg_partials >= 1, "L3 regularisation needs reg_partials");
      if (sms > o.n_reg_partials) sms = (int)o.n_reg_partials;
      KGE_CUDA_OK(cudaMemsetAsync(o.reg_partials, 0, sizeof(double) * o.n_reg_partials, st));
      e.reg_partials = o.reg_partials;
    }
  }
  bool two = nunits >= 64;                           // enough work per lane to split the row in two parts
  if (const char *pe = getenv("KGE_ENTITY_PARTS")) two = two && pe[0] != '1';   // (tuning: one warp per whole row)
  e.upp = two ? (nunits + 1) / 2 : nunits;
  // slots have the kernel's compile-time half stride: CH chunks of 32 float4 units, CH = (complex ? 8 : 16) / parts
  const size_t slotbytes = (size_t)(CPLX ? 2 : 1) * ((CPLX ? 8 : 16) / (two ? 2 : 1)) * 32 * 16;
  // warps x ring depth: as many slots in flight as the shared memory holds (KGE_ENTITY_WARPS / KGE_ENTITY_DEPTH tune)
  int We = entity_warps(two ? 2 : 1, fused), depth = 2;
  if (const char *w = getenv("KGE_ENTITY_WARPS")) { const int v = atoi(w); if (v >= 1 && v <= We) We = v; }
  while (depth < 4 && 16 + (size_t)We * (depth + 1) * (slotbytes + 16) <= 227 * 1024) ++depth;
  if (const char *dp = getenv("KGE_ENTITY_DEPTH")) { const int v = atoi(dp); if (v >= 2 && v <= depth) depth = v; }
  while (We > 1 && 16 + (size_t)We * depth * (slotbytes + 16) > 227 * 1024) --We;
  e.depth = depth;
  { const char *h = getenv("KGE_L2_HINTS"); e.l2_hints = h && h[0] == '1'; }   // measured: no gain at cfg 3, off by default
  const size_t esmem = 16 + (size_t)We * depth * (slotbytes + 16);
#define KGE_ENT_LAUNCH(S, F)                                                                         \
  do {                                                                                               \
    auto k = entity_kernel<MODEL, HEAD, S, F>;                                                       \
    KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));   \
    k<<<sms, We * 32, esmem, st>>>(e);                                                               \
  } while (0)
  if (two) { if (fused) KGE_ENT_LAUNCH(2, true); else KGE_ENT_LAUNCH(2, false); }
  else { if (fused) KGE_ENT_LAUNCH(1, true); else KGE_ENT_LAUNCH(1, false); }
#undef KGE_ENT_LAUNCH
  KGE_CUDA_OK(cudaGetLastError());
  if (fused && a.entity_adam_applied) *a.entity_adam_applied = 1;
  return KGE_OK;
}

unsigned long long *row_phase_counters();   // kge_train.cu: device counters when KGE_ROW_PHASES=1, else nullptr

// which register / shared-memory variant of row_kernel_split runs (KGE_SPLIT_VARIANT=0..4 overrides; see split_warps)
static int split_variant() {
  const char *v = getenv("KGE_SPLIT_VARIANT");
  if (v && v[0] >= '0' && v[0] <= '4' && !v[1]) return v[0] - '0';
  return 2;
}

template <int MODEL, bool HEAD>
static int launch_rows_v(const RowArgs &a, bool vec4, int threads, size_t smem, void *workspace, size_t workspace_bytes,
                         cudaStream_t st) {
  int grid = a.row_count;
  // TMA ring variants: rows are 16-byte multiples, one k-tile covers the row, and >= 4 warps get a double buffer
  constexpr bool CPLX = op_is_complex(op_of(MODEL, HEAD));
  const int nunits = a.d / 4;
  const bool want_adam = a.entity_adam != nullptr;
  if (vec4 && nunits <= 32 * (CPLX ? 8 : 16) && !getenv("KGE_NO_TMA")) {
    const size_t rowbytes = (size_t)a.De * 4;
    // entity-sharded multi-GPU step: the row kernel writes into the peer-visible arrays of a.shard_ws, the sort and the
    // entity pass belong to the owners (kge_train_entity_sharded)
    const bool sharded = a.mir.world > 1 && a.shard_ws != nullptr;
    const bool split = sharded ? (a.do_loss && a.pos_row_loss && a.loss_kind != KGE_LOSS_POSITIVE &&
                                  split_path_shape_ok(a.row_count, a.N, a.De, a.d, CPLX, a.nentity, true))
                               : (workspace && (a.gE || want_adam) && !(a.do_loss && a.loss_kind == KGE_LOSS_POSITIVE) &&
                                  workspace_bytes >= split_workspace_bytes(a.row_count, a.N, a.De, a.nentity) &&
                                  split_path_shape_ok(a.row_count, a.N, a.De, a.d, CPLX, a.nentity,
                                                      want_adam && a.pos_row_loss && !a.defer_entity));
    // ---- single-read path: row-major forward + dL/dq, counting sort, entity-major dL/dx -----------------
    if (split) {
      constexpr int Hs = CPLX ? 2 : 1;
      const int chunks = (nunits + 31) / 32;                 // 128-float chunks per half row
      const int nch = chunks <= 4 ? 4 : (chunks <= 8 ? 8 : 16);
      const int var = split_variant();
      const size_t hs = (size_t)Hs * 128 * nch;              // padded slot (floats)
      // q, dq, rot, the score arrays, scratch, and the double-buffered stage of the positive triple's rows (+ its 2 mbarriers)
      const size_t fixed_s = sizeof(float) * (hs + (size_t)((a.De + 3) & ~3) + 2 * (size_t)((a.d + 3) & ~3) +
                                              2 * (size_t)a.N + 32 + 2 * (2 * (size_t)a.De + (size_t)a.Dr)) + 16 + 16;
      // Ws = row groups per CTA (one warp each, or a pair of warps: variants 3 / 4); a group owns two slots, two
      // mbarriers and the pair's exchange words
      const int wpr = split_warps_per_row(var);
      const int wcap = split_warps(CPLX, nch, var) / wpr;
      int Ws = wcap;
      if (a.N < 4 * Ws) Ws = a.N >= 16 ? (a.N + 3) / 4 : 4;              // short candidate lists: fewer, busier warps
      if (Ws > wcap) Ws = wcap;
      // ring depth (KGE_SPLIT_RING=2..4, bounded by what fits next to q, dq and the score arrays)
      auto group_bytes = [&](int depth) { return (size_t)depth * (hs * sizeof(float) + 8) + 16 + 4; };
      // (measured on B200 at cfg 3: a third slot per group does not help -- 0.673 vs 0.650 ms per step -- the gather is
      // not short of bytes in flight; the ring stays at 2 unless KGE_SPLIT_RING asks for more)
      int ring = 2, ring_max = 2;
      while (ring_max < 4 && fixed_s + Ws * group_bytes(ring_max + 1) <= 227 * 1024) ++ring_max;
      if (const char *r = getenv("KGE_SPLIT_RING")) { const int v = atoi(r); if (v >= 2 && v <= ring_max) ring = v; }
      while (Ws > 1 && fixed_s + Ws * group_bytes(ring) > 227 * 1024) --Ws;
      // KGE_SPLIT_CTAS=2 (experiment): two CTAs of half the warps per SM, so that the block-wide phases of one row
      // (query vector, loss, fold, chain rule, positive triple: ~29 % of a row's cycles at cfg 3) run under the
      // candidate loop of the other CTA's row
      int ctas = 1;
      if (const char *c = getenv("KGE_SPLIT_CTAS")) {
        if (c[0] == '2' && wpr == 1 && Ws >= 8 && 2 * (fixed_s + (Ws / 2) * group_bytes(ring) + 1024) <= 227 * 1024) {
          ctas = 2;
          Ws /= 2;
        }
      }
      const size_t per_warp = group_bytes(ring);
      if (Ws >= (wpr == 2 ? 2 : 4) && fixed_s + Ws * per_warp <= 227 * 1024) {
        const size_t total = fixed_s + Ws * per_warp;
        RowArgs ar = a;
        ar.ring = ring;
        ar.phase_cycles = row_phase_counters();
        { const char *h = getenv("KGE_L2_HINTS"); ar.l2_hints = h && h[0] == '1'; }
        SplitWs ws = sharded ? *a.shard_ws : carve_split_ws(workspace, a.row_count, a.N, a.De, a.nentity);
        if (!sharded) {
          // the fused optimizer needs the positive triple in the same launch (its gradient rows reach the entity pass
          // through ws.Dvec); without it the entity-side rows of the positives go to gE with atomics
          if (!(want_adam && a.pos_row_loss && !a.defer_entity)) ws.Dvec = nullptr;
          KGE_REQUIRE(ws.Dvec || a.gE, "grad_entity is required when the entity optimizer is not fused");
          KGE_CUDA_OK(cudaMemsetAsync(ws.cnt, 0, ((size_t)(2 * a.nentity + 1) + 16) * 4, st));
        }
        // persistent CTAs (one per SM: the slots take the whole shared memory), rows are dealt round-robin
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int sgrid = grid < sms * ctas ? grid : sms * ctas;
#define KGE_SPLIT_LAUNCH2(NCH, VAR)                                                                    \
  do {                                                                                                 \
    auto k = row_kernel_split<MODEL, HEAD, NCH, VAR>;                                                  \
    KGE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));     \
    k<<<sgrid, Ws * wpr * 32, total, st>>>(ar, ws);                                                    \
  } while (0)
#define KGE_SPLIT_LAUNCH(NCH)                                                                          \
  do {                                                                                                 \
    if (var == 0) KGE_SPLIT_LAUNCH2(NCH, 0);                                                           \
    else if (var == 1) KGE_SPLIT_LAUNCH2(NCH, 1);                                                      \
    else if (var == 3) KGE_SPLIT_LAUNCH2(NCH, 3);                                                      \
    else if (var == 4) KGE_SPLIT_LAUNCH2(NCH, 4);                                                      \
    else KGE_SPLIT_LAUNCH2(NCH, 2);                                                                    \
  } while (0)
        if (nch == 4) KGE_SPLIT_LAUNCH(4);
        else if (nch == 8) KGE_SPLIT_LAUNCH(8);
        else {
          if constexpr (CPLX) { set_error("row too wide"); return KGE_ERR_INVALID; }      // unreachable: nunits <= 256
          else KGE_SPLIT_LAUNCH(16);
        }
#undef KGE_SPLIT_LAUNCH
#undef KGE_SPLIT_LAUNCH2
        KGE_CUDA_OK(cudaGetLastError());
        if (a.fused_positive && a.pos_row_loss) *a.fused_positive = 1;
        if (sharded) {
          if (a.entity_deferred) *a.entity_deferred = 1;
          return KGE_OK;
        }
        {
          const int tiles = (int)((a.nentity + 1023) / 1024);
          scan_tiles_kernel<<<tiles, 1024, 0, st>>>(ws.cnt, ws.cursor, ws.tile_tot, a.nentity);
          KGE_CUDA_OK(cudaGetLastError());
          scan_apply_kernel<<<tiles, 1024, 0, st>>>(ws.cnt, ws.cursor, ws.tile_tot, a.nentity);
          KGE_CUDA_OK(cudaGetLastError());
        }
        {
          const int64_t pairs = (int64_t)a.row_count * (a.N + (ws.Dvec ? 3 : 0));
          int g2 = (int)((pairs + 255) / 256);
          if (g2 > 148 * 16) g2 = 148 * 16;
          scatter_pairs_kernel<<<g2, 256, 0, st>>>(a.cand, a.cand_stride, a.row_begin, a.row_count, a.N, a.nentity,
                                                   ws.ids32, ws.G, ws.Dvec ? ws.dids : nullptr, ws.cursor, ws.perm,
                                                   ws.gsorted);
          KGE_CUDA_OK(cudaGetLastError());
        }
        if (a.defer_entity) {
          if (a.entity_deferred) *a.entity_deferred = 1;
          return KGE_OK;
        }
        return launch_entity_pass<MODEL, HEAD>(a, ws, 0, a.nentity, 0, st);
      }
    }
    KGE_REQUIRE(!sharded, "the entity-sharded step needs the single-read path (check kge_train_plan first)");
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
  KGE_REQUIRE(a.mir.world <= 1, "the entity-sharded step needs the single-read path (check kge_train_plan first)");
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

template <int MODEL>
int launch_rows_model(bool head, const RowArgs &a, bool vec4, int threads, size_t smem, void *workspace,
                      size_t workspace_bytes, cudaStream_t st) {
  return head ? launch_rows_v<MODEL, true>(a, vec4, threads, smem, workspace, workspace_bytes, st)
              : launch_rows_v<MODEL, false>(a, vec4, threads, smem, workspace, workspace_bytes, st);
}

template <int MODEL>
int launch_entity_model(bool head, const RowArgs &a, const SplitWs &ws, int64_t ent_begin, int64_t ent_end, int slot,
                        cudaStream_t st, int reserve_sms) {
  return head ? launch_entity_pass<MODEL, true>(a, ws, ent_begin, ent_end, slot, st, reserve_sms)
              : launch_entity_pass<MODEL, false>(a, ws, ent_begin, ent_end, slot, st, reserve_sms);
}

}  // namespace kge
