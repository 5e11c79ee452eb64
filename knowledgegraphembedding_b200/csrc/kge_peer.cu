// kge_peer.cu -- multi-GPU training exchange over NVLink peer memory: gradient reduce-scatter, Adam on the owned slice
// and parameter broadcast in ONE kernel (no NCCL on the data path).
//
// The reference trains on one device (run.py:241-242); batch-sharded data parallelism is this repository's extension
// (DESIGN.md section 6).  Every rank holds E, R and a gradient workspace [dE | dR | dM | row losses]; after the local
// train kernels the workspaces have to be summed and the dense Adam step (model.py:303) applied to every replica.
// Instead of all-reduce + replicated Adam, rank g owns the g-th contiguous slice of the parameter region:
//     for every 16-byte group of its slice:  load the group from all G workspaces (G-1 NVLink reads, fixed rank order),
//     sum, run torch.optim.Adam's update on the local p / exp_avg / exp_avg_sq, store p, and push the new parameter
//     values into the other ranks' workspaces (G-1 NVLink writes).
// A second, small kernel waits until every peer has finished pushing and copies the received slices into the local
// tables.  Wire traffic per rank is (G-1)/G of the region in each direction, overlapped (reads and writes use opposite
// NVLink directions); Adam's HBM traffic drops from 7 table passes to 7/G + 2(G-1)/G, and all replicas end bit-identical.
// Moments are valid only on the owning rank (the host gathers them when optimizer.state_dict() is taken).
//
// Cross-GPU ordering uses two flag channels per rank in peer-visible memory: channel 0 "my gradients are final",
// channel 1 "my pushes are done".  Flags carry the step epoch; waits are bounded (globaltimer) and report through
// err_flag instead of hanging the GPU.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "kge_common.cuh"

namespace kge {

constexpr int PEER_MAX = KGE_PEER_MAX_RANKS;
constexpr unsigned long long kPeerTimeoutNsDefault = 60ull * 1000ull * 1000ull * 1000ull;   // KGE_PEER_TIMEOUT_S

struct PeerTensor {
  float *p, *m, *v;
  int64_t off, n;                // position of the tensor's gradient inside the workspace (floats), element count
  float step_size, bc2_sqrt;
  int l3;                        // this tensor takes part in the L3 regulariser (model.py:290-297)
};
struct PeerArgs {
  float *grad[PEER_MAX];         // workspace base of every rank (peer-mapped); grad[rank] is local
  float *mc;                     // multicast mapping of the workspaces (NVSwitch multimem), or null
  int64_t rlo4, rhi4;            // region of the workspace this call exchanges (float4 units); [lo4, hi4) lies inside
  uint32_t *flags[PEER_MAX];     // flag block of every rank: [2][PEER_MAX] uint32
  int world, rank;
  uint32_t epoch;
  PeerTensor t[3];
  int nt;
  int64_t lo4, hi4;              // slice of the region this rank owns, in float4 units
  int64_t row_off, row_n;        // loss rows inside the workspace (floats); summed over ranks into rows_out
  float *rows_out;
  float w1, b2, w2, eps, l3x3;   // l3x3 = 3 * regularization (0: no L3 term)
  int32_t *err;
  unsigned long long timeout_ns; // bound on a barrier wait
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float *p) {          // system-scope load: never served from a stale L1
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_peer4(float *p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// NVSwitch in-fabric reduction / broadcast (NVLS): one load returns the sum of the G replicas of an address, one store
// writes all G replicas.  Halves the NVLink bytes of the exchange compared with G-1 unicast reads + G-1 unicast writes.
__device__ __forceinline__ float4 multimem_sum4(const float *mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ float multimem_sum1(const float *mc) {
  float r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(r) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_store4(float *mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// CTA 0 publishes `epoch` on `channel` to every rank; every CTA then waits until all ranks have published it here.
__device__ __forceinline__ void peer_barrier(const PeerArgs &a, int channel) {
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + channel * PEER_MAX + a.rank, a.epoch);
  }
  if (threadIdx.x < a.world) {
    const uint32_t *f = a.flags[a.rank] + channel * PEER_MAX + threadIdx.x;
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(f) - a.epoch) < 0) {
      if (global_ns() - t0 > a.timeout_ns) {             // a peer died or fell out of step: report, do not hang
        if (a.err) atomicExch(a.err, 2);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void peer_adam(float &p, float g, float &m, float &v, const PeerArgs &a, const PeerTensor &t) {
  if (t.l3) g = g + a.l3x3 * p * fabsf(p);                 // dense L3 gradient on the (replicated) parameter, added once
  m = m + (g - m) * a.w1;                                  // same operation order as adam_elem (kge_adam.cuh)
  v = v * a.b2;
  v = v + a.w2 * g * g;
  const float denom = sqrtf(v) / t.bc2_sqrt + a.eps;
  p = p + t.step_size * (m / denom);
}

__device__ __forceinline__ int tensor_of(const PeerArgs &a, int64_t i) {      // tensor holding workspace float i, or -1
  for (int k = 0; k < a.nt; ++k)
    if (i >= a.t[k].off && i < a.t[k].off + a.t[k].n) return k;
  return -1;
}

// W = compile-time bound on the number of ranks (2, 4, 8 or PEER_MAX): W float4 loads in flight per thread
// MC: the reduction and the broadcast go through the multicast mapping (a.mc) instead of W unicast accesses
// Adam on one 16-byte group of the owned slice; returns the new parameter values (zeros outside any tensor)
__device__ __forceinline__ float4 peer_update_group(const PeerArgs &a, int k, int64_t i, float4 g) {
  const PeerTensor &t = a.t[k];
  const int64_t j = i - t.off;
  if (j + 4 <= t.n && ((((uintptr_t)t.p | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0)) {
    float4 p = *reinterpret_cast<float4 *>(t.p + j), m = *reinterpret_cast<float4 *>(t.m + j),
           v = *reinterpret_cast<float4 *>(t.v + j);
    peer_adam(p.x, g.x, m.x, v.x, a, t);
    peer_adam(p.y, g.y, m.y, v.y, a, t);
    peer_adam(p.z, g.z, m.z, v.z, a, t);
    peer_adam(p.w, g.w, m.w, v.w, a, t);
    *reinterpret_cast<float4 *>(t.p + j) = p;
    *reinterpret_cast<float4 *>(t.m + j) = m;
    *reinterpret_cast<float4 *>(t.v + j) = v;
    return p;
  }
  float gs[4] = {g.x, g.y, g.z, g.w}, os[4] = {0.f, 0.f, 0.f, 0.f};      // tensor tail or unaligned storage
  for (int e = 0; e < 4; ++e)
    if (j + e < t.n) {
      float p = t.p[j + e], m = t.m[j + e], v = t.v[j + e];
      peer_adam(p, gs[e], m, v, a, t);
      t.p[j + e] = p; t.m[j + e] = m; t.v[j + e] = v;
      os[e] = p;
    }
  return make_float4(os[0], os[1], os[2], os[3]);
}

// W = compile-time bound on the number of ranks (2, 4, 8 or PEER_MAX): W unicast float4 loads in flight per thread.
// MCR / MCS: the reduction / the broadcast go through the multicast mapping (a.mc); with MCR 4 groups are in flight
// per thread.  CTAs are small (128 threads) so that one fits next to the persistent entity kernel of the next slice.
constexpr int PEER_THREADS = 128;
template <int W, bool MCR, bool MCS>
__global__ void __launch_bounds__(PEER_THREADS) peer_reduce_adam_kernel(const PeerArgs a) {
  peer_barrier(a, 0);                                      // every rank's gradients are final
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  if constexpr (MCR) {
    constexpr int U = 4;
    for (int64_t i4 = a.lo4 + tid; i4 < a.hi4; i4 += U * nth) {
      float4 g[U];
      int k[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = (i4 + u * nth) * 4;
        k[u] = (i4 + u * nth < a.hi4) ? tensor_of(a, i) : -1;
        if (k[u] >= 0) g[u] = multimem_sum4(a.mc + i);     // one load = the sum over all ranks, reduced in the switch
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = (i4 + u * nth) * 4;
        if (k[u] < 0) continue;
        const float4 out = peer_update_group(a, k[u], i, g[u]);
        if constexpr (MCS) {
          multimem_store4(a.mc + i, out);                  // every replica's slot (ours too) takes the new parameters
        } else {
          for (int r = 0; r < a.world; ++r)
            if (r != a.rank) st_peer4(a.grad[r] + i, out);
        }
      }
    }
  } else {
    for (int64_t i4 = a.lo4 + tid; i4 < a.hi4; i4 += nth) {
      const int64_t i = i4 * 4;
      float4 part[W];
#pragma unroll
      for (int r = 0; r < W; ++r)
        if (r < a.world) part[r] = ld_peer4(a.grad[r] + i);
      float4 g = part[0];
#pragma unroll
      for (int r = 1; r < W; ++r)
        if (r < a.world) { g.x += part[r].x; g.y += part[r].y; g.z += part[r].z; g.w += part[r].w; }
      const int k = tensor_of(a, i);
      if (k < 0) continue;
      const float4 out = peer_update_group(a, k, i, g);
      if constexpr (MCS) {
        multimem_store4(a.mc + i, out);
      } else {
#pragma unroll
        for (int r = 0; r < W; ++r)
          if (r < a.world && r != a.rank) st_peer4(a.grad[r] + i, out);    // the new parameters travel in the slot
      }
    }
  }
  // per-row losses: each row is non-zero on exactly one rank; every rank sums all of them for its own log line
  for (int64_t i = tid; i < a.row_n; i += nth) {
    float s = 0.f;
    if constexpr (MCR) {
      s = multimem_sum1(a.mc + a.row_off + i);
    } else {
      for (int r = 0; r < a.world; ++r) {
        float x;
        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x) : "l"(a.grad[r] + a.row_off + i) : "memory");
        s += x;
      }
    }
    a.rows_out[i] = s;
  }
}

__global__ void __launch_bounds__(PEER_THREADS) peer_finish_kernel(const PeerArgs a) {
  peer_barrier(a, 1);                                      // every rank has pushed its slice (and read ours)
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t own = a.hi4 - a.lo4;
  const float *local = a.grad[a.rank];
  for (int64_t c = tid; c < (a.rhi4 - a.rlo4) - own; c += nth) {    // every float4 group of the region we do not own
    const int64_t i4 = a.rlo4 + c < a.lo4 ? a.rlo4 + c : a.rlo4 + c + own;
    const int64_t i = i4 * 4;
    const int k = tensor_of(a, i);
    if (k < 0) continue;
    const PeerTensor &t = a.t[k];
    const int64_t j = i - t.off;
    const float4 val = ld_peer4(local + i);
    if (j + 4 <= t.n && (((uintptr_t)t.p & 15) == 0)) {
      *reinterpret_cast<float4 *>(t.p + j) = val;
    } else {
      const float vs[4] = {val.x, val.y, val.z, val.w};
      for (int e = 0; e < 4; ++e)
        if (j + e < t.n) t.p[j + e] = vs[e];
    }
  }
}

// Stand-alone cross-GPU barrier (entity-sharded step): one warp.  Thread r publishes `epoch` (and, with err exchange,
// this rank's error flag) in rank r's flag block, then waits for rank r's epoch here; a peer's error becomes ours, so
// every rank cancels the same update and raises.  Channels 0/1 belong to kge_peer_reduce_adam; the error words of
// channel c live in channel c + 4.
struct BarrierArgs {
  uint32_t *flags[PEER_MAX];
  int world, rank, channel, exchange_err;
  int phase;                       // 0: arrive and wait, 1: arrive only, 2: wait only (for an arrival published earlier)
  uint32_t epoch;
  int32_t *err;
  unsigned long long timeout_ns;
};
__global__ void __launch_bounds__(32) peer_barrier_kernel(const BarrierArgs a) {
  const int t = threadIdx.x;
  if (t >= a.world) return;
  if (a.phase != 2) {
    if (a.exchange_err) {
      const uint32_t mine = a.err ? (uint32_t)*reinterpret_cast<volatile int32_t *>(a.err) : 0u;
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.flags[t] + (a.channel + 4) * PEER_MAX + a.rank), "r"(mine)
                   : "memory");
    }
    __threadfence_system();
    st_release_sys(a.flags[t] + a.channel * PEER_MAX + a.rank, a.epoch);
  }
  if (a.phase == 1) return;
  const uint32_t *f = a.flags[a.rank] + a.channel * PEER_MAX + t;
  const unsigned long long t0 = global_ns();
  bool ok = true;
  while ((int32_t)(ld_acquire_sys(f) - a.epoch) < 0) {
    if (global_ns() - t0 > a.timeout_ns) {
      if (a.err) atomicExch(a.err, 2);
      ok = false;
      break;
    }
    __nanosleep(100);
  }
  if (ok && a.exchange_err && a.err) {
    uint32_t theirs;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(theirs) : "l"(a.flags[a.rank] + (a.channel + 4) * PEER_MAX + t)
                 : "memory");
    if (theirs) atomicCAS(a.err, 0, (int32_t)theirs);
  }
}

static unsigned long long peer_timeout_ns() {
  const char *t = getenv("KGE_PEER_TIMEOUT_S");            // a rank that is later than this aborts the step (err_flag 2)
  const double secs = t ? atof(t) : 0.0;
  return secs > 0.0 ? (unsigned long long)(secs * 1e9) : kPeerTimeoutNsDefault;
}

}  // namespace kge

using namespace kge;

extern "C" int kge_peer_barrier(const kge_peer_group_t *grp, int channel, uint32_t epoch, int exchange_err, int phase,
                                int32_t *err_flag, void *stream) {
  KGE_REQUIRE(grp && grp->world >= 2 && grp->world <= PEER_MAX && grp->rank >= 0 && grp->rank < grp->world,
              "bad peer group");
  KGE_REQUIRE(channel >= 2 && channel <= 3, "barrier channels 2 and 3 are free (0 and 1 belong to kge_peer_reduce_adam)");
  BarrierArgs a{};
  KGE_REQUIRE(phase >= 0 && phase <= 2, "phase: 0 = arrive + wait, 1 = arrive, 2 = wait");
  a.world = grp->world; a.rank = grp->rank; a.channel = channel; a.exchange_err = exchange_err; a.epoch = epoch;
  a.phase = phase;
  a.err = err_flag; a.timeout_ns = peer_timeout_ns();
  for (int r = 0; r < grp->world; ++r) {
    KGE_REQUIRE(grp->flags[r], "peer %d is not mapped", r);
    a.flags[r] = (uint32_t *)grp->flags[r];
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

extern "C" int kge_peer_alloc(int device, int64_t bytes, void **ptr) {
  KGE_REQUIRE(ptr && bytes > 0, "bad arguments");
  KGE_CUDA_OK(cudaSetDevice(device));
  KGE_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));             // a whole cudaMalloc block: exportable with cudaIpc
  KGE_CUDA_OK(cudaMemset(*ptr, 0, (size_t)bytes));
  return KGE_OK;
}

extern "C" int kge_peer_free(void *ptr) {
  if (ptr) KGE_CUDA_OK(cudaFree(ptr));
  return KGE_OK;
}

extern "C" int kge_peer_export(void *ptr, void *host_handle) {
  KGE_REQUIRE(ptr && host_handle, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == KGE_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  KGE_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(host_handle, &h, sizeof(h));
  return KGE_OK;
}

extern "C" int kge_peer_open(int device, const void *host_handle, void **peer_ptr) {
  KGE_REQUIRE(host_handle && peer_ptr, "bad arguments");
  KGE_CUDA_OK(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, host_handle, sizeof(h));
  KGE_CUDA_OK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KGE_OK;
}

extern "C" int kge_peer_close(void *peer_ptr) {
  if (peer_ptr) KGE_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
  return KGE_OK;
}

extern "C" int kge_peer_reduce_adam(const kge_peer_group_t *grp, uint32_t epoch, const kge_adam_tensor_t *ts,
                                    int nt, int64_t param_floats, int64_t region_begin4, int64_t region_end4,
                                    int64_t slice_begin4, int64_t slice_end4, int64_t row_offset, int64_t row_floats,
                                    float *rows_out, double lr, double beta1, double beta2, double eps,
                                    double l3_coefficient, int32_t *err_flag, void *stream) {
  KGE_REQUIRE(grp && ts && nt >= 1 && nt <= 3, "kge_peer_reduce_adam takes 1..3 tensors");
  KGE_REQUIRE(grp->world >= 2 && grp->world <= PEER_MAX && grp->rank >= 0 && grp->rank < grp->world,
              "peer group of %d ranks not supported (2..%d)", grp->world, PEER_MAX);
  KGE_REQUIRE(param_floats > 0 && param_floats % 4 == 0, "parameter region must be a multiple of 4 floats");
  KGE_REQUIRE(region_begin4 >= 0 && region_begin4 <= slice_begin4 && slice_begin4 <= slice_end4 &&
                  slice_end4 <= region_end4 && region_end4 * 4 <= param_floats,
              "bad region / slice");
  KGE_REQUIRE(row_floats == 0 || (rows_out && row_offset >= param_floats), "bad loss-row region");
  PeerArgs a{};
  a.world = grp->world; a.rank = grp->rank; a.epoch = epoch;
  for (int r = 0; r < grp->world; ++r) {
    KGE_REQUIRE(grp->grad[r] && grp->flags[r], "peer %d is not mapped", r);
    KGE_REQUIRE(((uintptr_t)grp->grad[r] & 15) == 0, "peer workspace %d is not 16-byte aligned", r);
    a.grad[r] = (float *)grp->grad[r];
    a.flags[r] = (uint32_t *)grp->flags[r];
  }
  const float *base = a.grad[a.rank];
  a.nt = nt;
  for (int i = 0; i < nt; ++i) {
    KGE_REQUIRE(ts[i].param && ts[i].grad && ts[i].exp_avg && ts[i].exp_avg_sq && ts[i].numel > 0 && ts[i].step >= 1,
                "bad Adam tensor %d", i);
    const int64_t off = ts[i].grad - base;
    KGE_REQUIRE(off >= 0 && off % 4 == 0 && off + ts[i].numel <= param_floats,
                "gradient of tensor %d is not a 16-byte aligned view of the local workspace", i);
    const double bc1 = 1.0 - pow(beta1, (double)ts[i].step);
    const double bc2 = 1.0 - pow(beta2, (double)ts[i].step);
    a.t[i] = PeerTensor{ts[i].param, ts[i].exp_avg, ts[i].exp_avg_sq, off, ts[i].numel, (float)(-(lr / bc1)),
                        (float)sqrt(bc2), (ts[i].l3 && l3_coefficient != 0.0) ? 1 : 0};
  }
  a.rlo4 = region_begin4; a.rhi4 = region_end4; a.lo4 = slice_begin4; a.hi4 = slice_end4;
  a.row_off = row_offset; a.row_n = row_floats; a.rows_out = rows_out;
  a.w1 = (float)(1.0 - beta1); a.b2 = (float)beta2; a.w2 = (float)(1.0 - beta2); a.eps = (float)eps;
  a.l3x3 = (float)(3.0 * l3_coefficient);
  a.err = err_flag;
  a.timeout_ns = peer_timeout_ns();
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = 148 * 12;
  // multicast use: bit 0 = reduce through the switch (multimem.ld_reduce), bit 1 = broadcast through it (multimem.st)
  int mc_mode = 0;
  if (grp->multicast) {
    KGE_REQUIRE(((uintptr_t)grp->multicast & 15) == 0, "multicast mapping is not 16-byte aligned");
    a.mc = (float *)grp->multicast;
    const char *e = getenv("KGE_PEER_MC_MODE");
    mc_mode = e ? (atoi(e) & 3) : 3;
  }
#define KGE_PEER_LAUNCH(W)                                                                               \
  do {                                                                                                   \
    if (mc_mode == 2) peer_reduce_adam_kernel<W, false, true><<<grid, PEER_THREADS, 0, st>>>(a);        \
    else peer_reduce_adam_kernel<W, false, false><<<grid, PEER_THREADS, 0, st>>>(a);                    \
  } while (0)
  if (mc_mode == 3) peer_reduce_adam_kernel<1, true, true><<<grid, PEER_THREADS, 0, st>>>(a);
  else if (mc_mode == 1) peer_reduce_adam_kernel<1, true, false><<<grid, PEER_THREADS, 0, st>>>(a);
  else if (a.world <= 2) KGE_PEER_LAUNCH(2);
  else if (a.world <= 4) KGE_PEER_LAUNCH(4);
  else if (a.world <= 8) KGE_PEER_LAUNCH(8);
  else KGE_PEER_LAUNCH(PEER_MAX);
#undef KGE_PEER_LAUNCH
  KGE_CUDA_OK(cudaGetLastError());
  peer_finish_kernel<<<grid, PEER_THREADS, 0, st>>>(a);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}
