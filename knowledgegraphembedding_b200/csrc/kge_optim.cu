// kge_optim.cu -- dense fused Adam (torch.optim.Adam semantics), L3 regulariser, loss finalisation.
#include <stdarg.h>
#include <string.h>

#include "kge_common.cuh"
#include "kge_adam.cuh"

namespace kge {

// ---- error plumbing -----------------------------------------------------------------------------------
static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_model(const kge_model_t *m) {
  KGE_REQUIRE(m != nullptr, "null model");
  KGE_REQUIRE(m->model >= KGE_TRANSE && m->model <= KGE_PROTATE, "model %d not supported", m->model);   // model.py:63
  KGE_REQUIRE(m->entity && m->relation, "model tables are null");
  KGE_REQUIRE(m->nentity > 0 && m->nrelation > 0 && m->hidden_dim > 0, "empty model");
  const int64_t d = m->hidden_dim;
  switch (m->model) {
    case KGE_ROTATE:       // model.py:66-67
      KGE_REQUIRE(m->entity_dim == 2 * d && m->relation_dim == d, "RotatE should use --double_entity_embedding");
      break;
    case KGE_COMPLEX:      // model.py:69-70
      KGE_REQUIRE(m->entity_dim == 2 * d && m->relation_dim == 2 * d,
                  "ComplEx should use --double_entity_embedding and --double_relation_embedding");
      break;
    default:               // broadcasting [B,1,D_e] with [B,1,D_r] in model.py:168,177,241 needs equal dims
      KGE_REQUIRE(m->entity_dim == m->relation_dim, "entity_dim %lld and relation_dim %lld must match for this model",
                  (long long)m->entity_dim, (long long)m->relation_dim);
  }
  KGE_REQUIRE(m->model != KGE_PROTATE || m->modulus, "pRotatE needs the modulus parameter");
  KGE_REQUIRE(m->entity_dim < (1 << 24), "entity_dim too large");
  return KGE_OK;
}

int DeviceGuard::enter(int device) {
  KGE_CUDA_OK(cudaGetDevice(&prev));
  if (prev != device) {
    KGE_CUDA_OK(cudaSetDevice(device));
    changed = true;
  }
  return KGE_OK;
}

// ---- sum of subsampling weights (model.py:285) --------------------------------------------------------------
__global__ void weight_sum_kernel(const float *__restrict__ w, int64_t B, float *__restrict__ out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) acc += w[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    out[0] = t;
  }
}

// ---- loss finalisation (model.py:281-288, 296) ----------------------------------------------------------------
__global__ void loss_finalize_kernel(const float *__restrict__ pos_row, const float *__restrict__ neg_row,
                                     const float *__restrict__ w, const float *__restrict__ wsum, int64_t B,
                                     float reg, const double *__restrict__ reg_partials, int64_t nparts,
                                     float *__restrict__ out) {
  __shared__ double sh[3][32];
  double p = 0, n = 0, r = 0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const float wi = w ? w[i] : 1.f;
    p += (double)(wi * pos_row[i]);
    n += (double)(wi * neg_row[i]);
  }
  if (reg_partials)
    for (int64_t i = threadIdx.x; i < nparts; i += blockDim.x) r += reg_partials[i];
  for (int o = 16; o > 0; o >>= 1) {
    p += __shfl_xor_sync(0xffffffffu, p, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = p; sh[1][threadIdx.x >> 5] = n; sh[2][threadIdx.x >> 5] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double P = 0, N = 0, Rg = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { P += sh[0][i]; N += sh[1][i]; Rg += sh[2][i]; }
    const double denom = w ? (double)wsum[0] : (double)B;
    const float pl = (float)(-P / denom), nl = (float)(-N / denom);
    const float regv = (float)((double)reg * Rg);
    out[0] = pl;
    out[1] = nl;
    out[2] = (pl + nl) / 2.f + regv;
    out[3] = regv;
  }
}

// ---- Adam --------------------------------------------------------------------------------------------------
struct AdamTensor {
  float *p, *g, *m, *v;
  int64_t n;
  AdamScalars s;
  int l3;
};
struct AdamArgs {
  AdamTensor t[4];
  int nt;
  double *reg_partials;
  const int32_t *skip_flag;      // device flag (or NULL): a non-zero value (bad index in this step) cancels the update
};

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
  if (a.skip_flag && *a.skip_flag) return;                 // model.py:86-146 would have raised before optimizer.step()
  double racc = 0.0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int ti = 0; ti < a.nt; ++ti) {
    const AdamTensor t = a.t[ti];
    const bool l3 = t.l3 != 0 && t.s.l3x3 != 0.f;
    const int64_t n4 = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0) ? t.n / 4 : 0;
    float4 *p4 = reinterpret_cast<float4 *>(t.p), *g4 = reinterpret_cast<float4 *>(t.g);
    float4 *m4 = reinterpret_cast<float4 *>(t.m), *v4 = reinterpret_cast<float4 *>(t.v);
    for (int64_t i = tid; i < n4; i += nth) {
      float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
      adam_elem(p.x, g.x, m.x, v.x, t.s, l3, racc);
      adam_elem(p.y, g.y, m.y, v.y, t.s, l3, racc);
      adam_elem(p.z, g.z, m.z, v.z, t.s, l3, racc);
      adam_elem(p.w, g.w, m.w, v.w, t.s, l3, racc);
      p4[i] = p; m4[i] = m; v4[i] = v;
      if (l3) g4[i] = g;
    }
    for (int64_t i = n4 * 4 + tid; i < t.n; i += nth) {
      float p = t.p[i], g = t.g[i], m = t.m[i], v = t.v[i];
      adam_elem(p, g, m, v, t.s, l3, racc);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      if (l3) t.g[i] = g;
    }
  }
  if (a.reg_partials) {
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = racc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0;
      for (int i = 0; i < 8; ++i) s += sh[i];
      a.reg_partials[blockIdx.x] = s;
    }
  }
}

// sum |x|^3 over up to 4 tensors -> per-block double partials (the value of the L3 regulariser, model.py:292-295, when
// the update itself runs elsewhere: the NVLink peer exchange applies the L3 gradient per owned slice)
__global__ void __launch_bounds__(256) l3_partials_kernel(const AdamArgs a) {
  double racc = 0.0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int ti = 0; ti < a.nt; ++ti) {
    const AdamTensor t = a.t[ti];
    if (!t.l3) continue;
    for (int64_t i = tid; i < t.n; i += nth) {
      const float ax = fabsf(t.p[i]);
      racc += (double)(ax * ax * ax);
    }
  }
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = racc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < 8; ++i) s += sh[i];
    a.reg_partials[blockIdx.x] = s;
  }
}

}  // namespace kge

using namespace kge;

extern "C" int kge_abi_version(void) { return KGE_ABI_VERSION; }
extern "C" const char *kge_last_error(void) { return g_error; }

extern "C" int kge_device_check(int device, int *sm_count, int64_t *l2_bytes) {
  cudaDeviceProp prop;
  KGE_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (l2_bytes) *l2_bytes = prop.l2CacheSize;
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libkge_b200 is built for sm_100a only (no fallback)", device, prop.major, prop.minor);
    return KGE_ERR_DEVICE;
  }
  return KGE_OK;
}

extern "C" int kge_zero(void *ptr, int64_t bytes, void *stream) {
  KGE_REQUIRE(ptr && bytes >= 0, "bad arguments");
  KGE_CUDA_OK(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
  return KGE_OK;
}

// Batch staging for train_step (model.py:263-266 `.cuda()` copies, :305-310 `.item()` read-backs) without a framework
// dispatch per tensor: plain cudaMemcpyAsync on the caller's stream.  Pinned host memory makes the H2D copy
// asynchronous (the caller keeps the source alive until the stream has passed it); pageable memory is staged by the
// driver before the call returns.
extern "C" int kge_copy_h2d(void *dst_device, const void *host_src, int64_t bytes, void *stream) {
  KGE_REQUIRE(dst_device && host_src && bytes >= 0, "bad arguments");
  if (bytes) KGE_CUDA_OK(cudaMemcpyAsync(dst_device, host_src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return KGE_OK;
}

extern "C" int kge_copy_d2h_sync(void *host_dst, const void *src_device, int64_t bytes, void *stream) {
  KGE_REQUIRE(host_dst && src_device && bytes >= 0, "bad arguments");
  if (bytes) KGE_CUDA_OK(cudaMemcpyAsync(host_dst, src_device, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  KGE_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return KGE_OK;
}

extern "C" int kge_weight_sum(const float *weight, int64_t B, float *out, void *stream) {
  KGE_REQUIRE(weight && out && B > 0, "bad arguments");
  weight_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(weight, B, out);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

extern "C" int kge_loss_finalize(const float *pos_row, const float *neg_row, const float *weight,
                                 const float *weight_sum, int64_t B, float regularization,
                                 const double *reg_partials, int64_t n_reg_partials, float *out, void *stream) {
  KGE_REQUIRE(pos_row && neg_row && out && B > 0, "bad arguments");
  KGE_REQUIRE(!weight || weight_sum, "subsampling weights need their sum");
  loss_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pos_row, neg_row, weight, weight_sum, B, regularization,
                                                           reg_partials, n_reg_partials, out);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

extern "C" int kge_adam_step(const kge_adam_tensor_t *ts, int nt, double lr, double beta1, double beta2, double eps,
                             double l3, double *reg_partials, int64_t n_reg_partials, const int32_t *skip_flag,
                             void *stream) {
  KGE_REQUIRE(ts && nt >= 1 && nt <= 4, "kge_adam_step takes 1..4 tensors");
  AdamArgs a{};
  a.nt = nt;
  a.skip_flag = skip_flag;
  int64_t total = 0;
  for (int i = 0; i < nt; ++i) {
    KGE_REQUIRE(ts[i].param && ts[i].grad && ts[i].exp_avg && ts[i].exp_avg_sq && ts[i].numel > 0 && ts[i].step >= 1,
                "bad Adam tensor %d", i);
    a.t[i] = AdamTensor{ts[i].param, ts[i].grad, ts[i].exp_avg, ts[i].exp_avg_sq, ts[i].numel,
                        adam_scalars(lr, beta1, beta2, eps, l3, ts[i].step), ts[i].l3};
    total += ts[i].numel;
  }
  int grid = (int)((total / 4 + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  if (l3 != 0.0) {
    KGE_REQUIRE(reg_partials && n_reg_partials >= 1, "L3 regularisation needs reg_partials");
    if (grid > n_reg_partials) grid = (int)n_reg_partials;
    KGE_CUDA_OK(cudaMemsetAsync(reg_partials, 0, sizeof(double) * n_reg_partials, (cudaStream_t)stream));
    a.reg_partials = reg_partials;
  }
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}

extern "C" int kge_l3_partials(const kge_adam_tensor_t *ts, int nt, double *reg_partials, int64_t n_reg_partials,
                               void *stream) {
  KGE_REQUIRE(ts && nt >= 1 && nt <= 4 && reg_partials && n_reg_partials >= 1, "bad arguments");
  AdamArgs a{};
  a.nt = nt;
  int64_t total = 0;
  for (int i = 0; i < nt; ++i) {
    KGE_REQUIRE(ts[i].param && ts[i].numel > 0, "bad tensor %d", i);
    a.t[i] = AdamTensor{ts[i].param, nullptr, nullptr, nullptr, ts[i].numel, AdamScalars{}, ts[i].l3};
    total += ts[i].numel;
  }
  int grid = (int)((total + 1023) / 1024);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid > n_reg_partials) grid = (int)n_reg_partials;
  if (grid < 1) grid = 1;
  KGE_CUDA_OK(cudaMemsetAsync(reg_partials, 0, sizeof(double) * n_reg_partials, (cudaStream_t)stream));
  a.reg_partials = reg_partials;
  l3_partials_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  KGE_CUDA_OK(cudaGetLastError());
  return KGE_OK;
}
