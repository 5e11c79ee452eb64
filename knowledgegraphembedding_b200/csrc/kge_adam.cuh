// kge_adam.cuh -- the element update of torch.optim.Adam (torch/optim/adam.py single-tensor path, defaults of
// run.py:266-269: no amsgrad, no weight decay), shared by the dense kernel (kge_optim.cu), the entity-major train pass
// (kge_train_split.cuh) and the NVLink exchange (kge_peer.cu) so that every path rounds identically.
#pragma once
#include "kge_common.cuh"

namespace kge {

struct AdamScalars {
  float w1, b2, w2, eps, l3x3;   // 1-beta1, beta2, 1-beta2, eps, 3*l3
  float step_size, bc2_sqrt;     // -(lr / (1 - beta1^t)),  sqrt(1 - beta2^t)
  float inv_bc2_sqrt;            // 1 / sqrt(1 - beta2^t)  (adam_pair_fast)
};

// host: the bias-correction scalars exactly as torch computes them (python floats = double, then fp32 kernels)
inline AdamScalars adam_scalars(double lr, double beta1, double beta2, double eps, double l3, int64_t step) {
  AdamScalars s;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  s.w1 = (float)(1.0 - beta1); s.b2 = (float)beta2; s.w2 = (float)(1.0 - beta2); s.eps = (float)eps;
  s.l3x3 = (float)(3.0 * l3);
  s.step_size = (float)(-(lr / bc1));
  s.bc2_sqrt = (float)sqrt(bc2);
  s.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  return s;
}

#if defined(__CUDACC__)
__device__ __forceinline__ void adam_elem(float &p, float &g, float &m, float &v, const AdamScalars &a, bool l3,
                                          double &racc) {
  if (l3) {
    const float ax = fabsf(p);
    racc += (double)(ax * ax * ax);
    g = g + a.l3x3 * p * ax;                               // d/dx of l3 * sum|x|^3
  }
  m = m + (g - m) * a.w1;                                  // exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.b2;                                            // exp_avg_sq.mul_(beta2)
  v = v + a.w2 * g * g;                                    //   .addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;       // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p + a.step_size * (m / denom);                       // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// Two elements at once for the entity-major pass (kge_train_split.cuh), where the update is fused into an issue-bound
// kernel: exp_avg / exp_avg_sq are updated with the same correctly rounded fp32 operations as adam_elem (they carry
// the optimizer's long-term state), but the step itself uses the MUFU approximations
//     denom = sqrt.approx(v) * (1 / bc2_sqrt) + eps,   p += (step_size * m) * rcp.approx(denom)
// instead of an IEEE square root and two IEEE divisions (~40 instructions per element).  Both approximations are within
// 2^-22 relative, so the parameter moves by the reference's step times (1 +- 5e-7): |delta p| <= lr * 5e-7 per step, five
// orders of magnitude inside the 1e-5 tolerance on updated embeddings, and p's error does not feed back into m or v.
__device__ __forceinline__ void adam_pair_fast(f2 &p, f2 g, f2 &m, f2 &v, const AdamScalars &a, bool l3, double &racc) {
  if (l3) {
    float p0, p1, g0, g1;
    unpack2(p, p0, p1);
    unpack2(g, g0, g1);
    const float a0 = fabsf(p0), a1 = fabsf(p1);
    racc += (double)(a0 * a0 * a0) + (double)(a1 * a1 * a1);
    g = pack2(g0 + a.l3x3 * p0 * a0, g1 + a.l3x3 * p1 * a1);
  }
  m = fma2(sub2(g, m), pack2(a.w1, a.w1), m);                                   // exp_avg.lerp_(grad, 1 - beta1)
  v = fma2(mul2(g, pack2(a.w2, a.w2)), g, mul2(v, pack2(a.b2, a.b2)));           // exp_avg_sq.mul_(beta2).addcmul_(...)
  float v0, v1;
  unpack2(v, v0, v1);
  const f2 denom = fma2(pack2(sqrt_approx(v0), sqrt_approx(v1)), pack2(a.inv_bc2_sqrt, a.inv_bc2_sqrt), pack2(a.eps, a.eps));
  float d0, d1;
  unpack2(denom, d0, d1);
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  p = fma2(mul2(m, pack2(a.step_size, a.step_size)), pack2(r0, r1), p);          // param.addcdiv_(exp_avg, denom, -step_size)
}
#endif

}  // namespace kge
