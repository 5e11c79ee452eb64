// kge_adam.cuh -- the element update of torch.optim.Adam (torch/optim/adam.py single-tensor path, defaults of
// run.py:266-269: no amsgrad, no weight decay), shared by the dense kernel (kge_optim.cu), the entity-major train pass
// (kge_train_split.cuh) and the NVLink exchange (kge_peer.cu) so that every path rounds identically.
#pragma once

namespace kge {

struct AdamScalars {
  float w1, b2, w2, eps, l3x3;   // 1-beta1, beta2, 1-beta2, eps, 3*l3
  float step_size, bc2_sqrt;     // -(lr / (1 - beta1^t)),  sqrt(1 - beta2^t)
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
#endif

}  // namespace kge
