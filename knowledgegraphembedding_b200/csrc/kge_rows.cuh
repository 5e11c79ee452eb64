// kge_rows.cuh -- element-level building blocks shared by the train-row kernel and the eval kernels.
//
// build_q:       fold the fixed side of a triple into the query vector q   (model.py:166-249)
// op_forward:    per-element contribution of one candidate row against q
// op_backward:   its derivative
// chain_q:       derivative of q w.r.t. the fixed entity row and the relation row
#pragma once
#include "kge_common.cuh"

namespace kge {

// ---- query vector ---------------------------------------------------------------------------------
// F = fixed entity row (head for tail-batch/single, tail for head-batch), Rr = relation row.
// Written with un-contractable primitives so that the train path, the eval path and the CPU oracle
// agree bit-for-bit on q.
// qd = position of the imaginary half inside q (d for a compact vector, a padded stride in the train kernel's smem)
template <int MODEL, bool HEAD>
__device__ __forceinline__ void build_q(const float *__restrict__ F, const float *__restrict__ Rr, int k, int d,
                                        float scale, float *__restrict__ q, int qd) {
  if constexpr (MODEL == KGE_TRANSE) {
    float f = F[k], r = Rr[k];
    q[k] = HEAD ? fsub(r, f) : fadd(f, r);                 // model.py:168 (r - t) / :170 (h + r)
  } else if constexpr (MODEL == KGE_DISTMULT) {
    float f = F[k], r = Rr[k];
    q[k] = HEAD ? fmul(r, f) : fmul(f, r);                 // model.py:177 (r * t) / :179 (h * r)
  } else if constexpr (MODEL == KGE_COMPLEX) {
    float fr = F[k], fi = F[d + k], rr = Rr[k], ri = Rr[d + k];
    if (HEAD) {                                            // model.py:190-191
      q[k] = fadd(fmul(rr, fr), fmul(ri, fi));
      q[qd + k] = fsub(fmul(rr, fi), fmul(ri, fr));
    } else {                                               // model.py:194-195
      q[k] = fsub(fmul(fr, rr), fmul(fi, ri));
      q[qd + k] = fadd(fmul(fr, ri), fmul(fi, rr));
    }
  } else if constexpr (MODEL == KGE_ROTATE) {
    float fr = F[k], fi = F[d + k];
    float s, c;
    sincos_rep(fdiv(Rr[k], scale), &s, &c);                // model.py:209-212
    if (HEAD) {                                            // model.py:215-216
      q[k] = fadd(fmul(c, fr), fmul(s, fi));
      q[qd + k] = fsub(fmul(c, fi), fmul(s, fr));
    } else {                                               // model.py:220-221
      q[k] = fsub(fmul(fr, c), fmul(fi, s));
      q[qd + k] = fadd(fmul(fr, s), fmul(fi, c));
    }
  } else {                                                 // pRotatE, model.py:236-243
    float pf = fdiv(F[k], scale), pr = fdiv(Rr[k], scale);
    q[k] = HEAD ? fsub(pr, pf) : fadd(pf, pr);
  }
}
template <int MODEL, bool HEAD>
__device__ __forceinline__ void build_q(const float *__restrict__ F, const float *__restrict__ Rr, int k, int d,
                                        float scale, float *__restrict__ q) {
  build_q<MODEL, HEAD>(F, Rr, k, d, scale, q, d);
}

// RotatE with the rotation (cos, sin of the relation phase) given: the persistent train kernel evaluates the phases of
// a row's relation once and reuses them for q, the chain rule and the positive triple (same values as build_q /
// chain_q, which call sincos_rep themselves).
template <bool HEAD>
__device__ __forceinline__ void build_q_rot(const float *__restrict__ F, float c, float s, int k, int d,
                                            float *__restrict__ q, int qd) {
  const float fr = F[k], fi = F[d + k];
  if (HEAD) {                                              // model.py:215-216
    q[k] = fadd(fmul(c, fr), fmul(s, fi));
    q[qd + k] = fsub(fmul(c, fi), fmul(s, fr));
  } else {                                                 // model.py:220-221
    q[k] = fsub(fmul(fr, c), fmul(fi, s));
    q[qd + k] = fadd(fmul(fr, s), fmul(fi, c));
  }
}
template <bool HEAD>
__device__ __forceinline__ void chain_q_rot(const float *__restrict__ F, float c, float s, const float *__restrict__ dq,
                                            int k, int d, float scale, float &dF0, float &dF1, float &dR0) {
  const float fr = F[k], fi = F[d + k], a = dq[k], b = dq[d + k];
  float dc, ds;
  if (HEAD) {      // q = conj(e^{i th}) * t
    dF0 = a * c - b * s;  dF1 = a * s + b * c;
    dc = a * fr + b * fi; ds = a * fi - b * fr;
  } else {         // q = h * e^{i th}
    dF0 = a * c + b * s;  dF1 = -a * s + b * c;
    dc = a * fr + b * fi; ds = -a * fi + b * fr;
  }
  dR0 = (-dc * s + ds * c) / scale;
}

// ---- forward element op (train path: compiler may contract, approximate sqrt allowed) ------------------
template <int OP>
__device__ __forceinline__ float op_forward(float q0, float q1, float x0, float x1, float scale) {
  if constexpr (OP == OP_SUBABS) return fabsf(q0 - x0);
  else if constexpr (OP == OP_ADDABS) return fabsf(x0 + q0);
  else if constexpr (OP == OP_MUL) return q0 * x0;
  else if constexpr (OP == OP_CMUL) return q0 * x0 + q1 * x1;
  else if constexpr (OP == OP_CDIST) {
    float a = q0 - x0, b = q1 - x1;
    return sqrt_approx(a * a + b * b);
  } else if constexpr (OP == OP_SUBSIN) return fabsf(sinf(q0 - x0 / scale));
  else return fabsf(sinf(x0 / scale + q0));
}

// derivative: go = dL/d(sum over k).  Returns the element value (pRotatE needs it for d/dmodulus).
template <int OP>
__device__ __forceinline__ float op_backward(float q0, float q1, float x0, float x1, float scale, float go,
                                             float &dq0, float &dq1, float &dx0, float &dx1) {
  if constexpr (OP == OP_SUBABS) {
    float e = q0 - x0;
    float de = e > 0.f ? go : (e < 0.f ? -go : 0.f);       // sign(0) = 0 like torch
    dq0 = de; dx0 = -de;
    return 0.f;
  } else if constexpr (OP == OP_ADDABS) {
    float e = x0 + q0;
    float de = e > 0.f ? go : (e < 0.f ? -go : 0.f);
    dq0 = de; dx0 = de;
    return 0.f;
  } else if constexpr (OP == OP_MUL) {
    dq0 = go * x0; dx0 = go * q0;
    return 0.f;
  } else if constexpr (OP == OP_CMUL) {
    dq0 = go * x0; dq1 = go * x1; dx0 = go * q0; dx1 = go * q1;
    return 0.f;
  } else if constexpr (OP == OP_CDIST) {
    float a = q0 - x0, b = q1 - x1;
    float m2 = a * a + b * b;
    float inv = m2 >= kFltMin ? go * rsqrt_fast(m2) : 0.f; // norm subgradient is 0 at the origin
    dq0 = a * inv; dq1 = b * inv; dx0 = -dq0; dx1 = -dq1;
    return 0.f;
  } else {
    float px = x0 / scale;
    float e = OP == OP_SUBSIN ? q0 - px : px + q0;
    float s, c;
    sincosf(e, &s, &c);
    float de = s > 0.f ? go * c : (s < 0.f ? -go * c : 0.f);
    dq0 = de;
    dx0 = (OP == OP_SUBSIN ? -de : de) / scale;
    return fabsf(s);
  }
}

// forward element value together with u = d(value)/dq, so that dL/dq = (dL/dsum) * u.  Used by the single-read
// path: sweep 1 computes value and u once, u is parked in the shared-memory slot, sweep 2 only does acc += c * u.
template <int OP>
__device__ __forceinline__ float op_unit(float q0, float q1, float x0, float x1, float scale, float &u0, float &u1) {
  u1 = 0.f;
  if constexpr (OP == OP_SUBABS || OP == OP_ADDABS) {
    const float e = OP == OP_SUBABS ? q0 - x0 : x0 + q0;
    u0 = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
    return fabsf(e);
  } else if constexpr (OP == OP_MUL) {
    u0 = x0;
    return q0 * x0;
  } else if constexpr (OP == OP_CMUL) {
    u0 = x0; u1 = x1;
    return q0 * x0 + q1 * x1;
  } else if constexpr (OP == OP_CDIST) {
    const float a = q0 - x0, b = q1 - x1;
    const float m2 = a * a + b * b;
    const float inv = m2 >= kFltMin ? rsqrt_fast(m2) : 0.f; // norm subgradient 0 at the origin
    u0 = a * inv; u1 = b * inv;
    return m2 * inv;                                       // = sqrt(m2)
  } else {
    const float px = x0 / scale;
    const float e = OP == OP_SUBSIN ? q0 - px : px + q0;
    float sn, cs;
    sincosf(e, &sn, &cs);
    u0 = sn > 0.f ? cs : (sn < 0.f ? -cs : 0.f);
    return fabsf(sn);
  }
}
__host__ __device__ constexpr bool op_unit_is_x(int op) { return op == OP_MUL || op == OP_CMUL; }

template <int MODEL>
__device__ __forceinline__ float finish_score(float acc, float gamma, float modulus) {
  if constexpr (MODEL == KGE_TRANSE || MODEL == KGE_ROTATE) return gamma - acc;      // model.py:172,228
  else if constexpr (MODEL == KGE_PROTATE) return gamma - acc * modulus;             // model.py:248
  else return acc;                                                                    // model.py:181,198
}
// dL/d(sum) given g = dL/dscore
template <int MODEL>
__device__ __forceinline__ float dsum_of(float g, float modulus) {
  if constexpr (MODEL == KGE_TRANSE || MODEL == KGE_ROTATE) return -g;
  else if constexpr (MODEL == KGE_PROTATE) return -g * modulus;
  else return g;
}

// ---- chain rule q -> (fixed entity row, relation row) at element k -------------------------------
// dF / dR are *accumulated with atomics* by the caller; this returns the values.
template <int MODEL, bool HEAD>
__device__ __forceinline__ void chain_q(const float *__restrict__ F, const float *__restrict__ Rr,
                                        const float *__restrict__ dq, int k, int d, float scale,
                                        float &dF0, float &dF1, float &dR0, float &dR1) {
  dF1 = 0.f; dR1 = 0.f;
  if constexpr (MODEL == KGE_TRANSE) {
    dR0 = dq[k];
    dF0 = HEAD ? -dq[k] : dq[k];
  } else if constexpr (MODEL == KGE_DISTMULT) {
    dF0 = dq[k] * Rr[k];
    dR0 = dq[k] * F[k];
  } else if constexpr (MODEL == KGE_COMPLEX) {
    float fr = F[k], fi = F[d + k], rr = Rr[k], ri = Rr[d + k], a = dq[k], b = dq[d + k];
    if (HEAD) {      // q = conj(r) * t
      dR0 = a * fr + b * fi;  dR1 = a * fi - b * fr;
      dF0 = a * rr - b * ri;  dF1 = a * ri + b * rr;
    } else {         // q = h * r
      dF0 = a * rr + b * ri;  dF1 = -a * ri + b * rr;
      dR0 = a * fr + b * fi;  dR1 = -a * fi + b * fr;
    }
  } else if constexpr (MODEL == KGE_ROTATE) {
    float fr = F[k], fi = F[d + k], a = dq[k], b = dq[d + k];
    float s, c;
    sincos_rep(fdiv(Rr[k], scale), &s, &c);
    float dc, ds;
    if (HEAD) {      // q = conj(e^{i th}) * t
      dF0 = a * c - b * s;  dF1 = a * s + b * c;
      dc = a * fr + b * fi; ds = a * fi - b * fr;
    } else {         // q = h * e^{i th}
      dF0 = a * c + b * s;  dF1 = -a * s + b * c;
      dc = a * fr + b * fi; ds = -a * fi + b * fr;
    }
    dR0 = (-dc * s + ds * c) / scale;
  } else {           // pRotatE: q = ph + pr  |  pr - pt
    float v = dq[k] / scale;
    dR0 = v;
    dF0 = HEAD ? -v : v;
  }
}

}  // namespace kge
