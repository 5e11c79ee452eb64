/*
 * kge_oracle.c -- CPU oracle for the KGEModel hot path, plain C.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates the reference algorithm (/root/reference/codes/model.py, codes/dataloader.py) in scalar C so
 * that full-size cases finish in seconds.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load it (through oracle/c_oracle.py); nothing in the product does.
 *
 * Parity pin: tests/test_c_oracle.py checks every function here against oracle/kge_oracle.py, which is
 * itself pinned to golden vectors produced by running the unmodified reference
 * (tests/golden/make_golden.py).  The reference ships no tests of its own ("parity unpinned" by the
 * reference; pinned here on its own outputs).
 *
 * Evaluation scores are computed with the same IEEE operation sequence the CUDA evaluation kernels
 * document (un-fused element ops in the reference's association, index-order fp32
 * accumulation over k, the same Cody-Waite sin/cos), so GPU scores and ranks can be compared bit-for-bit.
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no fused multiply-add unless written as fmaf).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

enum { TRANSE = 0, DISTMULT = 1, COMPLEX_ = 2, ROTATE = 3, PROTATE = 4 };

int ko_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- sin/cos: Cody-Waite reduction by pi/2 + degree-7/8 minimax polynomials (restated from the
 * published Cephes sinf/cosf scheme); every step is a correctly rounded IEEE op, so any conforming
 * platform gives the same bits. ---- */
static void sincos_cw(float x, float *sn, float *cs) {
  float r;
  int q;
  if (fabsf(x) < 40000.0f) {
    float k = rintf(x * 0.636619772367581343f);
    q = (int)k;
    r = fmaf(-k, 1.5703125f, x);
    r = fmaf(-k, 4.837512969970703125e-4f, r);
    r = fmaf(-k, 7.54978995489188e-8f, r);
  } else if (fabsf(x) < 1.0e15f) {
    double k = rint((double)x * 0.63661977236758134308);
    q = (int)((long long)k & 3);
    double rd = fma(-k, 1.57079632679489655800e+00, (double)x);
    rd = fma(-k, 6.12323399573676603587e-17, rd);
    r = (float)rd;
  } else {
    r = x - x;
    q = 0;
  }
  float z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(z, ps, -1.6666654611e-1f);
  float s = fmaf(r * z, ps, r);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(z, pc, 4.166664568298827e-2f);
  float c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
  switch (q & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}

void ko_sincos(const float *x, int64_t n, float *s, float *c) {
  for (int64_t i = 0; i < n; ++i) sincos_cw(x[i], s + i, c + i);
}

/* rho/pi as the fp32 scalar the reference divides by: model.py:209 (pi) and :236 (the pRotatE constant) */
static float phase_scale(int model, float rho) {
  const double pi = model == PROTATE ? 3.14159262358979323846 : 3.14159265358979323846;
  return (float)((double)rho / pi);
}

/* ---- query vector: the fixed side of a triple folded in the reference's association --------------------
 * F = fixed entity row (head for tail-batch/single, tail for head-batch), Rr = relation row. */
static void build_q(int model, int head, const float *F, const float *Rr, int d, float scale, float *q) {
  for (int k = 0; k < d; ++k) {
    switch (model) {
      case TRANSE:                       /* model.py:168 (relation - tail) / :170 (head + relation) */
        q[k] = head ? Rr[k] - F[k] : F[k] + Rr[k];
        break;
      case DISTMULT:                     /* model.py:177 / :179 */
        q[k] = head ? Rr[k] * F[k] : F[k] * Rr[k];
        break;
      case COMPLEX_: {                   /* model.py:190-191 / :194-195 */
        float fr = F[k], fi = F[d + k], rr = Rr[k], ri = Rr[d + k];
        if (head) { q[k] = rr * fr + ri * fi; q[d + k] = rr * fi - ri * fr; }
        else      { q[k] = fr * rr - fi * ri; q[d + k] = fr * ri + fi * rr; }
        break;
      }
      case ROTATE: {                     /* model.py:209-212, 215-216 / 220-221 */
        float fr = F[k], fi = F[d + k], s, c;
        sincos_cw(Rr[k] / scale, &s, &c);
        if (head) { q[k] = c * fr + s * fi; q[d + k] = c * fi - s * fr; }
        else      { q[k] = fr * c - fi * s; q[d + k] = fr * s + fi * c; }
        break;
      }
      default: {                         /* pRotatE model.py:236-243 */
        float pf = F[k] / scale, pr = Rr[k] / scale;
        q[k] = head ? pr - pf : pf + pr;
      }
    }
  }
}

/* Accumulation order of evaluation scores: blocks of KB consecutive k summed in index order from 0.0f, block
 * sums added in index order (documented in DESIGN.md section 4; the CUDA kernels do the same). */
#define KB 32

static float elem(int model, int head, const float *q, const float *x, int k, int d, float scale) {
  switch (model) {
    case TRANSE:   return head ? fabsf(x[k] + q[k]) : fabsf(q[k] - x[k]);           /* model.py:168-172 */
    case DISTMULT: return q[k] * x[k];                                              /* model.py:177-181 */
    case COMPLEX_: { float t0 = q[k] * x[k], t1 = q[d + k] * x[d + k]; return t0 + t1; }   /* model.py:192-198 */
    case ROTATE: {                                                                  /* model.py:217-226 */
      float a = q[k] - x[k], b = q[d + k] - x[d + k];
      return sqrtf(fmaf(b, b, a * a));                 /* stack + norm(dim=0): sqrt(a^2 + b^2) */
    }
    default: {                                                                      /* model.py:241-246 */
      float px = x[k] / scale, s, c;
      sincos_cw(head ? px + q[k] : q[k] - px, &s, &c);
      return fabsf(s);
    }
  }
}

/* score of candidate row x against q (model.py:172,181,198,228,248) */
static float score_row(int model, int head, const float *q, const float *x, int d, float gamma, float scale,
                       float modulus) {
  float acc = 0.f;
  for (int k0 = 0; k0 < d; k0 += KB) {
    float part = 0.f;
    const int k1 = k0 + KB < d ? k0 + KB : d;
    for (int k = k0; k < k1; ++k) part = part + elem(model, head, q, x, k, d, scale);
    acc = acc + part;
  }
  switch (model) {
    case TRANSE: case ROTATE: return gamma - acc;
    case PROTATE: return gamma - acc * modulus;
    default: return acc;
  }
}

static int kdim(int model, int De) { return (model == COMPLEX_ || model == ROTATE) ? De / 2 : De; }

void ko_query_vectors(int model, int head, const float *E, const float *R, int De, int Dr, float rho,
                      const int64_t *queries, int64_t Q, float *qvec) {
  const int d = kdim(model, De);
  const float scale = phase_scale(model, rho);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < Q; ++i) {
    const int64_t f = queries[3 * i + (head ? 2 : 0)], r = queries[3 * i + 1];
    build_q(model, head, E + f * De, R + r * Dr, d, scale, qvec + i * De);
  }
}

/* KGEModel.forward for (positive [B,3], candidates [B,N]) in head-/tail-batch mode, or 'single' (mode 0) */
void ko_forward(int model, int mode, const float *E, const float *R, int De, int Dr, float gamma, float rho,
                float modulus, const int64_t *positive, const int64_t *negative, int64_t B, int64_t N, float *score) {
  const int d = kdim(model, De);
  const int head = mode == 1;
  const float scale = phase_scale(model, rho);
#pragma omp parallel
  {
    float *q = (float *)malloc(sizeof(float) * De);
#pragma omp for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
      const int64_t f = positive[3 * b + (head ? 2 : 0)], r = positive[3 * b + 1];
      build_q(model, head, E + f * De, R + r * Dr, d, scale, q);
      for (int64_t n = 0; n < N; ++n) {
        const int64_t c = mode == 0 ? positive[3 * b + 2] : negative[b * N + n];
        score[b * N + n] = score_row(model, head, q, E + c * De, d, gamma, scale, modulus);
      }
    }
    free(q);
  }
}

/* model.py:392-393 for a chunk of queries: all-entity scores + filter_bias, with dataloader.py:137-151's
 * encoding (a filtered column holds the positive id with bias -1, i.e. score(positive) - 1). */
void ko_eval_scores(int model, int head, const float *E, const float *R, int64_t nentity, int De, int Dr, float gamma,
                    float rho, float modulus, const int64_t *queries, int64_t Q, const int64_t *csr_off,
                    const int32_t *csr_ent, float *scores /* [Q, nentity] */) {
  const int d = kdim(model, De);
  const float scale = phase_scale(model, rho);
#pragma omp parallel
  {
    float *q = (float *)malloc(sizeof(float) * De);
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < Q; ++i) {
      const int64_t f = queries[3 * i + (head ? 2 : 0)], r = queries[3 * i + 1];
      const int64_t pos = queries[3 * i + (head ? 0 : 2)];
      float *row = scores + i * nentity;
      build_q(model, head, E + f * De, R + r * Dr, d, scale, q);
      for (int64_t j = 0; j < nentity; ++j) row[j] = score_row(model, head, q, E + j * De, d, gamma, scale, modulus);
      const float filtered = row[pos] + (-1.0f);
      for (int64_t t = csr_off[i]; t < csr_off[i + 1]; ++t)
        if (csr_ent[t] != pos) row[csr_ent[t]] = filtered;
    }
    free(q);
  }
}

/* model.py:396-411 with a stable descending order: 1 + #greater + #equal at a lower column */
void ko_ranks_from_scores(const float *scores, int64_t Q, int64_t nentity, const int64_t *queries, int head,
                          int64_t *ranks) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < Q; ++i) {
    const int64_t pos = queries[3 * i + (head ? 0 : 2)];
    const float *row = scores + i * nentity;
    const float sp = row[pos];
    int64_t c = 0;
    for (int64_t j = 0; j < nentity; ++j) c += (row[j] > sp) || (row[j] == sp && j < pos);
    ranks[i] = 1 + c;
  }
}
