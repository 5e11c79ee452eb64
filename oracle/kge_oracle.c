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

void ko_set_num_threads(int n) {      /* bench.py: torchrun exports OMP_NUM_THREADS=1; the CPU arm asks for every core */
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
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

/* ================================================================================================
 * KGEModel.train_step (model.py:251-312) on one batch: scores, self-adversarial / uniform loss,
 * closed-form backward (what autograd computes at model.py:301; formulas of SURVEY.md 8a row G1),
 * optional L3 regulariser, dense Adam (torch/optim/adam.py defaults as built at run.py:266-269).
 * Parallel over positive rows with OpenMP; gradient scatter uses atomic adds like index_add_.
 * ================================================================================================ */
static float logsigmoidf_(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
static float sigmoidf_(float x) { float e = expf(-fabsf(x)); return x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e); }

static inline void atomic_addf(float *p, float v) {
#pragma omp atomic
  *p += v;
}

/* d(score)/d(sum over k) */
static float dsum(int model, float g, float modulus) {
  return (model == TRANSE || model == ROTATE) ? -g : (model == PROTATE ? -g * modulus : g);
}

/* backward of one candidate row: adds dL/dx into gx (atomically), accumulates dL/dq into dq; returns sum|sin| for
 * pRotatE's modulus gradient */
static float row_backward(int model, int head, const float *q, const float *x, int d, float scale, float go,
                          float *dq, float *gx) {
  float vs = 0.f;
  for (int k = 0; k < d; ++k) {
    switch (model) {
      case TRANSE: {
        float e = head ? x[k] + q[k] : q[k] - x[k];
        float de = e > 0.f ? go : (e < 0.f ? -go : 0.f);              /* sign(0) = 0 */
        dq[k] += de;
        atomic_addf(gx + k, head ? de : -de);
        break;
      }
      case DISTMULT:
        dq[k] += go * x[k];
        atomic_addf(gx + k, go * q[k]);
        break;
      case COMPLEX_:
        dq[k] += go * x[k]; dq[d + k] += go * x[d + k];
        atomic_addf(gx + k, go * q[k]); atomic_addf(gx + d + k, go * q[d + k]);
        break;
      case ROTATE: {
        float a = q[k] - x[k], b = q[d + k] - x[d + k];
        float m = sqrtf(a * a + b * b);
        float da = m > 0.f ? go * a / m : 0.f, db = m > 0.f ? go * b / m : 0.f;   /* norm subgradient 0 at 0 */
        dq[k] += da; dq[d + k] += db;
        atomic_addf(gx + k, -da); atomic_addf(gx + d + k, -db);
        break;
      }
      default: {
        float px = x[k] / scale;
        float e = head ? px + q[k] : q[k] - px;
        float s = sinf(e), c = cosf(e);
        float de = s > 0.f ? go * c : (s < 0.f ? -go * c : 0.f);
        dq[k] += de;
        atomic_addf(gx + k, (head ? de : -de) / scale);
        vs += fabsf(s);
      }
    }
  }
  return vs;
}

/* chain rule dL/dq -> fixed entity row F and relation row Rr (atomically added) */
static void chain_backward(int model, int head, const float *F, const float *Rr, const float *dq, int d, float scale,
                           float *gF, float *gR) {
  for (int k = 0; k < d; ++k) {
    switch (model) {
      case TRANSE:
        atomic_addf(gR + k, dq[k]); atomic_addf(gF + k, head ? -dq[k] : dq[k]);
        break;
      case DISTMULT:
        atomic_addf(gF + k, dq[k] * Rr[k]); atomic_addf(gR + k, dq[k] * F[k]);
        break;
      case COMPLEX_: {
        float fr = F[k], fi = F[d + k], rr = Rr[k], ri = Rr[d + k], a = dq[k], b = dq[d + k];
        if (head) {
          atomic_addf(gR + k, a * fr + b * fi); atomic_addf(gR + d + k, a * fi - b * fr);
          atomic_addf(gF + k, a * rr - b * ri); atomic_addf(gF + d + k, a * ri + b * rr);
        } else {
          atomic_addf(gF + k, a * rr + b * ri); atomic_addf(gF + d + k, -a * ri + b * rr);
          atomic_addf(gR + k, a * fr + b * fi); atomic_addf(gR + d + k, -a * fi + b * fr);
        }
        break;
      }
      case ROTATE: {
        float fr = F[k], fi = F[d + k], a = dq[k], b = dq[d + k], s, c, dc, ds;
        sincos_cw(Rr[k] / scale, &s, &c);
        if (head) {
          atomic_addf(gF + k, a * c - b * s); atomic_addf(gF + d + k, a * s + b * c);
          dc = a * fr + b * fi; ds = a * fi - b * fr;
        } else {
          atomic_addf(gF + k, a * c + b * s); atomic_addf(gF + d + k, -a * s + b * c);
          dc = a * fr + b * fi; ds = -a * fi + b * fr;
        }
        atomic_addf(gR + k, (-dc * s + ds * c) / scale);
        break;
      }
      default: {
        float v = dq[k] / scale;
        atomic_addf(gR + k, v); atomic_addf(gF + k, head ? -v : v);
      }
    }
  }
}

static void adam_dense(float *p, const float *g, float *m, float *v, int64_t n, int step, double lr, double b1,
                       double b2, double eps) {
  const float w1 = (float)(1.0 - b1), fb2 = (float)b2, w2 = (float)(1.0 - b2), feps = (float)eps;
  const float step_size = (float)(-(lr / (1.0 - pow(b1, step)))), bc2s = (float)sqrt(1.0 - pow(b2, step));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    m[i] = m[i] + (g[i] - m[i]) * w1;
    v[i] = v[i] * fb2 + w2 * g[i] * g[i];
    p[i] = p[i] + step_size * (m[i] / (sqrtf(v[i]) / bc2s + feps));
  }
}

/* One train step.  mode: 1 head-batch, 2 tail-batch.  weight may be NULL (--uni_weight).  Tables, moments and grads
 * are updated in place; out[0..3] = positive_sample_loss, negative_sample_loss, loss, regularization. */
void ko_train_step(int model, int mode, float *E, float *R, float *modulus, int64_t nentity, int64_t nrelation, int De,
                   int Dr, float gamma, float rho, const int64_t *positive, const int64_t *negative,
                   const float *weight, int64_t B, int64_t N, int adversarial, float alpha, double reg, double lr,
                   int step, float *mE, float *vE, float *mR, float *vR, float *mM, float *vM, float *gE, float *gR,
                   float *gM, float *out) {
  const int d = kdim(model, De);
  const int head = mode == 1;
  const float scale = phase_scale(model, rho);
  const float mod = modulus ? modulus[0] : 1.f;
  memset(gE, 0, sizeof(float) * nentity * De);
  memset(gR, 0, sizeof(float) * nrelation * Dr);
  if (gM) gM[0] = 0.f;
  double wsum = 0.0;
  for (int64_t b = 0; b < B; ++b) wsum += weight ? weight[b] : 1.0;
  double pos_acc = 0.0, neg_acc = 0.0, gmod_acc = 0.0;
#pragma omp parallel reduction(+ : pos_acc, neg_acc, gmod_acc)
  {
    float *q = (float *)malloc(sizeof(float) * De), *dq = (float *)malloc(sizeof(float) * De);
    float *s = (float *)malloc(sizeof(float) * N), *g = (float *)malloc(sizeof(float) * N);
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
      const int64_t h = positive[3 * b], r = positive[3 * b + 1], t = positive[3 * b + 2];
      const float u = (float)((weight ? weight[b] : 1.0) / wsum);
      /* negatives (model.py:268-275) */
      const int64_t f = head ? t : h;
      build_q(model, head, E + f * De, R + r * Dr, d, scale, q);
      float zmax = -INFINITY;
      for (int64_t n = 0; n < N; ++n) {
        s[n] = score_row(model, head, q, E + negative[b * N + n] * De, d, gamma, scale, mod);
        if (adversarial && s[n] * alpha > zmax) zmax = s[n] * alpha;
      }
      float zsum = 0.f;
      for (int64_t n = 0; n < N; ++n) { g[n] = adversarial ? expf(s[n] * alpha - zmax) : 1.f; zsum += g[n]; }
      float row = 0.f;
      for (int64_t n = 0; n < N; ++n) {
        const float w = g[n] / zsum;
        row += w * logsigmoidf_(-s[n]);
        g[n] = 0.5f * u * w * sigmoidf_(s[n]);
      }
      neg_acc += (double)((weight ? weight[b] : 1.f) * row);
      memset(dq, 0, sizeof(float) * De);
      for (int64_t n = 0; n < N; ++n) {
        const int64_t c = negative[b * N + n];
        const float vs = row_backward(model, head, q, E + c * De, d, scale, dsum(model, g[n], mod), dq, gE + c * De);
        gmod_acc += (double)(-g[n] * vs);
      }
      chain_backward(model, head, E + f * De, R + r * Dr, dq, d, scale, gE + f * De, gR + r * Dr);
      /* positive triple, 'single' mode = the non-head-batch association (model.py:277-279) */
      build_q(model, 0, E + h * De, R + r * Dr, d, scale, q);
      const float sp = score_row(model, 0, q, E + t * De, d, gamma, scale, mod);
      pos_acc += (double)((weight ? weight[b] : 1.f) * logsigmoidf_(sp));
      const float gp = -0.5f * u * sigmoidf_(-sp);
      memset(dq, 0, sizeof(float) * De);
      const float vs = row_backward(model, 0, q, E + t * De, d, scale, dsum(model, gp, mod), dq, gE + t * De);
      gmod_acc += (double)(-gp * vs);
      chain_backward(model, 0, E + h * De, R + r * Dr, dq, d, scale, gE + h * De, gR + r * Dr);
    }
    free(q); free(dq); free(s); free(g);
  }
  if (gM) gM[0] = (float)gmod_acc;
  const float pl = (float)(-pos_acc / wsum), nl = (float)(-neg_acc / wsum);
  float regv = 0.f;
  if (reg != 0.0) {                                  /* model.py:290-296 */
    double r3 = 0.0;
    const int64_t nE = nentity * De, nR = nrelation * Dr;
#pragma omp parallel for reduction(+ : r3) schedule(static)
    for (int64_t i = 0; i < nE; ++i) { float ax = fabsf(E[i]); r3 += (double)(ax * ax * ax); gE[i] += (float)(3.0 * reg) * E[i] * ax; }
#pragma omp parallel for reduction(+ : r3) schedule(static)
    for (int64_t i = 0; i < nR; ++i) { float ax = fabsf(R[i]); r3 += (double)(ax * ax * ax); gR[i] += (float)(3.0 * reg) * R[i] * ax; }
    regv = (float)(reg * r3);
  }
  out[0] = pl; out[1] = nl; out[2] = (pl + nl) / 2.f + regv; out[3] = regv;
  adam_dense(E, gE, mE, vE, nentity * De, step, lr, 0.9, 0.999, 1e-8);      /* model.py:303 */
  adam_dense(R, gR, mR, vR, nrelation * Dr, step, lr, 0.9, 0.999, 1e-8);
  if (modulus) adam_dense(modulus, gM, mM, vM, 1, step, lr, 0.9, 0.999, 1e-8);
}
