/*
 * cfmm_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See cfmm_oracle.h for the scope statement and the parity pin.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math
 * (no FMA contraction, no reassociation: Julia never fuses a*b+c and evaluates
 * a*b*c left to right, so the operation order below is the reference's).
 */
#include "cfmm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Julia's max(x, 0) for Float64: NaN propagates (Base.max), otherwise the
 * larger value; max(-0.0, 0) is +0.0. */
static inline double jl_max0(double x) {
  if (x != x) return x;
  return x > 0.0 ? x : 0.0;
}

/* ------------------------------------------------------------------------- */
/* ProductTwoCoin                                                             */
/* ------------------------------------------------------------------------- */

/* src/cfmms.jl:125  prod_arb_δ(m, r, k, γ) = max(sqrt(γ*m*k) - r, 0)/γ */
static inline double prod_arb_delta(double m, double r, double k, double g) {
  return jl_max0(sqrt((g * m) * k) - r) / g;
}
/* src/cfmms.jl:126  prod_arb_λ(m, r, k, γ) = max(r - sqrt(k/(m*γ)), 0) */
static inline double prod_arb_lambda(double m, double r, double k, double g) {
  return jl_max0(r - sqrt(k / (m * g)));
}

/* src/cfmms.jl:130-140 */
void oracle_product_arb(const double R[2], double gamma, const double v[2],
                        double Delta[2], double Lambda[2]) {
  double k = R[0] * R[1];
  Delta[0] = prod_arb_delta(v[1] / v[0], R[0], k, gamma);
  Delta[1] = prod_arb_delta(v[0] / v[1], R[1], k, gamma);
  Lambda[0] = prod_arb_lambda(v[0] / v[1], R[0], k, gamma);
  Lambda[1] = prod_arb_lambda(v[1] / v[0], R[1], k, gamma);
}

/* ------------------------------------------------------------------------- */
/* GeometricMeanTwoCoin                                                       */
/* ------------------------------------------------------------------------- */

/* src/cfmms.jl:180  geom_arb_δ(m,r1,r2,η,γ) = max((γ*m*η*r1*r2^η)^(1/(η+1)) - r2, 0)/γ */
static inline double geom_arb_delta(double m, double r1, double r2, double e,
                                    double g) {
  double base = (((g * m) * e) * r1) * pow(r2, e);
  return jl_max0(pow(base, 1.0 / (e + 1.0)) - r2) / g;
}
/* src/cfmms.jl:181  geom_arb_λ(m,r1,r2,η,γ) = max(r1 - ((r2*r1^(1/η))/(η*γ*m))^(η/(1+η)), 0) */
static inline double geom_arb_lambda(double m, double r1, double r2, double e,
                                     double g) {
  double base = (r2 * pow(r1, 1.0 / e)) / ((e * g) * m);
  return jl_max0(r1 - pow(base, e / (1.0 + e)));
}

/* src/cfmms.jl:185-196 */
void oracle_geomean_arb(const double R[2], const double w[2], double gamma,
                        const double v[2], double Delta[2], double Lambda[2]) {
  double eta = w[0] / w[1];
  Delta[0] = geom_arb_delta(v[1] / v[0], R[1], R[0], eta, gamma);
  Delta[1] = geom_arb_delta(v[0] / v[1], R[0], R[1], 1.0 / eta, gamma);
  Lambda[0] = geom_arb_lambda(v[0] / v[1], R[0], R[1], 1.0 / eta, gamma);
  Lambda[1] = geom_arb_lambda(v[1] / v[0], R[1], R[0], eta, gamma);
}

/* ------------------------------------------------------------------------- */
/* UniV3 / BoundedProduct                                                     */
/* ------------------------------------------------------------------------- */

typedef struct {
  double k, alpha, beta, R1, R2; /* src/cfmms.jl:272-278 */
} bounded_product;

/* src/cfmms.jl:233-235 */
int64_t oracle_univ3_current_tick(const double *lower_ticks, int64_t n_ticks,
                                  double current_price) {
  /* lower_ticks is sorted in decreasing order; searchsortedlast(...; rev=true)
   * returns the last index whose value is >= current_price. */
  int64_t lo = 0, hi = n_ticks; /* count of leading entries >= current_price */
  while (lo < hi) {
    int64_t mid = lo + (hi - lo) / 2;
    if (lower_ticks[mid] >= current_price)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* src/cfmms.jl:251-259 and :294-313; idx is 1-based */
static bounded_product compute_at_tick(double current_price,
                                       int64_t current_tick,
                                       const double *lower_ticks,
                                       const double *liquidity, int64_t n_ticks,
                                       int64_t idx) {
  bounded_product t;
  double k = liquidity[idx - 1];
  double pminus = (idx < n_ticks) ? lower_ticks[idx] : 0.0;
  double pplus = lower_ticks[idx - 1];
  double alpha = sqrt(k / pplus);
  double beta = sqrt(k * pminus);
  double p;
  if (idx > current_tick)
    p = pplus;
  else if (idx < current_tick)
    p = pminus;
  else
    p = current_price;
  t.k = k;
  t.alpha = alpha;
  t.beta = beta;
  t.R1 = sqrt(k / p) - alpha;
  t.R2 = sqrt(k * p) - beta;
  return t;
}

void oracle_univ3_tick(double current_price, int64_t current_tick,
                       const double *lower_ticks, const double *liquidity,
                       int64_t n_ticks, int64_t idx, double out[5]) {
  bounded_product t = compute_at_tick(current_price, current_tick, lower_ticks,
                                      liquidity, n_ticks, idx);
  out[0] = t.k;
  out[1] = t.alpha;
  out[2] = t.beta;
  out[3] = t.R1;
  out[4] = t.R2;
}

/* src/cfmms.jl:289 */
static inline bounded_product flip_sides(bounded_product t) {
  bounded_product f = {t.k, t.beta, t.alpha, t.R2, t.R1};
  return f;
}

/* src/cfmms.jl:321-337 */
static inline void find_arb_pos(bounded_product t, double price, double *delta,
                                double *lambda) {
  double d = sqrt(t.k / price) - (t.R1 + t.alpha);
  if (d <= 0.0) {
    *delta = 0.0;
    *lambda = 0.0;
    return;
  }
  double d_max = t.k / t.beta - (t.R1 + t.alpha);
  if (d >= d_max) {
    *delta = d_max;
    *lambda = t.R2;
    return;
  }
  *delta = d;
  *lambda = (t.R2 + t.beta) - sqrt(price * t.k);
}

/* src/cfmms.jl:339-395 */
void oracle_univ3_arb(double current_price, int64_t current_tick,
                      const double *lower_ticks, const double *liquidity,
                      int64_t n_ticks, double gamma, const double v[2],
                      double Delta[2], double Lambda[2]) {
  double p = v[0] / v[1];
  Delta[0] = Delta[1] = 0.0;
  Lambda[0] = Lambda[1] = 0.0;

  /* no-arb interval, :347 */
  if (gamma * current_price <= p && p <= current_price / gamma) return;

  if (p < gamma * current_price) {
    int initial = 1;
    double price = p / gamma;
    for (int64_t idx = current_tick; idx <= n_ticks; ++idx) {
      bounded_product pool = compute_at_tick(current_price, current_tick,
                                             lower_ticks, liquidity, n_ticks, idx);
      if (pool.k == 0.0) {
        initial = 0;
        continue;
      }
      double d, l;
      find_arb_pos(pool, price, &d, &l);
      if (!initial && (d == 0.0 || l == 0.0)) break;
      Delta[0] += d;
      Lambda[1] += l;
      initial = 0;
    }
    Delta[0] /= gamma;
  } else {
    int initial = 1;
    double price = 1.0 / (gamma * p);
    for (int64_t idx = current_tick; idx >= 1; --idx) {
      bounded_product pool = flip_sides(compute_at_tick(
          current_price, current_tick, lower_ticks, liquidity, n_ticks, idx));
      if (pool.k == 0.0) {
        initial = 0;
        continue;
      }
      double d, l;
      find_arb_pos(pool, price, &d, &l);
      if (!initial && (d == 0.0 || l == 0.0)) break;
      Delta[1] += d;
      Lambda[0] += l;
      initial = 0;
    }
    Delta[1] /= gamma;
  }
}

/* src/cfmms.jl:401-409 */
static inline double max_amount_pos(bounded_product t) {
  if (t.beta > 0.0) return t.k / t.beta - (t.R1 + t.alpha);
  if (t.alpha > 0.0) return INFINITY;
  return 0.0;
}
/* src/cfmms.jl:411-414 */
static inline double forward_amount(bounded_product t, double d) {
  double l = (t.R2 + t.beta) - t.k / (t.R1 + t.alpha + d);
  return t.R2 < l ? t.R2 : l;
}

/* src/cfmms.jl:417-449 (trade_through_pools + forward_trade) */
double oracle_univ3_forward_trade(double current_price, int64_t current_tick,
                                  const double *lower_ticks,
                                  const double *liquidity, int64_t n_ticks,
                                  double gamma, const double Delta[2]) {
  if (Delta[0] == 0.0 && Delta[1] == 0.0) return 0.0;
  double lambda = 0.0;
  if (Delta[0] > 0.0) {
    double d = gamma * Delta[0];
    for (int64_t idx = current_tick; idx <= n_ticks; ++idx) {
      bounded_product pool = compute_at_tick(current_price, current_tick,
                                             lower_ticks, liquidity, n_ticks, idx);
      double mx = max_amount_pos(pool);
      if (mx > d) return lambda + forward_amount(pool, d);
      lambda += pool.R2;
      d -= mx;
    }
  } else {
    double d = gamma * Delta[1];
    for (int64_t idx = current_tick; idx >= 1; --idx) {
      bounded_product pool = flip_sides(compute_at_tick(
          current_price, current_tick, lower_ticks, liquidity, n_ticks, idx));
      double mx = max_amount_pos(pool);
      if (mx > d) return lambda + forward_amount(pool, d);
      lambda += pool.R2;
      d -= mx;
    }
  }
  return lambda;
}

/* ------------------------------------------------------------------------- */
/* Router sweep, src/router.jl:38-42                                          */
/* ------------------------------------------------------------------------- */

void oracle_sweep_product(int64_t m, const double *R, const double *gamma,
                          const int64_t *Ai, const double *v, double *Delta,
                          double *Lambda, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
  for (int64_t i = 0; i < m; ++i) {
    double vi[2] = {v[Ai[2 * i] - 1], v[Ai[2 * i + 1] - 1]}; /* v[cfmm.Ai] */
    oracle_product_arb(R + 2 * i, gamma[i], vi, Delta + 2 * i, Lambda + 2 * i);
  }
}

void oracle_sweep_geomean(int64_t m, const double *R, const double *gamma,
                          const int64_t *Ai, const double *w, const double *v,
                          double *Delta, double *Lambda, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
  for (int64_t i = 0; i < m; ++i) {
    double vi[2] = {v[Ai[2 * i] - 1], v[Ai[2 * i + 1] - 1]};
    oracle_geomean_arb(R + 2 * i, w + 2 * i, gamma[i], vi, Delta + 2 * i,
                       Lambda + 2 * i);
  }
}

void oracle_sweep_univ3(int64_t m, const double *current_price,
                        const double *gamma, const int64_t *Ai,
                        const int64_t *tick_off, const double *lower_ticks,
                        const double *liquidity, const double *v,
                        double *Delta, double *Lambda, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
  for (int64_t i = 0; i < m; ++i) {
    double vi[2] = {v[Ai[2 * i] - 1], v[Ai[2 * i + 1] - 1]};
    const double *lt = lower_ticks + tick_off[i];
    const double *lq = liquidity + tick_off[i];
    int64_t T = tick_off[i + 1] - tick_off[i];
    int64_t ct = oracle_univ3_current_tick(lt, T, current_price[i]);
    oracle_univ3_arb(current_price[i], ct, lt, lq, T, gamma[i], vi,
                     Delta + 2 * i, Lambda + 2 * i);
  }
}

/* ------------------------------------------------------------------------- */
/* Router folds, src/router.jl:79-83 (acc) and :98-100 (G scatter)            */
/* ------------------------------------------------------------------------- */

void oracle_fold(int64_t m, const int64_t *Ai, const double *Delta,
                 const double *Lambda, const double *v, double *acc, double *G) {
  double a = *acc;
  for (int64_t i = 0; i < m; ++i) {
    int64_t a1 = Ai[2 * i] - 1, a2 = Ai[2 * i + 1] - 1;
    double v1 = v[a1], v2 = v[a2];
    /* dot(Λ, v[Ai]) - dot(Δ, v[Ai]); Julia's generic dot starts from zero and
     * adds the products left to right */
    double dl = (0.0 + Lambda[2 * i] * v1) + Lambda[2 * i + 1] * v2;
    double dd = (0.0 + Delta[2 * i] * v1) + Delta[2 * i + 1] * v2;
    a += dl - dd;
    if (G) {
      G[a1] += Lambda[2 * i] - Delta[2 * i];
      G[a2] += Lambda[2 * i + 1] - Delta[2 * i + 1];
    }
  }
  *acc = a;
}

static inline void neumaier_add(long double *s, long double *c, long double x) {
  long double t = *s + x;
  if (fabsl(*s) >= fabsl(x))
    *c += (*s - t) + x;
  else
    *c += (x - t) + *s;
  *s = t;
}

void oracle_fold_compensated(int64_t m, const int64_t *Ai, const double *Delta,
                             const double *Lambda, const double *v,
                             long double *acc, long double *G, double *abs_G) {
  long double s = *acc, c = 0.0L;
  for (int64_t i = 0; i < m; ++i) {
    int64_t a1 = Ai[2 * i] - 1, a2 = Ai[2 * i + 1] - 1;
    long double f1 = (long double)Lambda[2 * i] - (long double)Delta[2 * i];
    long double f2 =
        (long double)Lambda[2 * i + 1] - (long double)Delta[2 * i + 1];
    neumaier_add(&s, &c, f1 * (long double)v[a1]);
    neumaier_add(&s, &c, f2 * (long double)v[a2]);
    if (G) {
      G[a1] += f1; /* 64-bit mantissa: 11 extra bits over the fp64 inputs */
      G[a2] += f2;
    }
    if (abs_G) {
      abs_G[a1] += fabs(Lambda[2 * i]) + fabs(Delta[2 * i]);
      abs_G[a2] += fabs(Lambda[2 * i + 1]) + fabs(Delta[2 * i + 1]);
    }
  }
  *acc = s + c;
}

/* ------------------------------------------------------------------------- */
/* CPU timing baseline, "faithful layout" flavour                             */
/* ------------------------------------------------------------------------- */

typedef struct faithful_pool faithful_pool;
typedef void (*arb_fn)(double *D, double *L, const faithful_pool *c,
                       const double *v);
struct faithful_pool {
  arb_fn find_arb; /* dynamic dispatch on an abstractly typed vector */
  double *R;       /* heap Vector{T}(2), cfmms.jl:13 */
  double gamma;
  int64_t *Ai; /* heap Vector{Int}(2), cfmms.jl:15 */
  double *w;   /* geomean only */
};

struct oracle_faithful {
  int64_t n_tokens, m, cap;
  faithful_pool **cfmms; /* Vector{CFMM{T}} of boxed objects, router.jl:6 */
  double **Ds, **Ls;     /* Vector{AbstractVector{T}}, router.jl:7-8 */
};

static void faithful_product(double *D, double *L, const faithful_pool *c,
                             const double *v) {
  oracle_product_arb(c->R, c->gamma, v, D, L);
}
static void faithful_geomean(double *D, double *L, const faithful_pool *c,
                             const double *v) {
  oracle_geomean_arb(c->R, c->w, c->gamma, v, D, L);
}

oracle_faithful *oracle_faithful_create(int64_t n_tokens) {
  oracle_faithful *o = (oracle_faithful *)calloc(1, sizeof(*o));
  o->n_tokens = n_tokens;
  return o;
}

static void faithful_grow(oracle_faithful *o, int64_t extra) {
  if (o->m + extra <= o->cap) return;
  int64_t cap = o->m + extra;
  o->cfmms = (faithful_pool **)realloc(o->cfmms, cap * sizeof(*o->cfmms));
  o->Ds = (double **)realloc(o->Ds, cap * sizeof(*o->Ds));
  o->Ls = (double **)realloc(o->Ls, cap * sizeof(*o->Ls));
  o->cap = cap;
}

static void faithful_push(oracle_faithful *o, arb_fn fn, const double *R,
                          double gamma, const int64_t *Ai, const double *w) {
  faithful_pool *c = (faithful_pool *)malloc(sizeof(*c));
  c->find_arb = fn;
  c->R = (double *)malloc(2 * sizeof(double));
  c->R[0] = R[0];
  c->R[1] = R[1];
  c->gamma = gamma;
  c->Ai = (int64_t *)malloc(2 * sizeof(int64_t));
  c->Ai[0] = Ai[0];
  c->Ai[1] = Ai[1];
  c->w = NULL;
  if (w) {
    c->w = (double *)malloc(2 * sizeof(double));
    c->w[0] = w[0];
    c->w[1] = w[1];
  }
  o->cfmms[o->m] = c;
  o->Ds[o->m] = (double *)calloc(2, sizeof(double)); /* zerotrade, router.jl:23-26 */
  o->Ls[o->m] = (double *)calloc(2, sizeof(double));
  o->m++;
}

void oracle_faithful_add_product(oracle_faithful *o, int64_t m, const double *R,
                                 const double *gamma, const int64_t *Ai) {
  faithful_grow(o, m);
  for (int64_t i = 0; i < m; ++i)
    faithful_push(o, faithful_product, R + 2 * i, gamma[i], Ai + 2 * i, NULL);
}
void oracle_faithful_add_geomean(oracle_faithful *o, int64_t m, const double *R,
                                 const double *gamma, const int64_t *Ai,
                                 const double *w) {
  faithful_grow(o, m);
  for (int64_t i = 0; i < m; ++i)
    faithful_push(o, faithful_geomean, R + 2 * i, gamma[i], Ai + 2 * i,
                  w + 2 * i);
}

double oracle_faithful_sweep(oracle_faithful *o, const double *v, double *G,
                             int threads) {
  int64_t m = o->m;
  /* find_arb!(r, v): threaded loop, per-pool heap gather, router.jl:38-42 */
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
  for (int64_t i = 0; i < m; ++i) {
    const faithful_pool *c = o->cfmms[i];
    double *vi = (double *)malloc(2 * sizeof(double)); /* v[cfmm.Ai] allocates */
    vi[0] = v[c->Ai[0] - 1];
    vi[1] = v[c->Ai[1] - 1];
    c->find_arb(o->Ds[i], o->Ls[i], c, vi);
    free(vi);
  }
  /* fn: serial acc, router.jl:79-83 */
  double acc = 0.0;
  for (int64_t i = 0; i < m; ++i) {
    const faithful_pool *c = o->cfmms[i];
    double v1 = v[c->Ai[0] - 1], v2 = v[c->Ai[1] - 1];
    acc += ((0.0 + o->Ls[i][0] * v1) + o->Ls[i][1] * v2) -
           ((0.0 + o->Ds[i][0] * v1) + o->Ds[i][1] * v2);
  }
  /* g!: serial scatter, router.jl:90, 98-100 */
  memset(G, 0, (size_t)o->n_tokens * sizeof(double));
  for (int64_t i = 0; i < m; ++i) {
    const faithful_pool *c = o->cfmms[i];
    G[c->Ai[0] - 1] += o->Ls[i][0] - o->Ds[i][0];
    G[c->Ai[1] - 1] += o->Ls[i][1] - o->Ds[i][1];
  }
  return acc;
}

void oracle_faithful_destroy(oracle_faithful *o) {
  if (!o) return;
  for (int64_t i = 0; i < o->m; ++i) {
    free(o->cfmms[i]->R);
    free(o->cfmms[i]->Ai);
    free(o->cfmms[i]->w);
    free(o->cfmms[i]);
    free(o->Ds[i]);
    free(o->Ls[i]);
  }
  free(o->cfmms);
  free(o->Ds);
  free(o->Ls);
  free(o);
}

/* ------------------------------------------------------------------------- */
/* CPU timing baseline, SoA flavour                                           */
/* ------------------------------------------------------------------------- */

double oracle_soa_sweep_product(int64_t m, const double *R, const double *gamma,
                                const int64_t *Ai, const double *v,
                                int64_t n_tokens, double *G, int threads) {
  int nt = threads > 1 ? threads : 1;
  double *priv = (double *)calloc((size_t)nt * (size_t)n_tokens, sizeof(double));
  double acc = 0.0;
#pragma omp parallel num_threads(nt) reduction(+ : acc)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    double *g = priv + (size_t)t * (size_t)n_tokens;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < m; ++i) {
      int64_t a1 = Ai[2 * i] - 1, a2 = Ai[2 * i + 1] - 1;
      double vi[2] = {v[a1], v[a2]};
      double D[2], L[2];
      oracle_product_arb(R + 2 * i, gamma[i], vi, D, L);
      acc += ((0.0 + L[0] * vi[0]) + L[1] * vi[1]) -
             ((0.0 + D[0] * vi[0]) + D[1] * vi[1]);
      g[a1] += L[0] - D[0];
      g[a2] += L[1] - D[1];
    }
  }
  for (int64_t j = 0; j < n_tokens; ++j) {
    double s = 0.0;
    for (int t = 0; t < nt; ++t) s += priv[(size_t)t * (size_t)n_tokens + j];
    G[j] = s;
  }
  free(priv);
  return acc;
}

/* Host cores the timing flavours may use.  omp_get_num_procs(), not
 * omp_get_max_threads(): launchers such as torchrun export OMP_NUM_THREADS=1,
 * which must not turn the CPU baseline into a single-thread run (every
 * parallel region here passes an explicit num_threads clause). */
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
