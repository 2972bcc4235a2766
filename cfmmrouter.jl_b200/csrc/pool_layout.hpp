// pool_layout.hpp -- host-side device layout of one pool type (pure C++, no CUDA:
// unit-tested on CPU through cfmm_debug_product_layout).
//
//  * orientation: ProductTwoCoin is exactly symmetric under exchanging its two
//    tokens, so when hub tokens are detected every pool is stored with its
//    higher-degree token first (hubs on the register-accumulated run side);
//  * order: stable counting sort by the (oriented) first token a; for the TMA
//    kernel additionally grouped by bucket(b) = b / nb, each bucket padded to
//    whole tiles (padding = position with order -1), one bucket id per tile.
#pragma once
#include <cstdint>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace cfmm {

// Stable counting sort of the items 0..m-1 by key(item) in [0, K), over `src` (the current
// order; null = identity): out[pos] = item.  Parallel form: every thread counts and then
// scatters its own contiguous block, with per-(key, thread) offsets -- stable, no atomics.
template <class KeyFn>
inline void counting_sort_stable(const int64_t* src, int64_t m, int64_t K, KeyFn key,
                                 std::vector<int64_t>& out) {
  out.assign((size_t)m, 0);
  int T = 1;
#ifdef _OPENMP
  T = omp_get_max_threads();
  const int64_t cap = (int64_t)32 << 20;  // counters in flight
  if ((int64_t)T * K > cap) T = (int)(cap / (K > 0 ? K : 1));
  if (T < 1) T = 1;
  if (m < (1 << 16)) T = 1;
#endif
  std::vector<int64_t> cnt((size_t)T * (size_t)K, 0);
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    const int64_t lo = m * t / T, hi = m * (t + 1) / T;
    int64_t* c = cnt.data() + (size_t)t * (size_t)K;
    for (int64_t p = lo; p < hi; ++p) c[key(src ? src[p] : p)]++;
#pragma omp barrier
#pragma omp single
    {
      int64_t run = 0;  // offsets in (key, thread) order
      for (int64_t k = 0; k < K; ++k)
        for (int tt = 0; tt < T; ++tt) {
          int64_t& x = cnt[(size_t)tt * (size_t)K + (size_t)k];
          const int64_t v = x;
          x = run;
          run += v;
        }
    }
    for (int64_t p = lo; p < hi; ++p) {
      const int64_t item = src ? src[p] : p;
      out[(size_t)c[key(item)]++] = item;
    }
  }
}

struct TileShape {
  int64_t tile = 0;   // padding unit in pools (one warp-chunk of the TMA kernel); 0 = no bucketing
  int64_t nbmax = 0;  // capacity of the shared ν / Ψ slices, in tokens
  int64_t nb_align = 1;  // bucket width is a multiple of this
};

struct PoolLayout {
  std::vector<int64_t> order;    // device position -> insertion index within the type, -1 = padding
  std::vector<int> oa, ob;       // device orientation per insertion index (0-based tokens)
  std::vector<uint8_t> swapped;  // per insertion index: stored with its two tokens exchanged
  std::vector<int> tile_bucket;  // bucket of every chunk (bucketed layouts only)
  int64_t m_padded = 0;
  int64_t nb = 0;                // bucket width in tokens
  bool bucketed = false;
  bool skewed = false;           // hub tokens detected
  bool used_skew_shape = false;
};

// Ai: [2m] 1-based token ids (validated by the caller).  orient: -1 auto (orient
// only when hubs are detected), 0 never, 1 always; only honoured when `symmetric`
// (ProductTwoCoin).  `normal` / `skew`: tile shapes for uniform / hub-detected graphs.
inline PoolLayout build_pool_layout(const int64_t* Ai, int64_t m, int64_t n_tokens, int orient,
                                    bool symmetric, TileShape normal, TileShape skew) {
  PoolLayout lay;
  lay.oa.resize((size_t)m);
  lay.ob.resize((size_t)m);
  lay.swapped.assign((size_t)m, 0);
  lay.m_padded = m;
  std::vector<int64_t> deg;
  if (symmetric && orient != 0) {
    deg.assign((size_t)n_tokens, 0);
#pragma omp parallel for schedule(static) if (m > (1 << 16))
    for (int64_t i = 0; i < m; ++i) {
#pragma omp atomic
      deg[(size_t)Ai[2 * i] - 1]++;
#pragma omp atomic
      deg[(size_t)Ai[2 * i + 1] - 1]++;
    }
    // hub detection: some token sits in far more pools than the average token.
    // On uniform graphs orientation only perturbs the layout (measured -2.6 %),
    // so in auto mode it is applied to skewed graphs only.
    int64_t max_deg = 0;
    for (int64_t d : deg) max_deg = d > max_deg ? d : max_deg;
    const double mean_deg = 2.0 * (double)m / (double)n_tokens;
    lay.skewed = (double)max_deg > 4.0 * mean_deg + 64.0;
    if (orient < 0 && !lay.skewed) deg.clear();
  }
#pragma omp parallel for schedule(static) if (m > (1 << 16))
  for (int64_t i = 0; i < m; ++i) {
    const int a = (int)(Ai[2 * i] - 1), b = (int)(Ai[2 * i + 1] - 1);
    const bool sw = !deg.empty() && deg[(size_t)b] > deg[(size_t)a];
    lay.swapped[(size_t)i] = sw;
    lay.oa[(size_t)i] = sw ? b : a;
    lay.ob[(size_t)i] = sw ? a : b;
  }
  // stable counting sort by the first token
  counting_sort_stable(nullptr, m, n_tokens, [&](int64_t i) { return (int64_t)lay.oa[(size_t)i]; }, lay.order);
  lay.used_skew_shape = lay.skewed && skew.tile > 0;
  const TileShape shape = lay.used_skew_shape ? skew : normal;
  if (shape.tile <= 0 || m == 0) return lay;
  // b-bucketed order: (bucket(b), a), each bucket padded to whole tiles
  const int64_t tile = shape.tile;
  const int64_t B = (n_tokens + shape.nbmax - 1) / shape.nbmax;
  int64_t nb = (n_tokens + B - 1) / B;
  if (shape.nb_align > 1) {  // round up; stays within the slice capacity when nbmax is a multiple too
    const int64_t up = (nb + shape.nb_align - 1) / shape.nb_align * shape.nb_align;
    if (up <= shape.nbmax) nb = up;
  }
  std::vector<int64_t> cnt((size_t)B + 1, 0);
  std::vector<int64_t> grouped;  // the a-sorted order, stably regrouped by bucket(b)
  counting_sort_stable(lay.order.data(), m, B, [&](int64_t i) { return (int64_t)(lay.ob[(size_t)i] / nb); },
                       grouped);
  {
    // bucket sizes from the grouped order: bucket ids are non-decreasing along it
    std::vector<int64_t> c2((size_t)B, 0);
#pragma omp parallel for schedule(static) if (m > (1 << 16))
    for (int64_t i = 0; i < m; ++i) {
#pragma omp atomic
      c2[(size_t)(lay.ob[(size_t)i] / nb)]++;
    }
    for (int64_t k = 0; k < B; ++k) cnt[(size_t)k + 1] = c2[(size_t)k];
  }
  int64_t padded = 0;
  for (int64_t k = 0; k < B; ++k) padded += (cnt[(size_t)k + 1] + tile - 1) / tile * tile;
  if (padded > 2 * m + 8 * tile) return lay;  // too sparse per bucket: a-sorted layout only
  std::vector<int64_t> start((size_t)B + 1, 0);  // padded start of each bucket
  for (int64_t k = 0; k < B; ++k)
    start[(size_t)k + 1] = start[(size_t)k] + (cnt[(size_t)k + 1] + tile - 1) / tile * tile;
  std::vector<int64_t> order((size_t)padded, -1);
  {
    // grouped[] is the padded order minus the pads: bucket k's pools go to start[k] onwards
    std::vector<int64_t> first((size_t)B + 1, 0);  // unpadded start of each bucket
    for (int64_t k = 0; k < B; ++k) first[(size_t)k + 1] = first[(size_t)k] + cnt[(size_t)k + 1];
#pragma omp parallel for schedule(static) if (m > (1 << 16))
    for (int64_t q = 0; q < m; ++q) {
      const int64_t i = grouped[(size_t)q];
      const int64_t k = lay.ob[(size_t)i] / nb;
      order[(size_t)(start[(size_t)k] + (q - first[(size_t)k]))] = i;
    }
  }
  lay.order.swap(order);
  lay.m_padded = padded;
  lay.nb = nb;
  lay.bucketed = true;
  lay.tile_bucket.resize((size_t)(padded / tile));
  for (int64_t k = 0; k < B; ++k)
    for (int64_t t = start[(size_t)k] / tile; t < start[(size_t)k + 1] / tile; ++t)
      lay.tile_bucket[(size_t)t] = (int)k;
  return lay;
}

}  // namespace cfmm
