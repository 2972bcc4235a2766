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

namespace cfmm {

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
    for (int64_t i = 0; i < m; ++i) {
      deg[(size_t)Ai[2 * i] - 1]++;
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
  for (int64_t i = 0; i < m; ++i) {
    const int a = (int)(Ai[2 * i] - 1), b = (int)(Ai[2 * i + 1] - 1);
    const bool sw = !deg.empty() && deg[(size_t)b] > deg[(size_t)a];
    lay.swapped[(size_t)i] = sw;
    lay.oa[(size_t)i] = sw ? b : a;
    lay.ob[(size_t)i] = sw ? a : b;
  }
  // stable counting sort by the first token
  {
    std::vector<int64_t> head((size_t)n_tokens + 1, 0);
    for (int64_t i = 0; i < m; ++i) head[(size_t)lay.oa[(size_t)i] + 1]++;
    for (int64_t t = 0; t < n_tokens; ++t) head[(size_t)t + 1] += head[(size_t)t];
    lay.order.assign((size_t)m, 0);
    for (int64_t i = 0; i < m; ++i) lay.order[(size_t)head[(size_t)lay.oa[(size_t)i]]++] = i;
  }
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
  for (int64_t i = 0; i < m; ++i) cnt[(size_t)(lay.ob[(size_t)i] / nb) + 1]++;
  int64_t padded = 0;
  for (int64_t k = 0; k < B; ++k) padded += (cnt[(size_t)k + 1] + tile - 1) / tile * tile;
  if (padded > 2 * m + 8 * tile) return lay;  // too sparse per bucket: a-sorted layout only
  std::vector<int64_t> start((size_t)B + 1, 0);  // padded start of each bucket
  for (int64_t k = 0; k < B; ++k)
    start[(size_t)k + 1] = start[(size_t)k] + (cnt[(size_t)k + 1] + tile - 1) / tile * tile;
  std::vector<int64_t> order((size_t)padded, -1), fill(start.begin(), start.end() - 1);
  for (int64_t p = 0; p < m; ++p) {  // stable: keeps the a-order inside a bucket
    const int64_t i = lay.order[(size_t)p];
    order[(size_t)fill[(size_t)(lay.ob[(size_t)i] / nb)]++] = i;
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

// ---- tile schedule of the TMA kernel ----------------------------------------------
// The padded device order is a sequence of chunks (one warp's share of a tile);
// chunk_bucket[c] is non-decreasing.  CTA g of `grid` owns the chunk range
// [C*g/grid, C*(g+1)/grid) -- balanced to one chunk -- and walks it in tiles of up
// to `max_chunks` consecutive chunks that never straddle a bucket boundary.
struct TileSchedule {
  std::vector<int> desc;       // 4 ints per tile: first chunk, chunk count, bucket, 0
  std::vector<int> cta_start;  // [grid + 1] tile index ranges
  int grid = 0;
};

inline TileSchedule build_tile_schedule(const std::vector<int>& chunk_bucket, int grid, int max_chunks) {
  TileSchedule ts;
  const int64_t C = (int64_t)chunk_bucket.size();
  if (grid > C) grid = (int)C;
  if (grid < 1) grid = 1;
  ts.grid = grid;
  ts.cta_start.assign((size_t)grid + 1, 0);
  for (int g = 0; g < grid; ++g) {
    const int64_t lo = C * g / grid, hi = C * (g + 1) / grid;
    int64_t c = lo;
    while (c < hi) {
      const int bk = chunk_bucket[(size_t)c];
      int64_t e = c + 1;
      while (e < hi && e - c < max_chunks && chunk_bucket[(size_t)e] == bk) ++e;
      ts.desc.push_back((int)c);
      ts.desc.push_back((int)(e - c));
      ts.desc.push_back(bk);
      ts.desc.push_back(0);
      c = e;
    }
    ts.cta_start[(size_t)g + 1] = (int)(ts.desc.size() / 4);
  }
  return ts;
}

}  // namespace cfmm
