// cfmm_capi.cu -- libcfmm_b200.so: C ABI (include/cfmm_b200.h) over the sm_100a
// sweep kernels.  Host side of the drop-in boundary: pool ingest, token sort,
// SoA upload, sweep orchestration, trade read-back, multi-GPU exchange set-up.
//
// There is deliberately no CPU fallback in this file: every compute entry
// point needs a CUDA device and fails with CFMM_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cfmm_b200.h"
#include "peer_exchange.cuh"
#include "pool_layout.hpp"
#include "sweep_kernels.cuh"
#include "product_tma.cuh"
#include "solver.cuh"

namespace {

thread_local std::string g_create_error;

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;  // owns its allocation: freed on every exit path
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return copy_in(p, h.data(), h.size() * sizeof(T));
  }
  // H2D from pageable memory, COMPLETE on return.  cudaMemcpy alone is not: for pageable sources
  // it returns once the data sits in the driver's staging buffer, the DMA may still be in flight --
  // and the contexts' streams are non-blocking, so a kernel launched on them right afterwards does
  // not wait for the legacy stream the copy ran on (seen as a rare illegal address in
  // update_reserves_kernel reading a position array that had not landed yet).
  static cudaError_t copy_in(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? cudaStreamSynchronize(cudaStreamLegacy) : e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

// One pool type's shard: host staging until finalize, device SoA afterwards.
struct PoolSet {
  int64_t m = 0;
  // host staging (insertion order)
  std::vector<double> R, gamma, w;        // 2m, m, 2m
  std::vector<int64_t> Ai;                // 2m, 1-based, validated
  std::vector<int64_t> gidx;              // m, global insertion index
  std::vector<double> cp;                 // univ3: m
  std::vector<int64_t> tick_off;          // univ3: m+1
  std::vector<double> lower, liq;         // univ3: CSR
  // after finalize
  cfmm::HVec<int64_t> order;              // sorted position -> insertion index within type
  std::vector<int64_t> pos_of;            // lazily: insertion index -> sorted position
  DevBuf<double2> d_R, d_w, d_outD, d_outL;
  DevBuf<double> d_gam, d_tickdata;
  DevBuf<double2> d_first[4];             // univ3: the current tick of every pool, four streams (arb_math.cuh)
  DevBuf<int2> d_Ai, d_tick;
  DevBuf<int64_t> d_gidx;                 // sorted position -> global insertion index
  int64_t total_ticks = 0;
  int64_t m_padded = 0;        // product: arrays padded to whole 96-pool chunks, per b-bucket
  bool in_fast_range = false;  // every R, γ in [2^-100, 2^100] and γ <= 1
  bool tma_ok = false;         // b-bucketed layout built (product only)
  int nb = 0;                  // bucket width in tokens
  int64_t n_chunks = 0;        // product: 96-pool chunks of the padded device order
  cfmm::BucketTable buckets;   // product: first chunk of every b-bucket (travels as a kernel parameter)
  // derived data of the TMA gradient kernel (product_tma.cuh): the chunk-blocked packed
  // stream (built lazily for the mode a sweep asks for) and the per-token 2^-s_b table of
  // the fixed-point Ψ[b] slice
  DevBuf<unsigned char> d_packed;
  int packed_mode = -1;        // -1 = stale; else (econ ? 1 : 0) | (fixed ? 2 : 0) | (compact ? 4 : 0)
  // compact stream: γ dictionary (<= 256 distinct fees) and the per-pool codes (device order)
  DevBuf<unsigned short> d_gcode;
  DevBuf<double> d_gtab;       // [256] 1/γ by code, then [256] γ by code
  bool compact_ok = false;
  DevBuf<double> d_inv_scale, d_tok_sum;
  bool fixed_ok = false;       // every token fits the fixed-point rules (range, totals)
  // speed-weighted CTA ranges of the TMA kernel (product_tma.cuh): the table passed to the next
  // launch, the smoothed per-CTA speed it was derived from, and the mapped pinned words the CTAs
  // report their loop durations to (tagged with the table version they ran under)
  cfmm::RangeTable ranges;
  std::vector<double> speed;
  DevBuf<unsigned> d_dur;      // per-CTA loop durations of the latest launch (device)
  unsigned* h_dur = nullptr;   // pinned landing zone of the occasional D2H copy of d_dur
  cudaEvent_t ev_dur = nullptr;
  bool dur_pending = false;    // a copy is in flight (ev_dur)
  bool balancing = false;      // the last TMA launch ran with speed feedback on
  int since_copy = 0;
  unsigned range_version = 1;
  int range_updates = 0;
  int64_t tma_launches = 0;
  cfmm::HVec<uint8_t> swapped;  // product: pool stored with its two tokens exchanged (insertion index)
  bool skewed = false;          // product: hub tokens detected at finalize
  void release() {
    d_R.release(); d_w.release(); d_outD.release(); d_outL.release();
    d_gam.release(); d_tickdata.release();
    for (auto& f : d_first) f.release();
    d_Ai.release(); d_tick.release(); d_gidx.release();
    d_packed.release(); d_inv_scale.release(); d_tok_sum.release(); d_gcode.release(); d_gtab.release();
    if (h_dur) cudaFreeHost(h_dur);
    h_dur = nullptr;
    if (ev_dur) cudaEventDestroy(ev_dur);
    ev_dur = nullptr;
    d_dur.release();
  }
};

}  // namespace

struct cfmm_ctx {
  int device = 0;
  int64_t n_tokens = 0;
  int64_t n_pools = 0;
  bool finalized = false;
  bool has_trades = false;
  PoolSet sets[3];
  cudaStream_t stream = nullptr;
  cudaStream_t last_stream = nullptr;  // stream of the last sweep (cfmm_sweep_device* may use the caller's)
  cudaEvent_t ev_order = nullptr;      // orders work across a change of stream
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DevBuf<double> d_nu;  // n
  double* h_stage = nullptr;   // pinned, n+1
  // finalize uploads: device-order arrays are gathered block by block into two pinned bounce
  // buffers and copied from there (DMA at link rate, overlapped with the next block's gather)
  void* h_bounce[2] = {nullptr, nullptr};
  cudaEvent_t ev_bounce[2] = {nullptr, nullptr};
  int sm_count = 148;
  // options
  int exact = 0;
  int debug_skip = 0;  // measurement only (tools/explore.py)
  int tma_variant = 0; // 0: b-bucketed ProductTwoCoin layout + TMA kernel; -1: a-sorted layout, first-generation kernel only (fixed at finalize)
  int use_tma = 1;     // 0: run the first-generation kernel even when the layout exists
  int orient_by_degree = -1; // ProductTwoCoin: store each pool with its higher-degree token first: -1 auto (skewed graphs only), 0 never, 1 always (fixed at finalize)
  int psi_fixed_point = 1;   // Ψ[b] partials of the TMA kernel: 1 = 64-bit fixed point on native shared atomics when the pool set allows it, 0 = fp64 CAS adds
  int sweep_events = 0;      // 1 = record ev0/ev1 around every sweep (cfmm_last_sweep_ms); disables the sweep graphs
  bool events_recorded = false;
  int geomean_log2 = 1;   // gradient-only GeometricMean sweeps: power through exp2/log2 (1, default) or pow (0)
  int gradient_math = 1;  // gradient-only ProductTwoCoin sweeps: 1 = economized (few-ulp), 0 = reference order (bit-identical per pool)
  unsigned long long epoch = 0;  // sweeps enqueued so far
  DevBuf<double> d_accum[2];     // ping-pong [Ψ; acc] accumulators, zeroed one sweep ahead in-kernel
  double* zero_pending = nullptr; // accumulator the first kernel of the current sweep must zero
  DevBuf<unsigned long long> d_grid_done;  // fused exchange: CTAs arrived, summed over all sweeps
  unsigned long long grid_done_target = 0;
  int fused_exchange = 1;         // product-only sets: run the peer exchange in the sweep kernel's tail
  int grid_waves = -1;            // first-generation kernel: waves of CTAs (-1 = 1; 0 = one CTA per 512 pools; see launch_sweep)
  int exchange_protocol = 0;      // 0 = by world size, 1 = LL one-shot, 2 = LL two-shot, 3 = direct 8-byte push (peer_exchange.cuh)
  int coop_launch = 0;            // fused exchange: launch the sweep kernel cooperatively (measured at N = 2: +4 to +8 us per step)
  int exchange_bypass = 0;        // 1 = sweeps return this rank's partial [Ψ; acc] (no exchange); every rank must agree
  cfmm::FusedExchange fx_pending; // set by enqueue_sweep when the next TMA launch must carry the exchange
  int blocks_per_sm = 0;  // 0 = occupancy-derived
  DevBuf<unsigned long long> d_trace;  // option "trace": per-CTA phase timestamps of the last TMA sweep
  int trace_grid = 0;
  // cfmm_sweep as ONE graph launch {H2D ν, sweep kernels, D2H [Ψ; acc]} per (accumulator
  // parity, counter-set parity): captured the second time the same pinned host buffers are
  // passed with unchanged options; any option / reserve / comm change invalidates (version)
  struct SweepGraph {
    cudaGraphExec_t exec = nullptr;
    const double* v = nullptr;
    double* psi = nullptr;
    unsigned long long version = 0;
    const double* seen_v = nullptr;  // key of the last eager call (capture on the next match)
    double* seen_psi = nullptr;
    unsigned long long seen_version = ~0ull;
    int64_t launches = 0;            // kernel launches one replay stands for
  } graphs[2][4];
  unsigned long long state_version = 1;
  int use_graphs = 1;
  bool capturing = false;  // cfmm_sweep is recording a graph: no event queries / side copies
  bool calibrating = false;  // cfmm_finalize is running its calibration sweeps (range table feedback every launch)
  const void* pinned_ok[2] = {nullptr, nullptr};  // host pointers already verified as pinned
  int balance = 1;              // 1 = TMA kernel: CTA ranges sized by measured CTA speed (feedback), 0 = even split
  int geomean_tma = 1;          // gradient-only GeometricMean sweeps on the TMA kernel (0: first-generation kernel)
  int compact_stream = 1;       // ProductTwoCoin, economized math: 24-byte pool records (γ dictionary) when the set allows it
  // resident CTAs per SM of every kernel instantiation this context has launched.
  // Per context, not per process: cudaFuncSetAttribute (the > 48 KB dynamic shared
  // memory opt-in) acts on the current device only, and contexts of one process
  // may sit on different devices and be driven from different host threads.
  std::unordered_map<const void*, int> occupancy;
  int64_t launches = 0;
  std::string err;
  cfmm::PeerExchange comm;
  // optional per-kernel timing (option "profile"): event pairs per launch
  struct Prof {
    std::vector<cudaEvent_t> ev;  // 2 per recorded launch
    std::vector<int> type;        // pool type (3 = peer exchange)
    size_t used = 0;              // launches recorded
  } prof;
};

namespace {

int fail(cfmm_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define CU_TRY(ctx, expr)                                                      \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess)                                                     \
      return fail((ctx), CFMM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,        \
                  cudaGetErrorString(_e), __FILE__, __LINE__);                 \
  } while (0)

int check_common(cfmm_ctx* ctx, int64_t m, const double* R, const double* gamma,
                 const int64_t* Ai) {
  if (!ctx) return CFMM_ERR_INVALID;
  if (ctx->finalized)
    return fail(ctx, CFMM_ERR_STATE, "pools cannot be added after cfmm_finalize");
  if (m < 0) return fail(ctx, CFMM_ERR_INVALID, "negative pool count");
  if (m > 0 && (!gamma || !Ai || !R))
    return fail(ctx, CFMM_ERR_INVALID, "null array argument");
  // first offending pool, if any (parallel scan; the error names the lowest index like a serial one)
  int64_t bad = m;
  const int64_t nt = ctx->n_tokens;
#pragma omp parallel for schedule(static) reduction(min : bad) if (m > (1 << 16))
  for (int64_t i = 0; i < m; ++i) {
    const int64_t a = Ai[2 * i], b = Ai[2 * i + 1];
    if (a < 1 || a > nt || b < 1 || b > nt || a == b) bad = i < bad ? i : bad;
  }
  for (int64_t i = bad; i < m && i == bad; ++i) {
    const int64_t a = Ai[2 * i], b = Ai[2 * i + 1];
    if (a < 1 || a > ctx->n_tokens || b < 1 || b > ctx->n_tokens)
      return fail(ctx, CFMM_ERR_INVALID,
                  "pool %lld: token index (%lld, %lld) outside 1..%lld",
                  (long long)i, (long long)a, (long long)b,
                  (long long)ctx->n_tokens);
    if (a == b)
      return fail(ctx, CFMM_ERR_INVALID,
                  "pool %lld: Ai[1] == Ai[2] == %lld (a two-coin pool needs two "
                  "distinct tokens)",
                  (long long)i, (long long)a);
  }
  return CFMM_OK;
}

void append_common(cfmm_ctx* ctx, PoolSet& s, int64_t m, const double* R,
                   const double* gamma, const int64_t* Ai) {
  if (R) s.R.insert(s.R.end(), R, R + 2 * m);
  s.gamma.insert(s.gamma.end(), gamma, gamma + m);
  s.Ai.insert(s.Ai.end(), Ai, Ai + 2 * m);
  {
    const size_t old = s.gidx.size();
    s.gidx.resize(old + (size_t)m);
    int64_t* gi = s.gidx.data() + old;
    const int64_t base = ctx->n_pools;
#pragma omp parallel for schedule(static) if (m > (1 << 16))
    for (int64_t i = 0; i < m; ++i) gi[i] = base + i;
  }
  s.m += m;
  ctx->n_pools += m;
}

inline bool fast_range_ok(double v) { return v >= cfmm::kFastLo && v <= cfmm::kFastHi; }

// layout of one pool type (pool_layout.hpp): ProductTwoCoin gets the b-bucketed,
// chunk-padded layout of the TMA kernel; the other types are a-sorted only.
cfmm::PoolLayout layout_for(const cfmm_ctx* ctx, int type, const int64_t* Ai, int64_t m,
                            void (*mark)(const char*) = nullptr) {
  const bool product = type == CFMM_POOL_PRODUCT;
  cfmm::TileShape shape;
  if ((product || type == CFMM_POOL_GEOMEAN) && ctx->tma_variant >= 0) {
    shape.tile = cfmm::kTmaChunk;
    shape.nbmax = cfmm::kTmaNbMax;
  }
  return cfmm::build_pool_layout(Ai, m, ctx->n_tokens, ctx->orient_by_degree, product, shape, shape, mark);
}

int refresh_scale(cfmm_ctx* ctx, PoolSet& s);

// finalize timing (environment CFMM_TIMING): phase marks on stderr
struct PhaseClock {
  bool on = getenv("CFMM_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[cfmm]   %-36s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};
PhaseClock* g_layout_clock = nullptr;  // (finalize is serialised per process while timing is on)
void layout_mark(const char* what) {
  if (g_layout_clock) g_layout_clock->mark(what);
}

constexpr size_t kBounceBytes = (size_t)16 << 20;

// dst[p] = fill(p) for p in [0, count): gathered in parallel into the pinned bounce buffers,
// block by block, each block copied asynchronously on the context stream while the next one is
// gathered.  Returns with copies possibly still in flight (upload_set synchronises once).
template <class T, class Fill>
cudaError_t upload_streamed(cfmm_ctx* ctx, DevBuf<T>& dst, int64_t count, Fill fill) {
  cudaError_t e = dst.alloc((size_t)count);
  if (e != cudaSuccess || count == 0) return e;
  for (int k = 0; k < 2; ++k) {
    if (!ctx->h_bounce[k] && (e = cudaMallocHost(&ctx->h_bounce[k], kBounceBytes)) != cudaSuccess) return e;
    if (!ctx->ev_bounce[k] &&
        (e = cudaEventCreateWithFlags(&ctx->ev_bounce[k], cudaEventDisableTiming)) != cudaSuccess)
      return e;
  }
  const int64_t per = (int64_t)(kBounceBytes / sizeof(T));
  int k = 0;
  for (int64_t lo = 0; lo < count; lo += per, k ^= 1) {
    const int64_t hi = std::min(count, lo + per);
    if ((e = cudaEventSynchronize(ctx->ev_bounce[k])) != cudaSuccess) return e;
    T* out = (T*)ctx->h_bounce[k];
#pragma omp parallel for schedule(static) if (hi - lo > (1 << 14))
    for (int64_t p = lo; p < hi; ++p) out[p - lo] = fill(p);
    if ((e = cudaMemcpyAsync(dst.p + lo, out, (size_t)(hi - lo) * sizeof(T), cudaMemcpyHostToDevice,
                             ctx->stream)) != cudaSuccess)
      return e;
    if ((e = cudaEventRecord(ctx->ev_bounce[k], ctx->stream)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

int upload_set(cfmm_ctx* ctx, int type) {
  PoolSet& s = ctx->sets[type];
  if (s.m == 0) return CFMM_OK;
  const int64_t m = s.m;
  PhaseClock clock;
  g_layout_clock = clock.on ? &clock : nullptr;
  cfmm::PoolLayout lay = layout_for(ctx, type, s.Ai.data(), m, clock.on ? layout_mark : nullptr);
  g_layout_clock = nullptr;
  const cfmm::HVec<int>&oa = lay.oa, &ob = lay.ob;
  s.swapped.swap(lay.swapped);
  s.skewed = lay.skewed;
  s.order.swap(lay.order);
  s.m_padded = lay.m_padded;
  s.tma_ok = lay.bucketed;
  s.nb = (int)lay.nb;
  if (s.tma_ok) {
    // bucket boundaries in chunks, for the kernel-parameter table
    const int B = lay.tile_bucket.empty() ? 0 : lay.tile_bucket.back() + 1;
    if (B > cfmm::kTmaMaxBuckets) {
      s.tma_ok = false;  // more b-buckets than the parameter table holds: first-generation kernel
    } else {
      s.n_chunks = (int64_t)lay.tile_bucket.size();
      s.buckets.n_buckets = B;
      int b = 0;
      s.buckets.first_chunk[0] = 0;
      for (int64_t c = 0; c < s.n_chunks; ++c)
        while (b < lay.tile_bucket[(size_t)c]) s.buckets.first_chunk[++b] = (int)c;
      while (b < B) s.buckets.first_chunk[++b] = (int)s.n_chunks;
    }
  }
  const int64_t mp = s.m_padded;
  const int64_t* order = s.order.data();
  // padding pools trail their bucket (fewer than one chunk of them): keyed like the last real
  // pool before them (monotone a), zero reserves, unit fee
  const auto real_before = [order](int64_t p) -> int64_t {
    while (p >= 0 && order[p] < 0) --p;
    return p < 0 ? -1 : order[p];
  };
  CU_TRY(ctx, upload_streamed(ctx, s.d_gam, mp, [&](int64_t p) {
    const int64_t i = order[p];
    return i < 0 ? 1.0 : s.gamma[(size_t)i];
  }));
  CU_TRY(ctx, upload_streamed(ctx, s.d_Ai, mp, [&](int64_t p) {
    const int64_t i = real_before(p);
    return i < 0 ? make_int2(0, 1) : make_int2(oa[(size_t)i], ob[(size_t)i]);
  }));
  // bit 62 of the global index marks a pool stored with its tokens exchanged
  CU_TRY(ctx, upload_streamed(ctx, s.d_gidx, mp, [&](int64_t p) {
    const int64_t i = order[p];
    return i < 0 ? (int64_t)-1 : (s.gidx[(size_t)i] | (s.swapped[(size_t)i] ? (1ll << 62) : 0));
  }));
  clock.mark("gather + upload gamma, Ai, gidx");
  s.compact_ok = false;
  if (type == CFMM_POOL_PRODUCT && s.tma_ok) {
    // γ dictionary of the compact stream: fees are categorical in practice.  Pass 1 (serial,
    // cheap): the distinct values, through a 1024-slot open-addressing table on the bit pattern;
    // pass 2 (the upload's gather): the codes
    std::vector<double> vals;
    constexpr int kSlots = 1024;
    std::vector<long long> key(kSlots, -1);
    std::vector<int> val(kSlots, 0);
    auto slot_of = [&](double g, bool insert) -> int {
      long long bits;
      memcpy(&bits, &g, sizeof(bits));
      if (bits == -1) return -1;  // (a NaN pattern: no dictionary)
      unsigned h = (unsigned)((unsigned long long)bits * 0x9E3779B97F4A7C15ull >> 54);
      for (;;) {
        if (key[h] == bits) return val[h];
        if (key[h] == -1) {
          if (!insert || vals.size() == (size_t)cfmm::kTmaGammaCodes) return -1;
          key[h] = bits;
          val[h] = (int)vals.size();
          vals.push_back(g);
          return val[h];
        }
        h = (h + 1) & (kSlots - 1);
      }
    };
    bool ok = slot_of(1.0, true) >= 0;  // (the padding pools' fee)
    double last = 1.0;
    for (int64_t i = 0; i < m && ok; ++i) {
      const double g = s.gamma[(size_t)i];
      if (g != last) {
        ok = (g == g) && slot_of(g, true) >= 0;
        last = g;
      }
    }
    if (ok) {
      std::vector<double> tab(2 * cfmm::kTmaGammaCodes, 1.0);
      for (size_t k = 0; k < vals.size(); ++k) {
        volatile double inv = 1.0 / vals[k];  // IEEE division, as inv of the 32-byte stream (pack_chunks_kernel)
        tab[k] = inv;
        tab[cfmm::kTmaGammaCodes + k] = vals[k];
      }
      CU_TRY(ctx, upload_streamed(ctx, s.d_gcode, mp, [&](int64_t p) {
        const int64_t i = order[p];
        return (unsigned short)slot_of(i < 0 ? 1.0 : s.gamma[(size_t)i], false);
      }));
      CU_TRY(ctx, s.d_gtab.upload(tab));
      s.compact_ok = true;
    }
    clock.mark("fee dictionary + codes");
  }
  if (type != CFMM_POOL_UNIV3) {
    int bad_range = 0;
#pragma omp parallel for schedule(static) reduction(| : bad_range) if (m > (1 << 16))
    for (int64_t i = 0; i < m; ++i) {
      const bool ok = fast_range_ok(s.R[2 * i]) && fast_range_ok(s.R[2 * i + 1]) &&
                      fast_range_ok(s.gamma[(size_t)i]) && s.gamma[(size_t)i] <= 1.0;
      bad_range |= ok ? 0 : 1;
    }
    s.in_fast_range = bad_range == 0;
    CU_TRY(ctx, upload_streamed(ctx, s.d_R, mp, [&](int64_t p) {
      const int64_t i = order[p];
      if (i < 0) return make_double2(0.0, 0.0);
      return s.swapped[(size_t)i] ? make_double2(s.R[2 * i + 1], s.R[2 * i])
                                  : make_double2(s.R[2 * i], s.R[2 * i + 1]);
    }));
    clock.mark("range check, gather + upload R");
    if (s.tma_ok) {
      int rc = refresh_scale(ctx, s);
      if (rc != CFMM_OK) return rc;
      clock.mark("scale table (device)");
    }
  }
  if (type == CFMM_POOL_GEOMEAN) {
    CU_TRY(ctx, upload_streamed(ctx, s.d_w, mp, [&](int64_t p) {
      const int64_t i = order[p];  // (padding pools: any valid weights)
      return i < 0 ? make_double2(0.5, 0.5) : make_double2(s.w[2 * i], s.w[2 * i + 1]);
    }));
  }
  if (type == CFMM_POOL_UNIV3) {
    // compute_at_tick (src/cfmms.jl:294-313) for every tick, once, on the host:
    // IEEE sqrt / div / mul / sub in the reference's order (see arb_math.cuh)
    std::vector<double> td;
    std::vector<double2> first[4];
    for (auto& f : first) f.assign((size_t)m, make_double2(0.0, 0.0));
    std::vector<int2> tick((size_t)m);
    td.assign(s.lower.size() * cfmm::kTickStride, 0.0);
    int64_t n_ticks_total = 0;
    for (int64_t p = 0; p < m; ++p) {
      const int64_t i = s.order[(size_t)p];
      const int64_t b = s.tick_off[(size_t)i], e = s.tick_off[(size_t)i + 1];
      const double price = s.cp[(size_t)i];
      first[1][(size_t)p].y = price;
      // current_tick = searchsortedlast(lower_ticks, current_price; rev=true)
      // (src/cfmms.jl:235): number of leading ticks >= current_price
      int cur = 0;
      while (b + cur < e && s.lower[(size_t)(b + cur)] >= price) ++cur;
      tick[(size_t)p] = make_int2((int)n_ticks_total, cur);
      for (int64_t q = b; q < e; ++q) {
        const int idx = (int)(q - b) + 1;  // 1-based
        volatile double k = s.liq[(size_t)q];
        volatile double pplus = s.lower[(size_t)q];                      // tick_high_price :252
        volatile double pminus = (q + 1 < e) ? s.lower[(size_t)q + 1] : 0.0;  // tick_low_price :255-259
        volatile double alpha = std::sqrt(k / pplus);
        volatile double beta = std::sqrt(k * pminus);
        volatile double pp = idx > cur ? pplus : (idx < cur ? pminus : price);
        volatile double R1 = std::sqrt(k / pp) - alpha;
        volatile double R2 = std::sqrt(k * pp) - beta;
        volatile double ra = R1 + alpha;
        volatile double rb = R2 + beta;
        volatile double dmax_up = k / beta - ra;
        volatile double dmax_dn = k / alpha - rb;
        // two direction blocks per pool, one 32-byte record per tick in each (arb_math.cuh)
        const size_t base = (size_t)n_ticks_total * cfmm::kTickStride, nt = (size_t)(e - b), ti = (size_t)(q - b);
        const double up_rec[4] = {k, ra, dmax_up, R2}, dn_rec[4] = {k, rb, dmax_dn, R1};
        for (int c = 0; c < 4; ++c) {
          td[base + ti * 4 + c] = up_rec[c];
          td[base + nt * 4 + ti * 4 + c] = dn_rec[c];
        }
        if (idx == cur) {  // the tick a walk starts in: also per pool, in pool order
          first[0][(size_t)p] = make_double2(k, ra);
          first[1][(size_t)p].x = rb;
          first[2][(size_t)p] = make_double2(dmax_up, R2);
          first[3][(size_t)p] = make_double2(dmax_dn, R1);
        }
      }
      n_ticks_total += e - b;
    }
    s.total_ticks = n_ticks_total;
    for (int f = 0; f < 4; ++f) CU_TRY(ctx, s.d_first[f].upload(first[f]));
    CU_TRY(ctx, s.d_tick.upload(tick));
    CU_TRY(ctx, s.d_tickdata.upload(td));
  }
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // the bounce copies read the staging below
  clock.mark("other arrays, drain");
  // host staging is no longer needed (order is kept for update_reserves)
  std::vector<double>().swap(s.R);
  std::vector<double>().swap(s.gamma);
  std::vector<double>().swap(s.w);
  std::vector<int64_t>().swap(s.Ai);
  std::vector<int64_t>().swap(s.gidx);
  std::vector<double>().swap(s.cp);
  std::vector<int64_t>().swap(s.tick_off);
  std::vector<double>().swap(s.lower);
  std::vector<double>().swap(s.liq);
  clock.mark("free host staging");
  return CFMM_OK;
}

// Sweeps may be enqueued on a caller-supplied stream (cfmm_sweep_device*), everything else runs
// on the context's own stream; the ping-pong accumulators, the trade buffers and the reserves
// make consecutive operations dependent.  Whenever the stream changes, the new one waits for
// the work enqueued on the previous one.
int use_stream(cfmm_ctx* ctx, cudaStream_t st) {
  if (ctx->last_stream && ctx->last_stream != st) {
    cudaError_t e = cudaEventRecord(ctx->ev_order, ctx->last_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ctx->ev_order, 0);
    if (e != cudaSuccess)
      return fail(ctx, CFMM_ERR_CUDA, "stream ordering failed: %s", cudaGetErrorString(e));
  }
  ctx->last_stream = st;
  return CFMM_OK;
}

// the first kernel of a sweep zeroes the other ping-pong accumulator
inline double* take_zero_pending(cfmm_ctx* ctx) {
  double* p = ctx->zero_pending;
  ctx->zero_pending = nullptr;
  return p;
}

// profiling: bracket one launch with events on its stream
struct ProfScope {
  cfmm_ctx* ctx;
  cudaStream_t st;
  bool on;
  ProfScope(cfmm_ctx* c, int type, cudaStream_t s) : ctx(c), st(s) {
    on = c->prof.used < c->prof.type.size();
    if (on) {
      c->prof.type[c->prof.used] = type;
      cudaEventRecord(c->prof.ev[2 * c->prof.used], st);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(ctx->prof.ev[2 * ctx->prof.used + 1], st);
      ctx->prof.used++;
    }
  }
};

template <class P>
int launch_sweep(cfmm_ctx* ctx, int ptype, const P& pools, PoolSet& s,
                 const double* d_v, double* d_psi, bool mat, cudaStream_t st) {
  constexpr int U = 2;
  const int64_t per_block = (int64_t)cfmm::kSweepThreads * U;
  const int64_t m_all = s.m_padded;  // includes the zero-trade padding pools, if any
  int64_t blocks = (m_all + per_block - 1) / per_block;
  // persistent-style grid: one wave of resident CTAs (148 SMs x occupancy)
  int& occ = ctx->occupancy[mat ? reinterpret_cast<const void*>(&cfmm::sweep_kernel<P, true, U>)
                                 : reinterpret_cast<const void*>(&cfmm::sweep_kernel<P, false, U>)];
  if (occ == 0) {
    if (mat)
      CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                      &occ, cfmm::sweep_kernel<P, true, U>, cfmm::kSweepThreads, 0));
    else
      CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                      &occ, cfmm::sweep_kernel<P, false, U>, cfmm::kSweepThreads, 0));
    if (occ < 1) occ = 1;
  }
  const int per_sm = ctx->blocks_per_sm > 0 && ctx->blocks_per_sm < occ ? ctx->blocks_per_sm : occ;
  // grid_waves: 1 (default) = one wave of resident CTAs striding over the pools; 0 = one CTA per
  // 512 pools, handed out by the hardware block scheduler (measured on UniV3, whose tick walks
  // make the work per pool uneven: 47.0 us against 45.2 us for the strided wave -- kept as a knob)
  const int waves = ctx->grid_waves >= 0 ? ctx->grid_waves : 1;
  const int64_t cap = (int64_t)ctx->sm_count * per_sm * (waves > 0 ? waves : 1);
  if (waves > 0 && blocks > cap) blocks = cap;
  if (mat && s.d_outD.n != (size_t)m_all) {
    CU_TRY(ctx, s.d_outD.alloc((size_t)m_all));
    CU_TRY(ctx, s.d_outL.alloc((size_t)m_all));
  }
  ProfScope prof(ctx, ptype, st);
  if (mat) {
    cfmm::sweep_kernel<P, true, U><<<(unsigned)blocks, cfmm::kSweepThreads, 0, st>>>(
        pools, d_v, d_psi, (int)ctx->n_tokens, s.d_outD.p, s.d_outL.p, m_all,
        ctx->exact | (ctx->debug_skip << 1), take_zero_pending(ctx));
  } else {
    cfmm::sweep_kernel<P, false, U><<<(unsigned)blocks, cfmm::kSweepThreads, 0, st>>>(
        pools, d_v, d_psi, (int)ctx->n_tokens, nullptr, nullptr, m_all,
        ctx->exact | (ctx->debug_skip << 1) | (ctx->gradient_math ? 16 : 0), take_zero_pending(ctx));
  }
  ctx->launches++;
  CU_TRY(ctx, cudaGetLastError());
  return CFMM_OK;
}

// (Re)compute the per-token scale table of the fixed-point slice and whether the pool set
// qualifies for it (three small kernels, off the hot path: finalize and every reserve
// mutation).  Marks the packed stream stale.
int refresh_scale(cfmm_ctx* ctx, PoolSet& s) {
  s.fixed_ok = false;
  s.packed_mode = -1;
  if (!s.tma_ok || s.m_padded == 0) return CFMM_OK;
  const int64_t mp = s.m_padded;
  const int n = (int)ctx->n_tokens;
  cudaStream_t st = ctx->stream;
  const int threads = 256;
  const unsigned pblocks = (unsigned)((mp + threads - 1) / threads);
  if (s.d_inv_scale.n != (size_t)n) CU_TRY(ctx, s.d_inv_scale.alloc((size_t)n));
  if (s.d_tok_sum.n != (size_t)n + 2) CU_TRY(ctx, s.d_tok_sum.alloc((size_t)n + 2));  // + 2 flag words
  CU_TRY(ctx, cudaMemsetAsync(s.d_tok_sum.p, 0, ((size_t)n + 2) * sizeof(double), st));
  int* d_flags = reinterpret_cast<int*>(s.d_tok_sum.p + n);
  cfmm::token_reserve_sum_kernel<<<pblocks, threads, 0, st>>>(s.d_R.p, s.d_Ai.p, mp, s.d_tok_sum.p);
  cfmm::token_scale_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(
      s.d_tok_sum.p, n, s.d_inv_scale.p, d_flags);
  cfmm::scale_check_kernel<<<pblocks, threads, 0, st>>>(s.d_R.p, s.d_Ai.p, mp, s.d_tok_sum.p,
                                                      s.d_inv_scale.p, d_flags);
  ctx->launches += 3;
  int h_flags[2] = {0, 0};
  CU_TRY(ctx, cudaMemcpyAsync(h_flags, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  CU_TRY(ctx, cudaGetLastError());
  s.fixed_ok = h_flags[0] == 0 && h_flags[1] == 0;
  return CFMM_OK;
}

// the packed stream of the mode this sweep runs in (rebuilt when the mode or the reserves changed)
template <int POOL>
int ensure_packed(cfmm_ctx* ctx, PoolSet& s, bool econ, bool fixed, bool compact, cudaStream_t st) {
  const int mode = (econ ? 1 : 0) | (fixed ? 2 : 0) | (compact ? 4 : 0);
  if (s.packed_mode == mode) return CFMM_OK;
  const size_t bytes = (size_t)s.n_chunks * (compact ? cfmm::kTmaChunk * cfmm::kTmaCompactPoolBytes : cfmm::tma_chunk_bytes<POOL>());
  if (s.d_packed.n < bytes) CU_TRY(ctx, s.d_packed.alloc((size_t)s.n_chunks * cfmm::tma_chunk_bytes<POOL>()));
  const int threads = 256;
  if (compact) {
    cfmm::pack_chunks_compact_kernel<<<(unsigned)((s.m_padded + threads - 1) / threads), threads, 0, st>>>(
        s.d_R.p, s.d_Ai.p, s.d_gcode.p, s.m_padded, s.nb, fixed ? s.d_inv_scale.p : nullptr, s.d_packed.p);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    s.packed_mode = mode;
    return CFMM_OK;
  }
  // ProductTwoCoin's economized form streams 1/γ; GeometricMean always γ
  cfmm::pack_chunks_kernel<<<(unsigned)((s.m_padded + threads - 1) / threads), threads, 0, st>>>(
      s.d_R.p, s.d_gam.p, s.d_Ai.p, POOL == 1 ? s.d_w.p : nullptr, s.m_padded,
      fixed ? s.d_inv_scale.p : nullptr, (POOL == 0 && econ) ? 1 : 0, cfmm::tma_chunk_bytes<POOL>(), s.d_packed.p);
  ctx->launches++;
  CU_TRY(ctx, cudaGetLastError());
  s.packed_mode = mode;
  return CFMM_OK;
}

template <int POOL, bool ECON, bool SKEW, bool FIXED, bool COMPACT = false>
int launch_tma_cfg(cfmm_ctx* ctx, PoolSet& s, const double* d_v, double* d_psi, cudaStream_t st) {
  auto kern = cfmm::product_sweep_tma<POOL, ECON, SKEW, FIXED, COMPACT>;
  constexpr int kThreads = cfmm::tma_threads<POOL>(), kSmem = cfmm::tma_smem_bytes_c<POOL, COMPACT>();
  int& occ = ctx->occupancy[reinterpret_cast<const void*>(kern)];
  if (occ == 0) {
    CU_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, kSmem));
    if (occ < 1) return fail(ctx, CFMM_ERR_CUDA, "product_sweep_tma does not fit on an SM");
  }
  int rc = ensure_packed<POOL>(ctx, s, ECON, FIXED, COMPACT, st);
  if (rc != CFMM_OK) return rc;
  // (grid and range table)
  const int per_sm = ctx->blocks_per_sm > 0 && ctx->blocks_per_sm < occ ? ctx->blocks_per_sm : occ;
  int grid = ctx->sm_count * per_sm;
  if (grid > s.n_chunks) grid = (int)s.n_chunks;
  // ---- speed-weighted ranges (see product_tma.cuh) ---------------------------------------
  unsigned* d_dur = nullptr;
  const bool tabled = grid <= cfmm::kTmaMaxRanges;
  const bool balancing = ctx->balance && tabled && s.n_chunks >= (int64_t)grid * 32;
  auto set_buckets = [&]() {  // b-bucket of every range's first chunk
    int b = 0;
    for (int g = 0; g <= grid; ++g) {
      const int c = g < grid ? s.ranges.first[g] : (int)s.n_chunks - 1;
      while (b + 1 < s.buckets.n_buckets && s.buckets.first_chunk[b + 1] <= c) ++b;
      s.ranges.bucket[g] = (short)b;
    }
  };
  if (!tabled) {
    s.ranges.n = 0;
  } else if (s.ranges.n != grid || !balancing) {
    if (s.ranges.n != grid || s.range_updates != 0) {  // (first launch with this grid, or balancing switched off)
      s.ranges.n = grid;
      for (int g = 0; g <= grid; ++g) s.ranges.first[g] = (int)(s.n_chunks * g / grid);
      set_buckets();
      s.speed.assign((size_t)grid, 1.0);
      s.range_version = (s.range_version % 250) + 1;
      s.range_updates = 0;
      s.dur_pending = false;
    }
  }
  if (balancing) {
    if (!s.h_dur) {
      CU_TRY(ctx, cudaMallocHost((void**)&s.h_dur, (cfmm::kTmaMaxRanges + 1) * sizeof(unsigned)));
      memset(s.h_dur, 0, (cfmm::kTmaMaxRanges + 1) * sizeof(unsigned));
      CU_TRY(ctx, s.d_dur.alloc(cfmm::kTmaMaxRanges + 1));
      CU_TRY(ctx, cudaMemset(s.d_dur.p, 0, (cfmm::kTmaMaxRanges + 1) * sizeof(unsigned)));
      CU_TRY(ctx, cudaStreamSynchronize(cudaStreamLegacy));  // (memsets run on the legacy stream; ours do not wait for it)
      CU_TRY(ctx, cudaEventCreateWithFlags(&s.ev_dur, cudaEventDisableTiming));
    }
    d_dur = s.d_dur.p;
    // a copy of the durations has landed: were they all measured under the CURRENT table?
    if (s.dur_pending && !ctx->capturing && cudaEventQuery(s.ev_dur) == cudaSuccess) {
      s.dur_pending = false;
      bool complete = true;
      const unsigned* hd = s.h_dur;
      for (int g = 0; g < grid && complete; ++g) complete = (hd[g] >> 24) == s.range_version && (hd[g] & 0xffffffu) != 0;
      if (complete) {
        double total = 0.0;
        for (int g = 0; g < grid; ++g) {
          const double len = (double)(s.ranges.first[g + 1] - s.ranges.first[g]);
          const double dur = (double)(hd[g] & 0xffffffu);
          const double rel = len / dur;  // chunks per tick
          const double keep = ctx->calibrating ? 0.5 : 0.8;
          s.speed[(size_t)g] = s.range_updates == 0 ? rel : keep * s.speed[(size_t)g] + (1.0 - keep) * rel;
          total += s.speed[(size_t)g];
        }
        // only the persistent part of the speed differences is worth following (SM position on the
        // die); launch-to-launch noise is as large: heavy smoothing, and lengths within +-15 % of even
        {
          const double mean = total / grid;
          total = 0.0;
          for (int g = 0; g < grid; ++g) {
            double& sp = s.speed[(size_t)g];
            sp = std::min(1.15 * mean, std::max(0.85 * mean, sp));
            total += sp;
          }
        }
        // new boundaries: lengths proportional to speed, exact total
        double acc_len = 0.0;
        int prev = 0;
        for (int g = 0; g < grid; ++g) {
          acc_len += (double)s.n_chunks * s.speed[(size_t)g] / total;
          int end = g + 1 == grid ? (int)s.n_chunks : (int)(acc_len + 0.5);
          const int min_end = prev + 1, max_end = (int)s.n_chunks - (grid - 1 - g);
          end = end < min_end ? min_end : (end > max_end ? max_end : end);
          s.ranges.first[g + 1] = end;
          prev = end;
        }
        set_buckets();
        s.range_version = (s.range_version % 250) + 1;
        s.range_updates++;
      }
    }
  }
  s.ranges.version = s.range_version;
  s.balancing = balancing;
  s.tma_launches++;
  cfmm::FusedExchange fx = ctx->fx_pending;
  if (fx.mode != 0) {
    fx.target = ctx->grid_done_target + (unsigned long long)grid;
    fx.grid_done = ctx->d_grid_done.p;
    ctx->fx_pending.mode = 0;  // consumed
  }
  ProfScope prof(ctx, POOL == 0 ? CFMM_POOL_PRODUCT : CFMM_POOL_GEOMEAN, st);
  const unsigned char* a_packed = s.d_packed.p;
  const double* a_gam = COMPACT ? s.d_gtab.p : s.d_gam.p;  // compact stream: the γ dictionary instead of per-pool γ
  int a_nb = s.nb, a_n = (int)ctx->n_tokens, a_range = s.in_fast_range ? 1 : 0, a_flags = ctx->exact;
  const double* a_scale = FIXED ? s.d_inv_scale.p : nullptr;
  double* a_zero = take_zero_pending(ctx);
  unsigned long long* a_trace = ctx->d_trace.n ? ctx->d_trace.p : nullptr;
  if (fx.mode != 0 && ctx->coop_launch) {
    // the fused exchange meets at a grid-wide barrier: a COOPERATIVE launch makes the driver
    // guarantee that every CTA is resident (or fail the launch) instead of inferring it
    void* args[] = {&a_packed, &a_gam, &s.buckets, &a_nb, &d_v, &a_scale, &d_psi, &a_n, &a_zero,
                    &a_range, &a_flags, &fx, &s.ranges, &d_dur, &a_trace};
    CU_TRY(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(kThreads),
                                            args, kSmem, st));
  } else {
    kern<<<grid, kThreads, kSmem, st>>>(a_packed, a_gam, s.buckets, a_nb, d_v, a_scale, d_psi, a_n, a_zero,
                                        a_range, a_flags, fx, s.ranges, d_dur, a_trace);
  }
  if (ctx->d_trace.n) ctx->trace_grid = grid;
  ctx->launches++;
  CU_TRY(ctx, cudaGetLastError());
  // the grid-barrier target moves only once the launch is known to be accepted
  if (fx.mode != 0) ctx->grid_done_target = fx.target;
  // fetch the durations now and then: often while the table is still settling, rarely afterwards
  if (balancing && !ctx->capturing && !s.dur_pending &&
      ++s.since_copy >= (ctx->calibrating ? 1 : (s.range_updates < 8 ? 4 : 256))) {
    s.since_copy = 0;
    CU_TRY(ctx, cudaMemcpyAsync(s.h_dur, s.d_dur.p, (size_t)grid * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaEventRecord(s.ev_dur, st));
    s.dur_pending = true;
  }
  return CFMM_OK;
}

template <int POOL>
int launch_tma(cfmm_ctx* ctx, PoolSet& s, const double* d_v, double* d_psi, cudaStream_t st) {
  const bool econ = ctx->gradient_math != 0;
  const bool fixed = ctx->psi_fixed_point && s.fixed_ok;
  if constexpr (POOL == 0) {
    if (econ && ctx->compact_stream && s.compact_ok && !ctx->exact) {
      if (!s.skewed && fixed) return launch_tma_cfg<0, true, false, true, true>(ctx, s, d_v, d_psi, st);
      if (!s.skewed && !fixed) return launch_tma_cfg<0, true, false, false, true>(ctx, s, d_v, d_psi, st);
      if (s.skewed && fixed) return launch_tma_cfg<0, true, true, true, true>(ctx, s, d_v, d_psi, st);
      return launch_tma_cfg<0, true, true, false, true>(ctx, s, d_v, d_psi, st);
    }
  }
#define CFMM_TMA_CASE(E, K, F) \
  if (econ == E && s.skewed == K && fixed == F) return launch_tma_cfg<POOL, E, K, F>(ctx, s, d_v, d_psi, st);
  CFMM_TMA_CASE(true, false, true)
  CFMM_TMA_CASE(true, false, false)
  CFMM_TMA_CASE(false, false, true)
  CFMM_TMA_CASE(false, false, false)
  if constexpr (POOL == 0) {  // hub orientation exists for the symmetric pool type only
    CFMM_TMA_CASE(true, true, true)
    CFMM_TMA_CASE(true, true, false)
    CFMM_TMA_CASE(false, true, true)
    CFMM_TMA_CASE(false, true, false)
  }
#undef CFMM_TMA_CASE
  return fail(ctx, CFMM_ERR_INVALID, "unreachable");
}

// One sweep.  The kernels accumulate into the internal ping-pong accumulator
// of this sweep (already zero: the previous sweep's first kernel cleared it) and
// clear the other one for the next sweep.  The result goes to d_dst if given
// (peer exchange writes it there directly; single GPU: one D2D copy), else it
// stays in the accumulator; *view receives the device pointer that holds it.
int enqueue_sweep(cfmm_ctx* ctx, const double* d_v, double* d_dst, bool mat,
                  cudaStream_t st, const double** view) {
  {
    int rc0 = use_stream(ctx, st);
    if (rc0 != CFMM_OK) return rc0;
  }
  if (ctx->sweep_events) CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));
  ctx->events_recorded = ctx->sweep_events != 0;
  ctx->epoch++;
  double* d_psi = ctx->d_accum[ctx->epoch & 1].p;
  ctx->zero_pending = ctx->d_accum[(ctx->epoch + 1) & 1].p;
  const size_t acc_bytes = (size_t)(ctx->n_tokens + 1) * sizeof(double);
  // Fused compute+collective: when the TMA ProductTwoCoin kernel is the only
  // kernel of this sweep, it runs the peer exchange in its own tail.
  bool fused = false;
  {
    const PoolSet& ps = ctx->sets[CFMM_POOL_PRODUCT];
    if (ctx->comm.attached() && !ctx->exchange_bypass && ctx->fused_exchange && !mat && ps.m > 0 && ps.tma_ok && ctx->use_tma &&
        ctx->debug_skip == 0 && ctx->sets[CFMM_POOL_GEOMEAN].m == 0 && ctx->sets[CFMM_POOL_UNIV3].m == 0) {
      fused = true;
      ctx->fx_pending.view = ctx->comm.view();
      ctx->fx_pending.dst = d_dst ? d_dst : d_psi;
      ctx->fx_pending.epoch = ctx->comm.begin_fused(&ctx->fx_pending.mode);
    }
  }
  int rc;
  {
    PoolSet& s = ctx->sets[CFMM_POOL_PRODUCT];
    constexpr int PT = CFMM_POOL_PRODUCT;
    if (s.m > 0) {
      if (!mat && s.tma_ok && ctx->use_tma && ctx->debug_skip == 0) {
        if ((rc = launch_tma<0>(ctx, s, d_v, d_psi, st)) != CFMM_OK) return rc;
      } else {
        cfmm::ProductPools p{s.d_R.p, s.d_gam.p, s.d_Ai.p};
        if ((rc = launch_sweep(ctx, PT, p, s, d_v, d_psi, mat, st)) != CFMM_OK) return rc;
      }
    }
  }
  {
    PoolSet& s = ctx->sets[CFMM_POOL_GEOMEAN];
    constexpr int PT = CFMM_POOL_GEOMEAN;
    if (s.m > 0) {
      cfmm::GeomeanPools p{s.d_R.p, s.d_gam.p, s.d_Ai.p, s.d_w.p};
      if (!mat && s.tma_ok && ctx->use_tma && ctx->geomean_tma && ctx->geomean_log2 && ctx->debug_skip == 0) {
        if ((rc = launch_tma<1>(ctx, s, d_v, d_psi, st)) != CFMM_OK) return rc;
      } else if (ctx->geomean_log2 && !mat) {
        cfmm::GeomeanPoolsLog2 q;
        static_cast<cfmm::GeomeanPools&>(q) = p;
        if ((rc = launch_sweep(ctx, PT, q, s, d_v, d_psi, mat, st)) != CFMM_OK) return rc;
      } else if ((rc = launch_sweep(ctx, PT, p, s, d_v, d_psi, mat, st)) != CFMM_OK) {
        return rc;
      }
    }
  }
  {
    PoolSet& s = ctx->sets[CFMM_POOL_UNIV3];
    constexpr int PT = CFMM_POOL_UNIV3;
    if (s.m > 0) {
      cfmm::Univ3Pools p{s.d_first[0].p, s.d_first[1].p, s.d_first[2].p, s.d_first[3].p, s.d_gam.p, s.d_Ai.p, s.d_tick.p,
                         s.d_tickdata.p, s.m_padded, (int)s.total_ticks};
      if ((rc = launch_sweep(ctx, PT, p, s, d_v, d_psi, mat, st)) != CFMM_OK) return rc;
    }
  }
  if (ctx->zero_pending) {  // no kernel ran (empty pool set): clear the other accumulator here
    CU_TRY(ctx, cudaMemsetAsync(take_zero_pending(ctx), 0, acc_bytes, st));
  }
  const double* result = d_psi;
  if (fused) {
    result = d_dst ? d_dst : d_psi;
  } else if (ctx->comm.attached() && !ctx->exchange_bypass) {
    ProfScope prof(ctx, 3, st);
    double* dst = d_dst ? d_dst : d_psi;
    if (!ctx->comm.all_reduce(d_psi, dst, ctx->n_tokens + 1, st))
      return fail(ctx, CFMM_ERR_COMM, "peer exchange failed: %s",
                  ctx->comm.error().c_str());
    ctx->launches += ctx->comm.launches_per_reduce();
    result = dst;
  } else if (d_dst) {
    CU_TRY(ctx, cudaMemcpyAsync(d_dst, d_psi, acc_bytes, cudaMemcpyDeviceToDevice, st));
    result = d_dst;
  }
  if (view) *view = result;
  if (ctx->sweep_events) CU_TRY(ctx, cudaEventRecord(ctx->ev1, st));
  if (mat) ctx->has_trades = true;
  return CFMM_OK;
}

int ready(cfmm_ctx* ctx) {
  if (!ctx) return CFMM_ERR_INVALID;
  if (!ctx->finalized)
    return fail(ctx, CFMM_ERR_STATE, "cfmm_finalize has not been called");
  return CFMM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------

extern "C" {

const char* cfmm_version(void) { return "0.1.0"; }

const char* cfmm_last_error(const cfmm_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int cfmm_create(cfmm_ctx** out, int device, int64_t n_tokens) {
  if (!out) return fail(nullptr, CFMM_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_tokens < 1 || n_tokens > (int64_t)0x7ffffff0)
    return fail(nullptr, CFMM_ERR_INVALID, "n_tokens must be in 1..2^31-16");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, CFMM_ERR_CUDA,
                "no CUDA device available (%s); libcfmm_b200 has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= n_dev)
    return fail(nullptr, CFMM_ERR_INVALID, "device %d out of range 0..%d", device,
                n_dev - 1);
  cfmm_ctx* ctx = new (std::nothrow) cfmm_ctx();
  if (!ctx) return fail(nullptr, CFMM_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  ctx->n_tokens = n_tokens;
#define CREATE_TRY(expr)                                                    \
  do {                                                                      \
    cudaError_t _e = (expr);                                                \
    if (_e != cudaSuccess) {                                                \
      fail(nullptr, CFMM_ERR_CUDA, "%s failed: %s", #expr,                  \
           cudaGetErrorString(_e));                                         \
      cfmm_destroy(ctx);                                                    \
      return CFMM_ERR_CUDA;                                                 \
    }                                                                       \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  CREATE_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreate(&ctx->ev0));
  CREATE_TRY(cudaEventCreate(&ctx->ev1));
  CREATE_TRY(ctx->d_nu.alloc((size_t)n_tokens));
  CREATE_TRY(ctx->d_grid_done.alloc(1));
  CREATE_TRY(cudaMemset(ctx->d_grid_done.p, 0, sizeof(unsigned long long)));
  memset(&ctx->fx_pending, 0, sizeof(ctx->fx_pending));
  for (auto& a : ctx->d_accum) {
    CREATE_TRY(a.alloc((size_t)n_tokens + 1));
    CREATE_TRY(cudaMemset(a.p, 0, ((size_t)n_tokens + 1) * sizeof(double)));
  }
  CREATE_TRY(cudaStreamSynchronize(cudaStreamLegacy));  // the memsets above ran on the legacy stream; ctx->stream does not wait for it
  CREATE_TRY(cudaMallocHost((void**)&ctx->h_stage, (size_t)(n_tokens + 1) * sizeof(double)));
#undef CREATE_TRY
  *out = ctx;
  return CFMM_OK;
}

namespace {
void drop_graphs(cfmm_ctx* ctx);
}

void cfmm_destroy(cfmm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->comm.detach();
  drop_graphs(ctx);
  for (auto e : ctx->prof.ev)
    if (e) cudaEventDestroy(e);
  for (auto& s : ctx->sets) s.release();
  ctx->d_nu.release();
  ctx->d_grid_done.release();
  ctx->d_trace.release();
  ctx->d_accum[0].release();
  ctx->d_accum[1].release();
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  for (void* b : ctx->h_bounce)
    if (b) cudaFreeHost(b);
  for (cudaEvent_t e : ctx->ev_bounce)
    if (e) cudaEventDestroy(e);
  if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int cfmm_add_product(cfmm_ctx* ctx, int64_t m, const double* R,
                     const double* gamma, const int64_t* Ai) {
  int rc = check_common(ctx, m, R, gamma, Ai);
  if (rc != CFMM_OK) return rc;
  append_common(ctx, ctx->sets[CFMM_POOL_PRODUCT], m, R, gamma, Ai);
  return CFMM_OK;
}

int cfmm_add_geomean(cfmm_ctx* ctx, int64_t m, const double* R,
                     const double* gamma, const int64_t* Ai, const double* w) {
  int rc = check_common(ctx, m, R, gamma, Ai);
  if (rc != CFMM_OK) return rc;
  if (m > 0 && !w) return fail(ctx, CFMM_ERR_INVALID, "null weight array");
  PoolSet& s = ctx->sets[CFMM_POOL_GEOMEAN];
  append_common(ctx, s, m, R, gamma, Ai);
  s.w.insert(s.w.end(), w, w + 2 * m);
  return CFMM_OK;
}

int cfmm_add_univ3(cfmm_ctx* ctx, int64_t m, const double* current_price,
                   const double* gamma, const int64_t* Ai,
                   const int64_t* tick_off, const double* lower_ticks,
                   const double* liquidity) {
  static const double dummy = 0.0;
  int rc = check_common(ctx, m, m > 0 ? &dummy : nullptr, gamma, Ai);
  if (rc != CFMM_OK) return rc;
  if (m == 0) return CFMM_OK;
  if (!current_price || !tick_off || !lower_ticks || !liquidity)
    return fail(ctx, CFMM_ERR_INVALID, "null array argument");
  if (tick_off[0] != 0)
    return fail(ctx, CFMM_ERR_INVALID, "tick_off[0] must be 0");
  PoolSet& s = ctx->sets[CFMM_POOL_UNIV3];
  for (int64_t i = 0; i < m; ++i) {
    const int64_t b = tick_off[i], e = tick_off[i + 1];
    if (e <= b)
      return fail(ctx, CFMM_ERR_INVALID, "univ3 pool %lld has no ticks", (long long)i);
    for (int64_t t = b + 1; t < e; ++t)
      if (!(lower_ticks[t] < lower_ticks[t - 1]))
        return fail(ctx, CFMM_ERR_INVALID,
                    "univ3 pool %lld: lower_ticks must be strictly decreasing",
                    (long long)i);
    if (!(lower_ticks[b] >= current_price[i]))
      return fail(ctx, CFMM_ERR_INVALID,
                  "univ3 pool %lld: current_price above the first lower tick "
                  "(current_tick == 0; BoundsError in the reference)",
                  (long long)i);
  }
  const int64_t n_ticks = tick_off[m];
  if ((int64_t)s.lower.size() + n_ticks > (int64_t)0x7fffffff)
    return fail(ctx, CFMM_ERR_INVALID, "more than 2^31-1 ticks in one context");
  const int64_t base = (int64_t)s.lower.size();
  if (s.tick_off.empty()) s.tick_off.push_back(0);
  for (int64_t i = 1; i <= m; ++i) s.tick_off.push_back(base + tick_off[i]);
  s.cp.insert(s.cp.end(), current_price, current_price + m);
  s.lower.insert(s.lower.end(), lower_ticks, lower_ticks + n_ticks);
  s.liq.insert(s.liq.end(), liquidity, liquidity + n_ticks);
  append_common(ctx, s, m, nullptr, gamma, Ai);
  return CFMM_OK;
}

// ---- flat pool files (an ingest format that is not an array of heap
// objects).  Little-endian, 64-byte header {magic "CFMMPOOL", u32 version = 1, u32 pool
// type, i64 m, i64 n_tokens, zero pad}, then the SoA arrays exactly as cfmm_add_* take
// them: R [2m] f64, gamma [m] f64, Ai [2m] i64 (1-based), and w [2m] f64 for
// GeometricMeanTwoCoin.  The file is mmap-ed; cfmm_add_pool_file feeds it to cfmm_add_*.
namespace {
struct PoolFileHeader {
  char magic[8];
  uint32_t version, type;
  int64_t m, n_tokens;
  char pad[32];
};
static_assert(sizeof(PoolFileHeader) == 64, "pool file header");

size_t pool_file_bytes(int type, int64_t m) {
  return sizeof(PoolFileHeader) + (size_t)m * (16 + 8 + 16 + (type == CFMM_POOL_GEOMEAN ? 16 : 0));
}
}  // namespace

int cfmm_pool_file_write(const char* path, int type, int64_t n_tokens, int64_t m, const double* R,
                         const double* gamma, const int64_t* Ai, const double* w) {
  if (!path || m < 0 || n_tokens < 1 || (type != CFMM_POOL_PRODUCT && type != CFMM_POOL_GEOMEAN) ||
      (m > 0 && (!R || !gamma || !Ai || (type == CFMM_POOL_GEOMEAN && !w))))
    return CFMM_ERR_INVALID;
  FILE* f = fopen(path, "wb");
  if (!f) return CFMM_ERR_INVALID;
  PoolFileHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "CFMMPOOL", 8);
  h.version = 1;
  h.type = (uint32_t)type;
  h.m = m;
  h.n_tokens = n_tokens;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  ok = ok && fwrite(R, 16, (size_t)m, f) == (size_t)m && fwrite(gamma, 8, (size_t)m, f) == (size_t)m &&
       fwrite(Ai, 16, (size_t)m, f) == (size_t)m;
  if (type == CFMM_POOL_GEOMEAN) ok = ok && fwrite(w, 16, (size_t)m, f) == (size_t)m;
  ok = (fclose(f) == 0) && ok;
  return ok ? CFMM_OK : CFMM_ERR_INVALID;
}

int cfmm_pool_file_info(const char* path, int* type, int64_t* n_tokens, int64_t* m) {
  if (!path || !type || !n_tokens || !m) return CFMM_ERR_INVALID;
  FILE* f = fopen(path, "rb");
  if (!f) return CFMM_ERR_INVALID;
  PoolFileHeader h;
  const bool ok = fread(&h, sizeof(h), 1, f) == 1;
  fseek(f, 0, SEEK_END);
  const long size = ftell(f);
  fclose(f);
  if (!ok || memcmp(h.magic, "CFMMPOOL", 8) != 0 || h.version != 1 || h.m < 0 ||
      (h.type != CFMM_POOL_PRODUCT && h.type != CFMM_POOL_GEOMEAN) ||
      (size_t)size != pool_file_bytes((int)h.type, h.m))
    return CFMM_ERR_INVALID;
  *type = (int)h.type;
  *n_tokens = h.n_tokens;
  *m = h.m;
  return CFMM_OK;
}

int cfmm_add_pool_file(cfmm_ctx* ctx, const char* path) {
  if (!ctx) return CFMM_ERR_INVALID;
  int type = 0;
  int64_t n_tokens = 0, m = 0;
  if (cfmm_pool_file_info(path, &type, &n_tokens, &m) != CFMM_OK)
    return fail(ctx, CFMM_ERR_INVALID, "'%s' is not a readable CFMM pool file", path ? path : "(null)");
  if (n_tokens != ctx->n_tokens)
    return fail(ctx, CFMM_ERR_INVALID, "pool file is for %lld tokens, the context for %lld",
                (long long)n_tokens, (long long)ctx->n_tokens);
  if (m == 0) return CFMM_OK;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(ctx, CFMM_ERR_INVALID, "cannot open '%s'", path);
  const size_t bytes = pool_file_bytes(type, m);
  void* map = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return fail(ctx, CFMM_ERR_NOMEM, "mmap of '%s' failed", path);
  const char* base = (const char*)map + sizeof(PoolFileHeader);
  const double* R = (const double*)base;
  const double* gamma = R + 2 * m;
  const int64_t* Ai = (const int64_t*)(gamma + m);
  const double* w = (const double*)(Ai + 2 * m);
  const int rc = type == CFMM_POOL_PRODUCT ? cfmm_add_product(ctx, m, R, gamma, Ai)
                                           : cfmm_add_geomean(ctx, m, R, gamma, Ai, w);
  munmap(map, bytes);
  return rc;
}

int cfmm_finalize(cfmm_ctx* ctx) {
  if (!ctx) return CFMM_ERR_INVALID;
  if (ctx->finalized) return fail(ctx, CFMM_ERR_STATE, "already finalized");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const bool timing = getenv("CFMM_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < 3; ++t) {
    int rc = upload_set(ctx, t);
    if (rc != CFMM_OK) return rc;
    if (timing) {
      const auto t1 = std::chrono::steady_clock::now();
      fprintf(stderr, "[cfmm] finalize: pool type %d (%lld pools) %.3f s\n", t, (long long)ctx->sets[t].m,
              std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
    }
  }
  ctx->finalized = true;
  // Calibration: a dozen gradient sweeps at ν = 1 settle the speed-weighted CTA ranges of the TMA
  // kernels (and pay the one-time costs of the first launch: function attributes, stream packing)
  // here rather than in the caller's first sweeps.  ~1 ms.
  {
    bool any = false;
    for (int t : {CFMM_POOL_PRODUCT, CFMM_POOL_GEOMEAN}) any = any || (ctx->sets[t].m > 0 && ctx->sets[t].tma_ok);
    if (any && ctx->balance) {
      std::vector<double> ones((size_t)ctx->n_tokens, 1.0);
      CU_TRY(ctx, DevBuf<double>::copy_in(ctx->d_nu.p, ones.data(), ones.size() * sizeof(double)));
      ctx->calibrating = true;
      int rc = CFMM_OK;
      for (int it = 0; it < 12 && rc == CFMM_OK; ++it) {
        const double* view = nullptr;
        rc = enqueue_sweep(ctx, ctx->d_nu.p, nullptr, false, ctx->stream, &view);
        if (rc == CFMM_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
          rc = fail(ctx, CFMM_ERR_CUDA, "calibration sweep failed: %s", cudaGetErrorString(cudaGetLastError()));
      }
      ctx->calibrating = false;
      if (rc != CFMM_OK) return rc;
      if (timing)
        fprintf(stderr, "[cfmm] finalize: calibration %.3f s\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
  }
  return CFMM_OK;
}

int64_t cfmm_num_pools(const cfmm_ctx* ctx) { return ctx ? ctx->n_pools : -1; }
int64_t cfmm_num_tokens(const cfmm_ctx* ctx) { return ctx ? ctx->n_tokens : -1; }
int64_t cfmm_launch_count(const cfmm_ctx* ctx) { return ctx ? ctx->launches : -1; }

int cfmm_sweep_device(cfmm_ctx* ctx, const double* d_v, double* d_psi_acc,
                      int materialize, void* stream) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (!d_v || !d_psi_acc) return fail(ctx, CFMM_ERR_INVALID, "null device pointer");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  return enqueue_sweep(ctx, d_v, d_psi_acc, materialize != 0, st, nullptr);
}

int cfmm_sweep_device_view(cfmm_ctx* ctx, const double* d_v, int materialize, void* stream,
                           const double** d_psi_acc_out) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (!d_v || !d_psi_acc_out) return fail(ctx, CFMM_ERR_INVALID, "null pointer");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  return enqueue_sweep(ctx, d_v, nullptr, materialize != 0, st, d_psi_acc_out);
}

namespace {

bool host_pinned(cfmm_ctx* ctx, const void* p) {
  if (p == ctx->pinned_ok[0] || p == ctx->pinned_ok[1]) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (a.type != cudaMemoryTypeHost) return false;
  ctx->pinned_ok[1] = ctx->pinned_ok[0];
  ctx->pinned_ok[0] = p;
  return true;
}

void drop_graphs(cfmm_ctx* ctx) {
  for (auto& row : ctx->graphs)
    for (auto& g : row) {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      g = cfmm_ctx::SweepGraph();
    }
}

}  // namespace

int cfmm_sweep(cfmm_ctx* ctx, const double* v, double* psi_out, double* acc_out,
               int materialize) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (!v || !psi_out || !acc_out)
    return fail(ctx, CFMM_ERR_INVALID, "null host pointer");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t nb = (size_t)ctx->n_tokens * sizeof(double);
  cudaStream_t st = ctx->stream;
  const bool contiguous = acc_out == psi_out + ctx->n_tokens;

  // ---- graph path: the whole call is one cudaGraphLaunch ---------------------------------
  // (a graph freezes the kernel parameters, the range table among them: capture only once the
  // speed feedback of the TMA kernels has settled)
  bool settled = true;
  for (int t : {CFMM_POOL_PRODUCT, CFMM_POOL_GEOMEAN}) {
    const PoolSet& ps = ctx->sets[t];
    if (ps.m > 0 && ps.tma_ok && ctx->use_tma && ctx->balance && (ps.tma_launches == 0 || (ps.balancing && ps.range_updates < 6)))
      settled = false;
  }
  const bool graphable = ctx->use_graphs && !materialize && contiguous && !ctx->sweep_events &&
                         !ctx->comm.attached() && ctx->prof.type.empty() && !ctx->d_trace.n &&
                         ctx->debug_skip == 0 && settled;
  cfmm_ctx::SweepGraph* g = nullptr;
  if (graphable) {
    g = &ctx->graphs[(ctx->epoch + 1) & 1][0];
    if (g->exec && g->v == v && g->psi == psi_out && g->version == ctx->state_version) {
      ctx->epoch++;
      ctx->launches += g->launches;
      ctx->events_recorded = false;
      CU_TRY(ctx, cudaGraphLaunch(g->exec, st));
      CU_TRY(ctx, cudaStreamSynchronize(st));
      return CFMM_OK;
    }
    const bool second = g->seen_v == v && g->seen_psi == psi_out && g->seen_version == ctx->state_version;
    if (second && host_pinned(ctx, v) && host_pinned(ctx, psi_out)) {
      // same buffers, same options as the last call on this parity: capture
      if (g->exec) cudaGraphExecDestroy(g->exec);
      g->exec = nullptr;
      const int64_t l0 = ctx->launches;
      cudaGraph_t graph = nullptr;
      CU_TRY(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      ctx->capturing = true;
      cudaError_t e = cudaMemcpyAsync(ctx->d_nu.p, v, nb, cudaMemcpyHostToDevice, st);
      const double* res = nullptr;
      if (e == cudaSuccess) rc = enqueue_sweep(ctx, ctx->d_nu.p, nullptr, false, st, &res);
      if (e == cudaSuccess && rc == CFMM_OK)
        e = cudaMemcpyAsync(psi_out, res, nb + sizeof(double), cudaMemcpyDeviceToHost, st);
      cudaError_t e2 = cudaStreamEndCapture(st, &graph);
      ctx->capturing = false;
      if (e == cudaSuccess && rc == CFMM_OK && e2 == cudaSuccess && graph &&
          cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess) {
        g->v = v;
        g->psi = psi_out;
        g->version = ctx->state_version;
        g->launches = ctx->launches - l0;
        cudaGraphDestroy(graph);
        CU_TRY(ctx, cudaGraphLaunch(g->exec, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        return CFMM_OK;
      }
      // capture failed: the sweep state has advanced without any work being done -- undo and
      // take the eager path for good
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      g->exec = nullptr;
      ctx->use_graphs = 0;
      ctx->epoch--;
      ctx->launches = l0;
      if (rc != CFMM_OK) return rc;
    }
    g->seen_v = v;
    g->seen_psi = psi_out;
    g->seen_version = ctx->state_version;
  }

  CU_TRY(ctx, cudaMemcpyAsync(ctx->d_nu.p, v, nb, cudaMemcpyHostToDevice, st));
  const double* res = nullptr;
  rc = enqueue_sweep(ctx, ctx->d_nu.p, nullptr, materialize != 0, st, &res);
  if (rc != CFMM_OK) return rc;
  // A peer that died leaves the exchange polling until its timeout (seconds): only a sweep that
  // took that long looks at the error word.
  const auto t_wait = std::chrono::steady_clock::now();
  auto comm_ok = [&]() -> int {
    if (!ctx->comm.attached()) return CFMM_OK;
    if (std::chrono::steady_clock::now() - t_wait < std::chrono::seconds(1)) return CFMM_OK;
    if (ctx->comm.timed_out())
      return fail(ctx, CFMM_ERR_COMM, "peer exchange timed out: a rank of the group did not deliver its packets");
    return CFMM_OK;
  };
  if (contiguous) {
    // caller keeps [psi ; acc] contiguous: one D2H copy
    CU_TRY(ctx, cudaMemcpyAsync(psi_out, res, nb + sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    return comm_ok();
  }
  CU_TRY(ctx, cudaMemcpyAsync(psi_out, res, nb, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_stage, res + ctx->n_tokens,
                              sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  *acc_out = ctx->h_stage[0];
  return comm_ok();
}

// ---- cfmm_solve: the outer iteration of route! on the device (solver.cuh) ----------------
int cfmm_solve(cfmm_ctx* ctx, const double* lin, const double* lower, const double* upper,
               const double* v0, const cfmm_solve_opts* opts_in, double* v_out, cfmm_solve_info* info) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (!lower || !v_out) return fail(ctx, CFMM_ERR_INVALID, "cfmm_solve: lower and v_out are required");
  cfmm_solve_opts o;
  o.max_iter = 15000;
  o.max_fun = 15000;
  o.pgtol = 1e-5;
  o.factr = 1e1;
  if (opts_in) o = *opts_in;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t n = ctx->n_tokens;
  constexpr int M = cfmm::kSolverM, K = cfmm::kSolverK, NG = cfmm::kSolverGram;
  for (int64_t i = 0; i < n; ++i)
    if (!(lower[i] == lower[i]) || (upper && !(upper[i] >= lower[i])))
      return fail(ctx, CFMM_ERR_INVALID, "cfmm_solve: bad bounds at token %lld", (long long)(i + 1));
  // device state
  DevBuf<double> vec, hist, red, box;
  DevBuf<unsigned> ticket;
  DevBuf<unsigned long long> pgbits;
  CU_TRY(ctx, vec.alloc((size_t)6 * n));
  CU_TRY(ctx, hist.alloc((size_t)2 * M * n));
  CU_TRY(ctx, red.alloc((size_t)cfmm::kSolverMaxBlocks * NG + NG + 8));
  CU_TRY(ctx, box.alloc((size_t)3 * n));
  CU_TRY(ctx, ticket.alloc(1));
  CU_TRY(ctx, pgbits.alloc(1));
  CU_TRY(ctx, cudaMemsetAsync(hist.p, 0, (size_t)2 * M * n * sizeof(double), st));
  CU_TRY(ctx, cudaMemsetAsync(ticket.p, 0, sizeof(unsigned), st));
  CU_TRY(ctx, cudaMemsetAsync(vec.p, 0, (size_t)6 * n * sizeof(double), st));
  cfmm::SolverVecs q;
  q.n = n;
  q.x = vec.p;
  q.g = vec.p + n;
  q.xt = vec.p + 2 * n;
  q.gt = vec.p + 3 * n;
  q.d = vec.p + 4 * n;
  q.pg = vec.p + 5 * n;
  q.S = hist.p;
  q.Y = hist.p + (size_t)M * n;
  q.partials = red.p;
  q.scal = red.p + (size_t)cfmm::kSolverMaxBlocks * NG;
  q.ticket = ticket.p;
  CU_TRY(ctx, cudaMemcpyAsync(box.p, lower, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  q.lower = box.p;
  q.upper = nullptr;
  q.lin = nullptr;
  if (upper) {
    bool finite = false;
    for (int64_t i = 0; i < n && !finite; ++i) finite = upper[i] < 1.0e300;
    if (finite) {
      CU_TRY(ctx, cudaMemcpyAsync(box.p + n, upper, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
      q.upper = box.p + n;
    }
  }
  if (lin) {
    CU_TRY(ctx, cudaMemcpyAsync(box.p + 2 * n, lin, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    q.lin = box.p + 2 * n;
  }
  {
    std::vector<double> start((size_t)n, 1.0 / (double)n);  // route!'s default start, router.jl:62
    if (v0) start.assign(v0, v0 + n);
    CU_TRY(ctx, cudaMemcpyAsync(q.d, start.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));  // `start` leaves scope
  }
  int blocks = (int)((n + cfmm::kSolverThreads - 1) / cfmm::kSolverThreads);
  if (blocks > cfmm::kSolverMaxBlocks) blocks = cfmm::kSolverMaxBlocks;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  cudaEventRecord(t0, st);

  double h[NG + 8];
  int fevals = 0;
  const double epsmch = 2.220446049250313e-16;
  auto evaluate = [&](double* f_out) -> int {  // sweep at xt, gt = lin + Ψ, f = linᵀxt + acc
    const double* view = nullptr;
    int r = enqueue_sweep(ctx, q.xt, nullptr, false, st, &view);
    if (r != CFMM_OK) return r;
    cfmm::solver_grad_kernel<<<blocks, cfmm::kSolverThreads, 0, st>>>(q, view);
    ctx->launches++;
    cudaMemcpyAsync(h + NG, q.scal + NG, 5 * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h + NG + 5, view + n, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(ctx, CFMM_ERR_CUDA, "cfmm_solve: %s", cudaGetErrorString(e));
    ++fevals;
    *f_out = h[NG + 1] + h[NG + 5];
    return CFMM_OK;
  };
  double W[K][K];
  double pgnorm = 0.0;
  auto commit = [&](int slot, int store) -> int {  // accept xt; W, |pg|_inf back
    cudaMemsetAsync(pgbits.p, 0, sizeof(unsigned long long), st);
    cfmm::solver_commit_kernel<<<blocks, cfmm::kSolverThreads, 0, st>>>(q, slot, store, pgbits.p);
    ctx->launches++;
    unsigned long long bits = 0;
    cudaMemcpyAsync(h, q.scal, NG * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&bits, pgbits.p, sizeof(bits), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(ctx, CFMM_ERR_CUDA, "cfmm_solve: %s", cudaGetErrorString(e));
    int k = 0;
    for (int r = 0; r < K; ++r)
      for (int c = r; c < K; ++c) W[r][c] = W[c][r] = h[k++];
    memcpy(&pgnorm, &bits, sizeof(double));
    return CFMM_OK;
  };

  // x = P(v0); first evaluation
  cfmm::solver_init_kernel<<<blocks, cfmm::kSolverThreads, 0, st>>>(q, q.d);
  ctx->launches++;
  double f = 0.0;
  if ((rc = evaluate(&f)) != CFMM_OK) return rc;
  if ((rc = commit(0, 0)) != CFMM_OK) return rc;

  int age[M];      // history slots, oldest first
  int cnt = 0, head = 0, iter = 0, status = 2, small_steps = 0;
  while (true) {
    if (!(f == f)) { status = 5; break; }            // NaN objective
    if (pgnorm <= o.pgtol) { status = 0; break; }
    if (iter >= o.max_iter) { status = 2; break; }
    if (fevals >= o.max_fun) { status = 3; break; }
    // ---- two-loop recursion in coefficient space over B = [S Y pg] ------------------------
    cfmm::SolverCoef cf;
    for (int j = 0; j < K; ++j) cf.c[j] = 0.0;
    cf.c[K - 1] = 1.0;
    double t_init = 1.0;
    if (cnt == 0) {
      const double nrm = std::sqrt(W[K - 1][K - 1]);
      t_init = nrm > 0.0 ? std::min(1.0, 1.0 / nrm) : 1.0;   // first step of length <= 1, like L-BFGS-B
    } else {
      double alpha[M], rho[M];
      for (int a = cnt - 1; a >= 0; --a) {
        const int j = age[a];
        rho[a] = 1.0 / W[j][M + j];
        double sq = 0.0;
        for (int c = 0; c < K; ++c) sq += W[j][c] * cf.c[c];
        alpha[a] = rho[a] * sq;
        cf.c[M + j] -= alpha[a];
      }
      const int jn = age[cnt - 1];
      const double gamma = W[jn][M + jn] / W[M + jn][M + jn];
      for (int c = 0; c < K; ++c) cf.c[c] *= gamma;
      for (int a = 0; a < cnt; ++a) {
        const int j = age[a];
        double yr = 0.0;
        for (int c = 0; c < K; ++c) yr += W[M + j][c] * cf.c[c];
        cf.c[j] += alpha[a] - rho[a] * yr;
      }
    }
    cfmm::solver_direction_kernel<<<blocks, cfmm::kSolverThreads, 0, st>>>(q, cf);
    ctx->launches++;
    // ---- Armijo backtracking along the projected path ------------------------------------
    // f is a sum of ~m terms of mixed sign: differences below ~8 eps |f| are rounding noise, and
    // near the (very flat) optimum of a large market every useful step is that small -- a
    // plain Armijo test would reject them all.  The slack admits them; pgtol / factr decide when
    // to stop.
    double t = t_init, f_new = f;
    bool accepted = false, stalled = false;
    for (int ls = 0; ls < 30 && fevals < o.max_fun; ++ls) {
      cfmm::solver_trial_kernel<<<blocks, cfmm::kSolverThreads, 0, st>>>(q, t);
      ctx->launches++;
      if ((rc = evaluate(&f_new)) != CFMM_OK) return rc;
      const double gdx = h[NG + 0], step2 = h[NG + 2];
      if (step2 == 0.0) { stalled = true; break; }   // the projected step does not move
      const double noise = 8.0 * epsmch * std::max(std::max(std::fabs(f), std::fabs(f_new)), 1.0);
      if (gdx < 0.0 && f_new <= f + 1e-4 * gdx + noise) { accepted = true; break; }
      if (!(gdx < 0.0) && cnt > 0) break;             // not a descent direction: restart from −pg
      if (f_new == f_new && f_new < 1e300 && gdx < 0.0) {
        // minimiser of the quadratic through f, the slope gdx (per unit t) and f_new, kept in [0.1 t, 0.5 t]
        const double slope = gdx / t, denom = 2.0 * (f_new - f - gdx);
        double tq = denom > 0.0 ? -slope * t * t / denom : 0.5 * t;
        t = std::min(0.5 * t, std::max(0.1 * t, tq));
      } else {
        t *= 0.1;
      }
    }
    if (!accepted) {
      if (cnt > 0 && !stalled) {  // drop the history and retry with steepest descent
        cnt = 0;
        continue;
      }
      status = stalled ? 1 : 4;
      break;
    }
    // ---- accept: store (s, y), new Gram matrix / projected gradient --------------------------
    const int slot = head;
    if ((rc = commit(slot, 1)) != CFMM_OK) return rc;
    // drop the slot's old pair from the age list, append the new one if its curvature is usable
    int w = 0;
    for (int a = 0; a < cnt; ++a)
      if (age[a] != slot) age[w++] = age[a];
    cnt = w;
    const double sy = W[slot][M + slot], yy = W[M + slot][M + slot];
    if (sy > 1e-10 * yy && yy > 0.0) {
      age[cnt++] = slot;
      head = (head + 1) % M;
    }
    ++iter;
    const double f_old = f;
    f = f_new;
    // L-BFGS-B's factr test -- on two consecutive steps: one short quasi-Newton step (fresh history,
    // a bound just hit) is not yet evidence of convergence
    if (f_old - f <= o.factr * epsmch * std::max(std::max(std::fabs(f_old), std::fabs(f)), 1.0)) {
      if (++small_steps >= 2) {
        status = 1;
        break;
      }
    } else {
      small_steps = 0;
    }
  }
  // final ν, and the trades at it (router.jl:106-107)
  CU_TRY(ctx, cudaMemcpyAsync(v_out, q.x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  const double* view = nullptr;
  if ((rc = enqueue_sweep(ctx, q.x, nullptr, true, st, &view)) != CFMM_OK) return rc;
  cudaEventRecord(t1, st);
  CU_TRY(ctx, cudaStreamSynchronize(st));
  if (info) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    info->iterations = iter;
    info->fun_evals = fevals;
    info->status = status;
    info->f = f;
    info->pg_norm = pgnorm;
    info->solve_ms = ms;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return CFMM_OK;
}

int cfmm_last_sweep_ms(cfmm_ctx* ctx, float* ms_out) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (!ms_out) return fail(ctx, CFMM_ERR_INVALID, "ms_out is NULL");
  if (!ctx->events_recorded)
    return fail(ctx, CFMM_ERR_STATE, "the last sweep recorded no events: set option \"sweep_events\" to 1 first");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, cudaEventSynchronize(ctx->ev1));
  CU_TRY(ctx, cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
  return CFMM_OK;
}

int cfmm_get_trades(cfmm_ctx* ctx, double* Delta, double* Lambda) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (!Delta || !Lambda) return fail(ctx, CFMM_ERR_INVALID, "null host pointer");
  if (!ctx->has_trades)
    return fail(ctx, CFMM_ERR_STATE,
                "no materialising sweep has run (call cfmm_sweep with materialize=1)");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->n_pools == 0) return CFMM_OK;
  if ((rc = use_stream(ctx, ctx->stream)) != CFMM_OK) return rc;
  DevBuf<double2> allD, allL;
  CU_TRY(ctx, allD.alloc((size_t)ctx->n_pools));
  cudaError_t e = allL.alloc((size_t)ctx->n_pools);
  if (e != cudaSuccess) {
    allD.release();
    return fail(ctx, CFMM_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  for (auto& s : ctx->sets) {
    if (s.m == 0) continue;
    const int threads = 256;
    const int64_t blocks = (s.m_padded + threads - 1) / threads;
    cfmm::scatter_trades_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        s.d_outD.p, s.d_outL.p, s.d_gidx.p, allD.p, allL.p, s.m_padded);
    ctx->launches++;
  }
  const size_t bytes = (size_t)ctx->n_pools * sizeof(double2);
  cudaError_t e1 = cudaMemcpyAsync(Delta, allD.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e2 = cudaMemcpyAsync(Lambda, allL.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
  allD.release();
  allL.release();
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
    return fail(ctx, CFMM_ERR_CUDA, "trade read-back failed: %s",
                cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
  return CFMM_OK;
}

int cfmm_update_reserves(cfmm_ctx* ctx, int type, int64_t first, int64_t count,
                         const double* R) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (type != CFMM_POOL_PRODUCT && type != CFMM_POOL_GEOMEAN)
    return fail(ctx, CFMM_ERR_INVALID, "update_reserves: type must be PRODUCT or GEOMEAN");
  PoolSet& s = ctx->sets[type];
  ctx->state_version++;
  if (first < 0 || count < 0 || first + count > s.m)
    return fail(ctx, CFMM_ERR_INVALID, "update_reserves: range [%lld, %lld) outside 0..%lld",
                (long long)first, (long long)(first + count), (long long)s.m);
  if (count == 0) return CFMM_OK;
  if (!R) return fail(ctx, CFMM_ERR_INVALID, "null reserve array");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  if ((rc = use_stream(ctx, ctx->stream)) != CFMM_OK) return rc;
  if (s.pos_of.empty()) {
    s.pos_of.resize((size_t)s.m);
    for (int64_t p = 0; p < s.m_padded; ++p)
      if (s.order[(size_t)p] >= 0) s.pos_of[(size_t)s.order[(size_t)p]] = p;
  }
  std::vector<double2> newR((size_t)count);
  for (int64_t j = 0; j < count; ++j) {
    newR[(size_t)j] = s.swapped[(size_t)(first + j)] ? make_double2(R[2 * j + 1], R[2 * j])
                                                     : make_double2(R[2 * j], R[2 * j + 1]);
    if (!fast_range_ok(R[2 * j]) || !fast_range_ok(R[2 * j + 1])) s.in_fast_range = false;
  }
  DevBuf<double2> d_new;
  DevBuf<int64_t> d_pos;
  CU_TRY(ctx, d_new.upload(newR));
  cudaError_t e = d_pos.alloc((size_t)count);
  if (e == cudaSuccess)
    e = DevBuf<int64_t>::copy_in(d_pos.p, s.pos_of.data() + first, (size_t)count * sizeof(int64_t));
  if (e == cudaSuccess) {
    const int threads = 256;
    cfmm::update_reserves_kernel<<<(unsigned)((count + threads - 1) / threads), threads, 0,
                                   ctx->stream>>>(s.d_R.p, d_pos.p, d_new.p, count);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
  }
  d_new.release();
  d_pos.release();
  if (e != cudaSuccess)
    return fail(ctx, CFMM_ERR_CUDA, "update_reserves failed: %s", cudaGetErrorString(e));
  if (s.tma_ok) return refresh_scale(ctx, s);
  return CFMM_OK;
}

int cfmm_apply_trades(cfmm_ctx* ctx) {
  int rc = ready(ctx);
  if (rc != CFMM_OK) return rc;
  if (!ctx->has_trades)
    return fail(ctx, CFMM_ERR_STATE,
                "no materialising sweep has run (call cfmm_sweep with materialize=1)");
  if (ctx->sets[CFMM_POOL_UNIV3].m > 0)
    return fail(ctx, CFMM_ERR_INVALID,
                "cfmm_apply_trades: UniV3 pools have no explicit reserves (R + γΔ − Λ is defined for "
                "ProductTwoCoin / GeometricMeanTwoCoin only)");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  ctx->state_version++;
  if ((rc = use_stream(ctx, ctx->stream)) != CFMM_OK) return rc;
  DevBuf<int> flag;
  std::vector<int> zero(1, 0);
  for (int t : {CFMM_POOL_PRODUCT, CFMM_POOL_GEOMEAN}) {
    PoolSet& s = ctx->sets[t];
    if (s.m == 0) continue;
    if (s.d_outD.n != (size_t)s.m_padded)
      return fail(ctx, CFMM_ERR_STATE, "trades of this pool type were never materialised");
    CU_TRY(ctx, flag.upload(zero));
    const int threads = 256;
    cfmm::apply_trades_kernel<<<(unsigned)((s.m_padded + threads - 1) / threads), threads, 0, ctx->stream>>>(
        s.d_R.p, s.d_gam.p, s.d_outD.p, s.d_outL.p, s.d_gidx.p, s.m_padded, flag.p);
    ctx->launches++;
    int h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    flag.release();
    if (e != cudaSuccess) return fail(ctx, CFMM_ERR_CUDA, "apply_trades failed: %s", cudaGetErrorString(e));
    if (h) s.in_fast_range = false;  // later sweeps take the generic (guarded) form
    if (s.tma_ok) {
      int rc2 = refresh_scale(ctx, s);
      if (rc2 != CFMM_OK) return rc2;
    }
  }
  return CFMM_OK;
}

int cfmm_set_option(cfmm_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return CFMM_ERR_INVALID;
  ctx->state_version++;  // captured sweep graphs are stale
  if (!strcmp(key, "sweep_graphs")) {
    ctx->use_graphs = value != 0;
    return CFMM_OK;
  }
  if (!strcmp(key, "exact")) {
    ctx->exact = value != 0;
  } else if (!strcmp(key, "blocks_per_sm")) {
    if (value < 0 || value > 32) return fail(ctx, CFMM_ERR_INVALID, "blocks_per_sm out of range");
    ctx->blocks_per_sm = (int)value;
  } else if (!strcmp(key, "tma_variant")) {
    if (value < -1 || value > 0) return fail(ctx, CFMM_ERR_INVALID, "tma_variant must be 0 (bucketed layout, TMA kernel) or -1");
    if (ctx->finalized)
      return fail(ctx, CFMM_ERR_STATE, "tma_variant fixes the pool layout: set it before cfmm_finalize");
    ctx->tma_variant = (int)value;
  } else if (!strcmp(key, "orient_by_degree")) {
    if (ctx->finalized)
      return fail(ctx, CFMM_ERR_STATE, "orient_by_degree fixes the pool layout: set it before cfmm_finalize");
    ctx->orient_by_degree = value < 0 ? -1 : (value != 0);
  } else if (!strcmp(key, "psi_fixed_point")) {
    ctx->psi_fixed_point = value != 0;
  } else if (!strcmp(key, "compact_stream")) {
    ctx->compact_stream = value != 0;
  } else if (!strcmp(key, "geomean_tma")) {
    ctx->geomean_tma = value != 0;
  } else if (!strcmp(key, "balance")) {
    ctx->balance = value != 0;
  } else if (!strcmp(key, "trace")) {
    // measurement only: 1 = every TMA sweep records per-CTA phase timestamps (cfmm_debug_read_trace)
    cudaSetDevice(ctx->device);
    if (value) {
      CU_TRY(ctx, ctx->d_trace.alloc((size_t)8 * 4096));
      CU_TRY(ctx, cudaMemset(ctx->d_trace.p, 0, 8 * 4096 * sizeof(unsigned long long)));
      CU_TRY(ctx, cudaStreamSynchronize(cudaStreamLegacy));
    } else {
      cudaStreamSynchronize(ctx->stream);
      ctx->d_trace.release();
    }
  } else if (!strcmp(key, "fused_exchange")) {
    ctx->fused_exchange = value != 0;
  } else if (!strcmp(key, "grid_waves")) {
    ctx->grid_waves = (int)value;
    ctx->state_version++;
  } else if (!strcmp(key, "coop_launch")) {
    ctx->coop_launch = value != 0;
  } else if (!strcmp(key, "exchange_bypass")) {
    ctx->exchange_bypass = value != 0;
  } else if (!strcmp(key, "exchange_two_shot")) {  // the two LL forms (peer_exchange.cuh)
    ctx->exchange_protocol = value != 0 ? 2 : 1;
    ctx->comm.force_mode(ctx->exchange_protocol);
    ctx->state_version++;
  } else if (!strcmp(key, "exchange_protocol")) {
    if (value < 0 || value > 3) return fail(ctx, CFMM_ERR_INVALID, "exchange_protocol: 0 (auto), 1, 2 or 3");
    ctx->exchange_protocol = (int)value;
    ctx->comm.force_mode(ctx->exchange_protocol);
    ctx->state_version++;
  } else if (!strcmp(key, "sweep_events")) {
    ctx->sweep_events = value != 0;
  } else if (!strcmp(key, "gradient_math")) {
    ctx->gradient_math = value != 0;
  } else if (!strcmp(key, "geomean_log2")) {
    ctx->geomean_log2 = value != 0;
  } else if (!strcmp(key, "use_tma")) {
    ctx->use_tma = value != 0;
  } else if (!strcmp(key, "debug_skip")) {
    ctx->debug_skip = (int)(value & 7);
  } else if (!strcmp(key, "profile")) {
    // value = number of kernel launches to time with CUDA events (0 = off)
    if (value < 0 || value > (1 << 22)) return fail(ctx, CFMM_ERR_INVALID, "profile out of range");
    cudaSetDevice(ctx->device);
    for (auto e : ctx->prof.ev) cudaEventDestroy(e);
    ctx->prof.ev.assign((size_t)value * 2, nullptr);
    ctx->prof.type.assign((size_t)value, -1);
    ctx->prof.used = 0;
    for (auto& e : ctx->prof.ev) CU_TRY(ctx, cudaEventCreate(&e));
  } else {
    return fail(ctx, CFMM_ERR_INVALID, "unknown option '%s'", key);
  }
  return CFMM_OK;
}

int cfmm_profile_read(cfmm_ctx* ctx, int type, double* total_ms, int64_t* launches) {
  if (!ctx || !total_ms || !launches) return CFMM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  double sum = 0.0;
  int64_t cnt = 0;
  for (size_t i = 0; i < ctx->prof.used; ++i) {
    if (ctx->prof.type[i] != type) continue;
    CU_TRY(ctx, cudaEventSynchronize(ctx->prof.ev[2 * i + 1]));
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->prof.ev[2 * i], ctx->prof.ev[2 * i + 1]));
    sum += ms;
    ++cnt;
  }
  *total_ms = sum;
  *launches = cnt;
  return CFMM_OK;
}

int cfmm_profile_read_times(cfmm_ctx* ctx, int type, float* ms_out, int64_t cap, int64_t* n_out) {
  if (!ctx || !n_out || cap < 0 || (cap > 0 && !ms_out)) return CFMM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  int64_t cnt = 0;
  for (size_t i = 0; i < ctx->prof.used; ++i) {
    if (ctx->prof.type[i] != type) continue;
    if (cnt < cap) {
      CU_TRY(ctx, cudaEventSynchronize(ctx->prof.ev[2 * i + 1]));
      CU_TRY(ctx, cudaEventElapsedTime(&ms_out[cnt], ctx->prof.ev[2 * i], ctx->prof.ev[2 * i + 1]));
    }
    ++cnt;
  }
  *n_out = cnt;
  return CFMM_OK;
}

int cfmm_profile_reset(cfmm_ctx* ctx) {
  if (!ctx) return CFMM_ERR_INVALID;
  ctx->prof.used = 0;
  return CFMM_OK;
}

void* cfmm_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void cfmm_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// Test hook (no CUDA call): the device layout finalize would build for m
// ProductTwoCoin pools.  info[6] = {m_padded, nb, bucketed, skewed, chunk, variant};
// order_out [cap] (device position -> pool index, -1 = padding), chunk_bucket_out
// [cap / chunk], swapped_out [m] are filled when cap >= m_padded (call with cap = 0 first).
int cfmm_debug_product_layout(int64_t n_tokens, int64_t m, const int64_t* Ai, int orient,
                              int variant, int64_t cap, int64_t* order_out,
                              int32_t* chunk_bucket_out, uint8_t* swapped_out, int64_t* info) {
  if (!Ai || !info || m < 0 || n_tokens < 2 || variant < -1 || variant > 0)
    return CFMM_ERR_INVALID;
  for (int64_t i = 0; i < 2 * m; ++i)
    if (Ai[i] < 1 || Ai[i] > n_tokens) return CFMM_ERR_INVALID;
  cfmm_ctx fake;
  fake.n_tokens = n_tokens;
  fake.tma_variant = variant;
  fake.orient_by_degree = orient;
  const cfmm::PoolLayout lay = layout_for(&fake, CFMM_POOL_PRODUCT, Ai, m);
  info[0] = lay.m_padded;
  info[1] = lay.nb;
  info[2] = lay.bucketed;
  info[3] = lay.skewed;
  info[4] = lay.bucketed ? cfmm::kTmaChunk : 0;
  info[5] = variant;
  if (cap >= lay.m_padded && order_out) {
    for (int64_t p = 0; p < lay.m_padded; ++p) order_out[p] = lay.order[(size_t)p];
    if (chunk_bucket_out)
      for (size_t t = 0; t < lay.tile_bucket.size(); ++t) chunk_bucket_out[t] = lay.tile_bucket[t];
    if (swapped_out)
      for (int64_t i = 0; i < m; ++i) swapped_out[i] = lay.swapped[(size_t)i];
  }
  return CFMM_OK;
}

// Measurement hook: the per-CTA phase timestamps (ns, %globaltimer) of the last TMA sweep
// recorded under option "trace": out[8 * grid] = per CTA {entry, slice ready, own range done,
// all chunks done, partials flushed, exit, grid barrier passed (fused exchange), SM id << 32 |
// chunks processed}.
int cfmm_debug_read_trace(cfmm_ctx* ctx, uint64_t* out, int64_t cap_ctas, int64_t* grid_out) {
  if (!ctx || !grid_out) return CFMM_ERR_INVALID;
  if (!ctx->d_trace.n) return fail(ctx, CFMM_ERR_STATE, "option \"trace\" is off");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  *grid_out = ctx->trace_grid;
  if (out && cap_ctas >= ctx->trace_grid && ctx->trace_grid > 0) {
    CU_TRY(ctx, cudaDeviceSynchronize());
    CU_TRY(ctx, cudaMemcpy(out, ctx->d_trace.p, (size_t)ctx->trace_grid * 8 * sizeof(uint64_t),
                           cudaMemcpyDeviceToHost));
  }
  return CFMM_OK;
}

// Test hook: number of (a/b, sqrt a, sqrt b) results, over n host-provided
// operand pairs, where the guard-free in-range recurrences differ from the IEEE
// intrinsics.  Must be 0 for operands in [2^-100, 2^100].
int cfmm_selftest_inrange_math(cfmm_ctx* ctx, const double* a, const double* b, int64_t n,
                               int64_t* mismatches) {
  if (!ctx || !a || !b || !mismatches || n < 0) return CFMM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  DevBuf<double> da, db;
  DevBuf<unsigned long long> dm;
  std::vector<double> ha(a, a + n), hb(b, b + n);
  std::vector<unsigned long long> hz(1, 0);
  CU_TRY(ctx, da.upload(ha));
  CU_TRY(ctx, db.upload(hb));
  CU_TRY(ctx, dm.upload(hz));
  if (n > 0) {
    cfmm::inrange_math_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        da.p, db.p, n, dm.p);
    ctx->launches++;
  }
  unsigned long long out = 0;
  cudaError_t e = cudaMemcpyAsync(&out, dm.p, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  da.release();
  db.release();
  dm.release();
  if (e != cudaSuccess) return fail(ctx, CFMM_ERR_CUDA, "selftest failed: %s", cudaGetErrorString(e));
  *mismatches = (int64_t)out;
  return CFMM_OK;
}

// ---- multi-GPU -----------------------------------------------------------------

int cfmm_comm_export(cfmm_ctx* ctx, void* handle_out) {
  if (!ctx || !handle_out) return CFMM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  static_assert(sizeof(cfmm::PeerHandle) <= CFMM_COMM_HANDLE_BYTES, "handle size");
  memset(handle_out, 0, CFMM_COMM_HANDLE_BYTES);
  if (!ctx->comm.export_handle(ctx->n_tokens + 1, (cfmm::PeerHandle*)handle_out))
    return fail(ctx, CFMM_ERR_COMM, "comm export failed: %s", ctx->comm.error().c_str());
  return CFMM_OK;
}

int cfmm_comm_attach(cfmm_ctx* ctx, int world, int rank, const void* handles) {
  if (!ctx || !handles) return CFMM_ERR_INVALID;
  if (world < 1 || world > cfmm::kMaxPeers || rank < 0 || rank >= world)
    return fail(ctx, CFMM_ERR_INVALID, "bad world/rank %d/%d", rank, world);
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  ctx->state_version++;
  if (!ctx->comm.attach(world, rank, (const unsigned char*)handles,
                        CFMM_COMM_HANDLE_BYTES, ctx->sm_count))
    return fail(ctx, CFMM_ERR_COMM, "comm attach failed: %s", ctx->comm.error().c_str());
  ctx->comm.force_mode(ctx->exchange_protocol);  // an option set before the attach holds
  return CFMM_OK;
}

// For callers of the asynchronous entry points (cfmm_sweep_device*): did any exchange enqueued so
// far give up on a peer?  Synchronises the context's last stream first.
int cfmm_comm_check(cfmm_ctx* ctx) {
  if (!ctx) return CFMM_ERR_INVALID;
  if (!ctx->comm.attached()) return CFMM_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->last_stream ? ctx->last_stream : ctx->stream));
  if (ctx->comm.timed_out())
    return fail(ctx, CFMM_ERR_COMM, "peer exchange timed out: a rank of the group did not deliver its packets");
  return CFMM_OK;
}

int cfmm_comm_detach(cfmm_ctx* ctx) {
  if (!ctx) return CFMM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->comm.detach();
  return CFMM_OK;
}

}  // extern "C"
