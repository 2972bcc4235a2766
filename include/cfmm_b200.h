/*
 * cfmm_b200.h -- C ABI of libcfmm_b200.so: the B200-native replacement for the
 * dual-decomposition inner loop of CFMMRouter.jl (reference @ 5932e42).
 *
 * The reference has no FFI; the seam this ABI fills is INSIDE route!
 * (src/router.jl:58-108): the three call sites of find_arb!(r, v)
 * (router.jl:75, 93, 104, 107) and the two fold loops of the L-BFGS-B callback
 * (acc: router.jl:79-83, gradient scatter: router.jl:98-100).  One
 * cfmm_sweep() call = one find_arb!(r, v) over every pool + both folds.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - every function returning int returns CFMM_OK (0) or a negative
 *    cfmm_status; nothing throws across the boundary.  The message for the
 *    last failure is cfmm_last_error(ctx) (ctx may be NULL for failures of
 *    cfmm_create).  Reference behaviour being replaced: Julia exceptions
 *    (ArgumentError from the ctors, src/cfmms.jl:77-78; BoundsError on a bad
 *    token index).
 *  - pointer arguments are caller-owned and borrowed for the duration of the
 *    call only; the context owns all device memory, streams and staging.
 *  - token indices are 1-BASED int64, exactly as Julia's cfmm.Ai
 *    (src/cfmms.jl:15, 87, 108); valid range 1..n_tokens, Ai[1] != Ai[2].
 *  - all values are IEEE fp64 (the only eltype the reference's tests use).
 *  - one context = one GPU = one shard of the pools.  A context is not
 *    thread-safe; use one per Router (the reference's callback is called
 *    synchronously from one thread, src/router.jl:105).
 *  - the library has NO CPU fallback: without a CUDA device every compute
 *    entry point fails with CFMM_ERR_CUDA.
 */
#ifndef CFMM_B200_H
#define CFMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfmm_ctx cfmm_ctx;

typedef enum cfmm_status {
  CFMM_OK = 0,
  CFMM_ERR_INVALID = -1, /* bad argument (ArgumentError / BoundsError analogue) */
  CFMM_ERR_CUDA = -2,    /* CUDA runtime failure or no device */
  CFMM_ERR_STATE = -3,   /* call order violated (e.g. sweep before finalize) */
  CFMM_ERR_NOMEM = -4,
  CFMM_ERR_COMM = -5     /* multi-GPU exchange set-up failure */
} cfmm_status;

typedef enum cfmm_pool_type {
  CFMM_POOL_PRODUCT = 0, /* ProductTwoCoin,       src/cfmms.jl:101-111 */
  CFMM_POOL_GEOMEAN = 1, /* GeometricMeanTwoCoin, src/cfmms.jl:152-165 */
  CFMM_POOL_UNIV3 = 2    /* UniV3,                src/cfmms.jl:226-245 */
} cfmm_pool_type;

/* ---- lifetime ------------------------------------------------------------ */

/* Replaces Router(objective, cfmms, n_tokens) (src/router.jl:18-35) for the
 * device-side state: creates an empty pool set on CUDA device `device`. */
int cfmm_create(cfmm_ctx **out, int device, int64_t n_tokens);
void cfmm_destroy(cfmm_ctx *ctx);
const char *cfmm_last_error(const cfmm_ctx *ctx);
/* "major.minor.patch" of the library */
const char *cfmm_version(void);

/* ---- pool ingest (before cfmm_finalize) ----------------------------------- */
/* Pools are numbered in GLOBAL INSERTION ORDER across all cfmm_add_* calls;
 * that is the order of r.cfmms / r.Δs / r.Λs on the reference side and the
 * order cfmm_get_trades returns. */

/* m x ProductTwoCoin(R, γ, idx) (src/cfmms.jl:101-111).
 * R: [2m] pool-major (R1,R2 per pool); gamma: [m]; Ai: [2m] 1-based. */
int cfmm_add_product(cfmm_ctx *ctx, int64_t m, const double *R,
                     const double *gamma, const int64_t *Ai);

/* m x GeometricMeanTwoCoin(R, w, γ, idx) (src/cfmms.jl:152-165). w: [2m]. */
int cfmm_add_geomean(cfmm_ctx *ctx, int64_t m, const double *R,
                     const double *gamma, const int64_t *Ai, const double *w);

/* m x UniV3(current_price, lower_ticks, liquidity, γ, Ai) (src/cfmms.jl:226-245)
 * in CSR form: pool i owns ticks tick_off[i] .. tick_off[i+1]-1 of
 * lower_ticks / liquidity (tick_off: [m+1], tick_off[0] == 0); lower_ticks
 * strictly decreasing within a pool.  current_tick is computed by the library
 * exactly as the reference ctor does (searchsortedlast rev=true, cfmms.jl:235);
 * a pool whose current_price exceeds its first lower tick (current_tick == 0,
 * a BoundsError in the reference) is rejected with CFMM_ERR_INVALID. */
int cfmm_add_univ3(cfmm_ctx *ctx, int64_t m, const double *current_price,
                   const double *gamma, const int64_t *Ai,
                   const int64_t *tick_off, const double *lower_ticks,
                   const double *liquidity);

/* Flat pool files: the ingest format for large pool sets (replaces building a
 * Vector{CFMM} of heap objects, src/router.jl:18-35, src/cfmms.jl:101-111).  Little-endian,
 * 64-byte header {"CFMMPOOL", u32 version 1, u32 cfmm_pool_type, i64 m, i64 n_tokens, pad},
 * then R [2m] f64, gamma [m] f64, Ai [2m] i64 (1-based), and w [2m] f64 for
 * GeometricMeanTwoCoin -- the arrays cfmm_add_product / cfmm_add_geomean take.
 * cfmm_add_pool_file mmaps the file and adds its pools (same validation, same errors). */
int cfmm_pool_file_write(const char *path, int type, int64_t n_tokens, int64_t m, const double *R,
                         const double *gamma, const int64_t *Ai, const double *w /* geomean */);
int cfmm_pool_file_info(const char *path, int *type, int64_t *n_tokens, int64_t *m);
int cfmm_add_pool_file(cfmm_ctx *ctx, const char *path);

/* Sort each pool type by its first token, lay it out SoA and upload. */
int cfmm_finalize(cfmm_ctx *ctx);

int64_t cfmm_num_pools(const cfmm_ctx *ctx);
int64_t cfmm_num_tokens(const cfmm_ctx *ctx);

/* ---- the hot path --------------------------------------------------------- */

/* One dual-gradient sweep at price vector v (host, [n_tokens]):
 *   find_arb!(r, v)                         src/router.jl:38-42
 *   acc  = Σ_i dot(Λ_i, v[Ai]) - dot(Δ_i, v[Ai])      src/router.jl:79-83
 *   psi  = Σ_i A_i (Λ_i - Δ_i)              src/router.jl:98-100 (G minus grad!(objective))
 * psi_out: host [n_tokens]; acc_out: host scalar.  With materialize != 0 the
 * per-pool trades Δ_i, Λ_i are also kept on the device (the final sweep of
 * route!, router.jl:107) for cfmm_get_trades.  Blocking: results are valid on
 * return.  When the context belongs to a multi-GPU group (cfmm_comm_attach),
 * psi/acc are the sums over all ranks' shards and every rank must call it.
 * Buffers from cfmm_host_alloc (pinned) avoid a staging copy; if acc_out ==
 * psi_out + n_tokens (one contiguous [psi ; acc] buffer) a single copy is used. */
int cfmm_sweep(cfmm_ctx *ctx, const double *v, double *psi_out,
               double *acc_out, int materialize);

/* Same sweep with ν already resident in device memory and the result left
 * there: d_v [n_tokens], d_psi_acc [n_tokens + 1] = [psi ; acc], both device
 * pointers on the context's device.  Enqueued on `stream` (a cudaStream_t;
 * NULL = the context's own stream) WITHOUT synchronising.  Stream rule: consecutive
 * operations of a context depend on each other (ping-pong accumulators, trade buffers,
 * reserves), so whenever an operation runs on a different stream than the previous one --
 * another caller stream, or the context's own stream, which cfmm_sweep, cfmm_get_trades,
 * cfmm_apply_trades, cfmm_update_reserves and cfmm_solve use -- the library makes the new
 * stream wait (event) for the work enqueued on the previous one.  The caller only has to
 * order its OWN reads of d_psi_acc after the sweep on `stream`. */
int cfmm_sweep_device(cfmm_ctx *ctx, const double *d_v, double *d_psi_acc,
                      int materialize, void *stream);

/* Zero-copy form: the result stays in a context-owned device buffer and
 * *d_psi_acc_out receives its address ([n_tokens + 1] fp64).  The buffer is
 * valid until the next sweep on this context is enqueued (two internal
 * accumulators alternate; each sweep clears the other one in-kernel, so no
 * memset or copy is launched). */
int cfmm_sweep_device_view(cfmm_ctx *ctx, const double *d_v, int materialize,
                           void *stream, const double **d_psi_acc_out);

/* Trades of the last materialising sweep, in global insertion order:
 * Delta, Lambda: host [2 * cfmm_num_pools] pool-major (r.Δs[i], r.Λs[i]). */
int cfmm_get_trades(cfmm_ctx *ctx, double *Delta, double *Lambda);

/* Overwrite reserves of pools [first, first+count) of one type, counted in
 * that type's own insertion order; R: [2*count].  (The reference reads
 * cfmm.R live on every sweep, so a caller that mutates reserves between
 * route! calls must push them.)  PRODUCT and GEOMEAN only. */
int cfmm_update_reserves(cfmm_ctx *ctx, int type, int64_t first, int64_t count,
                         const double *R);

/* Apply the trades of the last materialising sweep to the device-resident
 * reserves: R <- R + γΔ − Λ for every ProductTwoCoin / GeometricMeanTwoCoin pool
 * (the update the reference intends with update_reserves!(r), src/router.jl:127-132,
 * whose per-CFMM method is defined nowhere; formula from test/cfmms.jl:10).
 * Lets a caller route repeatedly on evolving state without re-uploading pools.
 * Fails with CFMM_ERR_INVALID if the context holds UniV3 pools. */
int cfmm_apply_trades(cfmm_ctx *ctx);

/* ---- the outer iteration on the device ---------------------------
 * Minimises the dual g(nu) = lin' nu + sum_i arb_i(nu) over the box lower <= nu <= upper
 * -- route! (src/router.jl:58-108) for objectives of the form f(nu) = lin' nu on a box,
 * which both objectives of the reference are (src/objectives.jl:62-79: lin = 0, lower =
 * c + 1e-8; :106-129: lin = Delta_in with lin[i] = 0, lower = sqrt(eps), 1 + sqrt(eps) at
 * i).  nu, the gradient, the L-BFGS history (m = 5) and the search direction stay in device
 * memory; every function/gradient evaluation is one sweep plus vector kernels, and only a
 * few scalars cross PCIe per evaluation.  The optimizer is a projected L-BFGS with Armijo
 * backtracking, not the Fortran L-BFGS-B: same minimiser of the convex dual, different
 * iterates.  On return v_out holds the final nu and the trades at it are materialised
 * (cfmm_get_trades), as after route!.  lin and upper may be NULL (0 / +inf), v0 NULL =
 * ones/n (router.jl:62).  status: 0 projected gradient <= pgtol, 1 relative decrease <=
 * factr*eps or no further move, 2 max_iter, 3 max_fun, 4 line search failed, 5 NaN. */
typedef struct {
  int max_iter, max_fun;
  double pgtol, factr;
} cfmm_solve_opts;
typedef struct {
  int iterations, fun_evals, status;
  double f, pg_norm, solve_ms;
} cfmm_solve_info;
int cfmm_solve(cfmm_ctx *ctx, const double *lin, const double *lower, const double *upper,
               const double *v0, const cfmm_solve_opts *opts /* NULL = route!'s defaults */,
               double *v_out, cfmm_solve_info *info /* may be NULL */);

/* Tunables (none is needed for normal use).
 *   "exact"            1 = evaluate all four closed forms exactly as written in the
 *                      reference for every pool; 0 (default) = evaluate only the side that
 *                      can trade, falling back to the full form near ties (same bits).
 *   "gradient_math"    gradient-only sweeps, where per-pool trades are not observable:
 *                      1 (default) = economized closed forms (ProductTwoCoin: one rsqrt +
 *                      one rcp; GeometricMean: one pow), <= ~3 ulp of the reserve per pool;
 *                      0 = the reference's operation order, bit-identical per pool.
 *                      Materialising sweeps and "exact" are always bit-identical.
 *   "tma_variant"      0 (default) = b-bucketed ProductTwoCoin layout + TMA kernel for
 *                      gradient-only sweeps; -1 = token-sorted layout, first-generation
 *                      kernel only.  Fixes the pool layout: before cfmm_finalize.
 *   "psi_fixed_point"  Ψ[b] partial sums of the TMA kernel: 1 (default) = 64-bit fixed
 *                      point on native 32-bit shared atomics (quantum <= 2^-59 of the
 *                      token's total reserve; used when every token's pools span <= 2^40
 *                      in reserve and totals lie within 2^+-200), 0 = fp64 CAS adds.
 *   "orient_by_degree" store each ProductTwoCoin pool with its higher-degree token first:
 *                      -1 (default) = only when finalize detects hub tokens, 0 never,
 *                      1 always; before cfmm_finalize.
 *   "use_tma"          0 = run the first-generation kernel on the same layout.
 *   "blocks_per_sm"    resident CTAs per SM of the persistent kernels (measurement knob).
 *   "grid_waves"       sweep_kernel (UniV3, materialising sweeps): 1 (default) = one wave of
 *                      resident CTAs striding over the pools, 0 = one CTA per 512 pools (hardware
 *                      block scheduling), N = N waves (measurement knob).
 *   "fused_exchange"   multi-GPU: 1 (default) = product-only sweeps run the peer exchange
 *                      in the sweep kernel's tail; 0 = separate exchange launch.
 *   "coop_launch"      multi-GPU: 1 = the fused sweep+exchange kernel is launched with
 *                      cudaLaunchCooperativeKernel (the driver guarantees that every CTA of its
 *                      grid barrier is resident, or fails the launch); 0 (default) = plain launch,
 *                      residency follows from the occupancy query that sizes the persistent grid,
 *                      and a barrier that cannot complete ends in CFMM_ERR_COMM after the poll
 *                      bound instead of hanging.  Measured at N = 2 on B200: the cooperative
 *                      launch costs +4 to +8 us per step.
 *   "exchange_bypass"  multi-GPU: 1 = sweeps skip the exchange and return this rank's partial
 *                      [psi ; acc] (verification; every rank must set it alike).
 *   "exchange_protocol" multi-GPU: how [psi ; acc] is summed over NVLink peer memory: 3 = direct
 *                      push, one hop, 8 bytes per value (a receive slot is "empty" or a value);
 *                      1 = LL one-shot (16-byte epoch-tagged packets, one hop); 2 = LL two-shot
 *                      (reduce-scatter + all-gather, two hops, fewest bytes); 0 (default) = by
 *                      group size: 3 up to 4 ranks, 2 beyond.  Every rank of the group must use
 *                      the same one.
 *   "exchange_two_shot" multi-GPU: 0 / 1 = "exchange_protocol" 1 / 2.
 *   "sweep_events"     1 = record the two CUDA events cfmm_last_sweep_ms needs around every
 *                      sweep (default 0; turns the sweep graphs off).
 *   "sweep_graphs"     1 (default) = cfmm_sweep replays {H2D nu, sweep, D2H [psi; acc]} as one
 *                      CUDA graph once the same pinned host buffers (cfmm_host_alloc, psi and
 *                      acc contiguous) have been passed twice with unchanged options.
 *   "geomean_log2"     gradient-only GeometricMean sweeps take the power as exp2(e*log2 t)
 *                      (1, default; <= 12 ulp over the admitted range, validated against
 *                      pow on hardware) or as pow (0).
 *   "compact_stream"   1 (default) = economized ProductTwoCoin sweeps stream 24-byte pool records
 *                      (fee through a dictionary of <= 256 distinct values, second token relative
 *                      to its bucket) when the pool set allows it; 0 = 32-byte records.
 *   "geomean_tma"      1 (default) = gradient-only GeometricMeanTwoCoin sweeps run on the TMA
 *                      kernel too (48-byte records, same fixed-point slice); 0 = first-generation
 *                      kernel.
 *   "balance"          TMA kernel: 1 (default) = every CTA's chunk range is sized by its measured
 *                      speed (SMs differ by ~15 %; durations are fed back through mapped pinned
 *                      memory and the range table is re-derived between launches); 0 = even split.
 *   "trace"            1 = TMA sweeps record per-CTA phase timestamps (cfmm_debug_read_trace).
 *   "profile"          N = time the next N kernel launches (cfmm_profile_read). */
int cfmm_set_option(cfmm_ctx *ctx, const char *key, int64_t value);

/* Device time (ms, CUDA events on the sweep stream) of the kernels of the
 * last cfmm_sweep / cfmm_sweep_device call; synchronises the stream. */
int cfmm_last_sweep_ms(cfmm_ctx *ctx, float *ms_out);
/* Number of kernel launches issued by this context so far. */
int64_t cfmm_launch_count(const cfmm_ctx *ctx);

/* Per-kernel device timing.  cfmm_set_option(ctx, "profile", N) arms CUDA-event
 * pairs for the next N kernel launches (recorded on the launching stream,
 * around each sweep kernel / the peer exchange).  cfmm_profile_read sums the
 * durations recorded so far for one pool type (cfmm_pool_type, or 3 = the
 * multi-GPU exchange kernel); it synchronises on the recorded events.
 * cfmm_profile_reset re-arms the same N pairs. */
int cfmm_profile_read(cfmm_ctx *ctx, int type, double *total_ms, int64_t *launches);
/* The individual durations (ms) behind cfmm_profile_read, in launch order: fills
 * ms_out[0 .. min(cap, *n_out)) and sets *n_out to the number recorded for `type`. */
int cfmm_profile_read_times(cfmm_ctx *ctx, int type, float *ms_out, int64_t cap, int64_t *n_out);
int cfmm_profile_reset(cfmm_ctx *ctx);

/* Test hook: counts, over n operand pairs (host arrays), the results of the
 * kernels' guard-free in-range division / square root that differ from IEEE
 * a/b, sqrt(a), sqrt(b).  0 for operands in [2^-100, 2^100]. */
int cfmm_selftest_inrange_math(cfmm_ctx *ctx, const double *a, const double *b,
                               int64_t n, int64_t *mismatches);

/* Test hook, needs no device: the layout cfmm_finalize would give m ProductTwoCoin
 * pools (Ai 1-based [2m]) for "tma_variant" `variant` and orientation mode `orient`.
 * info[6] = {m_padded, bucket width, bucketed?, hubs detected?, chunk size, variant};
 * with cap >= m_padded also order_out [m_padded] (device position -> pool index, -1 =
 * padding), chunk_bucket_out [m_padded / chunk] and swapped_out [m]. */
int cfmm_debug_product_layout(int64_t n_tokens, int64_t m, const int64_t *Ai, int orient,
                              int variant, int64_t cap, int64_t *order_out,
                              int32_t *chunk_bucket_out, uint8_t *swapped_out, int64_t *info);
/* Measurement hook (option "trace" = 1): per-CTA phase timestamps of the last TMA gradient
 * sweep, ns of %globaltimer: out[8 * grid] = {entry, price slice ready, own range done,
 * all chunks done, partials flushed, exit, grid barrier passed (fused exchange, else 0),
 * SM id << 32 | chunks processed} per CTA. */
int cfmm_debug_read_trace(cfmm_ctx *ctx, uint64_t *out, int64_t cap_ctas, int64_t *grid_out);

/* ---- pinned host memory helpers ------------------------------------------- */
void *cfmm_host_alloc(size_t bytes);
void cfmm_host_free(void *p);

/* ---- multi-GPU: one context per GPU, one process per GPU ------------------ */
/* Pools shard across ranks (each rank adds only its shard); the only exchange
 * is the sum of [psi ; acc] (n_tokens+1 fp64) after each sweep.  The exchange
 * runs over NVLink peer memory: every rank exports a handle to its exchange
 * buffer, the host side all-gathers the handles (any transport; the Python
 * host uses torch.distributed), and each rank attaches the others'. */
#define CFMM_COMM_HANDLE_BYTES 128
int cfmm_comm_export(cfmm_ctx *ctx, void *handle_out /* CFMM_COMM_HANDLE_BYTES */);
int cfmm_comm_attach(cfmm_ctx *ctx, int world, int rank,
                     const void *handles /* world * CFMM_COMM_HANDLE_BYTES */);
int cfmm_comm_detach(cfmm_ctx *ctx);
/* After asynchronous sweeps (cfmm_sweep_device*): waits for the stream of the last sweep and
 * returns CFMM_ERR_COMM if any exchange so far hit its poll bound (a rank that never launched
 * the matching sweep, a dead peer); CFMM_OK otherwise and for contexts outside a group.
 * cfmm_sweep makes the same check itself. */
int cfmm_comm_check(cfmm_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* CFMM_B200_H */
