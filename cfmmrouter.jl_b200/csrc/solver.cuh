// solver.cuh -- the outer iteration of route! on the device.
//
// route! (src/router.jl:58-108) minimises the dual g(ν) = f(ν) + Σ_i arb_i(ν) over a box
// with L-BFGS-B; every function/gradient evaluation is one sweep.  With the objective and
// the optimizer on the host, each evaluation moves ν (8n bytes) down and Ψ (8n bytes) up
// over PCIe and costs a stream synchronisation: for BASELINE configs 2-4 that round trip IS
// the step.  Both objectives of the reference (src/objectives.jl:62-79, 106-129) have the
// form f(ν) = linᵀν on the box {lower <= ν <= upper} (LinearNonnegative: lin = 0, lower =
// c + 1e-8; BasketLiquidation: lin = Δin with lin[i] = 0, lower = sqrt(eps), 1 + sqrt(eps)
// at i), so f and grad! are one fused vector kernel, and the optimizer here is a projected
// L-BFGS (two-loop recursion in coefficient space over the basis [S Y pg], Armijo
// backtracking along the projected path): ν, g, the history and the search direction stay
// in HBM; per evaluation only a handful of scalars cross PCIe.
//
// This is NOT the Fortran L-BFGS-B of the reference (Cauchy point + subspace minimisation):
// iterates differ, the minimiser of the convex dual does not.  The reference's tests pin
// nothing at the optimizer boundary beyond feasibility of the resulting trades
// (test/arb.jl:3-28, test/swap.jl:2-46); tests/test_gpu_solver.py restates those and
// compares the optimal value with the host path (scipy L-BFGS-B).
//
// Every reduction is deterministic (fixed-order block sums, then a fixed-order sum of the
// block partials by the last block): replicated multi-GPU drivers that see bitwise-equal
// [Ψ; acc] take bitwise-equal decisions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cfmm {

constexpr int kSolverM = 5;               // L-BFGS history (the reference's default m = 5)
constexpr int kSolverK = 2 * kSolverM + 1;  // basis columns: s_1..s_m, y_1..y_m, projected gradient
constexpr int kSolverGram = kSolverK * (kSolverK + 1) / 2;
constexpr int kSolverThreads = 256;
constexpr int kSolverMaxBlocks = 512;

struct SolverVecs {
  double* x;        // current iterate
  double* g;        // its gradient (lin + Ψ)
  double* xt;       // trial point
  double* gt;       // gradient at the trial point
  double* d;        // search direction
  double* pg;       // projected gradient at x
  double* S;        // [kSolverM][n]
  double* Y;        // [kSolverM][n]
  const double* lin;    // linear objective term (null = 0)
  const double* lower;  // box
  const double* upper;  // null = +inf
  double* partials;     // [kSolverMaxBlocks][kSolverGram] block partial sums
  double* scal;         // [kSolverGram + 8] results
  unsigned* ticket;     // last-block election
  int64_t n;
};

// deterministic grid reduction of NV per-thread values: block tree sum, then the last block
// to arrive sums the block partials in block order into out[0..NV)
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double* partials, double* out, unsigned* ticket) {
  __shared__ double sm[kSolverThreads / 32][NV];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) sm[warp][k] = t;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kSolverThreads / 32; ++w) t += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * NV + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x < NV) {
      double t = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(partials + (size_t)b * NV + threadIdx.x);
      out[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// Result slots in scal: [0, kSolverGram) Gram matrix; +0 gᵀ(xt − x), +1 linᵀxt, +2 |xt − x|²
// (trial); +3 gᵀd, +4 |d|² (direction).
// xt = P(x + t d)
__global__ void __launch_bounds__(kSolverThreads) solver_trial_kernel(SolverVecs q, double t) {
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kSolverThreads + threadIdx.x; i < q.n; i += (int64_t)gridDim.x * kSolverThreads) {
    const double x = q.x[i];
    double y = fma(t, q.d[i], x);
    y = fmax(y, q.lower[i]);
    if (q.upper) y = fmin(y, q.upper[i]);
    q.xt[i] = y;
    const double dx = y - x;
    v[0] = fma(q.g[i], dx, v[0]);
    if (q.lin) v[1] = fma(q.lin[i], y, v[1]);
    v[2] = fma(dx, dx, v[2]);
  }
  grid_reduce<3>(v, q.partials, q.scal + kSolverGram, q.ticket);
}

// gt = lin + Ψ(xt)  (grad!, src/objectives.jl:70-77, 115-121, plus the scatter of router.jl:98-100,
// which the sweep has already folded into psi)
__global__ void __launch_bounds__(kSolverThreads) solver_grad_kernel(SolverVecs q, const double* __restrict__ psi) {
  for (int64_t i = (int64_t)blockIdx.x * kSolverThreads + threadIdx.x; i < q.n; i += (int64_t)gridDim.x * kSolverThreads)
    q.gt[i] = (q.lin ? q.lin[i] : 0.0) + psi[i];
}

// Accept the trial point: (s, y) -> history slot (when store != 0), x <- xt, g <- gt, the
// projected gradient pg, and the Gram matrix W = BᵀB of B = [S Y pg] (upper triangle, row
// major) -> scal[0..kSolverGram); scal[kSolverGram] = |pg|_inf goes through a max-reduction
// of its own (exact, order-independent).
__global__ void __launch_bounds__(kSolverThreads)
    solver_commit_kernel(SolverVecs q, int slot, int store, unsigned long long* pgmax_bits) {
  double w[kSolverGram];
#pragma unroll
  for (int k = 0; k < kSolverGram; ++k) w[k] = 0.0;
  double pgmax = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kSolverThreads + threadIdx.x; i < q.n; i += (int64_t)gridDim.x * kSolverThreads) {
    const double xn = q.xt[i], gn = q.gt[i];
    if (store) {
      q.S[(size_t)slot * q.n + i] = xn - q.x[i];
      q.Y[(size_t)slot * q.n + i] = gn - q.g[i];
    }
    q.x[i] = xn;
    q.g[i] = gn;
    // projected gradient (what L-BFGS-B's pgtol test uses): the move x − g clipped to the box
    double p = gn;
    if (xn <= q.lower[i] && gn > 0.0) p = 0.0;
    if (q.upper && xn >= q.upper[i] && gn < 0.0) p = 0.0;
    q.pg[i] = p;
    pgmax = fmax(pgmax, fabs(p));
    double b[kSolverK];
#pragma unroll
    for (int j = 0; j < kSolverM; ++j) {
      b[j] = q.S[(size_t)j * q.n + i];
      b[kSolverM + j] = q.Y[(size_t)j * q.n + i];
    }
    b[kSolverK - 1] = p;
    int k = 0;
#pragma unroll
    for (int r = 0; r < kSolverK; ++r)
#pragma unroll
      for (int c = r; c < kSolverK; ++c) w[k] = fma(b[r], b[c], w[k]), ++k;
  }
  // |pg|_inf: non-negative doubles order like their bit patterns
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pgmax = fmax(pgmax, __shfl_down_sync(0xffffffffu, pgmax, o));
  if ((threadIdx.x & 31) == 0) atomicMax(pgmax_bits, (unsigned long long)__double_as_longlong(pgmax));
  grid_reduce<kSolverGram>(w, q.partials, q.scal, q.ticket);
}

// d = −B c on the free variables (0 where the bound is active and the gradient pushes out);
// out: [0] = gᵀd, [1] = |d|²
struct SolverCoef {
  double c[kSolverK];
};
__global__ void __launch_bounds__(kSolverThreads) solver_direction_kernel(SolverVecs q, SolverCoef cf) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kSolverThreads + threadIdx.x; i < q.n; i += (int64_t)gridDim.x * kSolverThreads) {
    double r = cf.c[kSolverK - 1] * q.pg[i];
#pragma unroll
    for (int j = 0; j < kSolverM; ++j) {
      r = fma(cf.c[j], q.S[(size_t)j * q.n + i], r);
      r = fma(cf.c[kSolverM + j], q.Y[(size_t)j * q.n + i], r);
    }
    const double x = q.x[i], g = q.g[i];
    double d = -r;
    if (x <= q.lower[i] && g > 0.0) d = 0.0;
    if (q.upper && x >= q.upper[i] && g < 0.0) d = 0.0;
    q.d[i] = d;
    v[0] = fma(g, d, v[0]);
    v[1] = fma(d, d, v[1]);
  }
  grid_reduce<2>(v, q.partials, q.scal + kSolverGram + 3, q.ticket);
}

// x = P(v0), and linᵀx -> scal[kSolverGram + 1]
__global__ void __launch_bounds__(kSolverThreads) solver_init_kernel(SolverVecs q, const double* __restrict__ v0) {
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kSolverThreads + threadIdx.x; i < q.n; i += (int64_t)gridDim.x * kSolverThreads) {
    double y = fmax(v0[i], q.lower[i]);
    if (q.upper) y = fmin(y, q.upper[i]);
    q.x[i] = y;
    q.xt[i] = y;
    if (q.lin) v[1] = fma(q.lin[i], y, v[1]);
  }
  grid_reduce<3>(v, q.partials, q.scal + kSolverGram, q.ticket);
}

}  // namespace cfmm
