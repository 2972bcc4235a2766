/* A plain-C caller of the ABI in include/cfmm_b200.h (no Python, no torch): the
 * README quick-start of CFMMRouter.jl (README.md:25-39) swept at ν = (1, 1) and
 * the easy-arb known answer of test/cfmms.jl:82-86.  Loads the library with
 * dlopen so that the test can also run where no GPU exists (exit code 3).
 *   gcc capi_demo.c -I../../include -ldl -o capi_demo && ./capi_demo ../../cfmmrouter.jl_b200/libcfmm_b200.so */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include "cfmm_b200.h"

#define LOAD(name) \
  __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); \
  if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: capi_demo <libcfmm_b200.so>\n"); return 2; }
  void *lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  LOAD(cfmm_create) LOAD(cfmm_destroy) LOAD(cfmm_last_error) LOAD(cfmm_version)
  LOAD(cfmm_add_product) LOAD(cfmm_finalize) LOAD(cfmm_sweep) LOAD(cfmm_get_trades)
  LOAD(cfmm_num_pools)
  printf("version %s\n", p_cfmm_version());

  cfmm_ctx *ctx = NULL;
  int rc = p_cfmm_create(&ctx, 0, 2);
  if (rc == CFMM_ERR_CUDA) { printf("no device: %s\n", p_cfmm_last_error(NULL)); return 3; }
  if (rc != CFMM_OK) { fprintf(stderr, "create: %s\n", p_cfmm_last_error(NULL)); return 1; }

  /* pools: README equal pool, README unequal pool, and the unit pool of test/cfmms.jl:70 */
  const double R[6] = {1e6, 1e6, 1e3, 2e3, 1.0, 1.0};
  const double gamma[3] = {1.0, 1.0, 1.0};
  const int64_t Ai[6] = {1, 2, 1, 2, 1, 2};
  if (p_cfmm_add_product(ctx, 3, R, gamma, Ai) != CFMM_OK || p_cfmm_finalize(ctx) != CFMM_OK) {
    fprintf(stderr, "ingest: %s\n", p_cfmm_last_error(ctx));
    return 1;
  }
  /* a bad index must be refused with CFMM_ERR_INVALID and a message (BoundsError analogue) */
  const int64_t bad[2] = {1, 3};
  cfmm_ctx *ctx2 = NULL;
  p_cfmm_create(&ctx2, 0, 2);
  rc = p_cfmm_add_product(ctx2, 1, R, gamma, bad);
  printf("bad index rc=%d msg=%s\n", rc, p_cfmm_last_error(ctx2));
  p_cfmm_destroy(ctx2);

  const double v[2] = {2.0, 1.0};
  double psi[2], acc, D[6], L[6];
  if (p_cfmm_sweep(ctx, v, psi, &acc, 1) != CFMM_OK || p_cfmm_get_trades(ctx, D, L) != CFMM_OK) {
    fprintf(stderr, "sweep: %s\n", p_cfmm_last_error(ctx));
    return 1;
  }
  printf("pools %lld\n", (long long)p_cfmm_num_pools(ctx));
  printf("unit pool at v=(2,1): D=(%.17g, %.17g) L=(%.17g, %.17g)\n", D[4], D[5], L[4], L[5]);
  printf("psi=(%.17g, %.17g) acc=%.17g\n", psi[0], psi[1], acc);
  p_cfmm_destroy(ctx);
  return 0;
}
