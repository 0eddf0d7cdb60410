/* drive.c -- the C ABI of libtnuts.so exercised from plain C99 (no Python struct mirrors, no C++):
 *   create -> set_model -> set_option -> init -> run -> fetch_draws / fetch -> get_state ->
 *   set_state on a second engine -> run -> bit-identical continuation -> destroy.
 * Build:  gcc -std=c99 -I include tests/c_abi/drive.c -o drive -L turing.jl_b200 -ltnuts -lm
 * Run:    ./drive out.bin     (writes the kept-draw array and the linked-space draws of run 1)
 * tests/test_gpu_cabi.py compiles and runs it on the GPU box and compares out.bin with the oracle;
 * the CPU suite only checks that it compiles and links against every symbol it uses. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tnuts.h"

#define CHECK(e, call)                                                              \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != TNUTS_OK) {                                                          \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, tnuts_last_error(e));           \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

int main(int argc, char **argv) {
  /* eight schools, non-centred (BASELINE.json configs[1]): D = 10 */
  static const double y[8] = {28, 8, -3, 7, -1, 1, 18, 12};
  static const double sigma[8] = {15, 10, 16, 11, 9, 11, 10, 18};
  tnuts_model m;
  memset(&m, 0, sizeof(m));
  m.family = TNUTS_EIGHT_SCHOOLS;
  m.D = 10;
  m.N = 8;
  m.y = y;
  m.sigma = sigma;
  m.hyper[0] = 5.0; /* sd of mu */
  m.hyper[1] = 5.0; /* scale of the half-Cauchy on tau */

  tnuts_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.algorithm = TNUTS_ALG_NUTS;
  cfg.delta = 0.8;
  cfg.max_depth = 10;
  cfg.delta_max = 1000.0;
  cfg.n_adapts = 20;
  cfg.seed = 11;
  cfg.n_chains = 64;
  cfg.metric = TNUTS_METRIC_DIAG;

  if (tnuts_abi_version() != TNUTS_ABI_VERSION) {
    fprintf(stderr, "ABI version mismatch\n");
    return 1;
  }
  tnuts_engine *e = NULL;
  CHECK(NULL, tnuts_create(&cfg, &e));
  CHECK(e, tnuts_set_model(e, &m));
  CHECK(e, tnuts_set_option(e, "keep_theta", 1.0));
  if (tnuts_dimension(e) != 10 || tnuts_num_params(e) != 10) return 1;
  CHECK(e, tnuts_init(e, NULL));

  const int64_t n_tr = 40;
  int64_t kept = 0;
  CHECK(e, tnuts_run(e, n_tr, 1, 1, &kept));
  if (kept != n_tr) return 1;
  const int ncol = tnuts_num_columns(e), C = cfg.n_chains, D = 10;
  double *draws = (double *)malloc(sizeof(double) * (size_t)kept * ncol * C);
  double *theta = (double *)malloc(sizeof(double) * (size_t)kept * D * C);
  double *eps = (double *)malloc(sizeof(double) * C);
  CHECK(e, tnuts_fetch_draws(e, draws));
  CHECK(e, tnuts_fetch(e, NULL, NULL, NULL, theta));
  CHECK(e, tnuts_get_stepsize(e, eps));

  /* LogDensityProblems entry on one point: lp(0) of the model */
  double th0[10] = {0}, lp0 = 0, g0[10];
  CHECK(e, tnuts_logdensity_and_gradient(e, th0, 1, &lp0, g0));

  /* HMCState round trip: a second engine restored from the blob continues bit for bit */
  size_t nb = 0;
  CHECK(e, tnuts_get_state(e, NULL, &nb));
  void *blob = malloc(nb);
  CHECK(e, tnuts_get_state(e, blob, &nb));
  CHECK(e, tnuts_run(e, 10, 1, 1, &kept));
  double *cont = (double *)malloc(sizeof(double) * (size_t)kept * ncol * C);
  CHECK(e, tnuts_fetch_draws(e, cont));
  tnuts_engine *e2 = NULL;
  CHECK(NULL, tnuts_create(&cfg, &e2));
  CHECK(e2, tnuts_set_model(e2, &m));
  CHECK(e2, tnuts_set_state(e2, blob, nb));
  CHECK(e2, tnuts_run(e2, 10, 1, 1, &kept));
  double *cont2 = (double *)malloc(sizeof(double) * (size_t)kept * ncol * C);
  CHECK(e2, tnuts_fetch_draws(e2, cont2));
  /* NaN-safe bit comparison */
  const int same = memcmp(cont, cont2, sizeof(double) * (size_t)kept * ncol * C) == 0;

  printf("ncol %d kept %lld grad_evals %lld launches %lld lp0 %.17g resume_identical %d\n", ncol, (long long)n_tr,
         (long long)tnuts_num_grad_evals(e), (long long)tnuts_num_kernel_launches(e), lp0, same);
  if (argc > 1) {
    FILE *f = fopen(argv[1], "wb");
    if (!f) return 1;
    const int64_t hdr[4] = {n_tr, ncol, C, D};
    fwrite(hdr, sizeof(hdr), 1, f);
    fwrite(draws, sizeof(double), (size_t)n_tr * ncol * C, f);
    fwrite(theta, sizeof(double), (size_t)n_tr * D * C, f);
    fwrite(eps, sizeof(double), (size_t)C, f);
    fwrite(&lp0, sizeof(double), 1, f);
    fclose(f);
  }
  CHECK(e2, tnuts_destroy(e2));
  CHECK(e, tnuts_destroy(e));
  free(draws); free(theta); free(eps); free(blob); free(cont); free(cont2);
  return same ? 0 : 2;
}
