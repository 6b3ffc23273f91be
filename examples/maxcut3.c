/* maxcut3.c — the C ABI of libclrs_b200.so from plain C: the Goemans-Williamson MAX-CUT relaxation of the 3-cycle
 * (README.md:39-72 of the reference: optimum 9/4), uploaded once as dense constraint matrices and once as triplets
 * with the sparsity shortcut, solved to duality gap 1e-30 by calling clrs_iterate in a loop exactly like the Julia shim of
 * INTEGRATION.md.
 *
 *   gcc -std=c99 -I include examples/maxcut3.c -o maxcut3 -L clusteredlowranksolver.jl_b200/csrc -lclrs_b200 \
 *       -Wl,-rpath,$PWD/clusteredlowranksolver.jl_b200/csrc
 *
 * Wire numbers (include/clrs_b200.h): int64 exp; int32 sign; int32 0; uint64 limb[W]; value = sign * 0.limbs * 2^exp. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "clrs_b200.h"

enum { PREC = 256, W = PREC / 64, WS = 16 + 8 * W, N = 3 };

static void put(unsigned char* rec, double v) {          /* exact: a double has 53 mantissa bits */
  memset(rec, 0, WS);
  if (v == 0.0) return;
  int e; double m = frexp(fabs(v), &e);                    /* |v| = m * 2^e, m in [1/2, 1) */
  int64_t ex = e; int32_t sg = v < 0 ? -1 : 1;
  uint64_t top = (uint64_t)ldexp(m, 64);                   /* the leading 64 bits of the mantissa (top bit set) */
  memcpy(rec, &ex, 8); memcpy(rec + 8, &sg, 4); memcpy(rec + 16 + 8 * (W - 1), &top, 8);
}
static double get(const unsigned char* rec) {
  int64_t ex; int32_t sg; uint64_t top; memcpy(&ex, rec, 8); memcpy(&sg, rec + 8, 4); memcpy(&top, rec + 16 + 8 * (W - 1), 8);
  return sg == 0 ? 0.0 : sg * ldexp((double)top, (int)ex - 64);
}
#define CHECK(call) do { int rc_ = (call); if (rc_ != CLRS_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, clrs_last_error(h)); return 1; } } while (0)

static int solve(int sparse) {
  clrs_options opt; clrs_default_options(&opt);
  opt.prec = PREC; opt.duality_gap_threshold = 1e-30; opt.sparse_schur = sparse;
  clrs_handle* h = NULL;
  if (clrs_create(&opt, &h) != CLRS_OK) { fprintf(stderr, "clrs_create: %s\n", h ? clrs_last_error(h) : "?"); clrs_destroy(h); return 1; }
  unsigned char zero[WS], one[WS], C[N * N * WS], c[N * WS], A[N * N * WS];
  put(zero, 0.0); put(one, 1.0);
  for (int i = 0; i < N; i++) { put(c + i * WS, 1.0); for (int j = 0; j < N; j++) put(C + (i * N + j) * WS, i == j ? 0.5 : -0.25); }   /* L / 4 */
  CHECK(clrs_set_free(h, 0, zero, zero, /*maximize*/ 1));
  CHECK(clrs_add_cluster(h, 0, N, zero /* B is 3 x 0 */, c));
  CHECK(clrs_add_block(h, 0, 0, 1, N, /*high_rank*/ 1, C));
  for (int p = 0; p < N; p++) {                             /* <E_pp, X> = 1 */
    if (sparse) { int32_t r = p, cc = p; CHECK(clrs_add_sparse_term(h, 0, 0, p, 1, &r, &cc, one, 0)); }
    else { for (int k = 0; k < N * N; k++) put(A + k * WS, k == p * N + p ? 1.0 : 0.0); CHECK(clrs_add_dense_term(h, 0, 0, p, A)); }
  }
  CHECK(clrs_finalize(h));
  clrs_iter_info info; int it = 0;
  for (; it < 200; it++) { CHECK(clrs_iterate(h, &info)); if (info.stop != CLRS_CONTINUE) break; }
  unsigned char d[WS], p[WS], g[WS];
  CHECK(clrs_get_objectives(h, d, p, g));
  printf("%s upload: stop=%d after %d iterations, dual %.15f primal %.15f gap %.3e\n", sparse ? "triplet" : "dense", info.stop, it, get(d), get(p), get(g));
  const int ok = info.stop == CLRS_STOP_OPTIMAL && fabs(get(p) - 2.25) < 1e-14 && get(g) < 1e-30;
  clrs_destroy(h);
  return ok ? 0 : 2;
}
int main(void) { int rc = solve(0); if (rc == 0) rc = solve(1); if (rc == 0) puts("MAXCUT3 OK"); return rc; }
