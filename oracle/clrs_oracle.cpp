// clrs_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product path).
//
// A plain restatement, on MPFR, of the interior-point iteration of
// nanleij/ClusteredLowRankSolver.jl (src/solver.jl:100-744 and helpers
// :792-1693, src/tools.jl:11-107).  Only tests/, __graft_entry__.smoke() and
// the cpu_baseline / --impl reference legs of bench.py may load this library.
//
// Pinning.  Julia and FLINT/Arb (the reference's arithmetic, reached through
// Arblib.jl, Project.toml:7, unpinned: no Manifest) are absent from this
// image, so the reference itself cannot be run.  The oracle is pinned against
// the known-answer values the reference's own tests and README hold for this
// path: MAX-CUT of the 3-cycle = 9/4 (README.md:70-72),
// min(x^2+1) = 1 (README.md:146-150), Delsarte n=8 = 240
// (test/runtests_solver.jl:86-87), delsarte(3,10,1/2) = 13.158314 (:15),
// 2-radii sphere packing n=8 ~ pi^4/384 (:19-22); see tests/test_oracle_pins.py.
// Kernel-level results are not pinned by any reference test.
//
// Arithmetic: MPFR round-to-nearest at `prec` bits for every operation
// (Arb's approx_* kernels are floating point at prec bits with unspecified
// last-bit rounding).  MPFR headers are not installed;
// the few prototypes used are declared by hand against libmpfr.so.6 (4.2.1).
#include <omp.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <cmath>
#include <vector>
#include <map>
#include <tuple>
#include <string>
#include <algorithm>
#include <chrono>

extern "C" {
typedef struct { long _prec; int _sign; long _exp; unsigned long* _d; } mpfr_s;
typedef mpfr_s* mpfr_p; typedef const mpfr_s* mpfr_cp;
int mpfr_set4(mpfr_p, mpfr_cp, int, int);
int mpfr_set_d(mpfr_p, double, int);
int mpfr_set_si(mpfr_p, long, int);
double mpfr_get_d(mpfr_cp, int);
int mpfr_add(mpfr_p, mpfr_cp, mpfr_cp, int);
int mpfr_sub(mpfr_p, mpfr_cp, mpfr_cp, int);
int mpfr_mul(mpfr_p, mpfr_cp, mpfr_cp, int);
int mpfr_div(mpfr_p, mpfr_cp, mpfr_cp, int);
int mpfr_sqrt(mpfr_p, mpfr_cp, int);
int mpfr_neg(mpfr_p, mpfr_cp, int);
int mpfr_cmp3(mpfr_cp, mpfr_cp, int);
int mpfr_mul_2si(mpfr_p, mpfr_cp, long, int);
}
static const int RN = 0;                       // MPFR_RNDN
static const long EXP_ZERO = LONG_MIN + 1;     // __MPFR_EXP_ZERO

static int g_prec = 256;
static int g_w = 4;                            // 64-bit limbs per number

// ---------------------------------------------------------------------------
// numbers and matrices
// ---------------------------------------------------------------------------
struct Num {                                   // one owned number
  mpfr_s v; std::vector<unsigned long> d;
  Num() : d(g_w, 0) { v._prec = g_prec; v._sign = 1; v._exp = EXP_ZERO; v._d = d.data(); }
  Num(const Num& o) : d(o.d) { v = o.v; v._d = d.data(); }
  Num& operator=(const Num& o) { d = o.d; v = o.v; v._d = d.data(); return *this; }
  mpfr_p p() { return &v; } mpfr_cp p() const { return &v; }
};
static inline bool is_zero(mpfr_cp a) { return a->_exp == EXP_ZERO; }
static inline int sgn(mpfr_cp a) { return is_zero(a) ? 0 : a->_sign; }
static inline void set(mpfr_p a, mpfr_cp b) { mpfr_set4(a, b, RN, b->_sign); }
static inline void set_abs(mpfr_p a, mpfr_cp b) { mpfr_set4(a, b, RN, 1); }
static inline int cmp(mpfr_cp a, mpfr_cp b) { return mpfr_cmp3(a, b, 1); }

struct Mat {                                   // row-major matrix of numbers
  int r = 0, c = 0; std::vector<mpfr_s> e; std::vector<unsigned long> d;
  Mat() {}
  Mat(int r_, int c_) { init(r_, c_); }
  void init(int r_, int c_) {
    r = r_; c = c_; e.assign((size_t)r * c, mpfr_s()); d.assign((size_t)r * c * g_w, 0);
    for (size_t i = 0; i < e.size(); i++) { e[i]._prec = g_prec; e[i]._sign = 1; e[i]._exp = EXP_ZERO; e[i]._d = d.data() + i * g_w; }
  }
  Mat(const Mat& o) { *this = o; }
  Mat& operator=(const Mat& o) {
    r = o.r; c = o.c; e = o.e; d = o.d;
    for (size_t i = 0; i < e.size(); i++) e[i]._d = d.data() + i * g_w;
    return *this;
  }
  mpfr_p operator()(int i, int j) { return &e[(size_t)i * c + j]; }
  mpfr_cp operator()(int i, int j) const { return &e[(size_t)i * c + j]; }
  void zero() { for (auto& x : e) { x._exp = EXP_ZERO; x._sign = 1; } }
};

// wire format (include/clrs_b200.h): int64 exp; int32 sign; int32 0; uint64 limb[W]
static size_t wire_size() { return 16 + 8 * (size_t)g_w; }
static void from_wire(mpfr_p a, const void* src) {
  const char* s = (const char*)src; int64_t ex; int32_t sg;
  memcpy(&ex, s, 8); memcpy(&sg, s + 8, 4);
  if (sg == 0) { a->_exp = EXP_ZERO; a->_sign = 1; return; }
  a->_sign = sg < 0 ? -1 : 1; a->_exp = (long)ex; memcpy(a->_d, s + 16, 8 * (size_t)g_w);
}
static void to_wire(void* dst, mpfr_cp a) {
  char* s = (char*)dst; memset(s, 0, wire_size());
  if (is_zero(a)) return;
  int64_t ex = a->_exp; int32_t sg = a->_sign < 0 ? -1 : 1;
  memcpy(s, &ex, 8); memcpy(s + 8, &sg, 4); memcpy(s + 16, a->_d, 8 * (size_t)g_w);
}
static void mat_from_wire(Mat& m, const void* src) { for (size_t i = 0; i < m.e.size(); i++) from_wire(&m.e[i], (const char*)src + i * wire_size()); }
static void mat_to_wire(void* dst, const Mat& m) { for (size_t i = 0; i < m.e.size(); i++) to_wire((char*)dst + i * wire_size(), &m.e[i]); }

// ---------------------------------------------------------------------------
// dense kernels (restating the Arb calls of SURVEY.md §2a)
// ---------------------------------------------------------------------------
// C = A*B  (approx_mul!, matmul_threaded!  src/tools.jl:175-209)
static void gemm(Mat& C, const Mat& A, const Mat& B) {
  const int M = A.r, K = A.c, N = B.c;
  if (C.r != M || C.c != N) C.init(M, N);
#pragma omp parallel
  { Num t, acc;
#pragma omp for schedule(dynamic, 1) collapse(2)
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
      acc.v._exp = EXP_ZERO; acc.v._sign = 1;
      for (int k = 0; k < K; k++) { mpfr_mul(t.p(), A(i, k), B(k, j), RN); mpfr_add(acc.p(), acc.p(), t.p(), RN); }
      set(C(i, j), acc.p());
    } }
}
// C = A^T * B
static void gemm_tn(Mat& C, const Mat& A, const Mat& B) {
  const int M = A.c, K = A.r, N = B.c;
  if (C.r != M || C.c != N) C.init(M, N);
#pragma omp parallel
  { Num t, acc;
#pragma omp for schedule(dynamic, 1) collapse(2)
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
      acc.v._exp = EXP_ZERO; acc.v._sign = 1;
      for (int k = 0; k < K; k++) { mpfr_mul(t.p(), A(k, i), B(k, j), RN); mpfr_add(acc.p(), acc.p(), t.p(), RN); }
      set(C(i, j), acc.p());
    } }
}
// <A,B> (LinearAlgebra.dot, src/tools.jl:25-35)
static void dot(mpfr_p res, const Mat& A, const Mat& B) {
  Num t; res->_exp = EXP_ZERO; res->_sign = 1;
  for (size_t i = 0; i < A.e.size(); i++) { mpfr_mul(t.p(), &A.e[i], &B.e[i], RN); mpfr_add(res, res, t.p(), RN); }
}
// max |a_ij|  (compute_error, src/solver.jl:816-825)
static void max_abs(mpfr_p res, const Mat& A) {
  Num t;
  for (size_t i = 0; i < A.e.size(); i++) { set_abs(t.p(), &A.e[i]); if (cmp(t.p(), res) > 0) set(res, t.p()); }
}
// approx_cholesky!  (src/tools.jl:75-107): lower factor in place, strict upper
// zeroed; returns 0 when a pivot is not strictly positive.  The per-element
// operation order is the reference's (subtract k = 1..j-1 in order, divide);
// only the loop nest is column-oriented so rows can run in parallel.
static int cholesky(Mat& A) {
  const int n = A.r;
  for (int j = 0; j < n; j++) {
    { Num t;
      for (int k = 0; k < j; k++) { mpfr_mul(t.p(), A(j, k), A(j, k), RN); mpfr_sub(A(j, j), A(j, j), t.p(), RN); }
      if (sgn(A(j, j)) <= 0) return 0;
      mpfr_sqrt(A(j, j), A(j, j), RN); }
#pragma omp parallel
    { Num t;
#pragma omp for schedule(static)
      for (int i = j + 1; i < n; i++) {
        for (int k = 0; k < j; k++) { mpfr_mul(t.p(), A(i, k), A(j, k), RN); mpfr_sub(A(i, j), A(i, j), t.p(), RN); }
        mpfr_div(A(i, j), A(i, j), A(j, j), RN);
      } }
  }
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) { A(i, j)->_exp = EXP_ZERO; A(i, j)->_sign = 1; }
  return 1;
}
// X = L^-1 B  (approx_solve_tril!), columns independent
static void solve_lower(Mat& X, const Mat& L, const Mat& B) {
  const int n = L.r, m = B.c; if (&X != &B) X = B;
#pragma omp parallel
  { Num t;
#pragma omp for schedule(static)
    for (int c = 0; c < m; c++) for (int i = 0; i < n; i++) {
      for (int k = 0; k < i; k++) { mpfr_mul(t.p(), L(i, k), X(k, c), RN); mpfr_sub(X(i, c), X(i, c), t.p(), RN); }
      mpfr_div(X(i, c), X(i, c), L(i, i), RN);
    } }
}
// X = L^-T B  (approx_solve_triu! on the transposed factor, src/solver.jl:1567-1572)
static void solve_lower_t(Mat& X, const Mat& L, const Mat& B) {
  const int n = L.r, m = B.c; if (&X != &B) X = B;
#pragma omp parallel
  { Num t;
#pragma omp for schedule(static)
    for (int c = 0; c < m; c++) for (int i = n - 1; i >= 0; i--) {
      for (int k = i + 1; k < n; k++) { mpfr_mul(t.p(), L(k, i), X(k, c), RN); mpfr_sub(X(i, c), X(i, c), t.p(), RN); }
      mpfr_div(X(i, c), X(i, c), L(i, i), RN);
    } }
}
// X = (L L^T)^-1 B  (solve_cho_precomp!)
static void solve_cho(Mat& X, const Mat& L, const Mat& B) { solve_lower(X, L, B); solve_lower_t(X, L, X); }
// (L L^T)^-1  (inv_cho_precomp!)
static void inv_cho(Mat& X, const Mat& L) {
  Mat I(L.r, L.r); for (int i = 0; i < L.r; i++) mpfr_set_si(I(i, i), 1, RN);
  solve_cho(X, L, I);
}
static void symmetrize_half(Mat& A) {           // A = (A + A^T)/2  (src/solver.jl:1509-1511)
  const int n = A.r;
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) {
    mpfr_add(A(i, j), A(i, j), A(j, i), RN); mpfr_mul_2si(A(i, j), A(i, j), -1, RN); set(A(j, i), A(i, j));
  }
}

// smallest eigenvalue of a symmetric double matrix: Householder
// tridiagonalisation + Sturm bisection.  Stands in for KrylovKit's Float64
// Lanczos (src/solver.jl:1659): accurate to ~1e-13, well inside its tol 1e-5.
static double min_eig_sym(std::vector<double> a, int n) {
  if (n == 1) return a[0];
  std::vector<double> d(n), e(n, 0.0);
  for (int k = 0; k < n - 2; k++) {
    double alpha = 0; for (int i = k + 1; i < n; i++) alpha += a[i * n + k] * a[i * n + k];
    alpha = std::sqrt(alpha); if (alpha == 0) continue;
    if (a[(k + 1) * n + k] > 0) alpha = -alpha;
    std::vector<double> v(n, 0.0);
    for (int i = k + 1; i < n; i++) v[i] = a[i * n + k];
    v[k + 1] -= alpha;
    double vn = 0; for (int i = k + 1; i < n; i++) vn += v[i] * v[i];
    if (vn == 0) continue;
    std::vector<double> p(n, 0.0);
    for (int i = 0; i < n; i++) { double s = 0; for (int j = k + 1; j < n; j++) s += a[i * n + j] * v[j]; p[i] = 2 * s / vn; }
    double K = 0; for (int i = k + 1; i < n; i++) K += v[i] * p[i]; K /= vn;
    for (int i = 0; i < n; i++) p[i] -= K * v[i];
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i * n + j] -= v[i] * p[j] + p[i] * v[j];
  }
  for (int i = 0; i < n; i++) d[i] = a[i * n + i];
  for (int i = 0; i < n - 1; i++) e[i] = a[(i + 1) * n + i];
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < n; i++) { double r = (i > 0 ? std::fabs(e[i - 1]) : 0) + (i < n - 1 ? std::fabs(e[i]) : 0); lo = std::min(lo, d[i] - r); hi = std::max(hi, d[i] + r); }
  auto count_below = [&](double x) { int cnt = 0; double q = 1; for (int i = 0; i < n; i++) { double off = i > 0 ? e[i - 1] * e[i - 1] : 0; q = d[i] - x - (i > 0 ? off / q : 0); if (q == 0) q = 1e-300; if (q < 0) cnt++; } return cnt; };
  for (int it = 0; it < 200 && hi - lo > 1e-15 * std::max(1.0, std::max(std::fabs(lo), std::fabs(hi))); it++) { double mid = 0.5 * (lo + hi); if (count_below(mid) >= 1) hi = mid; else lo = mid; }
  return 0.5 * (lo + hi);
}

// ---------------------------------------------------------------------------
// problem container (ClusteredLowRankSDP, src/interface.jl:807-819)
// ---------------------------------------------------------------------------
struct LRTerm { int r, s, p, k; Num lambda; Mat v, w; int colV = -1, rowW = -1; };  // one rank-one piece lambda * v w^T of A[r,s][p]
struct Block {
  int m = 1, delta = 1, n = 1; bool high_rank = false; Mat C;
  std::vector<int> dense_p; std::vector<Mat> dense_A;
  std::vector<std::vector<int>> dense_cols;                   // nonzero columns of A_p   (zero-skipping mode)
  std::vector<std::vector<std::pair<int, int>>> dense_nz;     // nonzero entries of A_p, row-major
  std::vector<LRTerm> lr;                                   // insertion order
  std::vector<std::vector<std::vector<int>>> rs;            // rs[r][s] -> term indices in order
  std::map<std::tuple<int, int, int, int>, int> find;       // (r,s,p,k) -> term
  std::vector<Mat> V, W;                                    // V[r]: delta x u_r ; W[r]: u'_r x delta (deduplicated)
  std::vector<std::vector<Mat>> A_Y;                        // A_Y[r][s], s<=r
  // state
  Mat X, Y, Xinv /*chol(X)*/, R, P, dX, dY;
};
struct Cluster { int P = 0; Mat B, c; std::vector<Block> blocks; Mat S, LinvB; };

struct Options {
  Num beta_infeasible, beta_feasible, gamma, omega_p, omega_d, gap_thr, derr_thr, perr_thr, max_comp_gap, step_thr;
  bool need_dual = false, need_primal = false, safe_step = true, correctoronly = false;
};

struct Info {   // mirrors clrs_iter_info (include/clrs_b200.h)
  int32_t iter, stop, pd_feasible, reserved;
  double mu, d_obj, p_obj, gap, err_P, err_p, err_d, alpha_d, alpha_p, beta_c, d_obj_new, p_obj_new, gap_new;
  double phase_ms[17];
};

struct Oracle {
  Options opt; std::string err;
  bool maximize = true; Num constant; int N = 0; Mat b;
  std::vector<Cluster> cl;
  Mat x, y, d, p, dx, dy, Q;
  int K = 0, Ptot = 0; std::vector<int> off;
  int iter = 1; bool pd_feas = false;
  Num d_obj, p_obj, gap, dual_error, primal_error;
  bool finalized = false;
  // optional sampling of the dense Schur path for the bounded CPU baseline
  int dense_p_limit = -1;
  // Zero-skipping evaluation of the dense Schur path.  Bit-identical to the plain path: every
  // skipped operation is an exact `x + 0*y` (tests/test_oracle.py checks S bit for bit); it only
  // makes full-size MAX-CUT instances (A_p = E_pp stored dense) affordable on the CPU.
  bool dense_skip_zeros = false;
  // bench.py reference arm: rows p < sample_limit of the dense Schur path run the plain dense algorithm
  // (timed in t_plain), the others the bit-identical zero-skipping one (timed in t_skip).
  int sample_limit = -1; double t_plain = 0, t_skip = 0; int t_np = 0;

  template <class F> void for_blocks(F f) { for (auto& c : cl) for (auto& b : c.blocks) f(c, b); }

  // ---- objectives (src/solver.jl:792-860) ----
  void dual_objective(mpfr_p r) {
    Num t; r->_exp = EXP_ZERO; r->_sign = 1; int idx = 0;
    for (auto& c : cl) for (int i = 0; i < c.P; i++, idx++) { mpfr_mul(t.p(), c.c(i, 0), x(idx, 0), RN); mpfr_add(r, r, t.p(), RN); }
    if (!maximize) mpfr_neg(r, r, RN);
    mpfr_add(r, r, constant.p(), RN);
  }
  void primal_objective(mpfr_p r) {
    Num t, s; r->_exp = EXP_ZERO; r->_sign = 1;
    for_blocks([&](Cluster&, Block& b) { dot(t.p(), b.C, b.Y); mpfr_add(r, r, t.p(), RN); });
    if (N > 0) { dot(s.p(), b, y); mpfr_add(r, r, s.p(), RN); }
    mpfr_add(r, r, constant.p(), RN);
  }
  void duality_gap(mpfr_p g, mpfr_cp dobj, mpfr_cp pobj) {
    Num a, bsum, one; mpfr_sub(a.p(), dobj, pobj, RN); set_abs(a.p(), a.p());
    mpfr_add(bsum.p(), dobj, pobj, RN); set_abs(bsum.p(), bsum.p()); mpfr_set_si(one.p(), 1, RN);
    if (cmp(bsum.p(), one.p()) < 0) set(bsum.p(), one.p());
    mpfr_div(g, a.p(), bsum.p(), RN);
  }

  // ---- setup (precompute_matrices_bilinear_pairings, src/solver.jl:985-1059) ----
  static bool same_vec(const Mat& a, const Mat& b) {
    if (a.e.size() != b.e.size()) return false;
    for (size_t i = 0; i < a.e.size(); i++) if (cmp(&a.e[i], &b.e[i]) != 0) return false;
    return true;
  }
  int finalize() {
    off.assign(cl.size() + 1, 0); K = 0;
    for (size_t j = 0; j < cl.size(); j++) off[j + 1] = off[j] + cl[j].P;
    Ptot = off.back();
    for (auto& c : cl) for (auto& b : c.blocks) {
      K += b.n;
      if (b.high_rank) continue;
      b.rs.assign(b.m, std::vector<std::vector<int>>(b.m));
      for (size_t e = 0; e < b.lr.size(); e++) { auto& t = b.lr[e]; b.rs[t.r][t.s].push_back((int)e); b.find[{t.r, t.s, t.p, t.k}] = (int)e; }
      b.V.resize(b.m); b.W.resize(b.m);
      for (int r = 0; r < b.m; r++) {
        std::vector<int> uv, uw;                 // representatives of the unique vectors (unique_idx, src/tools.jl:128-145)
        for (int s = 0; s < b.m; s++) for (int e : b.rs[r][s]) {
          int f = -1; for (size_t u = 0; u < uv.size(); u++) if (same_vec(b.lr[uv[u]].v, b.lr[e].v)) { f = (int)u; break; }
          if (f < 0) { f = (int)uv.size(); uv.push_back(e); } b.lr[e].colV = f;
          f = -1; for (size_t u = 0; u < uw.size(); u++) if (same_vec(b.lr[uw[u]].w, b.lr[e].w)) { f = (int)u; break; }
          if (f < 0) { f = (int)uw.size(); uw.push_back(e); } b.lr[e].rowW = f;
        }
        b.V[r].init(b.delta, (int)uv.size()); b.W[r].init((int)uw.size(), b.delta);
        for (size_t u = 0; u < uv.size(); u++) for (int a = 0; a < b.delta; a++) set(b.V[r](a, (int)u), b.lr[uv[u]].v(a, 0));
        for (size_t u = 0; u < uw.size(); u++) for (int a = 0; a < b.delta; a++) set(b.W[r]((int)u, a), b.lr[uw[u]].w(a, 0));
      }
      // the solver needs the transposed subblock of every low-rank piece (src/solver.jl:1009)
      for (auto& t : b.lr) if (!b.find.count({t.s, t.r, t.p, t.k})) { err = "low-rank term without its transposed subblock"; return 1; }
      b.A_Y.assign(b.m, std::vector<Mat>(b.m));
      for (int r = 0; r < b.m; r++) for (int s = 0; s <= r; s++) b.A_Y[r][s].init((int)b.rs[r][s].size(), 1);
    }
    // init (src/solver.jl:187-201)
    x.init(Ptot, 1); y.init(N, 1); d.init(Ptot, 1); p.init(N, 1); dx.init(Ptot, 1); dy.init(N, 1); Q.init(N, N);
    for_blocks([&](Cluster&, Block& b) {
      b.X.init(b.n, b.n); b.Y.init(b.n, b.n); b.Xinv.init(b.n, b.n); b.R.init(b.n, b.n); b.P.init(b.n, b.n); b.dX.init(b.n, b.n); b.dY.init(b.n, b.n);
      for (int i = 0; i < b.n; i++) { set(b.X(i, i), opt.omega_p.p()); set(b.Y(i, i), opt.omega_d.p()); }
    });
    for (auto& c : cl) { c.S.init(c.P, c.P); c.LinvB.init(c.P, N); }
    finalized = true;
    refresh_initial();
    return 0;
  }
  // initial objectives / residuals / errors (src/solver.jl:319-333)
  void refresh_initial() {
    dual_objective(d_obj.p()); primal_objective(p_obj.p()); duality_gap(gap.p(), d_obj.p(), p_obj.p());
    compute_residuals(false);
    compute_errors();
    pd_feas = cmp(dual_error.p(), opt.derr_thr.p()) < 0 && cmp(primal_error.p(), opt.perr_thr.p()) < 0;
    iter = 1;
  }
  void compute_errors() {   // src/solver.jl:807-835
    Num z; dual_error = z; primal_error = z;
    max_abs(dual_error.p(), p); for_blocks([&](Cluster&, Block& b) { max_abs(dual_error.p(), b.P); });
    max_abs(primal_error.p(), d);
  }

  // ---- sum_p a_p A_p  (compute_weighted_A!, src/solver.jl:1410-1470) ----
  void weighted_A(Mat Block::*dst, const Mat& a) {
    for (size_t j = 0; j < cl.size(); j++) for (auto& b : cl[j].blocks) {
      Mat& M = b.*dst; M.zero(); Num t, cur, av;
      if (b.high_rank) {
        for (size_t i = 0; i < b.dense_p.size(); i++) { mpfr_cp ap = a(off[j] + b.dense_p[i], 0);
          for (size_t e = 0; e < M.e.size(); e++) { if (dense_skip_zeros && is_zero(&b.dense_A[i].e[e])) continue; mpfr_mul(t.p(), &b.dense_A[i].e[e], ap, RN); mpfr_add(&M.e[e], &M.e[e], t.p(), RN); } }
        continue;
      }
      for (int r = 0; r < b.m; r++) for (int s = 0; s <= r; s++) for (int e : b.rs[r][s]) {
        auto& tm = b.lr[e]; mpfr_mul(cur.p(), a(off[j] + tm.p, 0), tm.lambda.p(), RN);
        // block[r-rows, s-cols] += (a_p lambda v) w^T   (Q = (vecs_left * VD)^T, :1455-1456)
        for (int u = 0; u < b.delta; u++) { mpfr_mul(av.p(), tm.v(u, 0), cur.p(), RN);
          for (int v = 0; v < b.delta; v++) { mpfr_mul(t.p(), av.p(), tm.w(v, 0), RN); mpfr_p dstp = M(r * b.delta + u, s * b.delta + v); mpfr_add(dstp, dstp, t.p(), RN); } }
      }
      if (b.m > 1) for (int i = 0; i < b.n; i++) for (int jj = 0; jj < i; jj++) set(M(jj, i), M(i, jj));   // symmetric!(.., :L)
    }
  }
  // ---- <A_*, Z> with vectors (trace_A, src/solver.jl:1290-1366) ----
  void trace_A_vec(Mat& res, Mat Block::*Zm) {
    res.zero(); Num t, acc, zv;
    for (size_t j = 0; j < cl.size(); j++) for (auto& b : cl[j].blocks) {
      Mat& Z = b.*Zm;
      if (b.high_rank) { for (size_t i = 0; i < b.dense_p.size(); i++) { dot(t.p(), b.dense_A[i], Z); mpfr_p rp = res(off[j] + b.dense_p[i], 0); mpfr_add(rp, t.p(), rp, RN); } continue; }
      for (int r = 0; r < b.m; r++) for (int s = 0; s <= r; s++) {
        // per p: sum_k lambda_k * ws_k^T Z[r,s] vs_k, doubled off the diagonal
        std::map<int, Num> perp; std::vector<int> order;
        for (int e : b.rs[r][s]) { auto& tm = b.lr[e]; acc.v._exp = EXP_ZERO; acc.v._sign = 1;
          for (int u = 0; u < b.delta; u++) { zv.v._exp = EXP_ZERO; zv.v._sign = 1;
            for (int v = 0; v < b.delta; v++) { mpfr_mul(t.p(), Z(r * b.delta + u, s * b.delta + v), tm.v(v, 0), RN); mpfr_add(zv.p(), zv.p(), t.p(), RN); }
            mpfr_mul(t.p(), zv.p(), tm.w(u, 0), RN); mpfr_add(acc.p(), acc.p(), t.p(), RN); }
          mpfr_mul(t.p(), tm.lambda.p(), acc.p(), RN);
          if (!perp.count(tm.p)) { perp[tm.p] = Num(); order.push_back(tm.p); }
          mpfr_add(perp[tm.p].p(), perp[tm.p].p(), t.p(), RN); }
        for (int pp : order) { if (r != s) mpfr_mul_2si(perp[pp].p(), perp[pp].p(), 1, RN); mpfr_p rp = res(off[j] + pp, 0); mpfr_add(rp, perp[pp].p(), rp, RN); }
      }
    }
  }
  // ---- <A_*, Y> from the stored pairings (trace_A((Y,A_Y)), src/solver.jl:1368-1407) ----
  void trace_A_pairings(Mat& res) {
    res.zero(); Num t;
    for (size_t j = 0; j < cl.size(); j++) for (auto& b : cl[j].blocks) {
      if (b.high_rank) { for (size_t i = 0; i < b.dense_p.size(); i++) { dot(t.p(), b.dense_A[i], b.Y); mpfr_p rp = res(off[j] + b.dense_p[i], 0); mpfr_add(rp, t.p(), rp, RN); } continue; }
      for (int r = 0; r < b.m; r++) for (int s = 0; s <= r; s++) {
        std::map<int, Num> perp; std::vector<int> order; int idx = 0;
        for (int e : b.rs[r][s]) { auto& tm = b.lr[e]; mpfr_mul(t.p(), b.A_Y[r][s](idx++, 0), tm.lambda.p(), RN);
          if (!perp.count(tm.p)) { perp[tm.p] = Num(); order.push_back(tm.p); }
          mpfr_add(perp[tm.p].p(), perp[tm.p].p(), t.p(), RN); }
        for (int pp : order) { if (r != s) mpfr_mul_2si(perp[pp].p(), perp[pp].p(), 1, RN); mpfr_p rp = res(off[j] + pp, 0); mpfr_add(rp, perp[pp].p(), rp, RN); }
      }
    }
  }
  // ---- residuals P, p, d (compute_residuals!, src/solver.jl:863-918) ----
  void compute_residuals(bool use_pairings) {
    weighted_A(&Block::P, x);
    for_blocks([&](Cluster&, Block& b) { for (size_t e = 0; e < b.P.e.size(); e++) {
      mpfr_sub(&b.P.e[e], &b.P.e[e], &b.X.e[e], RN);
      if (maximize) mpfr_sub(&b.P.e[e], &b.P.e[e], &b.C.e[e], RN); else mpfr_add(&b.P.e[e], &b.P.e[e], &b.C.e[e], RN); } });
    // d = c - B y - <A_*, Y>
    Mat tr(Ptot, 1);
    if (use_pairings) trace_A_pairings(tr); else trace_A_vec(tr, &Block::Y);
    for (size_t j = 0; j < cl.size(); j++) { Mat By; if (N > 0) gemm(By, cl[j].B, y);
      for (int i = 0; i < cl[j].P; i++) { mpfr_p di = d(off[j] + i, 0);
        if (N > 0) { mpfr_neg(di, By(i, 0), RN); mpfr_add(di, di, cl[j].c(i, 0), RN); } else set(di, cl[j].c(i, 0));
        mpfr_sub(di, di, tr(off[j] + i, 0), RN); } }
    // p = +-b - sum_j B_j^T x_j
    p.zero();
    for (size_t j = 0; j < cl.size(); j++) { if (N == 0) break; Mat xj(cl[j].P, 1), pj; for (int i = 0; i < cl[j].P; i++) set(xj(i, 0), x(off[j] + i, 0));
      gemm_tn(pj, cl[j].B, xj); for (int i = 0; i < N; i++) mpfr_sub(p(i, 0), p(i, 0), pj(i, 0), RN); }
    for (int i = 0; i < N; i++) { if (maximize) mpfr_add(p(i, 0), p(i, 0), b(i, 0), RN); else mpfr_sub(p(i, 0), p(i, 0), b(i, 0), RN); }
  }
  // ---- R = mu I - X Y [- dX dY]  (compute_residual_R!, src/solver.jl:961-983) ----
  void residual_R(mpfr_cp mu, bool second_order) {
    for_blocks([&](Cluster&, Block& b) {
      Mat T; gemm(T, b.X, b.Y); b.R.zero();
      for (int i = 0; i < b.n; i++) set(b.R(i, i), mu);
      for (size_t e = 0; e < T.e.size(); e++) mpfr_sub(&b.R.e[e], &b.R.e[e], &T.e[e], RN);
      if (second_order) { gemm(T, b.dX, b.dY); for (size_t e = 0; e < T.e.size(); e++) mpfr_sub(&b.R.e[e], &b.R.e[e], &T.e[e], RN); }
    });
  }
  // ---- Schur complement (compute_S_integrated!, src/solver.jl:1062-1226) ----
  void compute_S() {
    for (size_t j = 0; j < cl.size(); j++) { Cluster& c = cl[j]; c.S.zero();
      for (auto& b : c.blocks) {
        if (b.high_rank) {   // dense path :1089-1104
          int np = (int)b.dense_p.size(); int lim = dense_p_limit >= 0 ? std::min(np, dense_p_limit) : np;
          if (dense_skip_zeros && b.dense_cols.empty()) {
            b.dense_cols.resize(np); b.dense_nz.resize(np);
            for (int i = 0; i < np; i++) { std::vector<char> used(b.n, 0);
              for (int a = 0; a < b.n; a++) for (int c2 = 0; c2 < b.n; c2++) if (!is_zero(b.dense_A[i](a, c2))) { b.dense_nz[i].push_back({a, c2}); used[c2] = 1; }
              for (int c2 = 0; c2 < b.n; c2++) if (used[c2]) b.dense_cols[i].push_back(c2); }
          }
          t_np = np;
          for (int i = 0; i < lim; i++) {
            int pi = b.dense_p[i];
            auto tt0 = std::chrono::steady_clock::now();
            struct Tm { double* dst; std::chrono::steady_clock::time_point t0; ~Tm() { *dst += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } };
            const bool skip = dense_skip_zeros && !(sample_limit >= 0 && i < sample_limit);
            Tm tm{skip ? &t_skip : &t_plain, tt0};
            if (skip) {
              const std::vector<int>& J = b.dense_cols[i]; const int nj = (int)J.size();
              Mat Ap(b.n, nj), T1; for (int a = 0; a < b.n; a++) for (int k = 0; k < nj; k++) set(Ap(a, k), b.dense_A[i](a, J[k]));
              solve_cho(T1, b.Xinv, Ap);                                    // the nonzero columns of X^-1 A_p
#pragma omp parallel for schedule(dynamic, 1)
              for (int q = 0; q < np; q++) { int qi = b.dense_p[q]; if (qi < pi) continue; Num t, u, acc, tot;
                for (auto& ab : b.dense_nz[q]) { acc.v._exp = EXP_ZERO; acc.v._sign = 1;
                  for (int k = 0; k < nj; k++) { mpfr_mul(u.p(), T1(ab.first, k), b.Y(J[k], ab.second), RN); mpfr_add(acc.p(), acc.p(), u.p(), RN); }   // (T1 Y)[a,b]
                  mpfr_mul(t.p(), b.dense_A[q](ab.first, ab.second), acc.p(), RN); mpfr_add(tot.p(), tot.p(), t.p(), RN); }
                mpfr_add(c.S(pi, qi), c.S(pi, qi), tot.p(), RN); }
              continue;
            }
            Mat T1, T2; solve_cho(T1, b.Xinv, b.dense_A[i]); gemm(T2, T1, b.Y);
#pragma omp parallel for schedule(dynamic, 1)
            for (int q = 0; q < np; q++) { int qi = b.dense_p[q]; if (qi < pi) continue; Num t; dot(t.p(), b.dense_A[q], T2); mpfr_add(c.S(pi, qi), c.S(pi, qi), t.p(), RN); } }
          continue;
        }
        // low-rank path :1105-1212
        Mat Xi; inv_cho(Xi, b.Xinv);
        std::vector<std::vector<Mat>> BY(b.m, std::vector<Mat>(b.m)), BX(b.m, std::vector<Mat>(b.m));
        for (int r = 0; r < b.m; r++) for (int pass = 0; pass < 2; pass++) {
          const Mat& Src = pass == 0 ? b.Y : Xi; Mat col(b.n, b.delta), part;
          for (int i = 0; i < b.n; i++) for (int u = 0; u < b.delta; u++) set(col(i, u), Src(i, r * b.delta + u));
          gemm(part, col, b.V[r]);                                           // n x u_r   (:1125, :1137)
          for (int s = 0; s < b.m; s++) { Mat sub(b.delta, part.c);
            for (int u = 0; u < b.delta; u++) for (int v = 0; v < part.c; v++) set(sub(u, v), part(s * b.delta + u, v));
            gemm(pass == 0 ? BY[s][r] : BX[s][r], b.W[s], sub); }            // W_s * part[s-rows]   (:1131, :1143)
        }
        // A_Y (:1152-1170)
        for (int r = 0; r < b.m; r++) for (int s = 0; s <= r; s++) { int idx = 0;
          for (int e : b.rs[r][s]) { auto& tm = b.lr[e]; int et = b.find[{tm.s, tm.r, tm.p, tm.k}];
            set(b.A_Y[r][s](idx++, 0), BY[r][s](tm.rowW, b.lr[et].colV)); } }
        // S accumulation (:1176-1212); raw keys are already compact here
        const int nt = (int)b.lr.size();
        std::vector<std::vector<int>> by_p(c.P);                           // rows of S are owned by one thread each
        for (int e = 0; e < nt; e++) by_p[b.lr[e].p].push_back(e);
#pragma omp parallel
        { Num tot;
#pragma omp for schedule(dynamic, 1)
          for (int pp = 0; pp < c.P; pp++) for (int e1 : by_p[pp]) { auto& t1 = b.lr[e1];
            int l1 = b.lr[b.find[{t1.s, t1.r, t1.p, t1.k}]].rowW;            // pointers_left[s1][(r1,p,k1)]
            for (int e2 = 0; e2 < nt; e2++) { auto& t2 = b.lr[e2]; if (t2.p < t1.p) continue;
              int l2 = b.lr[b.find[{t2.s, t2.r, t2.p, t2.k}]].rowW;          // pointers_left[s2][(r2,q,k2)]
              mpfr_mul(tot.p(), t1.lambda.p(), t2.lambda.p(), RN);
              mpfr_mul(tot.p(), tot.p(), BX[t1.s][t2.r](l1, t2.colV), RN);
              mpfr_mul(tot.p(), tot.p(), BY[t2.s][t1.r](l2, t1.colV), RN);
              mpfr_add(c.S(t1.p, t2.p), c.S(t1.p, t2.p), tot.p(), RN);
            } } }
      }
      for (int i = 0; i < c.P; i++) for (int q = i + 1; q < c.P; q++) set(c.S(q, i), c.S(i, q));   // symmetric!  (:1222)
    }
  }
  // ---- decomposition (compute_T_decomposition!, src/solver.jl:1229-1287) ----
  int decomposition() {
    compute_S();
    for (size_t j = 0; j < cl.size(); j++) if (!cholesky(cl[j].S)) { err = "S was not decomposed succesfully in block " + std::to_string(j + 1); return 11; }
    if (N > 0) {
      Q.zero();
      for (auto& c : cl) { solve_lower(c.LinvB, c.S, c.B); Mat Qj; gemm_tn(Qj, c.LinvB, c.LinvB); for (size_t e = 0; e < Q.e.size(); e++) mpfr_add(&Q.e[e], &Q.e[e], &Qj.e[e], RN); }
      if (!cholesky(Q)) { err = "Q was not decomposed correctly."; return 12; }
    }
    return 0;
  }
  // ---- search direction (compute_search_direction!, src/solver.jl:1474-1616) ----
  void search_direction() {
    for_blocks([&](Cluster&, Block& b) {                    // Z = sym(X^-1 (P Y - R)) stored in dY
      Mat T; gemm(T, b.P, b.Y); for (size_t e = 0; e < T.e.size(); e++) mpfr_sub(&T.e[e], &T.e[e], &b.R.e[e], RN);
      solve_cho(b.dY, b.Xinv, T); symmetrize_half(b.dY); });
    Mat tr(Ptot, 1); trace_A_vec(tr, &Block::dY);           // rhs_x = -d - <A_*, Z>
    for (int i = 0; i < Ptot; i++) { mpfr_neg(dx(i, 0), d(i, 0), RN); mpfr_sub(dx(i, 0), dx(i, 0), tr(i, 0), RN); }
    std::vector<Mat> tx(cl.size());
    for (int i = 0; i < N; i++) set(dy(i, 0), p(i, 0));
    for (size_t j = 0; j < cl.size(); j++) { Mat rj(cl[j].P, 1); for (int i = 0; i < cl[j].P; i++) set(rj(i, 0), dx(off[j] + i, 0));
      solve_lower(tx[j], cl[j].S, rj);
      if (N > 0) { Mat u; gemm_tn(u, cl[j].LinvB, tx[j]); for (int i = 0; i < N; i++) mpfr_sub(dy(i, 0), dy(i, 0), u(i, 0), RN); } }
    if (N > 0) solve_cho(dy, Q, dy);
    for (size_t j = 0; j < cl.size(); j++) {
      if (N > 0) { Mat g; gemm(g, cl[j].LinvB, dy); for (int i = 0; i < cl[j].P; i++) mpfr_add(tx[j](i, 0), tx[j](i, 0), g(i, 0), RN); }
      Mat dxj; solve_lower_t(dxj, cl[j].S, tx[j]); for (int i = 0; i < cl[j].P; i++) set(dx(off[j] + i, 0), dxj(i, 0)); }
    weighted_A(&Block::dX, dx);                             // dX = P + sum dx_p A_p
    for_blocks([&](Cluster&, Block& b) {
      for (size_t e = 0; e < b.dX.e.size(); e++) mpfr_add(&b.dX.e[e], &b.dX.e[e], &b.P.e[e], RN);
      Mat T; gemm(T, b.dX, b.Y); for (size_t e = 0; e < T.e.size(); e++) mpfr_sub(&T.e[e], &b.R.e[e], &T.e[e], RN);
      solve_cho(b.dY, b.Xinv, T); symmetrize_half(b.dY); });   // dY = sym(X^-1 (R - dX Y))
  }
  // ---- step length (compute_step_length, src/solver.jl:1620-1693) ----
  int step_length(mpfr_p alpha, Mat Block::*Mm, Mat Block::*dMm, bool unsafe_step) {
    Num min_eig; bool have = false; int fail = 0;
    for_blocks([&](Cluster&, Block& b) {
      if (fail) return; Num ev; Mat& M = b.*Mm; Mat& dM = b.*dMm;
      if (b.n == 1) mpfr_div(ev.p(), dM(0, 0), M(0, 0), RN);
      else { Mat L = M; if (!cholesky(L)) { fail = 13; return; }
        Mat T; solve_lower(T, L, dM); Mat Tt(b.n, b.n); for (int i = 0; i < b.n; i++) for (int jj = 0; jj < b.n; jj++) set(Tt(i, jj), T(jj, i));
        solve_lower(T, L, Tt);
        std::vector<double> a((size_t)b.n * b.n); for (size_t e = 0; e < a.size(); e++) a[e] = mpfr_get_d(&T.e[e], RN);
        for (int i = 0; i < b.n; i++) for (int jj = 0; jj < i; jj++) { double s = 0.5 * (a[i * b.n + jj] + a[jj * b.n + i]); a[i * b.n + jj] = a[jj * b.n + i] = s; }
        mpfr_set_d(ev.p(), min_eig_sym(a, b.n) - 1e-5, RN); }   // Lanczos value - 10^-5  (:1662)
      if (!have || cmp(ev.p(), min_eig.p()) < 0) { min_eig = ev; have = true; } });
    if (fail) { err = "The cholesky decomposition could not be computed during the computation of the step length."; return fail; }
    Num ng; mpfr_neg(ng.p(), opt.gamma.p(), RN);
    if (cmp(min_eig.p(), ng.p()) > 0 && !unsafe_step) mpfr_set_si(alpha, 1, RN);
    else mpfr_div(alpha, ng.p(), min_eig.p(), RN);
    return 0;
  }
  bool terminate(int& reason) {   // src/solver.jl:921-950
    bool gap_opt = cmp(gap.p(), opt.gap_thr.p()) < 0, dual_feas = cmp(dual_error.p(), opt.derr_thr.p()) < 0, primal_feas = cmp(primal_error.p(), opt.perr_thr.p()) < 0;
    if (opt.need_dual && dual_feas) { reason = 2; return true; }
    if (opt.need_primal && primal_feas) { reason = 3; return true; }
    if (!opt.correctoronly && dual_feas && primal_feas && gap_opt) { reason = 1; return true; }
    return false;
  }
  // ---- one iteration (loop body src/solver.jl:362-592) ----
  int iterate(Info* info) {
    using clk = std::chrono::steady_clock; auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    memset(info, 0, sizeof(*info)); info->iter = iter;
    info->d_obj = mpfr_get_d(d_obj.p(), RN); info->p_obj = mpfr_get_d(p_obj.p(), RN); info->gap = mpfr_get_d(gap.p(), RN);
    int reason = 0;
    if (terminate(reason)) { info->stop = reason; info->pd_feasible = pd_feas; return 0; }
    Num mu, mu_p, t, r, beta, beta_c, mu_c, one, Kn; mpfr_set_si(one.p(), 1, RN); mpfr_set_si(Kn.p(), K, RN);
    auto dotXY = [&](mpfr_p res, Mat Block::*A, Mat Block::*B) { Num tt; res->_exp = EXP_ZERO; res->_sign = 1; for_blocks([&](Cluster&, Block& b) { dot(tt.p(), b.*A, b.*B); mpfr_add(res, res, tt.p(), RN); }); };
    dotXY(t.p(), &Block::X, &Block::Y); mpfr_div(mu.p(), t.p(), Kn.p(), RN);
    info->mu = mpfr_get_d(mu.p(), RN);
    if (opt.correctoronly) mu_p = mu; else if (!pd_feas) mpfr_mul(mu_p.p(), opt.beta_infeasible.p(), mu.p(), RN);
    if (cmp(mu.p(), opt.max_comp_gap.p()) > 0) { info->stop = 4; return 0; }
    auto t0 = clk::now();
    residual_R(mu_p.p(), false);
    auto t1 = clk::now();
    int fail = 0;
    for_blocks([&](Cluster&, Block& b) { if (fail) return; b.Xinv = b.X; if (!cholesky(b.Xinv)) fail = 10; });
    if (fail) { err = "The cholesky decomposition of X was not computed correctly."; return fail; }
    auto t2 = clk::now();
    if (int rc = decomposition()) return rc;
    auto t3 = clk::now();
    compute_residuals(true);
    auto t4 = clk::now();
    search_direction();                                   // predictor
    auto t5 = clk::now();
    { Num a, bb, c2, dd; dotXY(a.p(), &Block::X, &Block::Y); dotXY(bb.p(), &Block::X, &Block::dY); dotXY(c2.p(), &Block::dX, &Block::Y); dotXY(dd.p(), &Block::dX, &Block::dY);
      mpfr_add(a.p(), a.p(), bb.p(), RN); mpfr_add(a.p(), a.p(), c2.p(), RN); mpfr_add(a.p(), a.p(), dd.p(), RN);
      mpfr_mul(t.p(), mu.p(), Kn.p(), RN); mpfr_div(r.p(), a.p(), t.p(), RN); }
    if (cmp(r.p(), one.p()) < 0) mpfr_mul(beta.p(), r.p(), r.p(), RN); else beta = r;
    if (pd_feas) { beta_c = cmp(opt.beta_feasible.p(), beta.p()) > 0 ? opt.beta_feasible : beta; if (cmp(beta_c.p(), one.p()) > 0) beta_c = one; }   // stale pd_feas (:431-433)
    else beta_c = cmp(opt.beta_infeasible.p(), beta.p()) > 0 ? opt.beta_infeasible : beta;
    mpfr_mul(mu_c.p(), beta_c.p(), mu.p(), RN);
    auto t6 = clk::now();
    residual_R(mu_c.p(), true);
    auto t7 = clk::now();
    compute_errors();
    pd_feas = cmp(dual_error.p(), opt.derr_thr.p()) < 0 && cmp(primal_error.p(), opt.perr_thr.p()) < 0;   // (:441-447)
    { Num e; max_abs(e.p(), p); info->err_p = mpfr_get_d(e.p(), RN); Num e2; for_blocks([&](Cluster&, Block& b) { max_abs(e2.p(), b.P); }); info->err_P = mpfr_get_d(e2.p(), RN); info->err_d = mpfr_get_d(primal_error.p(), RN); }
    search_direction();                                   // corrector
    auto t8 = clk::now();
    Num alpha_d, alpha_p;
    if (int rc = step_length(alpha_d.p(), &Block::X, &Block::dX, pd_feas && !opt.safe_step)) return rc;
    if (int rc = step_length(alpha_p.p(), &Block::Y, &Block::dY, pd_feas && !opt.safe_step)) return rc;
    auto t9 = clk::now();
    info->beta_c = mpfr_get_d(beta_c.p(), RN); info->pd_feasible = pd_feas;
    info->alpha_d = mpfr_get_d(alpha_d.p(), RN); info->alpha_p = mpfr_get_d(alpha_p.p(), RN);
    info->phase_ms[0] = ms(t2, t3); info->phase_ms[1] = ms(t4, t5); info->phase_ms[2] = ms(t7, t8); info->phase_ms[3] = ms(t8, t9);
    info->phase_ms[4] = ms(t1, t2); info->phase_ms[5] = ms(t0, t1) + ms(t6, t7); info->phase_ms[6] = ms(t3, t4);
    { Num mn = cmp(alpha_d.p(), alpha_p.p()) < 0 ? alpha_d : alpha_p; if (cmp(mn.p(), opt.step_thr.p()) < 0) { info->stop = 5; return 0; } }
    if (pd_feas && opt.safe_step) { if (cmp(alpha_d.p(), alpha_p.p()) < 0) alpha_p = alpha_d; else alpha_d = alpha_p; }
    // step (:485-495)
    Num tt;
    for (int i = 0; i < Ptot; i++) { mpfr_mul(tt.p(), dx(i, 0), alpha_d.p(), RN); mpfr_add(x(i, 0), x(i, 0), tt.p(), RN); }
    for (int i = 0; i < N; i++) { mpfr_mul(tt.p(), dy(i, 0), alpha_p.p(), RN); mpfr_add(y(i, 0), y(i, 0), tt.p(), RN); }
    for_blocks([&](Cluster&, Block& b) { Num u; for (size_t e = 0; e < b.X.e.size(); e++) {
      mpfr_mul(u.p(), &b.dX.e[e], alpha_d.p(), RN); mpfr_add(&b.X.e[e], &b.X.e[e], u.p(), RN);
      mpfr_mul(u.p(), &b.dY.e[e], alpha_p.p(), RN); mpfr_add(&b.Y.e[e], &b.Y.e[e], u.p(), RN); } });
    dual_objective(d_obj.p()); primal_objective(p_obj.p()); duality_gap(gap.p(), d_obj.p(), p_obj.p());   // (:585-589)
    info->d_obj_new = mpfr_get_d(d_obj.p(), RN); info->p_obj_new = mpfr_get_d(p_obj.p(), RN); info->gap_new = mpfr_get_d(gap.p(), RN);
    iter++;
    return 0;
  }
};

// ---------------------------------------------------------------------------
// C interface: same shapes as include/clrs_b200.h with the clrs_oracle_ prefix
// ---------------------------------------------------------------------------
struct OOptions {   // == clrs_options
  int32_t prec, matmul_prec; double beta_infeasible, beta_feasible, gamma, omega_p, omega_d, gap_thr, derr_thr, perr_thr, max_comp_gap, step_thr;
  int32_t need_dual, need_primal, safe_step, correctoronly, device, gemm_path, sparse_schur;
};
static Num from_d(double v) { Num n; mpfr_set_d(n.p(), v, RN); return n; }

extern "C" {
int clrs_oracle_create(const OOptions* o, Oracle** out) {
  g_prec = o->prec > 0 ? o->prec : 256; g_w = (g_prec + 63) / 64;
  Oracle* h = new Oracle();
  h->opt.beta_infeasible = from_d(o->beta_infeasible); h->opt.beta_feasible = from_d(o->beta_feasible); h->opt.gamma = from_d(o->gamma);
  h->opt.omega_p = from_d(o->omega_p); h->opt.omega_d = from_d(o->omega_d); h->opt.gap_thr = from_d(o->gap_thr); h->opt.derr_thr = from_d(o->derr_thr);
  h->opt.perr_thr = from_d(o->perr_thr); h->opt.max_comp_gap = from_d(o->max_comp_gap); h->opt.step_thr = from_d(o->step_thr);
  h->opt.need_dual = o->need_dual; h->opt.need_primal = o->need_primal; h->opt.safe_step = o->safe_step; h->opt.correctoronly = o->correctoronly;
  *out = h; return 0;
}
int clrs_oracle_set_option_num(Oracle* h, int which, const void* w) {
  Num* t[] = {&h->opt.beta_infeasible, &h->opt.beta_feasible, &h->opt.gamma, &h->opt.omega_p, &h->opt.omega_d, &h->opt.gap_thr, &h->opt.derr_thr, &h->opt.perr_thr, &h->opt.max_comp_gap, &h->opt.step_thr};
  if (which < 0 || which > 9) return 1; from_wire(t[which]->p(), w); return 0;
}
void clrs_oracle_destroy(Oracle* h) { delete h; }
const char* clrs_oracle_last_error(const Oracle* h) { return h->err.c_str(); }
size_t clrs_oracle_wire_size(const Oracle*) { return wire_size(); }
int clrs_oracle_set_free(Oracle* h, int32_t N, const void* b, const void* constant, int32_t maximize) {
  h->N = N; h->b.init(N, 1); if (N > 0) mat_from_wire(h->b, b); from_wire(h->constant.p(), constant); h->maximize = maximize != 0; return 0;
}
int clrs_oracle_add_cluster(Oracle* h, int32_t j, int32_t P, const void* B, const void* c) {
  if (j != (int)h->cl.size()) { h->err = "clusters must be added in order"; return 1; }
  h->cl.emplace_back(); Cluster& cl = h->cl.back(); cl.P = P; cl.B.init(P, h->N); if (h->N > 0) mat_from_wire(cl.B, B); cl.c.init(P, 1); mat_from_wire(cl.c, c); return 0;
}
int clrs_oracle_add_block(Oracle* h, int32_t j, int32_t l, int32_t m, int32_t delta, int32_t high_rank, const void* C) {
  if (j >= (int)h->cl.size() || l != (int)h->cl[j].blocks.size()) { h->err = "blocks must be added in order"; return 1; }
  h->cl[j].blocks.emplace_back(); Block& b = h->cl[j].blocks.back(); b.m = m; b.delta = delta; b.n = m * delta; b.high_rank = high_rank != 0;
  b.C.init(b.n, b.n); mat_from_wire(b.C, C); return 0;
}
int clrs_oracle_add_dense_term(Oracle* h, int32_t j, int32_t l, int32_t p, const void* A) {
  Block& b = h->cl[j].blocks[l]; b.dense_p.push_back(p); b.dense_A.emplace_back(b.n, b.n); mat_from_wire(b.dense_A.back(), A); return 0;
}
// triplet form of a dense-block constraint matrix (include/clrs_b200.h: clrs_add_sparse_term); the oracle simply densifies it
int clrs_oracle_add_sparse_term(Oracle* h, int32_t j, int32_t l, int32_t p, int32_t nnz, const int32_t* rows, const int32_t* cols, const void* vals, int32_t mirror) {
  Block& b = h->cl[j].blocks[l]; b.dense_p.push_back(p); b.dense_A.emplace_back(b.n, b.n); Mat& A = b.dense_A.back(); size_t ws_ = wire_size();
  for (int t = 0; t < nnz; t++) { from_wire(A(rows[t], cols[t]), (const char*)vals + t * ws_); if (mirror && rows[t] != cols[t]) from_wire(A(cols[t], rows[t]), (const char*)vals + t * ws_); }
  return 0;
}
int clrs_oracle_add_lowrank_term(Oracle* h, int32_t j, int32_t l, int32_t r, int32_t s, int32_t p, int32_t rank, const void* lambda, const void* vs, const void* ws) {
  Block& b = h->cl[j].blocks[l]; size_t ws_ = wire_size();
  for (int k = 0; k < rank; k++) { b.lr.emplace_back(); LRTerm& t = b.lr.back(); t.r = r; t.s = s; t.p = p; t.k = k;
    from_wire(t.lambda.p(), (const char*)lambda + k * ws_); t.v.init(b.delta, 1); t.w.init(b.delta, 1);
    mat_from_wire(t.v, (const char*)vs + (size_t)k * b.delta * ws_); mat_from_wire(t.w, (const char*)ws + (size_t)k * b.delta * ws_); }
  return 0;
}
int clrs_oracle_finalize(Oracle* h) { return h->finalize(); }
int64_t clrs_oracle_state_matrix_count(const Oracle* h) { int64_t n = 0; for (auto& c : h->cl) for (auto& b : c.blocks) n += (int64_t)b.n * b.n; return n; }
int clrs_oracle_set_state(Oracle* h, const void* x, const void* X, const void* y, const void* Y) {
  size_t w = wire_size();
  if (x) mat_from_wire(h->x, x);
  if (y && h->N > 0) mat_from_wire(h->y, y);
  size_t o = 0; for (auto& c : h->cl) for (auto& b : c.blocks) { if (X) mat_from_wire(b.X, (const char*)X + o * w); if (Y) mat_from_wire(b.Y, (const char*)Y + o * w); o += (size_t)b.n * b.n; }
  h->refresh_initial(); return 0;
}
int clrs_oracle_get_state(Oracle* h, void* x, void* X, void* y, void* Y) {
  size_t w = wire_size();
  if (x) mat_to_wire(x, h->x);
  if (y && h->N > 0) mat_to_wire(y, h->y);
  size_t o = 0; for (auto& c : h->cl) for (auto& b : c.blocks) { if (X) mat_to_wire((char*)X + o * w, b.X); if (Y) mat_to_wire((char*)Y + o * w, b.Y); o += (size_t)b.n * b.n; }
  return 0;
}
int clrs_oracle_iterate(Oracle* h, Info* info) { return h->iterate(info); }
int clrs_oracle_get_objectives(Oracle* h, void* d_obj, void* p_obj, void* gap) {
  // recomputed from the current iterate as at src/solver.jl:626-628
  h->dual_objective(h->d_obj.p()); h->primal_objective(h->p_obj.p()); h->duality_gap(h->gap.p(), h->d_obj.p(), h->p_obj.p());
  to_wire(d_obj, h->d_obj.p()); to_wire(p_obj, h->p_obj.p()); to_wire(gap, h->gap.p()); return 0;
}
// OpenMP threads of the oracle's loops (torchrun exports OMP_NUM_THREADS=1: bench.py sets the count explicitly and reports it)
void clrs_oracle_set_threads(int32_t n) { if (n > 0) omp_set_num_threads(n); }
int32_t clrs_oracle_get_threads() { return (int32_t)omp_get_max_threads(); }
void clrs_oracle_set_dense_p_limit(Oracle* h, int32_t lim) { h->dense_p_limit = lim; }
void clrs_oracle_set_dense_skip_zeros(Oracle* h, int32_t on) { h->dense_skip_zeros = on != 0; }
void clrs_oracle_set_sample_limit(Oracle* h, int32_t lim) { h->sample_limit = lim; h->t_plain = h->t_skip = 0; }
void clrs_oracle_get_sample_times(Oracle* h, double* out3) { out3[0] = h->t_plain; out3[1] = h->t_skip; out3[2] = h->t_np; h->t_plain = h->t_skip = 0; }
// standalone kernels for parity tests
int clrs_oracle_mp_gemm(Oracle*, int32_t M, int32_t N, int32_t K, const void* A, const void* B, void* C) {
  Mat a(M, K), b(K, N), c; mat_from_wire(a, A); mat_from_wire(b, B); gemm(c, a, b); mat_to_wire(C, c); return 0;
}
int clrs_oracle_mp_cholesky(Oracle*, int32_t n, const void* A, void* L) {
  Mat a(n, n); mat_from_wire(a, A); int ok = cholesky(a); mat_to_wire(L, a); return ok ? 0 : 10;
}
// column-pivoted QR by modified Gram-Schmidt (pivot = largest remaining column norm, first index on ties): the checker of
// clrs_mp_qr_pivot; the reference's preprocess! uses LinearAlgebra.qr(mpsd, ColumnNorm()) in BigFloat (src/pre_postprocessing.jl:36),
// whose R agrees with this one up to the signs of its rows
int clrs_oracle_mp_qr_pivot(Oracle*, int32_t m, int32_t n, const void* Aw, void* Rw, int32_t* perm) {
  Mat A(m, n); mat_from_wire(A, Aw); const int kmax = m < n ? m : n; Mat R(kmax, n); std::vector<Num> norms(n); Num t, r, ri;
  for (int j = 0; j < n; j++) { perm[j] = j; norms[j].v._exp = EXP_ZERO; norms[j].v._sign = 1; for (int i = 0; i < m; i++) { mpfr_mul(t.p(), A(i, j), A(i, j), RN); mpfr_add(norms[j].p(), norms[j].p(), t.p(), RN); } }
  for (int k = 0; k < kmax; k++) {
    int p = k; for (int j = k + 1; j < n; j++) if (cmp(norms[j].p(), norms[p].p()) > 0) p = j;
    if (p != k) { for (int i = 0; i < m; i++) { set(t.p(), A(i, k)); set(A(i, k), A(i, p)); set(A(i, p), t.p()); }
      for (int i = 0; i < k; i++) { set(t.p(), R(i, k)); set(R(i, k), R(i, p)); set(R(i, p), t.p()); }
      set(t.p(), norms[k].p()); set(norms[k].p(), norms[p].p()); set(norms[p].p(), t.p()); std::swap(perm[k], perm[p]); }
    if (sgn(norms[k].p()) > 0) { mpfr_sqrt(r.p(), norms[k].p(), RN); set(R(k, k), r.p()); for (int i = 0; i < m; i++) mpfr_div(A(i, k), A(i, k), r.p(), RN); }
    else for (int i = 0; i < m; i++) { A(i, k)->_exp = EXP_ZERO; A(i, k)->_sign = 1; }
    for (int j = k + 1; j < n; j++) {
      r.v._exp = EXP_ZERO; r.v._sign = 1; for (int i = 0; i < m; i++) { mpfr_mul(t.p(), A(i, k), A(i, j), RN); mpfr_add(r.p(), r.p(), t.p(), RN); }
      set(R(k, j), r.p()); norms[j].v._exp = EXP_ZERO; norms[j].v._sign = 1;
      for (int i = 0; i < m; i++) { mpfr_mul(t.p(), r.p(), A(i, k), RN); mpfr_sub(A(i, j), A(i, j), t.p(), RN); mpfr_mul(t.p(), A(i, j), A(i, j), RN); mpfr_add(norms[j].p(), norms[j].p(), t.p(), RN); } }
  }
  mat_to_wire(Rw, R); return 0;
}
int64_t clrs_oracle_debug_get(Oracle* h, const char* what, int32_t j, int32_t l, void* out, int64_t cap) {
  std::string w(what); const Mat* m = nullptr;
  if (w == "S") m = &h->cl[j].S; else if (w == "LinvB") m = &h->cl[j].LinvB; else if (w == "Q") m = &h->Q;
  else if (w == "d") m = &h->d; else if (w == "p") m = &h->p; else if (w == "dx") m = &h->dx; else if (w == "dy") m = &h->dy;
  else if (w == "x") m = &h->x; else if (w == "y") m = &h->y;
  else { Block& b = h->cl[j].blocks[l];
    if (w == "Xinv") m = &b.Xinv; else if (w == "R") m = &b.R; else if (w == "P") m = &b.P; else if (w == "dX") m = &b.dX; else if (w == "dY") m = &b.dY; else if (w == "X") m = &b.X; else if (w == "Y") m = &b.Y; }
  if (!m) return -1; if ((int64_t)m->e.size() > cap) return -(int64_t)m->e.size();
  mat_to_wire(out, *m); return (int64_t)m->e.size();
}
}
