// clrs_device.cu — host driver of the B200 hot path and the C ABI (include/clrs_b200.h).
//
// One handle = one SDP resident in HBM.  clrs_iterate runs one predictor-corrector
// iteration of src/solver.jl:362-592 as a stream of kernels; every scalar decision
// of the loop body (mu, beta_c, step lengths, the safe-step rule, objectives) is
// taken ON the device by small scalar kernels, so the host synchronises once per
// iteration, to read the report.  No CPU fallback exists in this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <unordered_map>
#include <tuple>
#include <algorithm>
#include <stdexcept>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include "../../include/clrs_b200.h"
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "lanes.cuh"
#include "wire_host.h"

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

static inline int grid_for(int64_t n, int bs = 256, int cap = 148 * 8) { int64_t g = (n + bs - 1) / bs; if (g < 1) g = 1; return (int)std::min<int64_t>(g, cap); }

// ---- NCCL, loaded at run time (only multi-GPU handles need it) ----
struct NcclUid { char b[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr; int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr; int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr; int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr; int (*CommDestroy)(void*) = nullptr; const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    for (const char* n : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = "libnccl.so.2 not found"; return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId"); CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather"); AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce"); Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast"); CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy"); GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllGather || !AllReduce || !Broadcast || !CommDestroy) { err = "NCCL symbols missing"; return false; }
    return true;
  }
};
static NcclApi g_nccl;

// LPT partition of clusters over ranks by weight (the reference balances its threads by the same
// P_j^3 / n^3 weights, src/threadinginfo.jl:88,97); deterministic on every rank
static void partition_clusters(const std::vector<double>& w, int R, std::vector<int>& owner) {
  std::vector<int> order(w.size()); for (size_t i = 0; i < w.size(); i++) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a] > w[b]; });
  std::vector<double> load(R, 0.0); owner.assign(w.size(), 0);
  for (int j : order) { int best = 0; for (int r = 1; r < R; r++) if (load[r] < load[best]) best = r; owner[j] = best; load[best] += w[j]; }
}

// The whole plan of a sharded solve (host only, deterministic on every rank).  Work items are whole clusters (weight cw[j] = P_j^3 +
// the block weights) — except that a cluster that alone outweighs a rank's fair share by 25 % and has several blocks is SPLIT: its
// blocks become items of their own (the three-point bound is ONE cluster of 49-61 blocks), while its Schur
// complement, factor and solves are replicated.  Clusters that take the column-split path (`bigc`, §8(e)(ii)) keep their blocks together.
// split_mode: 0 never, 1 every multi-block cluster, -1 by weight.
struct ShardPlan { std::vector<int> cluster_owner, split; std::vector<std::vector<int>> block_owner; };
static ShardPlan plan_shards(const std::vector<double>& p3, const std::vector<std::vector<double>>& bw, const std::vector<int>& bigc, int R, int split_mode) {
  const size_t J = p3.size(); ShardPlan pl; pl.cluster_owner.assign(J, -1); pl.split.assign(J, 0); pl.block_owner.resize(J);
  std::vector<double> cw(J); double total = 0; for (size_t j = 0; j < J; j++) { cw[j] = p3[j]; for (double w : bw[j]) cw[j] += w; total += cw[j]; }
  std::vector<double> wgt; std::vector<std::pair<int, int>> item;                    // (cluster, block or -1)
  for (size_t j = 0; j < J; j++) {
    pl.split[j] = R > 1 && !bigc[j] && bw[j].size() >= 2 && split_mode != 0 && (split_mode == 1 || cw[j] > 1.25 * total / R);
    pl.block_owner[j].assign(bw[j].size(), -1);
    if (!pl.split[j]) { wgt.push_back(cw[j]); item.push_back({(int)j, -1}); continue; }
    for (size_t l = 0; l < bw[j].size(); l++) { wgt.push_back(bw[j][l]); item.push_back({(int)j, (int)l}); } }
  std::vector<int> owner; partition_clusters(wgt, R, owner);
  for (size_t i = 0; i < item.size(); i++) { const int j = item[i].first, l = item[i].second;
    if (l < 0) { pl.cluster_owner[j] = owner[i]; for (auto& o : pl.block_owner[j]) o = owner[i]; }
    else { pl.block_owner[j][l] = owner[i]; if (pl.cluster_owner[j] < 0) pl.cluster_owner[j] = owner[i]; } }      // the lead of a split cluster: the rank of its first block
  for (size_t j = 0; j < J; j++) if (pl.cluster_owner[j] < 0) pl.cluster_owner[j] = 0;
  return pl;
}

// scalar slots in device memory
enum { SC_MU, SC_MUP, SC_MUC, SC_BETA, SC_BETAC, SC_ALPHAD, SC_ALPHAP, SC_D0, SC_D1, SC_D2, SC_D3, SC_DOBJ, SC_POBJ, SC_GAP,
       SC_ERRP, SC_ERRp, SC_ERRd, SC_CX, SC_CY, SC_BY, SC_K, SC_ONE, SC_BETA_INF, SC_BETA_FEAS, SC_GAMMA, SC_OMEGA_P, SC_OMEGA_D,
       SC_GAPTHR, SC_DERRTHR, SC_PERRTHR, SC_MAXGAP, SC_STEPTHR, SC_CONSTANT, SC_ZERO, SC_TMP, SC_COUNT };
enum { FL_PDFEAS, FL_STOP, FL_STATUS, FL_COUNT };
enum { INFO_MU, INFO_DOBJ, INFO_POBJ, INFO_GAP, INFO_ERRP, INFO_ERRp, INFO_ERRd, INFO_ALPHAD, INFO_ALPHAP, INFO_BETAC, INFO_COUNT };

struct ScalarCfg { int correctoronly, safe_step, maximize; };

// out-of-line arithmetic for the one-thread scalar kernel: inlined, its ~20 k straight-line instructions (306 KB at 8 limbs)
// are fetched from L2 on every launch (20 us per launch, nine launches per iteration); as calls the code is a few KB
template <int NL> __device__ __noinline__ void mpc_mul(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) { mpn<NL> t; mp_mul(t, a, b); r = t; }
template <int NL> __device__ __noinline__ void mpc_add(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) { mpn<NL> t; mp_add(t, a, b); r = t; }
template <int NL> __device__ __noinline__ void mpc_sub(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) { mpn<NL> t; mp_sub(t, a, b); r = t; }
// mp_div (mpf.cuh: Newton reciprocal from a double-double seed, one correction step on the quotient) with its
// multiplications out of line: same operations, same bits
template <int NL> __device__ __noinline__ void mpc_div(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) {
  if (b.sign == 0) { mp_zero(r); return; }
  mpn<NL> m = b; m.exp = 0; m.sign = 1;
  double mh, ml, x0, c; mp_mant_dd(b.l[NL - 1], b.l[NL - 2], b.l[NL - 3], 0, mh, ml); dd_recip_seed(mh, ml, x0, c);
  mpn<NL> x, t, q, two; mp_from_double(x, x0); mp_from_double(t, c); mpc_add<NL>(x, x, t); mp_set_i32(two, 2);
#pragma unroll 1
  for (int it = 0; it < mp_newton_steps<NL>(); it++) { mpc_mul<NL>(t, m, x); mpc_sub<NL>(t, two, t); mpc_mul<NL>(x, x, t); }
  x.exp -= b.exp; x.sign = b.sign;
  mpc_mul<NL>(q, a, x); mpc_mul<NL>(t, q, b); mpc_sub<NL>(t, a, t); mpc_mul<NL>(t, t, x); mpc_add<NL>(r, q, t);
}
// the scalar logic of the loop body, one thread
template <int NL> __global__ void k_scalar(int phase, mpn<NL>* sc, int* fl, double* info, ScalarCfg cfg,
                                           int nblocks, const int32_t* bn, const int64_t* boff, const mpn<NL>* M, const mpn<NL>* dM,
                                           const double* lam, int which) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  typedef mpn<NL> num;
  if (phase == 0) {            // mu, mu_p  (src/solver.jl:369-380)
    num mu; mpc_div<NL>(mu, sc[SC_D0], sc[SC_K]); sc[SC_MU] = mu;
    num mup; mp_zero(mup);
    if (cfg.correctoronly) mup = mu; else if (!fl[FL_PDFEAS]) mpc_mul<NL>(mup, sc[SC_BETA_INF], mu);
    sc[SC_MUP] = mup;
    info[INFO_MU] = mp_to_double(mu);
    if (mp_cmp(mu, sc[SC_MAXGAP]) > 0) fl[FL_STOP] = CLRS_STOP_MAX_COMPLEMENTARY_GAP;
  } else if (phase == 1) {     // beta, beta_c, mu_c with the STALE pd_feas, then the fresh errors/pd_feas (src/solver.jl:429-447)
    num s, t, r, beta, betac;
    mpc_add<NL>(s, sc[SC_D0], sc[SC_D1]); mpc_add<NL>(s, s, sc[SC_D2]); mpc_add<NL>(s, s, sc[SC_D3]);
    mpc_mul<NL>(t, sc[SC_MU], sc[SC_K]); mpc_div<NL>(r, s, t);
    if (mp_cmp(r, sc[SC_ONE]) < 0) mpc_mul<NL>(beta, r, r); else beta = r;
    if (fl[FL_PDFEAS]) { betac = mp_cmp(sc[SC_BETA_FEAS], beta) > 0 ? sc[SC_BETA_FEAS] : beta; if (mp_cmp(betac, sc[SC_ONE]) > 0) betac = sc[SC_ONE]; }
    else betac = mp_cmp(sc[SC_BETA_INF], beta) > 0 ? sc[SC_BETA_INF] : beta;
    sc[SC_BETA] = beta; sc[SC_BETAC] = betac; num muc; mpc_mul<NL>(muc, betac, sc[SC_MU]); sc[SC_MUC] = muc;
    num de = mp_cmp_abs(sc[SC_ERRP], sc[SC_ERRp]) > 0 ? sc[SC_ERRP] : sc[SC_ERRp];
    fl[FL_PDFEAS] = (mp_cmp(de, sc[SC_DERRTHR]) < 0) && (mp_cmp(sc[SC_ERRd], sc[SC_PERRTHR]) < 0);
    info[INFO_BETAC] = mp_to_double(betac);
    info[INFO_ERRP] = mp_to_double(sc[SC_ERRP]); info[INFO_ERRp] = mp_to_double(sc[SC_ERRp]); info[INFO_ERRd] = mp_to_double(sc[SC_ERRd]);
  } else if (phase == 2) {     // smallest eigenvalue over this rank's blocks (src/solver.jl:1684-1686) -> sc[SC_TMP]
    num mn; bool have = false;
    for (int b = 0; b < nblocks; b++) {
      num ev;
      if (bn[b] == 1) mpc_div<NL>(ev, dM[boff[b]], M[boff[b]]);
      else { num c; mp_from_double(ev, lam[b]); mp_from_double(c, 1e-5); mpc_sub<NL>(ev, ev, c); }
      if (!have || mp_cmp(ev, mn) < 0) { mn = ev; have = true; }
    }
    if (!have) { mp_set_i32(mn, 1); mn.exp = 1 << 28; }        // a rank without blocks does not constrain the step
    sc[SC_TMP] = mn;
  } else if (phase == 6) {     // step length from the global smallest eigenvalue (src/solver.jl:1688-1692)
    num mn = sc[SC_TMP];
    num ng = sc[SC_GAMMA]; ng.sign = -ng.sign; num alpha;
    const bool unsafe_step = fl[FL_PDFEAS] && !cfg.safe_step;
    if (mp_cmp(mn, ng) > 0 && !unsafe_step) alpha = sc[SC_ONE]; else mpc_div<NL>(alpha, ng, mn);
    sc[which] = alpha;
  } else if (phase == 3) {     // threshold test and the safe-step rule (src/solver.jl:470-483)
    num ad = sc[SC_ALPHAD], ap = sc[SC_ALPHAP];
    info[INFO_ALPHAD] = mp_to_double(ad); info[INFO_ALPHAP] = mp_to_double(ap);
    num mn = mp_cmp(ad, ap) < 0 ? ad : ap;
    if (mp_cmp(mn, sc[SC_STEPTHR]) < 0) { fl[FL_STOP] = CLRS_STOP_STEP_TOO_SHORT; mp_zero(sc[SC_ALPHAD]); mp_zero(sc[SC_ALPHAP]); }
    else if (fl[FL_PDFEAS] && cfg.safe_step) { sc[SC_ALPHAD] = mn; sc[SC_ALPHAP] = mn; }
    if (fl[FL_STOP] == CLRS_STOP_MAX_COMPLEMENTARY_GAP || fl[FL_STATUS] != 0) { mp_zero(sc[SC_ALPHAD]); mp_zero(sc[SC_ALPHAP]); }   // the reference throws before the step (:394,1248,1276,1645): the last good iterate is kept
  } else if (phase == 4) {     // objectives and gap (src/solver.jl:792-847)
    num d = sc[SC_CX]; if (!cfg.maximize) d.sign = -d.sign; mpc_add<NL>(d, d, sc[SC_CONSTANT]);
    num p; mpc_add<NL>(p, sc[SC_CY], sc[SC_BY]); mpc_add<NL>(p, p, sc[SC_CONSTANT]);
    num a, b; mpc_sub<NL>(a, d, p); a.sign = a.sign ? 1 : 0; mpc_add<NL>(b, d, p); b.sign = b.sign ? 1 : 0;
    if (mp_cmp(b, sc[SC_ONE]) < 0) b = sc[SC_ONE];
    num g; mpc_div<NL>(g, a, b);
    sc[SC_DOBJ] = d; sc[SC_POBJ] = p; sc[SC_GAP] = g;
    info[INFO_DOBJ + 10] = mp_to_double(d); info[INFO_POBJ + 10] = mp_to_double(p); info[INFO_GAP + 10] = mp_to_double(g);
  } else if (phase == 5) {     // pd_feas from the current errors (initialisation, src/solver.jl:326-333)
    num de = mp_cmp_abs(sc[SC_ERRP], sc[SC_ERRp]) > 0 ? sc[SC_ERRP] : sc[SC_ERRp];
    fl[FL_PDFEAS] = (mp_cmp(de, sc[SC_DERRTHR]) < 0) && (mp_cmp(sc[SC_ERRd], sc[SC_PERRTHR]) < 0);
    info[INFO_ERRP] = mp_to_double(sc[SC_ERRP]); info[INFO_ERRp] = mp_to_double(sc[SC_ERRp]); info[INFO_ERRd] = mp_to_double(sc[SC_ERRd]);
  }
}
// small vector kernels
template <int NL> __global__ void k_vec_rhs(int n, mpn<NL>* dx, const mpn<NL>* d, const mpn<NL>* tr) {   // dx = -d - tr
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return; mpn<NL> a = d[i], b = tr[i], r; a.sign = -a.sign; mp_sub(r, a, b); dx[i] = r;
}
template <int NL> __global__ void k_set_diag(int n, mpn<NL>* A, int ld, const mpn<NL>* v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return; A[(int64_t)i * ld + i] = *v;
}
// S[plist[a], plist[b]] += T[b,a] for b >= a
template <int NL> __global__ void k_scatter_upper(int np, const int32_t* plist, const mpn<NL>* T, mpn<NL>* S, int ldS) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (idx >= (int64_t)np * np) return;
  int a = (int)(idx / np), b = (int)(idx % np); if (b < a) return;
  int p = plist[a], q = plist[b]; mpn<NL>* dst = S + (int64_t)min(p, q) * ldS + max(p, q);
  mpn<NL> o = *dst, t = T[(int64_t)b * np + a]; mp_add(o, o, t); *dst = o;     // T holds the lower triangle T[q][p]
}

// out[i] = reduction over ranks of gathered[r*n + i], in rank order (deterministic, identical on every rank)
// op 0: sum, 1: max |.|, 2: min (signed)
template <int NL> __global__ void k_combine_gathered(int64_t n, int R, const mpn<NL>* gathered, mpn<NL>* out, int op) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    mpn<NL> acc = gathered[i];
    for (int r = 1; r < R; r++) { mpn<NL> v = gathered[(int64_t)r * n + i];
      if (op == 0) mp_add(acc, acc, v); else if (op == 1) { if (mp_cmp_abs(v, acc) > 0) acc = v; } else { if (mp_cmp(v, acc) < 0) acc = v; } }
    if (op == 1 && acc.sign < 0) acc.sign = 1;
    out[i] = acc;
  }
}
// ---- cross-rank sums of multi-limb values with NCCL's own reductions: lanes.cuh holds the arithmetic ----
template <int NL> __global__ void k_lane_exp(int64_t n, const mpn<NL>* v, int32_t* E) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) E[i] = mp_lane_exp(v[i]);
}
template <int NL> __global__ void k_to_lanes(int64_t n, const mpn<NL>* v, const int32_t* E, long long* lanes) {     // lanes[k][i]: plane-major, coalesced
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    long long t[NL + 1]; mp_to_lanes<NL>(v[i], E[i], t);
#pragma unroll
    for (int k = 0; k <= NL; k++) lanes[(int64_t)k * n + i] = t[k];
  }
}
template <int NL> __global__ void k_from_lanes(int64_t n, const long long* lanes, const int32_t* E, mpn<NL>* v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    long long t[NL + 1];
#pragma unroll
    for (int k = 0; k <= NL; k++) t[k] = lanes[(int64_t)k * n + i];
    mpn<NL> r; mp_from_lanes<NL>(r, t, E[i]); v[i] = r;
  }
}
__global__ void k_combine_flags(int n, int R, const int* gathered, int* out) {
  int i = threadIdx.x; if (i >= n) return; int m = gathered[i]; for (int r = 1; r < R; r++) m = max(m, gathered[r * n + i]); out[i] = m;
}

// Q[r][c] += Qs[r][c] for the first `cols` columns of an N-row slab (row pitches lds, ldq)
template <int NL> __global__ void k_add_slab(int N, int cols, const mpn<NL>* Qs, int lds, mpn<NL>* Q, int ldq) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)N * cols; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols); mpn<NL> a = Q[(int64_t)r * ldq + c], b = Qs[(int64_t)r * lds + c]; mp_add(a, a, b); Q[(int64_t)r * ldq + c] = a; }
}
// pseudo-random multi-limb numbers for kernel benchmarks (splitmix-style hash)
template <int NL> __global__ void k_fill_random(int64_t n, mpn<NL>* a, uint64_t seed, int spread) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1); mpn<NL> v;
#pragma unroll
    for (int k = 0; k < NL; k++) { z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31; v.l[k] = (uint32_t)z; z += 0x9E3779B97F4A7C15ull; }
    v.l[NL - 1] |= 0x80000000u; v.exp = spread ? (int32_t)((z >> 8) % (2 * spread + 1)) - spread : 0; v.sign = (z & 1) ? 1 : -1; a[i] = v;
  }
}

// ---------------------------------------------------------------------------
struct SolverBase {
  std::string err; clrs_options opt; int prec = 256;
  virtual ~SolverBase() {}
  virtual int set_option_num(int which, const void* w) = 0;
  virtual int set_free(int N, const void* b, const void* constant, int maximize) = 0;
  virtual int add_cluster(int j, int P, const void* B, const void* c) = 0;
  virtual int add_block(int j, int l, int m, int delta, int high_rank, const void* C) = 0;
  virtual int add_dense_term(int j, int l, int p, const void* A) = 0;
  virtual int add_sparse_term(int j, int l, int p, int nnz, const int32_t* rows, const int32_t* cols, const void* vals, int mirror) = 0;
  virtual int add_lowrank_term(int j, int l, int r, int s, int p, int rank, const void* lam, const void* vs, const void* ws) = 0;
  virtual int finalize() = 0;
  virtual int set_state(const void* x, const void* X, const void* y, const void* Y) = 0;
  virtual int get_state(void* x, void* X, void* y, void* Y) = 0;
  virtual int64_t matrix_count() const = 0;
  virtual int iterate(clrs_iter_info* info) = 0;
  virtual int get_objectives(void* d, void* p, void* g) = 0;
  virtual int mp_gemm(int M, int N, int K, const void* A, const void* B, void* C, int path, double* ms) = 0;
  virtual int mp_cholesky(int n, const void* A, void* L) = 0;
  virtual int mp_qr_pivot(int m, int n, const void* A, void* R, int32_t* perm) = 0;
  virtual int64_t debug_get(const char* what, int j, int l, void* out, int64_t cap) = 0;
  virtual void profile(int enable) = 0;
  virtual void profile_get(double* out) = 0;
  virtual double last_iteration_ms() = 0;
  virtual void use_graph(int enable) = 0;
  virtual int bench_gemm(int M, int N, int K, int reps, int path, double* out) = 0;
  virtual int comm_init(int rank, int nranks, const void* uid) = 0;
  virtual int selftest() = 0;
  virtual int owner_of(int j) = 0;
  virtual int block_owner_of(int j, int l) = 0;
  size_t wire_size() const { return 16 + 8 * (size_t)((prec + 63) / 64); }
};

template <int NL> struct Solver : SolverBase {
  typedef mpn<NL> num;
  static constexpr int NS = I8Cfg<NL>::NS, NSP = I8Cfg<NL>::NSP;
  cudaStream_t st = nullptr;
  std::vector<void*> allocs;

  template <class T> T* dalloc(size_t n) { void* p = nullptr; if (n == 0) n = 1; CK(cudaMalloc(&p, n * sizeof(T))); CK(cudaMemsetAsync(p, 0, n * sizeof(T), st)); allocs.push_back(p); return (T*)p; }
  void release(void* p) { for (size_t i = 0; i < allocs.size(); i++) if (allocs[i] == p) { allocs[i] = allocs.back(); allocs.pop_back(); break; } cudaFree(p); }
  template <class T> T* upload(const std::vector<T>& v) { T* p = dalloc<T>(v.size()); if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); return p; }
  int W() const { return (prec + 63) / 64; }
  void w2m(num& r, const void* src) const { wire_to_mpn<NL>(r, src, W()); }
  void m2w(void* dst, const num& a) const { mpn_to_wire<NL>(dst, a, W()); }
  // wire records cross PCIe as raw bytes through a device staging buffer; the conversion runs on the device
  unsigned char* wstage = nullptr; size_t wstage_cap = 0;
  unsigned char* stage(size_t bytes) { if (bytes > wstage_cap) { if (wstage) CK(cudaFree(wstage)); wstage_cap = bytes + bytes / 4 + 4096; CK(cudaMalloc((void**)&wstage, wstage_cap)); } return wstage; }
  void wire_to_device(num* dst, const void* w, size_t n) {
    if (!n) return; unsigned char* sbuf = stage(n * wire_size());
    CK(cudaMemcpyAsync(sbuf, w, n * wire_size(), cudaMemcpyHostToDevice, st));
    nlaunch++, k_wire_to_mpn<NL><<<grid_for((int64_t)n), 256, 0, st>>>((int64_t)n, sbuf, W(), dst);
    CK(cudaStreamSynchronize(st));                        // the staging buffer is reused by the next piece
  }
  void device_to_wire(void* w, const num* src, size_t n) {
    if (!n) return; unsigned char* sbuf = stage(n * wire_size());
    nlaunch++, k_mpn_to_wire<NL><<<grid_for((int64_t)n), 256, 0, st>>>((int64_t)n, src, W(), sbuf);
    CK(cudaMemcpyAsync(w, sbuf, n * wire_size(), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  }
  num* upload_wire(const void* w, size_t n) { num* p = dalloc<num>(n); wire_to_device(p, w, n); return p; }
  void download_wire(void* w, const num* d, size_t n) { device_to_wire(w, d, n); }

  // ---- sliced panels -------------------------------------------------------
  // lay 0: dp4a words sl[vec][K4][NSP];  lay 1: tc planes planes[t][vec][Kp] (gemm_tc.cuh)
  struct Sliced { int32_t* sl = nullptr; int32_t* E = nullptr; uint8_t* planes = nullptr; int nvec = 0, K = 0, K4 = 0, Kp = 0, lay = 0; size_t cap_w = 0, cap_v = 0, cap_b = 0;
                  int64_t vpitch = 0;   /* vectors per plane when this is a view of a larger panel (0: nvec) */
                  int64_t pitch() const { return vpitch ? vpitch : nvec; } };
  // vectors [v0, v0 + cnt) of a tensor-core panel as a panel of their own (no ownership)
  static Sliced view(const Sliced& f, int v0, int cnt) { Sliced s = f; s.planes = f.planes + (size_t)v0 * f.Kp; s.E = f.E + v0; s.nvec = cnt; s.vpitch = f.pitch(); s.cap_w = s.cap_v = s.cap_b = 0; return s; }
  void ensure(Sliced& s, int nvec, int K, int lay) {
    int K4 = (K + 3) / 4; int Kp = (K + 31) & ~31;
    if ((size_t)nvec > s.cap_v) { alloc_gen++; if (s.E) CK(cudaFreeAsync(s.E, st)); s.cap_v = nvec; CK(cudaMallocAsync((void**)&s.E, std::max<size_t>(s.cap_v, 1) * sizeof(int32_t), st)); }
    if (lay == 0) { size_t need = (size_t)nvec * K4 * NSP;
      if (need > s.cap_w) { alloc_gen++; if (s.sl) CK(cudaFreeAsync(s.sl, st)); s.cap_w = need; CK(cudaMallocAsync((void**)&s.sl, std::max<size_t>(s.cap_w, 4) * sizeof(int32_t), st)); } }
    else { size_t need = (size_t)NS * nvec * Kp;
      if (need > s.cap_b) { alloc_gen++; if (s.planes) CK(cudaFreeAsync(s.planes, st)); s.cap_b = need; CK(cudaMallocAsync((void**)&s.planes, std::max<size_t>(s.cap_b, 16), st)); } }
    s.nvec = nvec; s.K = K; s.K4 = K4; s.Kp = Kp; s.lay = lay; s.vpitch = 0;
  }
  std::vector<Sliced*> owned_sliced;
  static VecView rows_view(const num* A, int lda, int M, int K) { VecView v; v.base = A; v.bstride = 0; v.vper = M > 0 ? M : 1; v.sv = lda; v.sk = 1; v.nvec = M; v.K = K; return v; }
  static VecView cols_view(const num* B, int ldb, int K, int N) { VecView v; v.base = B; v.bstride = 0; v.vper = N > 0 ? N : 1; v.sv = 1; v.sk = ldb; v.nvec = N; v.K = K; return v; }
  // E[vec] = largest exponent among the entries of each vector
  void vec_exponents(const VecView& v, int32_t* E) {
    if (v.K >= 8192 && (int64_t)v.nvec * 32 < 148 * 2048) {           // few long vectors: several CTAs per vector
      const int ny = std::max(1, std::min((v.K + 2047) / 2048, (148 * 16 + v.nvec - 1) / v.nvec));
      nlaunch++, k_fill_i32<<<(v.nvec + 255) / 256, 256, 0, st>>>(v.nvec, E, I8_EXP_NONE);
      nlaunch++, k_vec_exp_long<NL><<<dim3(v.nvec, ny), 256, 0, st>>>(v, E);
    } else
    nlaunch++, k_vec_exp<NL><<<(unsigned)(((int64_t)v.nvec * 32 + 255) / 256), 256, 0, st>>>(v, E);
  }
  // panel over the packed upper triangles of `cnt` n x n matrices at M (stride n^2): vectors [v0, v0 + cnt) of the tensor-core
  // panel s (already sized); sym: entries M[a][b] + M[b][a]
  void split_tri(Sliced& s, int v0, const num* M, int cnt, int n, bool sym) {
    VecView v; v.base = M; v.bstride = 0; v.vper = std::max(cnt, 1); v.sv = (int64_t)n * n; v.sk = 1; v.nvec = cnt; v.K = n * n;
    vec_exponents(v, s.E + v0);
    const int64_t tot_ = (int64_t)cnt * (s.Kp / 4);
    if (sym) nlaunch++, k_exp_add<<<(cnt + 127) / 128, 128, 0, st>>>(cnt, s.E + v0, 1);
    nlaunch++, k_split_tc_tri<NL><<<(unsigned)((tot_ + 127) / 128), 128, 0, st>>>(M, (int64_t)n * n, cnt, n, sym ? 1 : 0, s.E + v0, s.Kp, s.pitch(), s.planes + (size_t)v0 * s.Kp);
  }
  void split(Sliced& s, const VecView& v, bool kfast, int lay = 0, bool is_view = false) {
    if (!is_view) ensure(s, v.nvec, v.K, lay);
    if (v.nvec == 0 || v.K == 0) return;
    if (lay == 0 && v.K <= 512) { nlaunch++, k_split_warp<NL><<<(unsigned)(((int64_t)v.nvec * 32 + 127) / 128), 128, 0, st>>>(v, s.E, s.K4, s.sl); return; }
    vec_exponents(v, s.E);
    if (lay == 0) { int64_t tot_ = (int64_t)v.nvec * s.K4;
      nlaunch++, k_split<NL><<<(unsigned)((tot_ + 127) / 128), 128, 0, st>>>(v, s.E, s.K4, s.sl, kfast ? 1 : 0); }
    else if (kfast) { int64_t tot_ = (int64_t)v.nvec * (s.Kp / 4);
      nlaunch++, k_split_tc<NL><<<(unsigned)((tot_ + 127) / 128), 128, 0, st>>>(v, s.E, s.Kp, s.pitch(), s.planes, 1); }
    else { constexpr int VT = SplitTCfg<NL>::VT; dim3 grid((v.nvec + VT - 1) / VT, s.Kp / 32);
      nlaunch++, k_split_tc_t<NL><<<grid, VT * 8, 0, st>>>(v, s.E, s.Kp, s.pitch(), s.planes); }
  }
  void split_rows(Sliced& s, const num* A, int lda, int M, int K, int lay = 0) { split(s, rows_view(A, lda, M, K), true, lay); }
  void split_cols(Sliced& s, const num* B, int ldb, int K, int N, int lay = 0) { split(s, cols_view(B, ldb, K, N), false, lay); }
  // does a product of this shape go to the tensor cores?
  bool use_tc(int M, int N, int K) const { if (opt.gemm_path == 1) return false; if (opt.gemm_path == 2) return M >= 1 && N >= 1 && K >= 1; return M >= 128 && N >= 32 && K >= 96 && ((int64_t)M * N > 256 * 256 || K > 512); }   // up to ~256^3 the CUDA-core kernel wins (measured: 130^3 0.06 vs 0.20 ms, 200^3 0.16 vs 0.24, 300^3 0.34 vs 0.25)

  // C (M x N) = op(D, A*B);  A, B sliced with vector offsets a0, b0
  void gemm(const Sliced& A, int a0, const Sliced& B, int b0, int M, int N, num* C, int ldc, int mode = 0, const num* D = nullptr, int ldd = 0,
            int batch = 1, int64_t a_bvec = 0, int64_t b_bvec = 0, int64_t c_bs = 0, int64_t d_bs = 0, int lower_only = 0) {
    if (M == 0 || N == 0) return;
    if (A.K != B.K || A.lay != B.lay) throw CudaError("gemm: operand panels do not match");
    if (A.K == 0) throw CudaError("gemm: K == 0 not supported");
    prof_begin();
    if (A.lay == 1) gemm_tc(A, a0, B, b0, M, N, C, ldc, mode, D, ldd, batch, a_bvec, b_bvec, c_bs, d_bs, lower_only);
    else {
      GemmArgs g; g.M = M; g.N = N; g.K4 = A.K4; g.batch = batch;
      g.Asl = A.sl + (int64_t)a0 * A.K4 * NSP; g.EA = A.E + a0; g.a_bvec = a_bvec;
      g.Bsl = B.sl + (int64_t)b0 * B.K4 * NSP; g.EB = B.E + b0; g.b_bvec = b_bvec;
      g.C = C; g.ldc = ldc; g.c_bstride = c_bs; g.D = D; g.ldd = ldd; g.d_bstride = d_bs; g.mode = mode; g.lower_only = lower_only;
      dim3 grid((N + 15) / 16, (M + 15) / 16, batch);
      nlaunch++, k_gemm_dp4a<NL><<<grid, 256, 0, st>>>(g);
    }
    prof_end(2.0 * M * N * (double)A.K * batch * (lower_only ? 0.5 : 1.0), A.lay == 0 ? 0 : ((double)M * N >= 1e6 ? 2 : 1));
  }
  // ---- tensor-core path ---------------------------------------------------------------
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn encode_fn = nullptr;
  uint8_t* tc_bytes = nullptr; int32_t* tc_top = nullptr; size_t tc_cap = 0; int32_t* tc_raw = nullptr; size_t tc_raw_cap = 0;
  long alloc_gen = 0;     // bumped whenever a scratch buffer used by the iteration is reallocated (a captured graph is stale then)
  CUtensorMap make_map(const Sliced& P, int box_rows) {
    if (!encode_fn) { void* fn = nullptr; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q)); if (!fn) throw CudaError("cuTensorMapEncodeTiled not available"); encode_fn = (EncodeFn)fn; }
    CUtensorMap m; cuuint64_t dims[3] = {(cuuint64_t)P.Kp, (cuuint64_t)P.nvec, (cuuint64_t)NS};
    cuuint64_t strides[2] = {(cuuint64_t)P.Kp, (cuuint64_t)P.Kp * (cuuint64_t)P.pitch()};
    cuuint32_t box[3] = {(cuuint32_t)tc::KC, (cuuint32_t)box_rows, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, P.planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
  }
  // column tiling of a product with N columns: 128-wide tiles with a narrower last one and four diagonals per group, or
  // tiles of up to 160 columns with three (both fill the 512 TMEM columns).  The slice-pair count is the same either way,
  // so the choice minimises the summed MMA time of the tiles of one row block: an M=128, K=32 int8 MMA costs
  // max(operand bytes / 128 B per clock of shared memory, tensor work) cycles (+ issue overhead).
  static double mma_cycles(int bn) { return std::max((4096.0 + 32.0 * bn) / 128.0, 0.73 * bn) + 10.0; }    // 0.73 cycles per column at the measured int8 rate (2 x 1.6 PFLOP/s bf16)
  static void pick_tiles(int N, int& BN, int& group, int M = 0, int lower_only = 0) {
    const int N16 = (N + 15) & ~15;
    // summed MMA time of the tiles that are computed (all of them, or for a lower-triangular output those touching i >= j)
    auto cost = [&](int bn) { double c = 0; const int nbm = M > 0 ? (M + tc::BM - 1) / tc::BM : 1;
      for (int by = 0; by < nbm; by++) for (int n0 = 0; n0 < N; n0 += bn) if (!lower_only || n0 <= by * tc::BM + tc::BM - 1) c += mma_cycles(std::min(bn, (N - n0 + 15) & ~15));
      return c; };
    const int bnA = std::min(128, N16);
    const int ntB = (N16 + tc::BNMAX - 1) / tc::BNMAX; int bnB = (((N16 + ntB - 1) / ntB) + 15) & ~15; if (bnB > tc::BNMAX) bnB = tc::BNMAX;
    static const int force = getenv("CLRS_TC_GROUP") ? atoi(getenv("CLRS_TC_GROUP")) : 0;
    // measured (profiles/r02_tile_bench.md): 90000 x 300 x 300 gains 7.6 % from 160 + 144 over 128 + 128 + 48, while with many
    // column tiles (N = 641) the three-diagonal groups lose to wave quantisation: wide tiles only for one or two of them
    if (force == 4 || bnB <= 128 || (force != 3 && (ntB > 2 || cost(bnA) <= cost(bnB) * 1.02))) { BN = bnA; group = 4; } else { BN = bnB; group = 3; }
  }
  // matmul_prec (src/solver.jl:93,125): the products of the bilinear pairings and the T Y product of the dense path may be formed at a lower
  // precision.  Here that means FEWER SLICE-PAIR DIAGONALS: only the ns_cur most significant diagonals D_s, s < ns_cur, are produced
  // (8 bits each; 22 guard + 2 headroom bits as at full precision), i.e. ns (ns + 1) / 2 instead of 630 int8 products per MAC at 256 bit.
  // Tensor-core path only; the CUDA-core kernel keeps all diagonals (at least as accurate as asked).
  int ns_cur = NS;
  int ns_matmul() const { if (opt.matmul_prec <= 0 || opt.matmul_prec >= prec) return NS; return std::max(4, std::min(NS, (opt.matmul_prec + 24 + 7) / 8)); }
  struct NsScope { Solver& s; NsScope(Solver& s_) : s(s_) { s.ns_cur = s.ns_matmul(); } ~NsScope() { s.ns_cur = NS; } };
  static int tc_epi() { static const int v = getenv("CLRS_TC_EPI") ? atoi(getenv("CLRS_TC_EPI")) : 1; return v; }   // epilogue variant of tc::k_gemm_tc (gemm_tc.cuh)
  void gemm_tc(const Sliced& A, int a0, const Sliced& B, int b0, int M, int N, num* C, int ldc, int mode, const num* D, int ldd,
               int batch, int64_t a_bvec, int64_t b_bvec, int64_t c_bs, int64_t d_bs, int lower_only, int trans = 0) {
    if (a0 != 0 || b0 != 0) throw CudaError("gemm_tc: panel offsets are not supported");
    int BN, group; pick_tiles(N, BN, group, M, lower_only);
    const int Npitch = (N + 15) & ~15;
    CUtensorMap mA = make_map(A, tc::BM), mB = make_map(B, BN);
    // K longer than the int32 headroom (35 slices * K * 2^14 < 2^31), or few tiles with a long K: split K over
    // blockIdx.z (partial results summed in k_tc_recombine), sized to fill whole waves of 148 CTAs
    const int KMAX = ((1 << 17) / NS / 128) * 128; int tiles = 0;      // int32 headroom: NS * K * 2^14 < 2^31
    for (int by = 0; by < (M + tc::BM - 1) / tc::BM; by++) for (int bx = 0; bx < (N + BN - 1) / BN; bx++) if (!lower_only || bx * BN <= by * tc::BM + tc::BM - 1) tiles++;
    int nch = (A.Kp + KMAX - 1) / KMAX;
    if (batch == 1 && tiles < 148 && A.Kp >= 256) { const int waves = (tiles * nch + 147) / 148; nch = std::max(nch, std::min(waves * 148 / tiles, A.Kp >= 2048 ? A.Kp / 512 : (A.Kp + 127) / 128)); }
    if (nch > 1 && batch != 1) throw CudaError("gemm_tc: split-K with a batch is not supported");
    int kch = ((A.Kp + nch - 1) / nch + 127) & ~127; nch = (A.Kp + kch - 1) / kch;
    tc::Args a; a.M = M; a.N = N; a.Kp = A.Kp; a.k0 = 0; a.BN = BN; a.group = group; a.dsplit = 0; a.oraw = nullptr; a.a_bvec = (int)a_bvec; a.b_bvec = (int)b_bvec; a.NS = ns_cur; a.Npitch = Npitch;
    a.lower_only = lower_only; a.Kp_total = A.Kp; a.dbg = nullptr; a.epi = tc_epi();
    // under-parallelised products (a few tiles on 148 SMs): one CTA per (tile, K range, diagonal group), raw int32 sums
    // accumulated with red.add, carries resolved in k_tc_recombine_raw
    static const int dsplit_off = getenv("CLRS_TC_DSPLIT") ? atoi(getenv("CLRS_TC_DSPLIT")) == 0 : 0;
    const int ngroups = (ns_cur + group - 1) / group;
    if (!dsplit_off && batch == 1 && tiles * nch < 74 && (int64_t)M * Npitch <= (1 << 19) && A.Kp <= KMAX) {
      const int nkc = (A.Kp + 127) / 128; int nzk = std::max(1, std::min(nkc, (296 + tiles * ngroups - 1) / (tiles * ngroups)));
      kch = ((nkc + nzk - 1) / nzk) * 128; nzk = (A.Kp + kch - 1) / kch;
      const size_t need = (size_t)NS * M * Npitch;
      if (need > tc_raw_cap) { if (tc_raw) CK(cudaFreeAsync(tc_raw, st)); tc_raw_cap = need; alloc_gen++; CK(cudaMallocAsync((void**)&tc_raw, tc_raw_cap * sizeof(int32_t), st)); }
      CK(cudaMemsetAsync(tc_raw, 0, need * sizeof(int32_t), st));
      a.dsplit = 1; a.oraw = tc_raw; a.batch = 1; a.obytes = nullptr; a.otop = nullptr; a.kz_stride = kch;
      dim3 grid((N + BN - 1) / BN, (M + tc::BM - 1) / tc::BM, nzk * ngroups);
      nlaunch++, tc::k_gemm_tc<<<grid, tc::NTHREADS, tc::SMEM_BYTES, st>>>(mA, mB, a);
      nlaunch++, k_tc_recombine_raw<NL><<<(unsigned)(((int64_t)M * N + 127) / 128), 128, 0, st>>>(M, N, Npitch, tc_raw, A.E, B.E, C, ldc, D, ldd, mode, lower_only, trans);
      return;
    }
    const size_t outs2 = (size_t)std::max(batch, nch) * M * Npitch;
    if (outs2 > tc_cap) { if (tc_bytes) CK(cudaFreeAsync(tc_bytes, st)); if (tc_top) CK(cudaFreeAsync(tc_top, st)); tc_cap = outs2; alloc_gen++;
      CK(cudaMallocAsync((void**)&tc_bytes, tc_cap * NS, st)); CK(cudaMallocAsync((void**)&tc_top, tc_cap * sizeof(int32_t), st)); }
    a.batch = nch > 1 ? nch : batch; a.obytes = tc_bytes; a.otop = tc_top; a.kz_stride = nch > 1 ? kch : 0;
    dim3 grid((N + BN - 1) / BN, (M + tc::BM - 1) / tc::BM, a.batch);
    nlaunch++, tc::k_gemm_tc<<<grid, tc::NTHREADS, tc::SMEM_BYTES, st>>>(mA, mB, a);
    const int64_t tot_ = (int64_t)batch * M * N;
    if (ns_cur < NS) nlaunch++, k_tc_recombine<NL, true><<<(unsigned)((tot_ + 127) / 128), 128, 0, st>>>(M, N, Npitch, a.batch, tc_bytes, tc_top, A.E, a_bvec, B.E, b_bvec, C, ldc, c_bs, D, ldd, d_bs, mode, lower_only, nch > 1 ? nch : 1, trans, ns_cur);
    else nlaunch++, k_tc_recombine<NL, false><<<(unsigned)((tot_ + 127) / 128), 128, 0, st>>>(M, N, Npitch, a.batch, tc_bytes, tc_top, A.E, a_bvec, B.E, b_bvec, C, ldc, c_bs, D, ldd, d_bs, mode, lower_only, nch > 1 ? nch : 1, trans, NS);
  }
  // ---- optional per-launch GEMM profile (bench.py roofline) ------------------------------------
  bool prof_on = false; cudaEvent_t pe0 = nullptr, pe1 = nullptr; double prof_ms[3] = {0, 0, 0}, prof_flops[3] = {0, 0, 0}; long prof_n[3] = {0, 0, 0};   // 0: CUDA-core path, 1: tcgen05 small outputs, 2: tcgen05 outputs >= 1e6 numbers
  void prof_begin() { if (prof_on) CK(cudaEventRecord(pe0, st)); }
  void prof_end(double mp_flops, int lay) { if (!prof_on) return; CK(cudaEventRecord(pe1, st)); CK(cudaEventSynchronize(pe1)); float t = 0; cudaEventElapsedTime(&t, pe0, pe1); prof_ms[lay] += t; prof_flops[lay] += mp_flops; prof_n[lay]++; }
  long nlaunch = 0;
  Sliced tA, tB;    // scratch panels
  // second execution context (stream + scratch) for work that is independent of the main chain: the Cholesky of Y
  // for the step length only needs the iterate, so it runs beside the Schur assembly.  swap_ctx() exchanges the
  // members the helpers use; kernels capture their pointers at enqueue time, so swapping while enqueuing is safe.
  struct Ctx { cudaStream_t st = nullptr; Sliced tA, tB; num* chol_W = nullptr; size_t chol_W_cap = 0; uint8_t* tc_bytes = nullptr; int32_t* tc_top = nullptr; size_t tc_cap = 0; int32_t* tc_raw = nullptr; size_t tc_raw_cap = 0;
               num* trsm_R = nullptr; size_t trsm_cap = 0; cudaEvent_t ev = nullptr; };
  Ctx side, side2, side3;                            // side: Cholesky of Y; side2: R = mu I - XY beside chol(X), and Y's step-length eigenvalue beside X's; side3: first stage of the dense Schur products
  cudaEvent_t evS0 = nullptr, evS1 = nullptr; bool stage1_pending = false;
  cudaEvent_t evR0 = nullptr, evR1 = nullptr, evE0 = nullptr, evE1 = nullptr; num *U2 = nullptr, *T1b = nullptr; double* Td2 = nullptr; double* eigV2 = nullptr; EigTask* eigT2 = nullptr;
  cudaEvent_t evY0 = nullptr, evY1 = nullptr; num *LY = nullptr, *MinvY = nullptr;
  void swap_with(Ctx& c) { std::swap(st, c.st); std::swap(tA, c.tA); std::swap(tB, c.tB); std::swap(chol_W, c.chol_W); std::swap(chol_W_cap, c.chol_W_cap);
    std::swap(tc_bytes, c.tc_bytes); std::swap(tc_top, c.tc_top); std::swap(tc_cap, c.tc_cap); std::swap(tc_raw, c.tc_raw); std::swap(tc_raw_cap, c.tc_raw_cap); std::swap(trsm_R, c.trsm_R); std::swap(trsm_cap, c.trsm_cap); }
  void swap_ctx() { swap_with(side); }
  // Independent items (PSD blocks, clusters) are enqueued round-robin on NCTX execution contexts, so that the many small
  // kernels of a many-block problem (one diagonal-block Cholesky CTA, 16-CTA GEMMs) overlap instead of queueing on one
  // stream.  Item i runs on context i mod NCTX (context 0 = the caller's); fork and join are events.
  static constexpr int NCTX = 8;
  Ctx pctx[NCTX]; cudaEvent_t evFork = nullptr; int par_depth = 0;
  template <class F> void par_clusters(F&& f) { std::vector<Clu*> o; for (auto& c0 : cl) if (c0.owned) o.push_back(&c0); par_for((int)o.size(), [&](int i) { f(*o[i]); }); }
  template <class F> void par_blocks(F&& f) { par_for((int)blk.size(), [&](int i) { f(blk[i]); }); }
  template <class F> void par_for(int n, F&& f) {
    const int nc = std::min(n, NCTX);
    if (nc <= 1 || par_depth > 0 || prof_on) { for (int i = 0; i < n; i++) f(i); return; }
    par_depth++;
    CK(cudaEventRecord(evFork, st));
    for (int k = 1; k < nc; k++) CK(cudaStreamWaitEvent(pctx[k].st, evFork, 0));
    for (int i = 0; i < n; i++) { const int k = i % nc;
      if (k) swap_with(pctx[k]);
      try { f(i); } catch (...) { if (k) swap_with(pctx[k]); par_depth--; throw; }
      if (k) swap_with(pctx[k]); }
    for (int k = 1; k < nc; k++) { CK(cudaEventRecord(pctx[k].ev, pctx[k].st)); CK(cudaStreamWaitEvent(st, pctx[k].ev, 0)); }
    par_depth--;
  }
  // C = op(D, A*B) for plain matrices A (M x K, lda), B (K x N, ldb)
  void mm(const num* A, int lda, const num* B, int ldb, int M, int N, int K, num* C, int ldc, int mode = 0, const num* D = nullptr, int ldd = 0) {
    const int lay = use_tc(M, N, K) ? 1 : 0;
    // a short left operand (a 32-row panel against a long right-hand side): swap the operands so that the long
    // dimension fills the 128 TMEM lanes and the short one becomes a narrow N tile; the result is written transposed
    if (!lay && opt.gemm_path != 1 && M >= 16 && M < 128 && N >= 128 && K >= 96) {
      split_cols(tA, B, ldb, K, N, 1); split_rows(tB, A, lda, M, K, 1);
      prof_begin(); gemm_tc(tA, 0, tB, 0, N, M, C, ldc, mode, D, ldd, 1, 0, 0, 0, 0, 0, 1); prof_end(2.0 * M * N * (double)K, 1);
      return; }
    split_rows(tA, A, lda, M, K, lay); split_cols(tB, B, ldb, K, N, lay); gemm(tA, 0, tB, 0, M, N, C, ldc, mode, D, ldd);
  }

  // ---- multi-GPU: clusters sharded over ranks, partial results combined by all-gather + ordered sum ----
  int rank = 0, nranks = 1; void* comm = nullptr; num* gbuf = nullptr; size_t gcap = 0; int* gflags = nullptr;
  int comm_init(int rank_, int nranks_, const void* uid) override {
    if (finalized) { err = "clrs_comm_init must precede clrs_finalize"; return CLRS_ERR_ARG; }
    if (nranks_ <= 1) { rank = 0; nranks = 1; return CLRS_OK; }
    if (!g_nccl.load(err)) return CLRS_ERR_CUDA;
    NcclUid id; memcpy(id.b, uid, 128);
    int rc = g_nccl.CommInitRank(&comm, nranks_, id, rank_);
    if (rc != 0) { err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); return CLRS_ERR_CUDA; }
    rank = rank_; nranks = nranks_; gflags = dalloc<int>((size_t)nranks * FL_COUNT); return CLRS_OK;
  }
  int owner_of(int j) override { return (j >= 0 && j < (int)cl.size()) ? cl[j].owner : -1; }
  int block_owner_of(int j, int l) override { return (j >= 0 && j < (int)cl.size() && l >= 0 && l < (int)cl[j].blocks.size()) ? cl[j].blocks[l].brank : -1; }
  long long* lane_buf = nullptr; int32_t* lane_E = nullptr; size_t lane_cap = 0;
  void allreduce(num* v, int64_t n, int op) {      // op 0 sum, 1 max-abs, 2 min
    if (nranks == 1 || n == 0) return;
    // vectors and matrices (Q, sum_j u_j, p): exponent max + int64-lane sum with NCCL's own all-reduce; the handful of scalar
    // reductions (max-abs, min, dot products) keep the all-gather + ordered combine below
    static const int lanes_min = getenv("CLRS_LANES_MIN") ? atoi(getenv("CLRS_LANES_MIN")) : 16;
    if (op == 0 && n >= lanes_min) {
      constexpr int LN = NL + 1;
      if ((size_t)n > lane_cap) { if (lane_buf) CK(cudaFreeAsync(lane_buf, st)); if (lane_E) CK(cudaFreeAsync(lane_E, st)); lane_cap = (size_t)n; alloc_gen++;
        CK(cudaMallocAsync((void**)&lane_buf, lane_cap * LN * sizeof(long long), st)); CK(cudaMallocAsync((void**)&lane_E, lane_cap * sizeof(int32_t), st)); }
      nlaunch++, k_lane_exp<NL><<<grid_for(n), 256, 0, st>>>(n, v, lane_E);
      if (g_nccl.AllReduce(lane_E, lane_E, (size_t)n, /*ncclInt32*/ 2, /*ncclMax*/ 2, comm, st) != 0) throw CudaError("ncclAllReduce (max) failed");
      nlaunch++, k_to_lanes<NL><<<grid_for(n), 256, 0, st>>>(n, v, lane_E, lane_buf);
      if (g_nccl.AllReduce(lane_buf, lane_buf, (size_t)n * LN, /*ncclInt64*/ 4, /*ncclSum*/ 0, comm, st) != 0) throw CudaError("ncclAllReduce (sum) failed");
      nlaunch++, k_from_lanes<NL><<<grid_for(n), 256, 0, st>>>(n, lane_buf, lane_E, v);
      return;
    }
    if ((size_t)n * nranks > gcap) { if (gbuf) CK(cudaFreeAsync(gbuf, st)); gcap = (size_t)n * nranks; alloc_gen++; CK(cudaMallocAsync((void**)&gbuf, gcap * sizeof(num), st)); }
    int rc = g_nccl.AllGather(v, gbuf, (size_t)n * sizeof(num), /*ncclChar*/ 0, comm, st);
    if (rc != 0) throw CudaError("ncclAllGather failed");
    nlaunch++, k_combine_gathered<NL><<<grid_for(n), 256, 0, st>>>(n, nranks, gbuf, v, op);
  }
  void allreduce_flags() {
    if (nranks == 1) return;
    int rc = g_nccl.AllGather(flags, gflags, FL_COUNT * sizeof(int), 0, comm, st); if (rc != 0) throw CudaError("ncclAllGather failed");
    nlaunch++, k_combine_flags<<<1, 32, 0, st>>>(FL_COUNT, nranks, gflags, flags);
  }

  // ---- flat helpers ------------------------------------------------------------
  void zero(num* a, int64_t n) { if (n) nlaunch++, k_zero<NL><<<grid_for(n), 256, 0, st>>>(n, a); }
  void copy(num* r, const num* a, int64_t n) { if (n) nlaunch++, k_copy<NL><<<grid_for(n), 256, 0, st>>>(n, r, a); }
  void addsub(num* r, const num* a, int sa, const num* b, int sb, int64_t n) { if (n) nlaunch++, k_addsub<NL><<<grid_for(n), 256, 0, st>>>(n, r, a, sa, b, sb); }
  num* partial = nullptr;
  void reduce(const num* a, const num* b, int64_t n, num* out, int mode) {   // b == null: max-abs
    if (n == 0) { if (mode == 0) zero(out, 1); return; }
    int g = grid_for(n, 128, 256);
    nlaunch++, k_reduce_partial<NL><<<g, 128, 128 * sizeof(num), st>>>(n, a, b, partial);
    nlaunch++, k_reduce_final<NL><<<1, 128, 128 * sizeof(num), st>>>(g, partial, b ? 0 : 1, out, mode);
  }

  // ---- blocked Cholesky with explicit factor inverse -----------------------------------
  // A (n x n, lower part read) -> L in place (strict upper zeroed); Minv = L^-1 (lower triangular, full n x n buffer)
  num* chol_W = nullptr; size_t chol_W_cap = 0;
  // full_inverse: also assemble L^-1 below the diagonal blocks (X and Y blocks: products with L^-1 and X^-1 follow);
  // otherwise only the inverses of the 32 x 32 diagonal blocks are formed and solves go by block substitution.
  // Two-level blocking: 128-column outer panels, 32-column steps inside a panel.  A step factors its 32 x 32 diagonal block
  // (k_potrf_diag), solves the rows below it and updates only the REST OF THE PANEL (K = 32, CUDA cores); the trailing
  // matrix is updated once per panel with K = 128, which is a tensor-core shape (src/tools.jl:69-107 is the unblocked loop).
  static constexpr int PNL = 128;
  long long* potrf_dbg = nullptr;   // kernel tuning only: clock64 timeline of the first diagonal-block factorisation (CLRS_POTRF_TIMELINE)
  void chol(num* A, int lda, int n, num* Minv, int ldm, int code, bool full_inverse = true) {
    if (n == 0) return;
    if (ldm == n) zero(Minv, (int64_t)n * n); else for (int r = 0; r < n; r++) zero(Minv + (int64_t)r * ldm, n);
    static const int pnl_env = getenv("CLRS_CHOL_PANEL") ? std::max(32, atoi(getenv("CLRS_CHOL_PANEL")) / 32 * 32) : 0;
    const int pnl = pnl_env ? pnl_env : (n >= 448 ? PNL : 32);        // measured: n = 640 gains (L^-1 B 7.8 -> 4.4 ms at 16 limbs), n <= 400 does not
    // (Measured and not kept, round 2: a look-ahead variant that factored the next diagonal block on a second stream beside the rest of the
    //  trailing update — bit-identical factor, tools/gpu_chol_check.py — did not shorten the chain: the 32-step substitution of k_trsm32 has the
    //  same ~100 us latency for 32 rows as for 600, so splitting it into "next block" + "rest" put it on the critical path twice:
    //  chol S of the P = 640 cluster 8.5 -> 9.1 ms, profiles/README.md.  What would shorten the chain is a panel kernel that factors the
    //  diagonal block together with the 32 rows below it and applies the update of the next diagonal block itself.)
    for (int K0 = 0; K0 < n; K0 += pnl) {
      const int KE = std::min(n, K0 + pnl);
      for (int k0 = K0; k0 < KE; k0 += 32) {
        const int nb = std::min(32, KE - k0), rem = n - k0 - nb;
        nlaunch++, k_potrf_diag<NL><<<1, POTRF_THREADS, POTRF_SMEM(NL), st>>>(nb, A + (int64_t)k0 * lda + k0, lda, Minv + (int64_t)k0 * ldm + k0, ldm, flags + FL_STATUS, code, full_inverse ? 1 : 0, k0 == 0 ? potrf_dbg : nullptr);
        if (rem <= 0) continue;
        num* A21 = A + (int64_t)(k0 + nb) * lda + k0;
        if (full_inverse) { split_rows(tA, A21, lda, rem, nb); split_rows(tB, Minv + (int64_t)k0 * ldm + k0, ldm, nb, nb);
          gemm(tA, 0, tB, 0, rem, nb, A21, lda); }                    // L21 = A21 * inv(L11)^T
        else nlaunch++, k_trsm32<NL><<<(rem + 7) / 8, 256, 0, st>>>(nb, A + (int64_t)k0 * lda + k0, lda, Minv + (int64_t)k0 * ldm + k0, ldm, A21, lda, 1, rem, A21, lda, 1);   // rows of L21 by substitution
        const int pc = KE - k0 - nb;                                  // columns of this panel still to be factored
        if (pc > 0) { num* Ap = A + (int64_t)(k0 + nb) * lda + k0 + nb;
          split_rows(tA, A21, lda, rem, nb);
          gemm(tA, 0, tA, 0, rem, pc, Ap, lda, 1, Ap, lda); }           // A[k0+nb:, k0+nb:KE] -= L21 L21[0:pc]^T  (the strict upper part of A is never read)
      }
      const int remO = n - KE;
      if (remO > 0) { const int kw = KE - K0; num* Lp = A + (int64_t)KE * lda + K0; num* A22 = A + (int64_t)KE * lda + KE;
        split_rows(tA, Lp, lda, remO, kw, use_tc(remO, remO, kw) ? 1 : 0);
        gemm(tA, 0, tA, 0, remO, remO, A22, lda, 1, A22, lda, 1, 0, 0, 0, 0, 1); }   // A22 -= L21 L21^T (lower), K = panel width
    }
    nlaunch++, k_zero_upper<NL><<<grid_for((int64_t)n * n), 256, 0, st>>>(n, A, lda);
    // rows of the inverse below the diagonal blocks: M[i,0:k0] = -inv(L_ii) * (L[i,0:k0] * M[0:k0,0:k0])
    if (full_inverse && n > 32) {
      size_t need = (size_t)32 * n; if (need > chol_W_cap) { chol_W = dalloc<num>(need); chol_W_cap = need; alloc_gen++; }
      for (int k0 = 32; k0 < n; k0 += 32) {
        const int nb = std::min(32, n - k0);
        mm(A + (int64_t)k0 * lda, lda, Minv, ldm, nb, k0, k0, chol_W, k0);
        mm(Minv + (int64_t)k0 * ldm + k0, ldm, chol_W, k0, nb, k0, nb, Minv + (int64_t)k0 * ldm, ldm, 3);
      }
    }
  }
  // X = L^-1 B by block forward substitution (approx_solve_tril!, src/solver.jl:1258), right-looking with the same two-level
  // blocking: a 32-row block is solved against its diagonal block (k_trsm32) and updates the rest of its 128-row panel
  // (K = 32); the rows below the panel are updated once per panel, B[KE:, :] -= L[KE:, K0:KE] X[K0:KE, :] (K = 128, tensor cores)
  num* trsm_R = nullptr; size_t trsm_cap = 0;   // (per-context scratch slot, kept for the context swap)
  void trsm_lower(const num* Lf, int ldl, int n, const num* Minv, int ldm, const num* B, int ldb, int ncols, num* Xo, int ldx) {
    if (n == 0 || ncols == 0) return;
    if (ldb == ncols && ldx == ncols) copy(Xo, B, (int64_t)n * ncols); else for (int r = 0; r < n; r++) copy(Xo + (int64_t)r * ldx, B + (int64_t)r * ldb, ncols);
    static const int pnl_env = getenv("CLRS_CHOL_PANEL") ? std::max(32, atoi(getenv("CLRS_CHOL_PANEL")) / 32 * 32) : 0;
    const int pnl = pnl_env ? pnl_env : (n >= 448 ? PNL : 32);
    for (int K0 = 0; K0 < n; K0 += pnl) {
      const int KE = std::min(n, K0 + pnl);
      for (int k0 = K0; k0 < KE; k0 += 32) {
        const int nb = std::min(32, KE - k0), pr = KE - k0 - nb;
        num* Xk = Xo + (int64_t)k0 * ldx;
        nlaunch++, k_trsm32<NL><<<(ncols + 7) / 8, 256, 0, st>>>(nb, Lf + (int64_t)k0 * ldl + k0, ldl, Minv + (int64_t)k0 * ldm + k0, ldm, Xk, 1, ldx, ncols, Xk, 1, ldx);
        if (pr > 0) { num* Xr = Xo + (int64_t)(k0 + nb) * ldx;
          mm(Lf + (int64_t)(k0 + nb) * ldl + k0, ldl, Xk, ldx, pr, ncols, nb, Xr, ldx, 1, Xr, ldx); }
      }
      const int remO = n - KE;
      if (remO > 0) { num* Xr = Xo + (int64_t)KE * ldx;
        mm(Lf + (int64_t)KE * ldl + K0, ldl, Xo + (int64_t)K0 * ldx, ldx, remO, ncols, KE - K0, Xr, ldx, 1, Xr, ldx); }
    }
  }
  // x <- L^-1 x (forward) or L^-T x (backward) by block substitution in one launch (k_trsv_fused: one CTA per 32-row block,
  // blocks chained through `ready` flags); CLRS_TRSV_FUSED=0 selects the two-launches-per-block form (same bits)
  void trsv(const num* Lf, int ldl, int n, const num* Minv, int ldm, num* xv, bool transposed, unsigned* ready) {
    const int nblk = (n + 31) / 32; if (nblk == 0) return;
    static const int fused = getenv("CLRS_TRSV_FUSED") ? atoi(getenv("CLRS_TRSV_FUSED")) : 1;
    if (fused && nblk > 1 && nblk <= 148) {
      CK(cudaMemsetAsync(ready, 0, nblk * sizeof(unsigned), st));
      nlaunch++, k_trsv_fused<NL><<<nblk, 1024, 0, st>>>(n, Lf, ldl, Minv, ldm, xv, transposed ? 1 : 0, ready);
      return;
    }
    for (int bi = 0; bi < nblk; bi++) {
      const int b = transposed ? nblk - 1 - bi : bi, k0 = b * 32, nb = std::min(32, n - k0);
      nlaunch++, k_trsv_block<NL><<<1, 1024, 0, st>>>(nb, Lf + (int64_t)k0 * ldl + k0, ldl, Minv + (int64_t)k0 * ldm + k0, ldm, xv + k0, transposed ? 1 : 0);
      if (!transposed) { const int rem = n - k0 - nb;                                           // rows below: L[k0+nb.., k0..k0+nb)
        if (rem > 0) nlaunch++, k_trsv_update<NL><<<(rem + 7) / 8, 256, 0, st>>>(rem, nb, Lf + (int64_t)(k0 + nb) * ldl + k0, (int64_t)ldl, 1, xv + k0, xv + k0 + nb); }
      else if (k0 > 0) nlaunch++, k_trsv_update<NL><<<(k0 + 7) / 8, 256, 0, st>>>(k0, nb, Lf + (int64_t)k0 * ldl, 1, (int64_t)ldl, xv + k0, xv);   // columns left: L[k0..k0+nb, 0..k0)^T
    }
  }

  // ---- problem description -------------------------------------------------------------
  struct HTerm { int r, s, p, k; num lam; std::vector<num> v, w; int colV = -1, rowW = -1; };
  struct Block {
    int j = 0, l = 0, m = 1, delta = 1, n = 1; bool high_rank = false; int64_t off = 0;   // offset in the flat block storage (owned blocks)
    bool mine = true; int brank = 0;                                                      // split clusters: the rank that holds this block
    int64_t goff = 0;                                                                      // offset in the global (all ranks) block order
    std::vector<num> hC;
    // dense
    std::vector<int> dense_p; size_t Aall_cap = 0;
    int np = 0; int32_t* d_plist = nullptr; num* Aall = nullptr; int32_t* nz_start = nullptr; int32_t* nz_idx = nullptr; int32_t* nzT_start = nullptr; int32_t* nzT_p = nullptr; int64_t nnz = 0; Sliced AallB, AallV; num* T1 = nullptr; num* T2 = nullptr; num* Sd = nullptr; Sliced T1S, T2V; bool sparse = false;   /* sparse: S from the nonzero lists of the A_p (k_schur_sparse), no products */ bool tri = false;   /* tri: Schur inner products over the packed upper triangle (all A_p symmetric) */
    // low rank
    std::vector<HTerm> lr;
    int nP = 0; int32_t* lr_plist = nullptr; int32_t* lr_tstart = nullptr; LRTermDev* lr_terms = nullptr; num* lr_lam = nullptr;
    std::vector<int> u_r, ul_r; std::vector<num*> V, W; std::vector<Sliced> Vs, Ws;        // V_r: delta x u_r ; W_r: ul_r x delta
    std::vector<num*> BX, BY, ZV; MatRef *d_BX = nullptr, *d_BY = nullptr, *d_ZV = nullptr, *d_W = nullptr;
    num* part = nullptr; int umax = 0;
    struct RS { int cnt = 0; int32_t* elist = nullptr; num* H = nullptr; Sliced Hs; num* G = nullptr; };
    std::vector<RS> rs;                                                                     // index r*m+s, s <= r
    // per-iteration cached panels (layout `lay`: 1 = tensor-core panels for large blocks)
    Sliced YS, XiS, MS, MSY; int lay = 0;
  };
  struct Clu { int owner = 0; bool owned = true; bool split = false, lead = true;   /* split: the BLOCKS of this cluster are spread over the ranks; S_j, x_j, B_j, the factor and the solves are replicated, `lead` (= the owner) alone adds the cluster-level terms to cross-rank sums */ int P = 0; std::vector<num> hB, hc; std::vector<Block> blocks; num *B = nullptr, *S = nullptr, *Minv = nullptr, *LinvB = nullptr, *t = nullptr; unsigned* ready = nullptr; int off = 0;
               bool big = false; num *Bc = nullptr, *Gc = nullptr, *Qs = nullptr; };   /* big: L^-1 B and its Q contribution are split by COLUMNS over all ranks: Bc = this rank's columns of B (P x ncr), Gc = all column chunks of L^-1 B ([rank][P][ncr]), Qs = this rank's N x ncr slab of G^T G */
  std::vector<Clu> cl; std::vector<Block*> blk;   // blk: all blocks in (j,l) order
  int N = 0, Ptot = 0, Ksum = 0, maximize = 1; std::vector<num> hb; num hconst;
  int64_t tot = 0;   // numbers in the flat block storage of the blocks this rank owns
  int64_t gtot = 0;  // numbers in all blocks of the SDP
  // device state
  num *X = nullptr, *Y = nullptr, *Cm = nullptr, *L = nullptr, *Minv = nullptr, *Xi = nullptr, *R = nullptr, *P = nullptr, *dX = nullptr, *dY = nullptr, *T1 = nullptr, *TXY = nullptr, *U = nullptr;
  num* tmpU = nullptr; num* LinvBall = nullptr; int Pown = 0; unsigned* q_ready = nullptr; int ncr = 0;   /* ncr: columns of B per rank for big clusters */
  num *x = nullptr, *y = nullptr, *c = nullptr, *cobj = nullptr, *b = nullptr, *d = nullptr, *p = nullptr, *dx = nullptr, *dy = nullptr, *tr = nullptr, *Q = nullptr, *QMinv = nullptr, *tmpN = nullptr;
  num* sc = nullptr; int* flags = nullptr; double* dinfo = nullptr; double* Td = nullptr; double* lamX = nullptr; double* lamY = nullptr; double* eigV = nullptr; EigTask* eigT = nullptr;
  int64_t* d_boff = nullptr; int32_t* d_bn = nullptr; BlockTab bt;
  num hopt[10]; bool hopt_set[10] = {false};
  bool finalized = false; int iter = 1;
  double h_gap = 1, h_derr = 1e300, h_perr = 1e300, h_dobj = 0, h_pobj = 0; int h_pdfeas = 0;
  cudaEvent_t ev[24];

  Solver(const clrs_options& o) {
    opt = o; prec = o.prec;
    int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaError("no CUDA device: libclrs_b200 has no CPU fallback");
    CK(cudaSetDevice(o.device)); cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, o.device));
    if (pr.major < 10) throw CudaError("an sm_100 device is required");
    { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&st, cudaStreamDefault, hi)); }      // the main chain outranks the side streams
    CK(cudaStreamCreate(&side.st)); CK(cudaStreamCreate(&side2.st)); CK(cudaStreamCreate(&side3.st)); CK(cudaEventCreateWithFlags(&evS0, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&evS1, cudaEventDisableTiming)); for (cudaEvent_t* e : {&evR0, &evR1, &evE0, &evE1}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&evY0, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&evY1, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&evFork, cudaEventDisableTiming)); for (int k = 1; k < NCTX; k++) { CK(cudaStreamCreate(&pctx[k].st)); CK(cudaEventCreateWithFlags(&pctx[k].ev, cudaEventDisableTiming)); }
    for (auto& e : ev) CK(cudaEventCreate(&e)); for (auto& r : evD) for (auto& e : r) CK(cudaEventCreate(&e));
    CK(cudaFuncSetAttribute(k_potrf_diag<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM(NL)));
    CK(cudaFuncSetAttribute(tc::k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaEventCreate(&pe0)); CK(cudaEventCreate(&pe1));
    mp_zero(hconst);
    double dv[10] = {o.beta_infeasible, o.beta_feasible, o.gamma, o.omega_p, o.omega_d, o.duality_gap_threshold, o.dual_error_threshold, o.primal_error_threshold, o.max_complementary_gap, o.step_length_threshold};
    for (int i = 0; i < 10; i++) mp_from_double(hopt[i], dv[i]);
    partial = dalloc<num>(256); sc = dalloc<num>(SC_COUNT); flags = dalloc<int>(FL_COUNT); dinfo = dalloc<double>(32);
  }
  ~Solver() {
    cudaSetDevice(opt.device); cudaDeviceSynchronize();
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    owned_sliced.push_back(&tA); owned_sliced.push_back(&tB); owned_sliced.push_back(&side.tA); owned_sliced.push_back(&side.tB); owned_sliced.push_back(&side2.tA); owned_sliced.push_back(&side2.tB);
    owned_sliced.push_back(&side3.tA); owned_sliced.push_back(&side3.tB); if (side3.tc_bytes) cudaFree(side3.tc_bytes); if (side3.tc_top) cudaFree(side3.tc_top); if (side3.tc_raw) cudaFree(side3.tc_raw); cudaStreamDestroy(side3.st); cudaEventDestroy(evS0); cudaEventDestroy(evS1);
    if (side2.tc_bytes) cudaFree(side2.tc_bytes); if (side2.tc_top) cudaFree(side2.tc_top); if (side2.tc_raw) cudaFree(side2.tc_raw); cudaStreamDestroy(side2.st); for (cudaEvent_t e : {evR0, evR1, evE0, evE1}) if (e) cudaEventDestroy(e);
    if (side.tc_bytes) cudaFree(side.tc_bytes); if (side.tc_top) cudaFree(side.tc_top); if (side.tc_raw) cudaFree(side.tc_raw); cudaEventDestroy(evY0); cudaEventDestroy(evY1); cudaStreamDestroy(side.st);
    for (int k = 1; k < NCTX; k++) { Ctx& c = pctx[k]; owned_sliced.push_back(&c.tA); owned_sliced.push_back(&c.tB); if (c.tc_bytes) cudaFree(c.tc_bytes); if (c.tc_top) cudaFree(c.tc_top); if (c.tc_raw) cudaFree(c.tc_raw);
      if (c.ev) cudaEventDestroy(c.ev); if (c.st) cudaStreamDestroy(c.st); }
    if (evFork) cudaEventDestroy(evFork);
    for (Sliced* s : owned_sliced) { if (s->sl) cudaFree(s->sl); if (s->E) cudaFree(s->E); if (s->planes) cudaFree(s->planes); }
    if (tc_bytes) cudaFree(tc_bytes); if (tc_top) cudaFree(tc_top); if (tc_raw) cudaFree(tc_raw); if (pe0) cudaEventDestroy(pe0); if (pe1) cudaEventDestroy(pe1);
    for (void* p : allocs) cudaFree(p);
    if (wstage) cudaFree(wstage); if (gbuf) cudaFree(gbuf); if (lane_buf) cudaFree(lane_buf); if (lane_E) cudaFree(lane_E);
    drop_graph(); for (auto& e : ev) cudaEventDestroy(e); for (auto& r : evD) for (auto& e : r) cudaEventDestroy(e);
    cudaStreamDestroy(st);
  }
  int set_option_num(int which, const void* w) override { if (which < 0 || which > 9) return CLRS_ERR_ARG; w2m(hopt[which], w); return 0; }
  int set_free(int N_, const void* b_, const void* constant, int maximize_) override {
    N = N_; hb.resize(N); for (int i = 0; i < N; i++) w2m(hb[i], (const char*)b_ + i * wire_size()); w2m(hconst, constant); maximize = maximize_ != 0; return 0;
  }
  int add_cluster(int j, int P, const void* B, const void* c_) override {
    if (j != (int)cl.size()) { err = "clusters must be added in order"; return CLRS_ERR_ARG; }
    cl.emplace_back(); Clu& c0 = cl.back(); c0.P = P; c0.hB.resize((size_t)P * N); c0.hc.resize(P);
    for (size_t i = 0; i < c0.hB.size(); i++) w2m(c0.hB[i], (const char*)B + i * wire_size());
    for (int i = 0; i < P; i++) w2m(c0.hc[i], (const char*)c_ + i * wire_size());
    return 0;
  }
  int add_block(int j, int l, int m, int delta, int high_rank, const void* C) override {
    if (j < 0 || j >= (int)cl.size() || l != (int)cl[j].blocks.size()) { err = "blocks must be added in order"; return CLRS_ERR_ARG; }
    if (m < 1 || delta < 1) { err = "clrs_add_block: m and delta must be positive"; return CLRS_ERR_ARG; }
    if (high_rank && m != 1) { err = "dense blocks have one subblock"; return CLRS_ERR_ARG; }
    cl[j].blocks.emplace_back(); Block& b0 = cl[j].blocks.back(); b0.j = j; b0.l = l; b0.m = m; b0.delta = delta; b0.n = m * delta; b0.high_rank = high_rank != 0;
    b0.hC.resize((size_t)b0.n * b0.n); for (size_t i = 0; i < b0.hC.size(); i++) w2m(b0.hC[i], (const char*)C + i * wire_size());
    return 0;
  }
  // next n x n slot of the block's dense constraint buffer (grown geometrically, at most P_j slots)
  num* dense_slot(int j, Block& b0) {
    const size_t nn = (size_t)b0.n * b0.n, have = b0.dense_p.size();
    if (have + 1 > b0.Aall_cap) { const size_t cap = std::max<size_t>(8, std::min<size_t>(2 * b0.Aall_cap, (size_t)cl[j].P)); const size_t ncap = std::max(cap, have + 1);
      num* nb_ = nullptr; CK(cudaMalloc((void**)&nb_, ncap * nn * sizeof(num))); allocs.push_back(nb_);
      if (b0.Aall) { CK(cudaMemcpyAsync(nb_, b0.Aall, have * nn * sizeof(num), cudaMemcpyDeviceToDevice, st)); CK(cudaStreamSynchronize(st)); release(b0.Aall); }
      b0.Aall = nb_; b0.Aall_cap = ncap; }
    return b0.Aall + have * nn;
  }
  int add_dense_term(int j, int l, int p_, const void* A) override {
    if (j < 0 || j >= (int)cl.size() || l < 0 || l >= (int)cl[j].blocks.size()) { err = "clrs_add_dense_term: no such block"; return CLRS_ERR_ARG; }
    if (p_ < 0 || p_ >= cl[j].P) { err = "clrs_add_dense_term: constraint row out of range"; return CLRS_ERR_ARG; }
    Block& b0 = cl[j].blocks[l]; if (!b0.high_rank) { err = "dense term on a low-rank block"; return CLRS_ERR_ARG; }
    // the matrix goes to the device as raw wire bytes and is converted there (27e6 numbers for the BASELINE workload: a host
    // loop over them cost more than the whole upload)
    wire_to_device(dense_slot(j, b0), A, (size_t)b0.n * b0.n);
    b0.dense_p.push_back(p_);
    return 0;
  }
  // the same constraint matrix given by its nonzero entries (SDPA-sparse input, src/SDPAtoCLRS.jl:3-31): only the triplets cross
  // PCIe (MAX-CUT n = 300: 300 numbers instead of 27e6), the dense slot is zeroed and filled on the device
  int add_sparse_term(int j, int l, int p_, int nnz, const int32_t* rows, const int32_t* cols, const void* vals, int mirror) override {
    if (j < 0 || j >= (int)cl.size() || l < 0 || l >= (int)cl[j].blocks.size()) { err = "clrs_add_sparse_term: no such block"; return CLRS_ERR_ARG; }
    if (p_ < 0 || p_ >= cl[j].P || nnz < 0) { err = "clrs_add_sparse_term: constraint row out of range"; return CLRS_ERR_ARG; }
    Block& b0 = cl[j].blocks[l]; if (!b0.high_rank) { err = "sparse term on a low-rank block"; return CLRS_ERR_ARG; }
    for (int t = 0; t < nnz; t++) if (rows[t] < 0 || rows[t] >= b0.n || cols[t] < 0 || cols[t] >= b0.n) { err = "clrs_add_sparse_term: entry out of range"; return CLRS_ERR_ARG; }
    { std::vector<int64_t> pos; pos.reserve(2 * (size_t)nnz);                 // every position at most once (the scatter kernel writes without ordering)
      for (int t = 0; t < nnz; t++) { pos.push_back((int64_t)rows[t] * b0.n + cols[t]); if (mirror && rows[t] != cols[t]) pos.push_back((int64_t)cols[t] * b0.n + rows[t]); }
      std::sort(pos.begin(), pos.end()); if (std::adjacent_find(pos.begin(), pos.end()) != pos.end()) { err = "clrs_add_sparse_term: a position is given twice"; return CLRS_ERR_ARG; } }
    num* slot = dense_slot(j, b0); const size_t nn = (size_t)b0.n * b0.n;
    CK(cudaMemsetAsync(slot, 0, nn * sizeof(num), st));                  // all-zero bytes are the number 0 (sign = 0)
    if (nnz > 0) { int32_t* drc = nullptr; num* dv = nullptr; CK(cudaMalloc((void**)&drc, 2 * (size_t)nnz * sizeof(int32_t))); CK(cudaMalloc((void**)&dv, (size_t)nnz * sizeof(num)));
      CK(cudaMemcpyAsync(drc, rows, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(drc + nnz, cols, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      wire_to_device(dv, vals, (size_t)nnz);
      nlaunch++, k_scatter_triplets<NL><<<(nnz + 127) / 128, 128, 0, st>>>(nnz, drc, drc + nnz, dv, b0.n, mirror, slot);
      CK(cudaStreamSynchronize(st)); CK(cudaFree(drc)); CK(cudaFree(dv)); }
    b0.dense_p.push_back(p_);
    return 0;
  }
  int add_lowrank_term(int j, int l, int r, int s, int p_, int rank, const void* lam, const void* vs, const void* ws) override {
    if (j < 0 || j >= (int)cl.size() || l < 0 || l >= (int)cl[j].blocks.size()) { err = "clrs_add_lowrank_term: no such block"; return CLRS_ERR_ARG; }
    Block& b0 = cl[j].blocks[l]; if (b0.high_rank) { err = "low-rank term on a dense block"; return CLRS_ERR_ARG; }
    if (p_ < 0 || p_ >= cl[j].P || r < 0 || r >= b0.m || s < 0 || s >= b0.m || rank < 0) { err = "clrs_add_lowrank_term: index out of range"; return CLRS_ERR_ARG; }
    for (int k = 0; k < rank; k++) { b0.lr.emplace_back(); HTerm& t = b0.lr.back(); t.r = r; t.s = s; t.p = p_; t.k = k;
      w2m(t.lam, (const char*)lam + k * wire_size()); t.v.resize(b0.delta); t.w.resize(b0.delta);
      for (int a = 0; a < b0.delta; a++) { w2m(t.v[a], (const char*)vs + ((size_t)k * b0.delta + a) * wire_size()); w2m(t.w[a], (const char*)ws + ((size_t)k * b0.delta + a) * wire_size()); } }
    return 0;
  }
  static bool same_vec(const std::vector<num>& a, const std::vector<num>& b) {
    for (size_t i = 0; i < a.size(); i++) { if (a[i].sign != b[i].sign) return false; if (a[i].sign == 0) continue; if (a[i].exp != b[i].exp || memcmp(a[i].l, b[i].l, sizeof(a[i].l))) return false; }
    return true;
  }
  static uint64_t vec_hash(const std::vector<num>& a) {       // FNV-1a over sign, exponent and limbs (zeros hash alike whatever their other fields)
    uint64_t h = 1469598103934665603ull; auto mix = [&](uint32_t w) { h ^= w; h *= 1099511628211ull; };
    for (auto& x : a) { mix((uint32_t)x.sign); if (x.sign == 0) continue; mix((uint32_t)x.exp); for (int q = 0; q < NL; q++) mix(x.l[q]); }
    return h; }
  Sliced& own(Sliced& s) { owned_sliced.push_back(&s); return s; }

  // ---- finalize: build tables (precompute_matrices_bilinear_pairings, src/solver.jl:985-1059), allocate, initialise ----
  int finalize() override {
    Ptot = 0; Ksum = 0; tot = 0; gtot = 0; blk.clear();
    { // clusters (and the blocks of split clusters) -> ranks: plan_shards above; weights P^3 + sum n^3 (src/threadinginfo.jl:88,97)
      static const int split_env = getenv("CLRS_SPLIT_BLOCKS") ? atoi(getenv("CLRS_SPLIT_BLOCKS")) : -1;      // 0 never, 1 every multi-block cluster, default by weight
      static const int bigp0 = getenv("CLRS_BIG_CLUSTER") ? atoi(getenv("CLRS_BIG_CLUSTER")) : 512;
      auto bwf = [](const Block& b0) { const double n3 = (double)b0.n * b0.n * b0.n; return n3 * (b0.high_rank ? 2.0 * b0.dense_p.size() + 15 : 15); };
      std::vector<double> p3; std::vector<std::vector<double>> bw; std::vector<int> bigc;
      for (auto& c0 : cl) { p3.push_back((double)c0.P * c0.P * c0.P); bw.emplace_back(); for (auto& b0 : c0.blocks) bw.back().push_back(bwf(b0));
        bigc.push_back(nranks > 1 && N > 0 && bigp0 > 0 && c0.P >= bigp0); }                                   // column-split clusters keep their blocks together
      const ShardPlan pl = plan_shards(p3, bw, bigc, nranks, split_env);
      for (size_t j = 0; j < cl.size(); j++) { Clu& c0 = cl[j]; c0.split = pl.split[j] != 0; c0.owner = pl.cluster_owner[j]; c0.lead = c0.owner == rank; c0.owned = c0.split || c0.lead;
        for (size_t l = 0; l < c0.blocks.size(); l++) { c0.blocks[l].brank = pl.block_owner[j][l]; c0.blocks[l].mine = pl.block_owner[j][l] == rank; } } }
    for (auto& c0 : cl) { c0.off = Ptot; Ptot += c0.P;
      for (auto& b0 : c0.blocks) { Ksum += b0.n; b0.goff = gtot; gtot += (int64_t)b0.n * b0.n; if (c0.owned && b0.mine) { b0.off = tot; tot += (int64_t)b0.n * b0.n; blk.push_back(&b0); } } }
    std::vector<int64_t> boff; std::vector<int32_t> bn; for (Block* b0 : blk) { boff.push_back(b0->off); bn.push_back(b0->n); } boff.push_back(tot);
    d_boff = upload(boff); d_bn = upload(bn); bt.off = d_boff; bt.n = d_bn; bt.nblocks = (int)blk.size();
    num** flat[] = {&X, &Y, &Cm, &L, &Minv, &Xi, &R, &P, &dX, &dY, &T1, &TXY, &U, &LY, &MinvY};
    for (auto f : flat) *f = dalloc<num>(tot);
    { std::vector<num> hC(tot); for (Block* b0 : blk) std::copy(b0->hC.begin(), b0->hC.end(), hC.begin() + b0->off); CK(cudaMemcpyAsync(Cm, hC.data(), tot * sizeof(num), cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); }
    x = dalloc<num>(Ptot); d = dalloc<num>(Ptot); dx = dalloc<num>(Ptot); tr = dalloc<num>(Ptot);
    y = dalloc<num>(N); p = dalloc<num>(N); dy = dalloc<num>(N); tmpN = dalloc<num>(N); Q = dalloc<num>((size_t)N * N); QMinv = dalloc<num>((size_t)N * N);
    { std::vector<num> hc, hco; num z; mp_zero(z); for (auto& c0 : cl) { if (c0.owned) hc.insert(hc.end(), c0.hc.begin(), c0.hc.end()); else hc.insert(hc.end(), c0.P, z);
        if (c0.owned && c0.lead) hco.insert(hco.end(), c0.hc.begin(), c0.hc.end()); else hco.insert(hco.end(), c0.P, z); }
      c = upload(hc); cobj = upload(hco); b = upload(hb); }   // c of clusters owned elsewhere reads as 0; cobj: <c,x> is summed over the ranks, a split cluster counts once
    tmpU = dalloc<num>(N); q_ready = dalloc<unsigned>((size_t)(N + 31) / 32 + 1);
    Td = dalloc<double>(tot); lamX = dalloc<double>(blk.size()); lamY = dalloc<double>(blk.size());
    { std::vector<EigTask> et; size_t vtot = 0; for (Block* b0 : blk) vtot += (size_t)b0->n * (std::min(b0->n, EIG_MMAX) + 1); eigV = dalloc<double>(vtot); size_t o = 0;
      for (Block* b0 : blk) { EigTask t; t.T = Td + b0->off; t.n = b0->n; t.V = eigV + o; o += (size_t)b0->n * (std::min(b0->n, EIG_MMAX) + 1); et.push_back(t); } eigT = upload(et);
      Td2 = dalloc<double>(tot); eigV2 = dalloc<double>(vtot); U2 = dalloc<num>(tot); T1b = dalloc<num>(tot);
      for (auto& t : et) { t.T = Td2 + (t.T - Td); t.V = eigV2 + (t.V - eigV); } eigT2 = upload(et); }
    // the L_j^-1 B_j of the owned clusters are the row blocks of ONE matrix, so Q = (vcat LinvB)^T (vcat LinvB) is one product (:1268-1269)
    // Big clusters of a multi-rank solve: the owner factors S_j and broadcasts L_j; every rank solves
    // L_j^-1 B_j for ITS columns of B_j, the column chunks are all-gathered, and every rank forms its N x ncr slab of
    // G_j^T G_j, which it adds into its partial Q before the one all-reduce of Q.  No collective beyond broadcast + all-gather.
    static const int bigp = getenv("CLRS_BIG_CLUSTER") ? atoi(getenv("CLRS_BIG_CLUSTER")) : 512;
    ncr = nranks > 1 ? (N + nranks - 1) / nranks : N;
    for (auto& c0 : cl) c0.big = nranks > 1 && N > 0 && bigp > 0 && c0.P >= bigp;
    Pown = 0; for (auto& c0 : cl) if (c0.owned && c0.lead && !c0.big) Pown += c0.P;
    LinvBall = dalloc<num>((size_t)Pown * N);
    { int64_t o = 0; for (auto& c0 : cl) if (c0.owned && c0.lead && !c0.big) { c0.LinvB = LinvBall + o * N; o += c0.P; } }
    for (auto& c0 : cl) if (c0.owned && !c0.lead && !c0.big) c0.LinvB = dalloc<num>((size_t)c0.P * N);
    for (auto& c0 : cl) {
      if (c0.big) {                                                    // on every rank: the factor, this rank's columns of B, all chunks of G, the Q slab
        std::vector<num> hBc((size_t)c0.P * ncr); num z; mp_zero(z);
        for (int i = 0; i < c0.P; i++) for (int c = 0; c < ncr; c++) { const int col = rank * ncr + c; hBc[(size_t)i * ncr + c] = col < N ? c0.hB[(size_t)i * N + col] : z; }
        c0.Bc = upload(hBc); c0.Gc = dalloc<num>((size_t)nranks * c0.P * ncr); c0.Qs = dalloc<num>((size_t)nranks * ncr * ncr);
        if (!c0.owned) { c0.S = dalloc<num>((size_t)c0.P * c0.P); c0.Minv = dalloc<num>((size_t)c0.P * c0.P); }
      }
      if (!c0.owned) { for (auto& b0 : c0.blocks) if (b0.Aall) { release(b0.Aall); b0.Aall = nullptr; } continue; }   // uploaded before the partition was known
      c0.B = upload(c0.hB); c0.S = dalloc<num>((size_t)c0.P * c0.P); c0.Minv = dalloc<num>((size_t)c0.P * c0.P); c0.t = dalloc<num>(c0.P); c0.ready = dalloc<unsigned>((size_t)(c0.P + 31) / 32 + 1);
      for (auto& b0 : c0.blocks) { if (!b0.mine) { if (b0.Aall) { release(b0.Aall); b0.Aall = nullptr; } continue; } if (int rc = finalize_block(c0, b0)) return rc; }
    }
    // scalars
    { std::vector<num> h(SC_COUNT); for (auto& v : h) mp_zero(v);
      mp_set_i32(h[SC_K], Ksum); mp_set_i32(h[SC_ONE], 1);
      h[SC_BETA_INF] = hopt[0]; h[SC_BETA_FEAS] = hopt[1]; h[SC_GAMMA] = hopt[2]; h[SC_OMEGA_P] = hopt[3]; h[SC_OMEGA_D] = hopt[4];
      h[SC_GAPTHR] = hopt[5]; h[SC_DERRTHR] = hopt[6]; h[SC_PERRTHR] = hopt[7]; h[SC_MAXGAP] = hopt[8]; h[SC_STEPTHR] = hopt[9]; h[SC_CONSTANT] = hconst;
      CK(cudaMemcpyAsync(sc, h.data(), SC_COUNT * sizeof(num), cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); }
    // x = 0, y = 0, X = omega_p I, Y = omega_d I  (src/solver.jl:187-201)
    for (Block* b0 : blk) { nlaunch++, k_set_diag<NL><<<(b0->n + 127) / 128, 128, 0, st>>>(b0->n, X + b0->off, b0->n, sc + SC_OMEGA_P); nlaunch++, k_set_diag<NL><<<(b0->n + 127) / 128, 128, 0, st>>>(b0->n, Y + b0->off, b0->n, sc + SC_OMEGA_D); }
    finalized = true;
    initial_quantities();
    return 0;
  }
  int finalize_block(Clu& c0, Block& b0) {
    const int n = b0.n, m = b0.m, dl = b0.delta;
    own(b0.YS); own(b0.XiS); own(b0.MS); own(b0.MSY); b0.lay = use_tc(n, n, n) ? 1 : 0;
    if (b0.high_rank) {
      b0.np = (int)b0.dense_p.size();
      if (use_tc(b0.np * n, n, n)) b0.lay = 1;          // the (np n) x n x n Schur products decide the panel layout of a dense block (n = 100: 7.1 ms on the CUDA cores -> 5.0 ms)
      std::vector<int32_t> pl(b0.dense_p.begin(), b0.dense_p.end()); b0.d_plist = upload(pl);
      { // nonzero structure of the A_p (their zero entries contribute exact zeros to <A_p,Z> and sum_p x_p A_p): per p the list of
        // nonzero positions, and per position the list of p that are nonzero there.  The mask is computed on the device; the host only
        // turns the bytes into the two index lists.
        const size_t nn = (size_t)n * n, cnt = (size_t)b0.np * nn; std::vector<uint8_t> mask(cnt);
        if (cnt) { uint8_t* dm = nullptr; CK(cudaMalloc((void**)&dm, cnt)); nlaunch++, k_nonzero_mask<NL><<<grid_for((int64_t)cnt), 256, 0, st>>>((int64_t)cnt, b0.Aall, dm);
          CK(cudaMemcpyAsync(mask.data(), dm, cnt, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); CK(cudaFree(dm)); }
        std::vector<int32_t> st_(b0.np + 1, 0), idx, tcount(nn + 1, 0);
        for (int i = 0; i < b0.np; i++) { const uint8_t* mrow = mask.data() + (size_t)i * nn; for (size_t e = 0; e < nn; e++) if (mrow[e]) { idx.push_back((int32_t)e); tcount[e + 1]++; } st_[i + 1] = (int32_t)idx.size(); }
        for (size_t e = 0; e < nn; e++) tcount[e + 1] += tcount[e];
        std::vector<int32_t> tp(idx.size()), fill(tcount.begin(), tcount.end() - 1);
        for (int i = 0; i < b0.np; i++) for (int32_t q = st_[i]; q < st_[i + 1]; q++) tp[fill[idx[q]]++] = i;       // ascending p per position
        b0.nnz = (int64_t)idx.size(); b0.nz_start = upload(st_); b0.nz_idx = upload(idx); b0.nzT_start = upload(tcount); b0.nzT_p = upload(tp); }
      if (!b0.Aall) b0.Aall = dalloc<num>(1);
      b0.Sd = dalloc<num>((size_t)b0.np * b0.np);
      { // opt-in sparsity shortcut: 3 multiplications per pair of nonzero entries on the CUDA cores against the two
        // n^3 products per constraint + the inner products on the tensor cores (roughly 4 x the multi-precision multiply rate)
        static const int sp_env = getenv("CLRS_SPARSE_SCHUR") ? atoi(getenv("CLRS_SPARSE_SCHUR")) : -1;
        const bool want = sp_env >= 0 ? sp_env != 0 : opt.sparse_schur != 0;
        const double dense_macs = 2.0 * b0.np * (double)n * n * n + 0.5 * (double)b0.np * b0.np * n * n, sparse_macs = 1.5 * (double)b0.nnz * (double)b0.nnz;
        b0.sparse = want && b0.np > 0 && 4.0 * sparse_macs < dense_macs; }
      if (b0.sparse) return 0;
      b0.T1 = dalloc<num>((size_t)b0.np * n * n); b0.T2 = dalloc<num>((size_t)b0.np * n * n);
      own(b0.AallB); own(b0.AallV); own(b0.T1S); own(b0.T2V);
      if (b0.np > 0) {
        VecView v; v.base = b0.Aall; v.bstride = (int64_t)n * n; v.vper = n; v.sv = 1; v.sk = n; v.nvec = b0.np * n; v.K = n; split(b0.AallB, v, false, b0.lay);   // columns of every A_p
        // A_q flattened for the inner products <A_q, T_p>.  The reference requires symmetric constraint matrices (src/checks.jl); when
        // they are (checked bit for bit here) and the product is a tensor-core shape, the inner products run over the packed
        // upper triangle with T_p symmetrised on the fly: K = n (n + 1) / 2 instead of n^2
        static const int tri_off = getenv("CLRS_SCHUR_TRI") ? atoi(getenv("CLRS_SCHUR_TRI")) == 0 : 0;
        const bool tcv = use_tc(b0.np, b0.np, n * n);
        if (tcv && !tri_off) { int* fl = dalloc<int>(1); nlaunch++, k_check_symmetric<NL><<<grid_for((int64_t)b0.np * n * n), 256, 0, st>>>(b0.Aall, b0.np, n, fl);
          int h = 1; CK(cudaMemcpyAsync(&h, fl, sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); b0.tri = (h == 0); }
        if (b0.tri) { ensure(b0.AallV, b0.np, n * (n + 1) / 2, 1); split_tri(b0.AallV, 0, b0.Aall, b0.np, n, false); }
        else { VecView w; w.base = b0.Aall; w.bstride = 0; w.vper = b0.np; w.sv = (int64_t)n * n; w.sk = 1; w.nvec = b0.np; w.K = n * n; split(b0.AallV, w, true, tcv ? 1 : 0); }
      }
      return 0;
    }
    // low rank: deduplicated bases per subblock row r
    std::map<std::tuple<int, int, int, int>, int> find;
    for (size_t e = 0; e < b0.lr.size(); e++) { auto& t = b0.lr[e]; find[{t.r, t.s, t.p, t.k}] = (int)e; if (t.r >= m || t.s >= m || t.p >= c0.P) { err = "low-rank term out of range"; return CLRS_ERR_ARG; } }
    b0.u_r.assign(m, 0); b0.ul_r.assign(m, 0); b0.V.assign(m, nullptr); b0.W.assign(m, nullptr); b0.Vs.resize(m); b0.Ws.resize(m);
    for (int r = 0; r < m; r++) {
      std::vector<int> uv, uw;
      // unique vectors in order of first appearance (unique_idx, src/tools.jl:128-145), found through a hash of the limbs
      std::unordered_multimap<uint64_t, int> hv, hw;
      auto find_or_add = [&](std::unordered_multimap<uint64_t, int>& tab, std::vector<int>& uniq, const std::vector<num>& vec, bool isv) {
        const uint64_t key = vec_hash(vec); auto range = tab.equal_range(key);
        for (auto it = range.first; it != range.second; ++it) if (same_vec(isv ? b0.lr[uniq[it->second]].v : b0.lr[uniq[it->second]].w, vec)) return it->second;
        return -1; };
      for (int s = 0; s < m; s++) for (size_t e = 0; e < b0.lr.size(); e++) { auto& t = b0.lr[e]; if (t.r != r || t.s != s) continue;
        int f = find_or_add(hv, uv, t.v, true);
        if (f < 0) { f = (int)uv.size(); uv.push_back((int)e); hv.emplace(vec_hash(t.v), f); } t.colV = f;
        f = find_or_add(hw, uw, t.w, false);
        if (f < 0) { f = (int)uw.size(); uw.push_back((int)e); hw.emplace(vec_hash(t.w), f); } t.rowW = f; }
      b0.u_r[r] = (int)uv.size(); b0.ul_r[r] = (int)uw.size(); b0.umax = std::max(b0.umax, std::max(b0.u_r[r], b0.ul_r[r]));
      std::vector<num> hV((size_t)dl * uv.size()), hW((size_t)uw.size() * dl);
      for (size_t u = 0; u < uv.size(); u++) for (int a = 0; a < dl; a++) hV[(size_t)a * uv.size() + u] = b0.lr[uv[u]].v[a];
      for (size_t u = 0; u < uw.size(); u++) for (int a = 0; a < dl; a++) hW[u * dl + a] = b0.lr[uw[u]].w[a];
      b0.V[r] = upload(hV); b0.W[r] = upload(hW);
      own(b0.Vs[r]); own(b0.Ws[r]);
      split_cols(b0.Vs[r], b0.V[r], b0.u_r[r], dl, b0.u_r[r]); split_rows(b0.Ws[r], b0.W[r], dl, b0.ul_r[r], dl);
    }
    // constraint-sorted term table
    std::vector<int> order(b0.lr.size()); for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b_) { return b0.lr[a].p < b0.lr[b_].p; });
    std::vector<int> pos(b0.lr.size()); for (size_t i = 0; i < order.size(); i++) pos[order[i]] = (int)i;
    std::vector<LRTermDev> terms(order.size()); std::vector<num> lam(order.size()); std::vector<int32_t> plist, tstart;
    for (size_t i = 0; i < order.size(); i++) { auto& t = b0.lr[order[i]]; auto it = find.find({t.s, t.r, t.p, t.k});
      if (it == find.end()) { err = "low-rank term without its transposed subblock (src/solver.jl:1009)"; return CLRS_ERR_ARG; }
      auto& tt = b0.lr[it->second];
      LRTermDev d0; d0.p = t.p; d0.r = t.r; d0.s = t.s; d0.lam = (int)i; d0.colV = t.colV; d0.rowW = t.rowW; d0.rowW_t = tt.rowW; d0.colV_t = tt.colV; terms[i] = d0; lam[i] = t.lam;
      if (plist.empty() || plist.back() != t.p) { plist.push_back(t.p); tstart.push_back((int)i); } }
    tstart.push_back((int)order.size());
    b0.nP = (int)plist.size(); b0.lr_plist = upload(plist); b0.lr_tstart = upload(tstart); b0.lr_terms = upload(terms); b0.lr_lam = upload(lam);
    // pairing matrices
    b0.BX.assign(m * m, nullptr); b0.BY.assign(m * m, nullptr); b0.ZV.assign(m * m, nullptr);
    std::vector<MatRef> rBX(m * m), rBY(m * m), rZV(m * m), rW(m);
    for (int s = 0; s < m; s++) for (int r = 0; r < m; r++) { size_t sz = (size_t)b0.ul_r[s] * b0.u_r[r];
      b0.BX[s * m + r] = dalloc<num>(sz); b0.BY[s * m + r] = dalloc<num>(sz); rBX[s * m + r] = {b0.BX[s * m + r], b0.u_r[r]}; rBY[s * m + r] = {b0.BY[s * m + r], b0.u_r[r]}; }
    for (int r = 0; r < m; r++) { rW[r] = {b0.W[r], dl}; for (int s = 0; s <= r; s++) { b0.ZV[r * m + s] = dalloc<num>((size_t)dl * b0.u_r[r]); rZV[r * m + s] = {b0.ZV[r * m + s], b0.u_r[r]}; } }
    b0.d_BX = upload(rBX); b0.d_BY = upload(rBY); b0.d_ZV = upload(rZV); b0.d_W = upload(rW);
    b0.part = dalloc<num>((size_t)n * std::max(b0.umax, 1));
    // per (r,s<=r): piece lists for sum_p a_p A_p
    b0.rs.resize(m * m);
    for (int r = 0; r < m; r++) for (int s = 0; s <= r; s++) { auto& q = b0.rs[r * m + s]; std::vector<int32_t> el; std::vector<num> H;
      for (size_t e = 0; e < b0.lr.size(); e++) { auto& t = b0.lr[e]; if (t.r != r || t.s != s) continue; el.push_back(pos[e]); for (int a = 0; a < dl; a++) H.push_back(t.w[a]); }
      q.cnt = (int)el.size(); if (q.cnt == 0) continue;
      q.elist = upload(el); q.H = upload(H); q.G = dalloc<num>((size_t)dl * q.cnt); own(q.Hs); split_cols(q.Hs, q.H, dl, q.cnt, dl); }
    return 0;
  }

  // ---- pieces of the iteration ---------------------------------------------------------
  // dst_b = sum_p a_p A_p per block  (compute_weighted_A!, src/solver.jl:1410-1470)
  void weighted_A(num* dst, const num* a) {
    par_blocks([&](Block* bp) { Block& b0 = *bp; Clu& c0 = cl[b0.j];
      num* M = dst + b0.off; const int n = b0.n, m = b0.m, dl = b0.delta; const num* aj = a + c0.off;
      if (b0.high_rank) { nlaunch++, k_weighted_dense<NL><<<grid_for((int64_t)n * n), 256, 0, st>>>(b0.d_plist, b0.Aall, (int64_t)n * n, b0.nzT_start, b0.nzT_p, aj, M); return; }
      zero(M, (int64_t)n * n);
      for (int r = 0; r < m; r++) for (int s = 0; s <= r; s++) { auto& q = b0.rs[r * m + s]; if (q.cnt == 0) continue;
        nlaunch++, k_weighted_cols<NL><<<(q.cnt * dl + 127) / 128, 128, 0, st>>>(q.cnt, q.elist, b0.lr_terms, b0.lr_lam, aj, MatRef{b0.V[r], b0.u_r[r]}, dl, q.G, q.cnt);
        split_rows(tA, q.G, q.cnt, dl, q.cnt);
        gemm(tA, 0, q.Hs, 0, dl, dl, M + (int64_t)r * dl * n + (int64_t)s * dl, n); }
      if (m > 1) nlaunch++, k_mirror<NL><<<grid_for((int64_t)n * n), 256, 0, st>>>(n, M, n, 0);
    });
  }
  // out[p] = <A_p, Z>  (trace_A with vectors, src/solver.jl:1290-1366)
  void trace_vectors(num* out, const num* Z) {
    zero(out, Ptot);
    par_blocks([&](Block* bp) { Block& b0 = *bp;                                                   // the products Z V_r of all blocks side by side
      const num* Zb = Z + b0.off; const int n = b0.n, m = b0.m, dl = b0.delta;
      if (b0.high_rank || b0.nP == 0) return;
      for (int r = 0; r < m; r++) for (int s = 0; s <= r; s++) { if (b0.rs[r * m + s].cnt == 0) continue;
        split_rows(tA, Zb + (int64_t)r * dl * n + (int64_t)s * dl, n, dl, dl); gemm(tA, 0, b0.Vs[r], 0, dl, b0.u_r[r], b0.ZV[r * m + s], b0.u_r[r]); } });
    par_clusters([&](Clu& c0) { for (auto& b0 : c0.blocks) { if (!b0.mine) continue;               // blocks of one cluster add into the same rows: in order
      const num* Zb = Z + b0.off; const int n = b0.n, m = b0.m, dl = b0.delta; num* oj = out + c0.off;
      if (b0.high_rank) { if (b0.np) nlaunch++, k_trace_dense<NL><<<b0.np, 128, 128 * sizeof(num), st>>>(b0.np, b0.d_plist, b0.Aall, (int64_t)n * n, b0.nz_start, b0.nz_idx, Zb, oj); continue; }
      if (b0.nP == 0) continue;
      nlaunch++, k_trace_vectors<NL><<<(b0.nP + 63) / 64, 64, 0, st>>>(b0.nP, b0.lr_plist, b0.lr_tstart, b0.lr_terms, b0.lr_lam, b0.d_W, b0.d_ZV, m, dl, oj);
    } });
    sum_split_rows(out);
  }
  // out[p] = <A_p, Y> from the stored pairings (src/solver.jl:1368-1407)
  void trace_pairings(num* out) {
    zero(out, Ptot);
    par_clusters([&](Clu& c0) { for (auto& b0 : c0.blocks) { if (!b0.mine) continue; num* oj = out + c0.off;
      if (b0.high_rank) { if (b0.np) nlaunch++, k_trace_dense<NL><<<b0.np, 128, 128 * sizeof(num), st>>>(b0.np, b0.d_plist, b0.Aall, (int64_t)b0.n * b0.n, b0.nz_start, b0.nz_idx, Y + b0.off, oj); continue; }
      if (b0.nP) nlaunch++, k_trace_pairings<NL><<<(b0.nP + 63) / 64, 64, 0, st>>>(b0.nP, b0.lr_plist, b0.lr_tstart, b0.lr_terms, b0.lr_lam, b0.d_BY, b0.m, oj); } });
    sum_split_rows(out);
  }
  // rows of a split cluster hold this rank's blocks only: summed over the ranks (main stream, cluster order: the same order on every rank)
  void sum_split_rows(num* out) { if (nranks > 1) for (auto& c0 : cl) if (c0.split && c0.P) allreduce(out + c0.off, c0.P, 0); }
  // P, d, p  (compute_residuals!, src/solver.jl:863-918); `tr` must hold <A_*, Y>
  void residuals() {
    weighted_A(P, x);
    k_residual_P<NL><<<grid_for(tot), 256, 0, st>>>(tot, P, X, Cm, maximize);
    // d = c - B y - tr   (segments of clusters owned by other ranks stay 0)
    addsub(d, c, 1, tr, -1, Ptot);
    if (N > 0) for (auto& c0 : cl) if (c0.owned && c0.P) nlaunch++, k_gemv_n<NL><<<(c0.P * 32 + 255) / 256, 256, 0, st>>>(c0.P, N, c0.B, N, y, d + c0.off, -1, 1);
    // p = +-b - sum_j B_j^T x_j   (partial sums over the owned clusters, combined over the ranks)
    if (N > 0) { zero(p, N);
      for (auto& c0 : cl) if (c0.owned && c0.lead && c0.P) nlaunch++, k_gemv_t<NL><<<(N + 31) / 32, 256, 0, st>>>(c0.P, N, c0.B, N, x + c0.off, p, -1, 1);
      allreduce(p, N, 0);
      addsub(p, p, 1, b, maximize ? 1 : -1, N); }
  }
  void errors() { reduce(P, nullptr, tot, sc + SC_ERRP, 0); allreduce(sc + SC_ERRP, 1, 1); reduce(p, nullptr, N, sc + SC_ERRp, 0); reduce(d, nullptr, Ptot, sc + SC_ERRd, 0); allreduce(sc + SC_ERRd, 1, 1); }
  void objectives() {
    reduce(cobj, x, Ptot, sc + SC_CX, 0); reduce(Cm, Y, tot, sc + SC_CY, 0); allreduce(sc + SC_CX, 2, 0);   // SC_CX, SC_CY are adjacent
    reduce(b, y, N, sc + SC_BY, 0);
    scalar(4);
  }
  void scalar(int phase, const num* M = nullptr, const num* dM = nullptr, const double* lam = nullptr, int which = 0) {
    ScalarCfg cfg{opt.correctoronly, opt.safe_step, maximize};
    nlaunch++, k_scalar<NL><<<1, 32, 0, st>>>(phase, sc, flags, dinfo, cfg, (int)blk.size(), d_bn, d_boff, M, dM, lam, which);
  }
  void pull_info() {
    double h[32]; int f[FL_COUNT];
    CK(cudaMemcpyAsync(h, dinfo, sizeof(h), cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(f, flags, sizeof(f), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    memcpy(hinfo, h, sizeof(h)); memcpy(hflags, f, sizeof(f));
  }
  double hinfo[32]; int hflags[FL_COUNT];
  // initial objectives / residuals / errors  (src/solver.jl:319-333)
  void initial_quantities() {
    CK(cudaMemsetAsync(flags, 0, FL_COUNT * sizeof(int), st));
    objectives();
    trace_vectors(tr, Y);
    residuals(); errors(); scalar(5);
    reduce(X, Y, tot, sc + SC_D0, 0); allreduce(sc + SC_D0, 1, 0);
    pull_info();
    h_dobj = hinfo[INFO_DOBJ + 10]; h_pobj = hinfo[INFO_POBJ + 10]; h_gap = hinfo[INFO_GAP + 10];
    h_derr = std::max(hinfo[INFO_ERRP], hinfo[INFO_ERRp]); h_perr = hinfo[INFO_ERRd]; h_pdfeas = hflags[FL_PDFEAS];
    fetch_thresholds();
    iter = 1;
  }
  // terminate() compares full-precision values; the host copy keeps them as device numbers compared on demand
  num h_gapn, h_derrn, h_perrn, h_thr[3];
  void fetch_thresholds() {
    num h[SC_COUNT]; CK(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    h_gapn = h[SC_GAP]; h_derrn = mp_cmp_abs(h[SC_ERRP], h[SC_ERRp]) > 0 ? h[SC_ERRP] : h[SC_ERRp]; h_perrn = h[SC_ERRd];
    h_thr[0] = h[SC_GAPTHR]; h_thr[1] = h[SC_DERRTHR]; h_thr[2] = h[SC_PERRTHR];
  }

  // Schur complement S_j and its factorisation  (compute_T_decomposition!, src/solver.jl:1229-1287)
  void pairings_lowrank(Block& b0) {
    const int n = b0.n, m = b0.m, dl = b0.delta;
    if (b0.nP == 0) return;
    NsScope mp(*this);                                                                                // pairings at matmul_prec (:1125-1143)
    for (int pass = 0; pass < 2; pass++) {
      const num* Src = (pass == 0 ? Y : Xi) + b0.off; auto& Bout = pass == 0 ? b0.BY : b0.BX;
      for (int r = 0; r < m; r++) { if (b0.u_r[r] == 0) continue;
        split_rows(tA, Src + (int64_t)r * dl, n, n, dl);
        gemm(tA, 0, b0.Vs[r], 0, n, b0.u_r[r], b0.part, b0.u_r[r]);                                   // part = Src[:, r-cols] V_r   (:1125,:1137)
        for (int s = 0; s < m; s++) { if (b0.ul_r[s] == 0) continue;
          split_cols(tB, b0.part + (int64_t)s * dl * b0.u_r[r], b0.u_r[r], dl, b0.u_r[r]);
          gemm(b0.Ws[s], 0, tB, 0, b0.ul_r[s], b0.u_r[r], Bout[s * m + r], b0.u_r[r]); } }                // W_s part[s-rows]            (:1131,:1143)
    }
  }
  void schur_add_lowrank(Clu& c0, Block& b0) {
    if (b0.nP == 0) return; const int m = b0.m;
    int64_t np2 = (int64_t)b0.nP * b0.nP;
    nlaunch++, k_schur_lowrank<NL><<<(unsigned)((np2 + 127) / 128), 128, 0, st>>>(b0.nP, b0.lr_plist, b0.lr_tstart, b0.lr_terms, b0.lr_lam, b0.d_BX, b0.d_BY, m, c0.S, c0.P);
  }
  // ---- dense Schur path in two stages (blocks whose A_p are all symmetric, tensor-core shapes) --------------------------------
  // <A_q, X^-1 A_p Y> = <A_q, (X^-1 W_p)^T> with W_p = A_p Y.  W_p needs only Y, so the first of the two 90000-row products
  // is issued BEFORE the Cholesky factorisation of X, on a low-priority side stream, and fills the SMs that the single-CTA
  // panel chain of chol(X) leaves idle (the main stream has the highest priority, so its kernels take the next SM that frees).
  bool staged(const Block& b0) const { static const int off = getenv("CLRS_SCHUR_STAGED") ? atoi(getenv("CLRS_SCHUR_STAGED")) == 0 : 0; return !off && b0.high_rank && !b0.sparse && b0.tri && b0.lay == 1 && b0.np > 0; }
  int dense_chunk(const Block& b0, int nsm = 148) {   // constraints per chunk: two rounds of tiles on nsm SMs (n = 300: 63 on 148 SMs)
    static const int chunks_off = getenv("CLRS_SCHUR_CHUNKS") ? atoi(getenv("CLRS_SCHUR_CHUNKS")) == 1 : 0;
    const int n = b0.n, np = b0.np;
    if (chunks_off || b0.lay != 1 || b0.AallV.lay != 1 || (int64_t)np * n < 32768) return np;
    int BNt, grp; pick_tiles(n, BNt, grp); const int ntn = (n + BNt - 1) / BNt;
    return std::max(1, ((2 * nsm / ntn) * 128) / n);
  }
  void dense_stage1(Block& b0) {                      // W_p = A_p Y, stored as W[(p,i)][b], and its split by columns (p,b) for stage 2
    const int n = b0.n, np = b0.np; const int64_t nn = (int64_t)n * n; const int pc = dense_chunk(b0), nchunk = (np + pc - 1) / pc;
    ensure(b0.T1S, np * n, n, 1);
    par_for(nchunk, [&](int c) {
      const int p0 = c * pc, cnt = std::min(pc, np - p0); if (cnt <= 0) return;
      Sliced Ac = view(b0.AallB, p0 * n, cnt * n), T1c = view(b0.T1S, p0 * n, cnt * n); num* T1p = b0.T1 + (int64_t)p0 * nn;
      { NsScope mp(*this); gemm(Ac, 0, b0.YS, 0, cnt * n, n, T1p, n); }                                // (columns of the symmetric A_p) x (columns of Y): the T Y product, at matmul_prec (:1097)
      VecView v; v.base = T1p; v.bstride = nn; v.vper = n; v.sv = 1; v.sk = n; v.nvec = cnt * n; v.K = n; split(T1c, v, false, 1, true); });
  }
  void dense_stage2(Block& b0) {                      // T_p^T[(p,b)][a] = sum_i W_p[i][b] X^-1[a][i]; S[p,q] += <A_q, T_p^T> over the packed triangle
    const int n = b0.n, np = b0.np; const int64_t nn = (int64_t)n * n; const int pc = dense_chunk(b0), nchunk = (np + pc - 1) / pc;
    ensure(b0.T2V, np, n * (n + 1) / 2, 1);
    par_for(nchunk, [&](int c) {
      const int p0 = c * pc, cnt = std::min(pc, np - p0); if (cnt <= 0) return;
      Sliced T1c = view(b0.T1S, p0 * n, cnt * n); num* T2p = b0.T2 + (int64_t)p0 * nn;
      gemm(T1c, 0, b0.XiS, 0, cnt * n, n, T2p, n);
      split_tri(b0.T2V, p0, T2p, cnt, n, true); });
    gemm(b0.AallV, 0, b0.T2V, 0, np, np, b0.Sd, np, 0, nullptr, 0, 1, 0, 0, 0, 0, 1);          // Sd[q][p], q >= p only
  }
  void pairings_dense(Block& b0) {                    // T = X^-1 A_p Y, S[p,q] += <A_q, T>   (src/solver.jl:1089-1104)
    const int n = b0.n, np = b0.np; if (np == 0) return; const int64_t nn = (int64_t)n * n;
    // chunk of constraints = a whole number of waves of the product's CTAs on the 148 SMs (n = 300: 2 column tiles x 148 row
    // tiles = 63 constraints = exactly 2 waves), so chunking costs no extra partial wave
    static const int chunks_off = getenv("CLRS_SCHUR_CHUNKS") ? atoi(getenv("CLRS_SCHUR_CHUNKS")) == 1 : 0;
    int pc = np;
    if (!chunks_off && b0.lay == 1 && b0.AallV.lay == 1 && (int64_t)np * n >= 32768) {
      int BNt, grp; pick_tiles(n, BNt, grp); const int ntn = (n + BNt - 1) / BNt;       // column tiles of the product
      pc = std::max(1, ((2 * 148 / ntn) * 128) / n); }                                    // two waves of CTAs per chunk
    const int nchunk = (np + pc - 1) / pc;
    if (nchunk > 1) {
      // the constraints are processed in chunks on different execution contexts: while one chunk's products occupy the
      // tensor pipe, the other's recombination / exponent / split kernels (HBM and load-store bound) run beside them
      ensure(b0.T1S, np * n, n, 1); ensure(b0.T2V, np, b0.tri ? n * (n + 1) / 2 : n * n, 1);
      par_for(nchunk, [&](int c) {
        const int p0 = c * pc, cnt = std::min(pc, np - p0); if (cnt <= 0) return;
        Sliced Ac = view(b0.AallB, p0 * n, cnt * n), T1c = view(b0.T1S, p0 * n, cnt * n), T2c = view(b0.T2V, p0, cnt);
        num* T1p = b0.T1 + (int64_t)p0 * nn; num* T2p = b0.T2 + (int64_t)p0 * nn;
        gemm(Ac, 0, b0.XiS, 0, cnt * n, n, T1p, n);
        { VecView v; v.base = T1p; v.bstride = nn; v.vper = n; v.sv = 1; v.sk = n; v.nvec = cnt * n; v.K = n; split(T1c, v, false, 1, true); }
        { NsScope mp(*this); gemm(T1c, 0, b0.YS, 0, cnt * n, n, T2p, n); }
        if (b0.tri) split_tri(b0.T2V, p0, T2p, cnt, n, true);
        else { VecView v; v.base = T2p; v.bstride = 0; v.vper = cnt; v.sv = nn; v.sk = 1; v.nvec = cnt; v.K = n * n; split(T2c, v, true, 1, true); }
      });
    } else {
    // T1t[(p,j)][i] = sum_k A_p[k][j] X^-1[i][k]   (= (X^-1 A_p)^T; the 90000-row operand sits on the 128-lane M side)
    gemm(b0.AallB, 0, b0.XiS, 0, np * n, n, b0.T1, n);
    // T2[(p,i)][b] = sum_j T1_p[i][j] Y[j][b]
    { VecView v; v.base = b0.T1; v.bstride = nn; v.vper = n; v.sv = 1; v.sk = n; v.nvec = np * n; v.K = n; split(b0.T1S, v, false, b0.lay); }
    { NsScope mp(*this); gemm(b0.T1S, 0, b0.YS, 0, np * n, n, b0.T2, n); }
    if (b0.tri) { ensure(b0.T2V, np, n * (n + 1) / 2, 1); split_tri(b0.T2V, 0, b0.T2, np, n, true); }
    else { VecView v; v.base = b0.T2; v.bstride = 0; v.vper = np; v.sv = nn; v.sk = 1; v.nvec = np; v.K = n * n; split(b0.T2V, v, true, b0.AallV.lay); }
    }
    // S[p,q] += sum_ab T2_p[ab] A_q[ab]
    gemm(b0.AallV, 0, b0.T2V, 0, np, np, b0.Sd, np, 0, nullptr, 0, 1, 0, 0, 0, 0, 1);          // Sd[q][p], q >= p only
  }
  void schur_sparse(Block& b0) {                     // Sd[q][p], q >= p, from the nonzero entries of the A_p (k_schur_sparse)
    const int np = b0.np; if (np == 0) return;
    nlaunch++, k_schur_sparse<NL><<<(unsigned)(((int64_t)np * np * 32 + 255) / 256), 256, 0, st>>>(np, b0.n, b0.nz_start, b0.nz_idx, b0.Aall, Xi + b0.off, Y + b0.off, b0.Sd);
  }
  void schur_add_dense(Clu& c0, Block& b0) {
    const int np = b0.np; if (np == 0) return;
    nlaunch++, k_scatter_upper<NL><<<(unsigned)(((int64_t)np * np + 127) / 128), 128, 0, st>>>(np, b0.d_plist, b0.Sd, c0.S, c0.P);
  }
  void decomposition(int e0) {
    if (stage1_pending) { CK(cudaStreamWaitEvent(st, evS1, 0)); stage1_pending = false; }
    par_blocks([&](Block* b0) { if (b0->sparse) schur_sparse(*b0); else if (staged(*b0)) dense_stage2(*b0); else if (b0->high_rank) pairings_dense(*b0); else pairings_lowrank(*b0); });
    par_clusters([&](Clu& c0) { zero(c0.S, (int64_t)c0.P * c0.P);
      for (auto& b0 : c0.blocks) { if (!b0.mine) continue; if (b0.high_rank) schur_add_dense(c0, b0); else schur_add_lowrank(c0, b0); }
      if (c0.P) nlaunch++, k_mirror<NL><<<grid_for((int64_t)c0.P * c0.P), 256, 0, st>>>(c0.P, c0.S, c0.P, 1); });
    if (nranks > 1) for (auto& c0 : cl) if (c0.split && c0.P) allreduce(c0.S, (int64_t)c0.P * c0.P, 0);   // S_j = sum over the ranks' blocks; every rank factors the same matrix
    rec(ev[e0]);
    par_clusters([&](Clu& c0) { chol(c0.S, c0.P, c0.P, c0.Minv, c0.P, CLRS_ERR_CHOL_S, false); });
    rec(ev[e0 + 1]);
    if (N > 0) {
      // big clusters: the factor goes to every rank (the stream order of the collectives is the cluster order on every rank)
      for (auto& c0 : cl) if (c0.big) { const size_t bytes = (size_t)c0.P * c0.P * sizeof(num);
        if (g_nccl.Broadcast(c0.S, c0.S, bytes, /*ncclChar*/ 0, c0.owner, comm, st) != 0 || g_nccl.Broadcast(c0.Minv, c0.Minv, bytes, 0, c0.owner, comm, st) != 0) throw CudaError("ncclBroadcast failed"); }
      par_clusters([&](Clu& c0) { if (c0.P && !c0.big) trsm_lower(c0.S, c0.P, c0.P, c0.Minv, c0.P, c0.B, N, N, c0.LinvB, N); });     // LinvB = L^-1 B  (:1258)
      for (auto& c0 : cl) if (c0.big) {                                                                  // this rank's columns of L_j^-1 B_j, then all chunks
        num* mine = c0.Gc + (size_t)rank * c0.P * ncr;
        trsm_lower(c0.S, c0.P, c0.P, c0.Minv, c0.P, c0.Bc, ncr, ncr, mine, ncr);
        if (g_nccl.AllGather(mine, c0.Gc, (size_t)c0.P * ncr * sizeof(num), 0, comm, st) != 0) throw CudaError("ncclAllGather failed"); }
      rec(ev[e0 + 2]);
      if (Pown == 0) zero(Q, (int64_t)N * N);
      else { split_cols(tA, LinvBall, N, Pown, N, use_tc(N, N, Pown) ? 1 : 0);
        gemm(tA, 0, tA, 0, N, N, Q, N, 0, nullptr, 0, 1, 0, 0, 0, 0, 1);                                  // Q = (vcat LinvB)^T (vcat LinvB), lower triangle  (:1268-1269)
        nlaunch++, k_mirror<NL><<<grid_for((int64_t)N * N), 256, 0, st>>>(N, Q, N, 0); }
      for (auto& c0 : cl) if (c0.big) {                                                                  // Q[:, my columns] += G_j^T G_j[:, my columns]
        const int lay = use_tc(N, ncr, c0.P) ? 1 : 0;
        VecView va; va.base = c0.Gc; va.bstride = (int64_t)c0.P * ncr; va.vper = ncr; va.sv = 1; va.sk = ncr; va.nvec = nranks * ncr; va.K = c0.P;     // all columns of G_j, chunk by chunk
        split(tA, va, false, lay);
        split_cols(tB, c0.Gc + (size_t)rank * c0.P * ncr, ncr, c0.P, ncr, lay);
        gemm(tA, 0, tB, 0, nranks * ncr, ncr, c0.Qs, ncr);
        const int mine = std::min(ncr, N - rank * ncr);
        if (mine > 0) nlaunch++, k_add_slab<NL><<<grid_for((int64_t)N * mine), 256, 0, st>>>(N, mine, c0.Qs, ncr, Q + (size_t)rank * ncr, N); }
      allreduce(Q, (int64_t)N * N, 0);                                                                  // the only cross-cluster coupling
      rec(ev[e0 + 3]);
      chol(Q, N, N, QMinv, N, CLRS_ERR_CHOL_Q, false);
    } else { rec(ev[e0 + 2]); rec(ev[e0 + 3]); }
    rec(ev[e0 + 4]);
  }
  // search direction  (compute_search_direction!, src/solver.jl:1474-1616)
  void direction(int which) {                                          // which: 0 predictor, 1 corrector (phase timers 13-17)
    rec(evD[which][0]);
    par_blocks([&](Block* b0) { const int n = b0->n; split_rows(tA, P + b0->off, n, n, n, b0->lay); gemm(tA, 0, b0->YS, 0, n, n, T1 + b0->off, n); });     // P Y
    addsub(T1, T1, 1, R, -1, tot);
    par_blocks([&](Block* b0) { const int n = b0->n; split_cols(tB, T1 + b0->off, n, n, n, b0->lay); gemm(b0->XiS, 0, tB, 0, n, n, dY + b0->off, n); });    // Z = X^-1 (P Y - R)
    nlaunch++, k_symmetrize<NL><<<grid_for(tot), 256, 0, st>>>(bt, tot, dY);
    rec(evD[which][1]);
    trace_vectors(tr, dY);
    if (Ptot) nlaunch++, k_vec_rhs<NL><<<(Ptot + 127) / 128, 128, 0, st>>>(Ptot, dx, d, tr);                                                          // rhs_x = -d - <A_*, Z>
    rec(evD[which][2]);
    // block elimination  (:1527-1582)
    if (N > 0) zero(tmpU, N);
    par_clusters([&](Clu& c0) { if (c0.P) { copy(c0.t, dx + c0.off, c0.P); trsv(c0.S, c0.P, c0.P, c0.Minv, c0.P, c0.t, false, c0.ready); } });                 // t_j = L_j^-1 rhs_j
    if (N > 0) for (auto& c0 : cl) { if (!c0.owned || !c0.lead || c0.P == 0) continue;
      if (c0.big) { for (int s = 0; s < nranks; s++) { const int ncs = std::min(ncr, N - s * ncr); if (ncs > 0) nlaunch++, k_gemv_t<NL><<<(ncs + 31) / 32, 256, 0, st>>>(c0.P, ncs, c0.Gc + (size_t)s * c0.P * ncr, ncr, c0.t, tmpU + s * ncr, 1, 1); } continue; }
      nlaunch++, k_gemv_t<NL><<<(N + 31) / 32, 256, 0, st>>>(c0.P, N, c0.LinvB, N, c0.t, tmpU, 1, 1); }                                        // u += LinvB_j^T t_j
    if (N > 0) { allreduce(tmpU, N, 0); addsub(dy, p, 1, tmpU, -1, N);                                                                       // dy = p - sum_j u_j
      trsv(Q, N, N, QMinv, N, dy, false, q_ready); trsv(Q, N, N, QMinv, N, dy, true, q_ready); }                                                              // dy = Q^-1 dy
    par_clusters([&](Clu& c0) { if (c0.P == 0) return;
      if (N > 0 && c0.big) { for (int s = 0; s < nranks; s++) { const int ncs = std::min(ncr, N - s * ncr); if (ncs > 0) nlaunch++, k_gemv_n<NL><<<(c0.P * 32 + 255) / 256, 256, 0, st>>>(c0.P, ncs, c0.Gc + (size_t)s * c0.P * ncr, ncr, dy + s * ncr, c0.t, 1, 1); } }
      else if (N > 0) nlaunch++, k_gemv_n<NL><<<(c0.P * 32 + 255) / 256, 256, 0, st>>>(c0.P, N, c0.LinvB, N, dy, c0.t, 1, 1);                                 // t_j += LinvB_j dy
      trsv(c0.S, c0.P, c0.P, c0.Minv, c0.P, c0.t, true, c0.ready); copy(dx + c0.off, c0.t, c0.P); });                                                 // dx_j = L_j^-T t_j
    rec(evD[which][3]);
    weighted_A(dX, dx); addsub(dX, dX, 1, P, 1, tot);                                                                                       // dX = P + sum dx_p A_p
    rec(evD[which][4]);
    par_blocks([&](Block* b0) { const int n = b0->n; split_rows(tA, dX + b0->off, n, n, n, b0->lay); gemm(tA, 0, b0->YS, 0, n, n, T1 + b0->off, n); });     // dX Y
    addsub(T1, R, 1, T1, -1, tot);
    par_blocks([&](Block* b0) { const int n = b0->n; split_cols(tB, T1 + b0->off, n, n, n, b0->lay); gemm(b0->XiS, 0, tB, 0, n, n, dY + b0->off, n); });    // dY = X^-1 (R - dX Y)
    nlaunch++, k_symmetrize<NL><<<grid_for(tot), 256, 0, st>>>(bt, tot, dY);
    rec(evD[which][5]);
  }
  // lambda_min( L^-1 dM L^-T ) per block in Float64  (compute_step_length, src/solver.jl:1620-1693); Mi holds L^-1
  void step_eigs(const num* Mi, const num* dM, double* lam, bool forY) {
    num* Ub = forY ? U2 : U; num* Tb = forY ? T1b : T1; double* Tdb = forY ? Td2 : Td; EigTask* et = forY ? eigT2 : eigT;    // own buffers: the two run side by side
    par_blocks([&](Block* b0) { const int n = b0->n; if (n == 1) return;
      Sliced& ms = forY ? b0->MSY : b0->MS;                                                             // rows of L^-1 (Y's are split on the side stream)
      if (!forY) split_rows(ms, Mi + b0->off, n, n, n, b0->lay);
      split_cols(tB, dM + b0->off, n, n, n, b0->lay); gemm(ms, 0, tB, 0, n, n, Ub + b0->off, n);         // U = L^-1 dM
      split_rows(tA, Ub + b0->off, n, n, n, b0->lay); gemm(tA, 0, ms, 0, n, n, Tb + b0->off, n); });       // T = U L^-T
    nlaunch++, k_to_double_sym<NL><<<grid_for(tot), 256, 0, st>>>(bt, tot, Tb, Tdb);
    if (!blk.empty()) { int maxn = 1; for (Block* b0 : blk) maxn = std::max(maxn, b0->n);
      nlaunch++, k_min_eig<<<(unsigned)blk.size(), maxn <= 32 ? 128 : (maxn <= 128 ? 256 : EIG_THREADS), 0, st>>>(et, lam, flags + FL_STATUS, CLRS_ERR_EIG); }
  }

  int check_status() { return hflags[FL_STATUS]; }
  bool host_terminate(int& reason) {   // terminate(), src/solver.jl:921-950, on full-precision values
    bool gap_opt = mp_cmp(h_gapn, h_thr[0]) < 0, dual_feas = mp_cmp(h_derrn, h_thr[1]) < 0, primal_feas = mp_cmp(h_perrn, h_thr[2]) < 0;
    if (opt.need_dual_feasible && dual_feas) { reason = CLRS_STOP_DUAL_FEASIBLE; return true; }
    if (opt.need_primal_feasible && primal_feas) { reason = CLRS_STOP_PRIMAL_FEASIBLE; return true; }
    if (!opt.correctoronly && dual_feas && primal_feas && gap_opt) { reason = CLRS_STOP_OPTIMAL; return true; }
    return false;
  }

  // ---- one iteration  (loop body src/solver.jl:362-592) -----------------------------------
  // timing events: recorded as external event nodes while the iteration is being captured into a CUDA graph
  bool capturing = false; cudaEvent_t evD[2][6];
  void rec(cudaEvent_t e) { if (capturing) CK(cudaEventRecordWithFlags(e, st, cudaEventRecordExternal)); else CK(cudaEventRecord(e, st)); }
  // everything the iteration enqueues, from the flag reset to the last reduction: pure stream work with no host
  // decision in between (scalars live on the device), so it can be replayed as a graph
  void enqueue_iteration() {
    CK(cudaMemsetAsync(flags + FL_STOP, 0, 2 * sizeof(int), st));
    rec(ev[0]);
    // side stream: Cholesky of Y and L_Y^-1 for the step length (Y does not change until the step at the end)
    CK(cudaEventRecord(evY0, st)); swap_ctx(); CK(cudaStreamWaitEvent(st, evY0, 0));
    copy(LY, Y, tot);
    par_blocks([&](Block* b0) { const int n = b0->n; if (n > 1) { chol(LY + b0->off, n, n, MinvY + b0->off, n, CLRS_ERR_CHOL_STEP); split_rows(b0->MSY, MinvY + b0->off, n, n, n, b0->lay); } });
    CK(cudaEventRecord(evY1, st)); swap_ctx();
    scalar(0);                                                        // mu, mu_p  (SC_D0 = <X,Y> is kept current)
    // R = mu_p I - X Y
    par_blocks([&](Block* b0) { const int n = b0->n; split_cols(b0->YS, Y + b0->off, n, n, n, b0->lay); });      // Y panels: used by R, the Schur products and the directions
    { bool any = false; for (Block* b0 : blk) any = any || staged(*b0);                                       // first stage of the dense Schur products: needs only Y
      if (any && !prof_on) { CK(cudaEventRecord(evS0, st)); swap_with(side3); CK(cudaStreamWaitEvent(st, evS0, 0));
        for (Block* b0 : blk) if (staged(*b0)) dense_stage1(*b0);
        CK(cudaEventRecord(evS1, st)); swap_with(side3); stage1_pending = true; }
      else if (any) for (Block* b0 : blk) if (staged(*b0)) dense_stage1(*b0); }
    CK(cudaEventRecord(evR0, st)); swap_with(side2); CK(cudaStreamWaitEvent(st, evR0, 0));                      // R is first needed by the predictor: beside chol(X) and the Schur assembly
    par_blocks([&](Block* b0) { const int n = b0->n; split_rows(tA, X + b0->off, n, n, n, b0->lay); gemm(tA, 0, b0->YS, 0, n, n, TXY + b0->off, n); });
    nlaunch++, k_residual_R<NL><<<grid_for(tot), 256, 0, st>>>(bt, tot, R, TXY, (const num*)nullptr, sc + SC_MUP);
    CK(cudaEventRecord(evR1, st)); swap_with(side2);
    rec(ev[1]);
    // Cholesky of X, L^-1, X^-1  (src/solver.jl:388-399, 1117)
    copy(L, X, tot);
    par_blocks([&](Block* b0) { const int n = b0->n; chol(L + b0->off, n, n, Minv + b0->off, n, CLRS_ERR_CHOL_X);
      split_cols(tA, Minv + b0->off, n, n, n, b0->lay); gemm(tA, 0, tA, 0, n, n, Xi + b0->off, n);               // X^-1 = L^-T L^-1
      split_rows(b0->XiS, Xi + b0->off, n, n, n, b0->lay); });
    rec(ev[2]);
    decomposition(3);                                                 // events 3..7
    trace_pairings(tr); residuals();
    CK(cudaStreamWaitEvent(st, evR1, 0));
    rec(ev[8]);
    direction(0);                                                     // predictor
    rec(ev[9]);
    reduce(X, dY, tot, sc + SC_D1, 0); reduce(dX, Y, tot, sc + SC_D2, 0); reduce(dX, dY, tot, sc + SC_D3, 0); allreduce(sc + SC_D1, 3, 0);   // D1..D3 adjacent
    errors(); scalar(1);
    rec(ev[10]);
    par_blocks([&](Block* b0) { const int n = b0->n; split_rows(tA, dX + b0->off, n, n, n, b0->lay); split_cols(tB, dY + b0->off, n, n, n, b0->lay); gemm(tA, 0, tB, 0, n, n, T1 + b0->off, n); });
    nlaunch++, k_residual_R<NL><<<grid_for(tot), 256, 0, st>>>(bt, tot, R, TXY, T1, sc + SC_MUC);                     // R = mu_c I - XY - dXdY
    rec(ev[11]);
    direction(1);                                                     // corrector
    rec(ev[12]);
    // step lengths: X reuses its factor of this iteration (X is unchanged); Y is factored here
    CK(cudaEventRecord(evE0, st)); swap_with(side2); CK(cudaStreamWaitEvent(st, evE0, 0)); CK(cudaStreamWaitEvent(st, evY1, 0));
    step_eigs(MinvY, dY, lamY, true);                                                                           // Y's eigenvalue beside X's
    CK(cudaEventRecord(evE1, st)); swap_with(side2);
    step_eigs(Minv, dX, lamX, false); scalar(2, X, dX, lamX); allreduce(sc + SC_TMP, 1, 2); scalar(6, nullptr, nullptr, nullptr, SC_ALPHAD);
    CK(cudaStreamWaitEvent(st, evY1, 0)); CK(cudaStreamWaitEvent(st, evE1, 0));
    scalar(2, Y, dY, lamY); allreduce(sc + SC_TMP, 1, 2); scalar(6, nullptr, nullptr, nullptr, SC_ALPHAP);
    allreduce_flags();                                                 // a failed factorisation on any rank stops every rank: no step is taken then
    scalar(3);
    rec(ev[13]);
    // the step  (src/solver.jl:485-495); alpha = 0 after a failure or a stop, so the iterate stays the last good one (:594-628)
    if (Ptot) nlaunch++, k_axpy<NL><<<grid_for(Ptot), 256, 0, st>>>(Ptot, x, dx, sc + SC_ALPHAD);
    if (N) nlaunch++, k_axpy<NL><<<grid_for(N), 256, 0, st>>>(N, y, dy, sc + SC_ALPHAP);
    nlaunch++, k_axpy<NL><<<grid_for(tot), 256, 0, st>>>(tot, X, dX, sc + SC_ALPHAD);
    nlaunch++, k_axpy<NL><<<grid_for(tot), 256, 0, st>>>(tot, Y, dY, sc + SC_ALPHAP);
    objectives();
    reduce(X, Y, tot, sc + SC_D0, 0); allreduce(sc + SC_D0, 1, 0);     // <X,Y> of the new iterate for the next mu
    rec(ev[14]);
  }
  // CUDA graph of the iteration: captured on the second iteration of a handle (the first one runs eagerly and sizes every
  // scratch buffer), replayed afterwards; re-captured if a scratch buffer was reallocated in between (alloc_gen)
  cudaGraphExec_t gexec = nullptr; long graph_gen = -1; long graph_nodes_launches = 0; int eager_iters = 0; bool graph_off = false;
  void drop_graph() { if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; } }
  void run_iteration() {
    static const int env_graph = getenv("CLRS_GRAPH") ? atoi(getenv("CLRS_GRAPH")) : 1;
    // multi-rank handles launch eagerly: replaying a graph that contains the NCCL collectives of two processes hung on the
    // first replay (2 x B200, NCCL 2.28.9; profiles/README.md), while the eager sharded path is the one validated in round 1
    static const int env_graph_mr = getenv("CLRS_GRAPH_MULTIRANK") ? atoi(getenv("CLRS_GRAPH_MULTIRANK")) : 0;
    // handles with a staged dense Schur path launch eagerly: the main stream's priority over the side stream of the first product
    // is what lets the panel chain of chol(X) progress beside it, and neither stream nor kernel-node priorities had any effect
    // inside a replayed graph (measured: 22.3 ms per iteration replayed, 20.4 ms eager; profiles/README.md)
    static const int staged_eager = getenv("CLRS_STAGED_EAGER") ? atoi(getenv("CLRS_STAGED_EAGER")) : 1;
    bool any_staged = false; if (staged_eager) for (Block* b0 : blk) any_staged = any_staged || staged(*b0);
    if (!env_graph || graph_off || prof_on || eager_iters < 1 || (nranks > 1 && !env_graph_mr) || any_staged) { enqueue_iteration(); eager_iters++; return; }
    if (gexec && graph_gen != alloc_gen) drop_graph();
    if (!gexec) {
      const long gen0 = alloc_gen, nl0 = nlaunch; cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      capturing = true;
      try { enqueue_iteration(); }
      catch (...) { capturing = false; cudaStreamEndCapture(st, &g); if (g) cudaGraphDestroy(g); cudaGetLastError(); graph_off = true; throw; }
      capturing = false;
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (e != cudaSuccess || !g || alloc_gen != gen0) {           // a buffer grew during the capture: run this iteration eagerly, capture the next one
        if (g) cudaGraphDestroy(g); cudaGetLastError(); nlaunch = nl0;
        if (e != cudaSuccess) graph_off = true;
        enqueue_iteration(); return; }
      graph_nodes_launches = nlaunch - nl0; nlaunch = nl0;
      e = cudaGraphInstantiate(&gexec, g, 0); cudaGraphDestroy(g);
      if (e != cudaSuccess) { gexec = nullptr; cudaGetLastError(); graph_off = true; enqueue_iteration(); return; }
      graph_gen = alloc_gen;
    }
    CK(cudaGraphLaunch(gexec, st)); nlaunch += graph_nodes_launches;
  }
  int iterate(clrs_iter_info* info) override {
    if (!finalized) { err = "clrs_finalize has not been called"; return CLRS_ERR_ARG; }
    memset(info, 0, sizeof(*info)); info->iter = iter; info->d_obj = h_dobj; info->p_obj = h_pobj; info->gap = h_gap; info->pd_feasible = h_pdfeas;
    int reason = 0; if (host_terminate(reason)) { info->stop = reason; last_ms = 0; return 0; }
    run_iteration();
    pull_info();
    if (int s = check_status()) {
      const char* msg = s == CLRS_ERR_CHOL_X ? "The cholesky decomposition of X was not computed correctly. Try again with higher precision"
                      : s == CLRS_ERR_CHOL_S ? "S was not decomposed succesfully, try again with higher precision."
                      : s == CLRS_ERR_CHOL_Q ? "Q was not decomposed correctly. Try restarting with a higher precision."
                      : s == CLRS_ERR_EIG ? "The eigenvalues could not be computed during the computation of the step length."
                      : "The cholesky decomposition could not be computed during the computation of the step length.";
      CK(cudaMemsetAsync(flags + FL_STATUS, 0, sizeof(int), st));
      err = msg; last_ms = 0; return s;
    }
    info->stop = hflags[FL_STOP]; info->pd_feasible = hflags[FL_PDFEAS];
    info->mu = hinfo[INFO_MU]; info->err_P = hinfo[INFO_ERRP]; info->err_p = hinfo[INFO_ERRp]; info->err_d = hinfo[INFO_ERRd];
    info->alpha_d = hinfo[INFO_ALPHAD]; info->alpha_p = hinfo[INFO_ALPHAP]; info->beta_c = hinfo[INFO_BETAC];
    info->d_obj_new = hinfo[INFO_DOBJ + 10]; info->p_obj_new = hinfo[INFO_POBJ + 10]; info->gap_new = hinfo[INFO_GAP + 10];
    auto ms = [&](int a, int b_) { float t = 0; if (cudaEventElapsedTime(&t, ev[a], ev[b_]) != cudaSuccess) { cudaGetLastError(); t = 0; } return (double)t; };
    auto msd = [&](int k) { float t0 = 0, t1 = 0; if (cudaEventElapsedTime(&t0, evD[0][k], evD[0][k + 1]) != cudaSuccess || cudaEventElapsedTime(&t1, evD[1][k], evD[1][k + 1]) != cudaSuccess) { cudaGetLastError(); return 0.0; } return (double)t0 + (double)t1; };
    last_ms = ms(0, 14);
    info->phase_ms[0] = ms(2, 7); info->phase_ms[1] = ms(8, 9); info->phase_ms[2] = ms(11, 12); info->phase_ms[3] = ms(12, 13); info->phase_ms[4] = ms(1, 2);
    info->phase_ms[5] = ms(0, 1) + ms(10, 11); info->phase_ms[6] = ms(7, 8);
    info->phase_ms[7] = ms(2, 3); info->phase_ms[8] = ms(3, 4); info->phase_ms[9] = ms(4, 5); info->phase_ms[10] = ms(5, 6); info->phase_ms[11] = ms(6, 7);
    for (int k = 0; k < 5; k++) info->phase_ms[12 + k] = msd(k);        // Z, rhs_x, solve, dX, dY of predictor + corrector (src/solver.jl:540)
    h_pdfeas = hflags[FL_PDFEAS];
    h_derr = std::max(info->err_P, info->err_p); h_perr = info->err_d;
    if (info->stop == 0) { h_dobj = info->d_obj_new; h_pobj = info->p_obj_new; h_gap = info->gap_new; iter++; }
    fetch_thresholds();
    return 0;
  }
  int get_objectives(void* d_, void* p_, void* g_) override {
    objectives(); num h[SC_COUNT]; CK(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    m2w(d_, h[SC_DOBJ]); m2w(p_, h[SC_POBJ]); m2w(g_, h[SC_GAP]); return 0;
  }
  int64_t matrix_count() const override { return gtot; }
  // x: all constraints; X, Y: all blocks of the SDP in (j,l) order (global layout).  A sharded handle reads/writes the
  // parts it owns; get_state leaves the others untouched (the host merges the ranks' outputs).
  // The four arrays cross PCIe as raw wire bytes through ONE device staging buffer: every contiguous run of owned
  // numbers is one copy + one conversion kernel, and the call synchronises once (pinned caller buffers make the copies true DMA).
  struct Run { num* dev; size_t first; size_t n; int arr; };            // device destination, first record in caller array `arr`, count
  void state_runs(std::vector<Run>& runs, bool hx, bool hX, bool hy, bool hY) {
    auto push = [&](num* dev, size_t first, size_t n, int arr) { if (!n) return;
      if (!runs.empty() && runs.back().arr == arr && runs.back().first + runs.back().n == first && runs.back().dev + runs.back().n == dev) runs.back().n += n; else runs.push_back({dev, first, n, arr}); };
    if (hx) for (auto& c0 : cl) if (c0.owned) push(x + c0.off, (size_t)c0.off, (size_t)c0.P, 0);
    if (hy) push(y, 0, (size_t)N, 2);
    if (hX) for (Block* b0 : blk) push(X + b0->off, (size_t)b0->goff, (size_t)b0->n * b0->n, 1);
    if (hY) for (Block* b0 : blk) push(Y + b0->off, (size_t)b0->goff, (size_t)b0->n * b0->n, 3);
  }
  int set_state(const void* x_, const void* X_, const void* y_, const void* Y_) override {
    std::vector<Run> runs; state_runs(runs, x_ != nullptr, X_ != nullptr, y_ != nullptr, Y_ != nullptr);
    size_t total = 0; for (auto& r : runs) total += r.n;
    const void* src[4] = {x_, X_, y_, Y_}; const size_t ws = wire_size();
    if (total) { unsigned char* sbuf = stage(total * ws); size_t o = 0;
      for (auto& r : runs) { CK(cudaMemcpyAsync(sbuf + o * ws, (const char*)src[r.arr] + r.first * ws, r.n * ws, cudaMemcpyHostToDevice, st));
        nlaunch++, k_wire_to_mpn<NL><<<grid_for((int64_t)r.n), 256, 0, st>>>((int64_t)r.n, sbuf + o * ws, W(), r.dev); o += r.n; } }
    initial_quantities(); return 0;                     // (synchronises once, reading the report)
  }
  int get_state(void* x_, void* X_, void* y_, void* Y_) override {
    std::vector<Run> runs; state_runs(runs, x_ != nullptr, X_ != nullptr, y_ != nullptr && N > 0, Y_ != nullptr);
    size_t total = 0; for (auto& r : runs) total += r.n;
    void* dst[4] = {x_, X_, y_, Y_}; const size_t ws = wire_size();
    if (!total) return 0;
    unsigned char* sbuf = stage(total * ws); size_t o = 0;
    for (auto& r : runs) { nlaunch++, k_mpn_to_wire<NL><<<grid_for((int64_t)r.n), 256, 0, st>>>((int64_t)r.n, r.dev, W(), sbuf + o * ws);
      CK(cudaMemcpyAsync((char*)dst[r.arr] + r.first * ws, sbuf + o * ws, r.n * ws, cudaMemcpyDeviceToHost, st)); o += r.n; }
    CK(cudaStreamSynchronize(st)); return 0;
  }
  // ---- standalone kernels -----------------------------------------------------------------
  // scratch of one call: freed when the call returns (the handle's `allocs` list only holds the SDP's own buffers)
  struct Scratch { std::vector<void*> p; ~Scratch() { for (void* q : p) cudaFree(q); } };
  template <class T> T* talloc(Scratch& sc_, size_t n) { void* p = nullptr; CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); sc_.p.push_back(p); CK(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), st)); return (T*)p; }
  int mp_gemm(int M, int N_, int K, const void* A, const void* B, void* C, int path, double* ms) override {
    Scratch tmp; num* dA = talloc<num>(tmp, (size_t)M * K); num* dB = talloc<num>(tmp, (size_t)K * N_); num* dC = talloc<num>(tmp, (size_t)M * N_);
    wire_to_device(dA, A, (size_t)M * K); wire_to_device(dB, B, (size_t)K * N_);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int saved = opt.gemm_path; if (path) opt.gemm_path = path;
    CK(cudaEventRecord(e0, st)); mm(dA, K, dB, N_, M, N_, K, dC, N_); CK(cudaEventRecord(e1, st)); opt.gemm_path = saved; CK(cudaStreamSynchronize(st));
    float t = 0; cudaEventElapsedTime(&t, e0, e1); if (ms) *ms = t; cudaEventDestroy(e0); cudaEventDestroy(e1);
    CK(cudaGetLastError());
    download_wire(C, dC, (size_t)M * N_); return 0;
  }
  // column-pivoted QR of an m x n matrix (modified Gram-Schmidt, pivot = largest remaining column norm): R is min(m,n) x n in the pivoted
  // column order, perm[k] = original index of pivoted column k  (the core of preprocess!, src/pre_postprocessing.jl:36,56,104)
  int mp_qr_pivot(int m, int n, const void* Aw, void* Rw, int32_t* permw) override {
    if (m <= 0 || n <= 0) { err = "clrs_mp_qr_pivot: empty matrix"; return CLRS_ERR_ARG; }
    Scratch tmp; const int kmax = std::min(m, n);
    num* dA = talloc<num>(tmp, (size_t)m * n); num* dR = talloc<num>(tmp, (size_t)kmax * n); num* dq = talloc<num>(tmp, (size_t)m); num* dn = talloc<num>(tmp, (size_t)n);
    int32_t* dperm = talloc<int32_t>(tmp, (size_t)n); int* dpiv = talloc<int>(tmp, 1);
    wire_to_device(dA, Aw, (size_t)m * n);
    { std::vector<int32_t> hp(n); for (int i = 0; i < n; i++) hp[i] = i; CK(cudaMemcpyAsync(dperm, hp.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); }
    nlaunch++, k_qr_colnorm2<NL><<<(n + 7) / 8, 256, 0, st>>>(m, n, dA, n, dn);
    for (int k0 = 0; k0 < kmax; k0++) {
      nlaunch++, k_qr_pivot<NL><<<1, 32, 0, st>>>(n, k0, dn, dpiv);
      nlaunch++, k_qr_swap<NL><<<grid_for((int64_t)m + k0 + 1), 256, 0, st>>>(m, k0, dpiv, dA, n, dR, n, dn, dperm);
      nlaunch++, k_qr_scale<NL><<<grid_for((int64_t)m, 256, 64), 256, 0, st>>>(m, k0, dA, n, dn, dR, n, dq);
      if (k0 + 1 < n) nlaunch++, k_qr_project<NL><<<(n - k0 - 1 + 7) / 8, 256, 0, st>>>(m, n, k0, dA, n, dq, dR, n, dn);
    }
    CK(cudaStreamSynchronize(st)); CK(cudaGetLastError());
    download_wire(Rw, dR, (size_t)kmax * n);
    CK(cudaMemcpyAsync(permw, dperm, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    return 0;
  }
  int mp_cholesky(int n, const void* A, void* Lw) override {
    Scratch tmp; num* dA = talloc<num>(tmp, (size_t)n * n); num* dM = talloc<num>(tmp, (size_t)n * n);
    wire_to_device(dA, A, (size_t)n * n);
    CK(cudaMemsetAsync(flags, 0, FL_COUNT * sizeof(int), st));
    const char* tl = getenv("CLRS_POTRF_TIMELINE");
    if (tl) potrf_dbg = talloc<long long>(tmp, 300);
    static const int noinv = getenv("CLRS_MP_CHOL_NOINV") ? atoi(getenv("CLRS_MP_CHOL_NOINV")) : 0;     // the variant without the inverse (S and Q blocks), for tools/gpu_chol_check.py
    chol(dA, n, n, dM, n, CLRS_ERR_CHOL_X, !noinv && (!tl || atoi(tl) != 2));       // CLRS_POTRF_TIMELINE=2: also without the inverse
    if (tl) { long long hh[300]; CK(cudaMemcpyAsync(hh, potrf_dbg, sizeof(hh), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); potrf_dbg = nullptr;
      fprintf(stderr, "potrf timeline (cycles): col  pivot  update  diag  | column total\n");
      for (int c = 0; c < 32; c++) { long long* q = hh + 1 + c * 8; fprintf(stderr, "%2d: %7lld %7lld %7lld | %7lld\n", c, q[1] - q[0], q[3] - q[2], q[5] - q[4], q[6] - (c ? q[-2] : hh[0])); }
      fprintf(stderr, "phase A %lld  phase B %lld  store %lld\n", hh[1 + 256] - hh[0], hh[2 + 256] - hh[1 + 256], hh[3 + 256] - hh[2 + 256]); }
    int f[FL_COUNT]; CK(cudaMemcpyAsync(f, flags, sizeof(f), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); CK(cudaGetLastError());
    download_wire(Lw, dA, (size_t)n * n);
    if (f[FL_STATUS]) { err = "non-positive pivot"; return f[FL_STATUS]; }
    return 0;
  }
  double last_ms = 0;
  void profile(int enable) override { prof_on = enable != 0; for (int i = 0; i < 3; i++) { prof_ms[i] = 0; prof_flops[i] = 0; prof_n[i] = 0; } nlaunch = 0; }
  void profile_get(double* o) override { for (int i = 0; i < 3; i++) { o[3 * i] = prof_ms[i]; o[3 * i + 1] = prof_flops[i]; o[3 * i + 2] = (double)prof_n[i]; } o[9] = (double)nlaunch; }
  double last_iteration_ms() override { return last_ms; }
  void use_graph(int enable) override { graph_off = !enable; if (!enable) drop_graph(); }
  // device self-test: warp-cooperative arithmetic (mpw.cuh) against the single-thread routines on random operands; returns mismatches
  int selftest() override {
    if constexpr (NL != 8 && NL != 16) return 0;                       // 10 limbs: the pivot chain stays on one thread
    constexpr int WL = (NL == 16) ? 16 : 8;
    const int n = 4096; mpn<WL>* a = dalloc<mpn<WL>>(n); mpn<WL>* b2 = dalloc<mpn<WL>>(n); int* mm = dalloc<int>(1);
    nlaunch++, k_fill_random<WL><<<16, 256, 0, st>>>(n, a, 77, 40); nlaunch++, k_fill_random<WL><<<16, 256, 0, st>>>(n, b2, 5, 40);
    nlaunch++, k_selftest_mpw<WL><<<n * 32 / 256, 256, 0, st>>>(n, a, b2, mm);
    int h = -1; CK(cudaMemcpyAsync(&h, mm, sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); CK(cudaGetLastError());
    if (getenv("CLRS_WOPS_BENCH")) { long long* o = dalloc<long long>(8); mpn<WL>* sink = dalloc<mpn<WL>>(1);
      for (int rep = 0; rep < 2; rep++) k_bench_wops<WL><<<1, 32, 0, st>>>(a, b2, o, sink);
      long long ho[8]; CK(cudaMemcpyAsync(ho, o, sizeof(ho), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
      fprintf(stderr, "cycles per op (%d limbs, one warp alone): w_mul %lld  w_addsub %lld  w_rsqrt_c %lld  w_mul_c %lld | one thread: mp_mul %lld  mp_add %lld  mp_rsqrt %lld\n", WL, ho[0], ho[1], ho[2], ho[3], ho[4], ho[5], ho[6]); }
    return h;
  }
  // kernel-only timing of C = A*B on device-generated operands: out = {split ms, gemm ms per rep (kernel + recombine), kernel-only ms per rep}
  int bench_gemm(int M, int N_, int K, int reps, int path, double* out) override {
    Scratch tmp; num* dA = talloc<num>(tmp, (size_t)M * K); num* dB = talloc<num>(tmp, (size_t)K * N_); num* dC = talloc<num>(tmp, (size_t)M * N_);
    nlaunch++, k_fill_random<NL><<<grid_for((int64_t)M * K), 256, 0, st>>>((int64_t)M * K, dA, 1234, 4);
    nlaunch++, k_fill_random<NL><<<grid_for((int64_t)K * N_), 256, 0, st>>>((int64_t)K * N_, dB, 99, 4);
    const int lay = path == 2 ? 1 : (path == 1 ? 0 : (use_tc(M, N_, K) ? 1 : 0));
    Sliced sa, sb; cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    CK(cudaEventRecord(e0, st)); split_rows(sa, dA, K, M, K, lay); split_cols(sb, dB, N_, K, N_, lay); CK(cudaEventRecord(e1, st));
    gemm(sa, 0, sb, 0, M, N_, dC, N_);                      // warm-up
    CK(cudaEventRecord(e1, st));
    for (int r = 0; r < reps; r++) gemm(sa, 0, sb, 0, M, N_, dC, N_);
    CK(cudaEventRecord(e2, st)); CK(cudaStreamSynchronize(st)); CK(cudaGetLastError());
    float t01 = 0, t12 = 0; cudaEventElapsedTime(&t01, e0, e1); cudaEventElapsedTime(&t12, e1, e2);
    out[0] = t01; out[1] = t12 / reps; out[2] = 0;
    if (lay == 1) {      // the product kernel alone, one K range (no recombination)
      int BN, group; pick_tiles(N_, BN, group);
      CUtensorMap mA = make_map(sa, tc::BM), mB = make_map(sb, BN);
      tc::Args a; a.M = M; a.N = N_; a.Kp = std::min(((1 << 17) / NS / 128) * 128, sa.Kp); a.k0 = 0; a.BN = BN; a.group = group; a.dsplit = 0; a.oraw = nullptr; a.a_bvec = 0; a.b_bvec = 0; a.NS = NS; a.Npitch = (N_ + 15) & ~15; a.batch = 1; a.obytes = tc_bytes; a.otop = tc_top; a.lower_only = 0; a.kz_stride = 0; a.Kp_total = a.Kp; a.dbg = nullptr; a.epi = tc_epi();
      if ((size_t)M * a.Npitch > tc_cap) throw CudaError("bench_gemm: byte planes smaller than one product");
      dim3 grid((N_ + BN - 1) / BN, (M + tc::BM - 1) / tc::BM, 1);
      CK(cudaEventRecord(e1, st));
      for (int r = 0; r < reps; r++) nlaunch++, tc::k_gemm_tc<<<grid, tc::NTHREADS, tc::SMEM_BYTES, st>>>(mA, mB, a);
      CK(cudaEventRecord(e2, st)); CK(cudaStreamSynchronize(st)); CK(cudaGetLastError());
      cudaEventElapsedTime(&t12, e1, e2); out[2] = t12 / reps;
      if (getenv("CLRS_TC_TIMELINE")) { long long* dd = talloc<long long>(tmp, 256); a.dbg = dd; tc::k_gemm_tc<<<grid, tc::NTHREADS, tc::SMEM_BYTES, st>>>(mA, mB, a); long long hh[256]; CK(cudaMemcpyAsync(hh, dd, sizeof(hh), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        fprintf(stderr, "group: wait_epi  mma_issue  | epilogue: wait_mma  work   (cycles, CTA 0)\n"); for (int gg = 0; gg < (NS + group - 1) / group; gg++) { long long* q = hh + gg * 8; fprintf(stderr, "%2d: %8lld %8lld | %8lld %8lld   t0=%lld\n", gg, q[1] - q[0], q[2] - q[1], q[4] - q[3], q[5] - q[4], q[0] - hh[0]); } }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    CK(cudaStreamSynchronize(st));
    if (sa.sl) cudaFree(sa.sl); if (sa.E) cudaFree(sa.E); if (sa.planes) cudaFree(sa.planes); if (sb.sl) cudaFree(sb.sl); if (sb.E) cudaFree(sb.E); if (sb.planes) cudaFree(sb.planes);
    return 0;
  }
  int64_t debug_get(const char* what, int j, int l, void* out, int64_t cap) override {
    std::string w(what); const num* src = nullptr; int64_t n = 0;
    if (w == "sparse?") return (j >= 0 && j < (int)cl.size() && l >= 0 && l < (int)cl[j].blocks.size() && cl[j].blocks[l].sparse) ? 1 : 0;   // which Schur path the block takes (tests)
    if (j >= 0 && j < (int)cl.size() && !cl[j].owned && (w == "S" || w == "LinvB" || w.size() > 2 || w == "X" || w == "Y" || w == "R" || w == "P" || w == "L")) return -1;
    if (w == "S") { src = cl[j].S; n = (int64_t)cl[j].P * cl[j].P; } else if (w == "LinvB") { if (cl[j].big) return -1; src = cl[j].LinvB; n = (int64_t)cl[j].P * N; }
    else if (w == "Q") { src = Q; n = (int64_t)N * N; } else if (w == "d") { src = d; n = Ptot; } else if (w == "p") { src = p; n = N; }
    else if (w == "dx") { src = dx; n = Ptot; } else if (w == "dy") { src = dy; n = N; } else if (w == "x") { src = x; n = Ptot; } else if (w == "y") { src = y; n = N; }
    else { Block& b0 = cl[j].blocks[l]; if (!b0.mine) return -1; n = (int64_t)b0.n * b0.n; const num* base = nullptr;
      if (w == "Xinv") base = Xi; else if (w == "R") base = R; else if (w == "P") base = P; else if (w == "dX") base = dX; else if (w == "dY") base = dY; else if (w == "X") base = X; else if (w == "Y") base = Y; else if (w == "L") base = L;
      if (!base) return -1; src = base + b0.off; }
    if (n > cap) return -n; if (n) download_wire(out, src, n); return n;
  }
};

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
struct clrs_handle { SolverBase* s; std::string err; };
#define GUARD(h, body) try { body } catch (const std::exception& e) { (h)->err = e.what(); if ((h)->s) (h)->s->err = e.what(); return CLRS_ERR_CUDA; }

extern "C" {
void clrs_default_options(clrs_options* o) {
  memset(o, 0, sizeof(*o)); o->prec = 256; o->beta_infeasible = 0.3; o->beta_feasible = 0.1; o->gamma = 0.9; o->omega_p = 1e10; o->omega_d = 1e10;
  o->duality_gap_threshold = 1e-15; o->dual_error_threshold = 1e-30; o->primal_error_threshold = 1e-30; o->max_complementary_gap = 1e100; o->step_length_threshold = 1e-7; o->safe_step = 1;
}
int clrs_create(const clrs_options* opt, clrs_handle** out) {
  clrs_handle* h = new clrs_handle(); h->s = nullptr; *out = h;
  try {
    if (opt->prec <= 0 || opt->prec > 512) { h->err = "this build supports prec <= 512 bits (8, 10 or 16 limbs of 32 bits)"; return CLRS_ERR_UNSUPPORTED; }
    if (opt->prec <= 256) h->s = new Solver<8>(*opt); else if (opt->prec <= 320) h->s = new Solver<10>(*opt); else h->s = new Solver<16>(*opt);
  } catch (const std::exception& e) { h->err = e.what(); return CLRS_ERR_CUDA; }
  return CLRS_OK;
}
void clrs_destroy(clrs_handle* h) { if (!h) return; delete h->s; delete h; }
const char* clrs_last_error(const clrs_handle* h) { if (!h) return ""; if (h->s && !h->s->err.empty()) return h->s->err.c_str(); return h->err.c_str(); }
size_t clrs_wire_size(const clrs_handle* h) { return h->s->wire_size(); }
int clrs_set_option_num(clrs_handle* h, int which, const void* w) { GUARD(h, return h->s->set_option_num(which, w);) }
int clrs_set_free(clrs_handle* h, int32_t N, const void* b, const void* c, int32_t mx) { GUARD(h, return h->s->set_free(N, b, c, mx);) }
int clrs_add_cluster(clrs_handle* h, int32_t j, int32_t P, const void* B, const void* c) { GUARD(h, return h->s->add_cluster(j, P, B, c);) }
int clrs_add_block(clrs_handle* h, int32_t j, int32_t l, int32_t m, int32_t delta, int32_t hr, const void* C) { GUARD(h, return h->s->add_block(j, l, m, delta, hr, C);) }
int clrs_add_dense_term(clrs_handle* h, int32_t j, int32_t l, int32_t p, const void* A) { GUARD(h, return h->s->add_dense_term(j, l, p, A);) }
int clrs_add_sparse_term(clrs_handle* h, int32_t j, int32_t l, int32_t p, int32_t nnz, const int32_t* rows, const int32_t* cols, const void* vals, int32_t mirror) { GUARD(h, return h->s->add_sparse_term(j, l, p, nnz, rows, cols, vals, mirror);) }
int clrs_add_lowrank_term(clrs_handle* h, int32_t j, int32_t l, int32_t r, int32_t s, int32_t p, int32_t rank, const void* lam, const void* vs, const void* ws) { GUARD(h, return h->s->add_lowrank_term(j, l, r, s, p, rank, lam, vs, ws);) }
int clrs_finalize(clrs_handle* h) { GUARD(h, return h->s->finalize();) }
int clrs_set_state(clrs_handle* h, const void* x, const void* X, const void* y, const void* Y) { GUARD(h, return h->s->set_state(x, X, y, Y);) }
int clrs_get_state(clrs_handle* h, void* x, void* X, void* y, void* Y) { GUARD(h, return h->s->get_state(x, X, y, Y);) }
int64_t clrs_state_matrix_count(const clrs_handle* h) { return h->s->matrix_count(); }
int clrs_iterate(clrs_handle* h, clrs_iter_info* info) { GUARD(h, return h->s->iterate(info);) }
int clrs_get_objectives(clrs_handle* h, void* d, void* p, void* g) { GUARD(h, return h->s->get_objectives(d, p, g);) }
int clrs_comm_init(clrs_handle* h, int32_t rank, int32_t nranks, const void* uid) { GUARD(h, return h->s->comm_init(rank, nranks, uid);) }
int clrs_comm_unique_id(void* out128) { std::string e; memset(out128, 0, 128); if (!g_nccl.load(e)) return CLRS_ERR_CUDA; return g_nccl.GetUniqueId(out128) == 0 ? CLRS_OK : CLRS_ERR_CUDA; }
int clrs_cluster_owner(clrs_handle* h, int32_t j) { return h->s->owner_of(j); }
int clrs_block_owner(clrs_handle* h, int32_t j, int32_t l) { return h->s->block_owner_of(j, l); }
int clrs_debug_selftest(clrs_handle* h) { try { return h->s->selftest(); } catch (const std::exception& e) { h->err = e.what(); return -1; } }
int clrs_partition_clusters(int32_t J, const double* weight, int32_t nranks, int32_t* owner) {
  if (J < 0 || nranks < 1) return CLRS_ERR_ARG; std::vector<double> w(weight, weight + J); std::vector<int> o; partition_clusters(w, nranks, o); for (int j = 0; j < J; j++) owner[j] = o[j]; return CLRS_OK;
}
int clrs_plan_shards(int32_t J, const double* p3, const int32_t* nblocks, const double* block_weight, const int32_t* column_split, int32_t nranks, int32_t split_mode,
                     int32_t* cluster_owner, int32_t* split, int32_t* block_owner) {
  if (J < 0 || nranks < 1) return CLRS_ERR_ARG;
  std::vector<double> p(p3, p3 + J); std::vector<std::vector<double>> bw(J); std::vector<int> big(J, 0); size_t o = 0;
  for (int j = 0; j < J; j++) { bw[j].assign(block_weight + o, block_weight + o + nblocks[j]); o += nblocks[j]; if (column_split) big[j] = column_split[j]; }
  const ShardPlan pl = plan_shards(p, bw, big, nranks, split_mode); o = 0;
  for (int j = 0; j < J; j++) { cluster_owner[j] = pl.cluster_owner[j]; split[j] = pl.split[j]; for (int l = 0; l < nblocks[j]; l++) block_owner[o++] = pl.block_owner[j][l]; }
  return CLRS_OK;
}
int clrs_mp_gemm(clrs_handle* h, int32_t M, int32_t N, int32_t K, const void* A, const void* B, void* C, int32_t path, double* ms) { GUARD(h, return h->s->mp_gemm(M, N, K, A, B, C, path, ms);) }
int clrs_mp_cholesky(clrs_handle* h, int32_t n, const void* A, void* L) { GUARD(h, return h->s->mp_cholesky(n, A, L);) }
int clrs_mp_qr_pivot(clrs_handle* h, int32_t m, int32_t n, const void* A, void* R, int32_t* perm) { GUARD(h, return h->s->mp_qr_pivot(m, n, A, R, perm);) }
void clrs_profile(clrs_handle* h, int32_t enable) { h->s->profile(enable); }
void clrs_profile_get(clrs_handle* h, double* out10) { h->s->profile_get(out10); }
double clrs_last_iteration_ms(clrs_handle* h) { return h->s->last_iteration_ms(); }
void clrs_use_graph(clrs_handle* h, int32_t enable) { h->s->use_graph(enable); }
int clrs_bench_gemm(clrs_handle* h, int32_t M, int32_t N, int32_t K, int32_t reps, int32_t path, double* out3) { GUARD(h, return h->s->bench_gemm(M, N, K, reps, path, out3);) }
int64_t clrs_debug_get(clrs_handle* h, const char* what, int32_t j, int32_t l, void* out, int64_t cap) { try { return h->s->debug_get(what, j, l, out, cap); } catch (const std::exception& e) { h->err = e.what(); return -1; } }
}
