/*
 * clrs_b200.h — C ABI of libclrs_b200.so
 *
 * B200-native replacement for the per-iteration linear algebra of the
 * primal-dual interior-point method in nanleij/ClusteredLowRankSolver.jl.
 * The reference has no FFI seam for this path (it calls Arb op by op from
 * Julia); this boundary is cut at the Julia function seams of
 * src/solver.jl, as laid out in SURVEY.md §8(b):
 *
 *   clrs_create            replaces  src/solver.jl:137-139  (option conversion)
 *   clrs_set_free/add_*    replace   src/solver.jl:169-184, 249-283 and the
 *                                    precompute at :985-1059 (dedup/pointers)
 *   clrs_finalize          replaces  src/solver.jl:187-201, 298-333 (init
 *                                    X=omega_p I, Y=omega_d I, preallocation,
 *                                    initial objectives/residuals/errors)
 *   clrs_set_state         replaces  src/solver.jl:202-239 (warm start)
 *   clrs_iterate           replaces  the loop body src/solver.jl:362-592
 *   clrs_get_state         replaces  src/solver.jl:523-526, 626-634
 *   clrs_last_error        carries the SolverFailure texts of
 *                                    src/solver.jl:396,1249,1277,1646,1672
 *
 * Plain C: pointers and sizes only, no CUDA / torch / Julia types.
 * There is NO CPU fallback: clrs_create fails with CLRS_ERR_CUDA when no
 * sm_100 device is present.
 *
 * Wire number format ("wire numbers").  Every multi-precision value crosses
 * the boundary as a fixed-size little-endian record of 16 + 8*W bytes with
 * W = ceil(prec/64):
 *
 *     int64  exp;        binary exponent
 *     int32  sign;       -1, 0 (value is zero, other fields ignored), +1
 *     int32  reserved;   must be 0
 *     uint64 limb[W];    mantissa, limb[W-1] most significant, top bit set
 *
 * value = sign * (limb as a W*64-bit integer) / 2^(64 W) * 2^exp, i.e. the
 * (sign, exp, d) triple of an MPFR / Julia BigFloat number at `prec` bits
 * (the reference converts every Arb midpoint to BigFloat at
 * src/solver.jl:747-750).  Matrices are row-major arrays of such records.
 *
 * Ownership: the handle owns all device and pinned memory.  Every pointer
 * argument is borrowed for the duration of the call only.
 * Threading: a handle may be used from any single OS thread at a time; the
 * library never calls back into the host language.
 */
#ifndef CLRS_B200_H
#define CLRS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clrs_handle clrs_handle;

/* Return codes.  10..14 map to the reference's SolverFailure sites; the host
 * shim turns them into `throw(SolverFailure(clrs_last_error(h)))`. */
enum {
  CLRS_OK = 0,
  CLRS_ERR_ARG = 1,          /* bad argument / inconsistent problem description */
  CLRS_ERR_CHOL_X = 10,      /* src/solver.jl:394-398  */
  CLRS_ERR_CHOL_S = 11,      /* src/solver.jl:1248-1250 */
  CLRS_ERR_CHOL_Q = 12,      /* src/solver.jl:1276-1278 */
  CLRS_ERR_CHOL_STEP = 13,   /* src/solver.jl:1645-1646 */
  CLRS_ERR_EIG = 14,         /* src/solver.jl:1671-1673 */
  CLRS_ERR_CUDA = 20,        /* CUDA / NCCL failure, or no sm_100 device */
  CLRS_ERR_UNSUPPORTED = 21
};

/* Loop exit reasons reported in clrs_iter_info.stop (mirrors error_code of
 * src/solver.jl:362-366, 376-380, 470-475 and `terminate`, :921-950). */
enum {
  CLRS_CONTINUE = 0,
  CLRS_STOP_OPTIMAL = 1,          /* dual & primal feasible and gap < threshold */
  CLRS_STOP_DUAL_FEASIBLE = 2,    /* need_dual_feasible satisfied   */
  CLRS_STOP_PRIMAL_FEASIBLE = 3,  /* need_primal_feasible satisfied */
  CLRS_STOP_MAX_COMPLEMENTARY_GAP = 4,  /* error_code 3 */
  CLRS_STOP_STEP_TOO_SHORT = 5          /* error_code 4 */
};

/* Options: 1:1 with the keyword arguments of solvesdp (src/solver.jl:100-127).
 * maxiterations, verbose, save_settings, preprocess and testing stay on the
 * host side of the boundary (the loop returns to the caller every iteration).
 * Numeric options are doubles here and converted exactly; any of them can be
 * overridden by a full-precision wire number with clrs_set_option_num (the
 * reference converts rationals such as gamma = 9//10 at full precision,
 * src/solver.jl:138-139). */
typedef struct clrs_options {
  int32_t prec;                      /* bits; 256 default (precision(BigFloat)); <= 512 in this build (8, 10 or 16 limbs) */
  int32_t matmul_prec;               /* bits for the products of the bilinear pairings and the T Y product of the dense path (solver.jl:93,125,1097,1125-1143); 0 = prec.
                                        Lower values produce fewer slice-pair diagonals on the tensor cores (ceil((matmul_prec + 24) / 8) of the 35 at 256 bit) */
  double  beta_infeasible;           /* 3//10 */
  double  beta_feasible;             /* 1//10 */
  double  gamma;                     /* 9//10 */
  double  omega_p;                   /* 1e10  */
  double  omega_d;                   /* 1e10  */
  double  duality_gap_threshold;     /* 1e-15 */
  double  dual_error_threshold;      /* 1e-30 */
  double  primal_error_threshold;    /* 1e-30 */
  double  max_complementary_gap;     /* 1e100 */
  double  step_length_threshold;     /* 1e-7  */
  int32_t need_dual_feasible;        /* false */
  int32_t need_primal_feasible;      /* false */
  int32_t safe_step;                 /* true  */
  int32_t correctoronly;             /* false */
  int32_t device;                    /* CUDA device ordinal for this handle */
  int32_t gemm_path;                 /* 0 auto, 1 force CUDA-core int8 (dp4a), 2 force tcgen05 */
  int32_t sparse_schur;              /* 0 (default): dense blocks use the GEMM pipeline of src/solver.jl:1089-1104 whatever the A_p look like (the
                                        reference exploits no sparsity there, :1088).  1: a dense block whose constraint matrices are sparse enough
                                        forms its Schur entries from their nonzero lists (SDPA's "F3" formula; MAX-CUT becomes X^-1 o Y) */
} clrs_options;

enum {
  CLRS_OPT_BETA_INFEASIBLE = 0, CLRS_OPT_BETA_FEASIBLE = 1, CLRS_OPT_GAMMA = 2,
  CLRS_OPT_OMEGA_P = 3, CLRS_OPT_OMEGA_D = 4, CLRS_OPT_DUALITY_GAP_THRESHOLD = 5,
  CLRS_OPT_DUAL_ERROR_THRESHOLD = 6, CLRS_OPT_PRIMAL_ERROR_THRESHOLD = 7,
  CLRS_OPT_MAX_COMPLEMENTARY_GAP = 8, CLRS_OPT_STEP_LENGTH_THRESHOLD = 9
};

/* Per-iteration report: the values of one row of the reference's verbose
 * table (src/solver.jl:566-582) plus the 17 phase timers of the `testing`
 * kwarg (src/solver.jl:531-540, 664-718), here in device milliseconds. */
typedef struct clrs_iter_info {
  int32_t iter;          /* 1-based index of the iteration just executed     */
  int32_t stop;          /* CLRS_CONTINUE or a CLRS_STOP_* reason            */
  int32_t pd_feasible;   /* pd_feas after this iteration's update (:443)     */
  int32_t reserved;
  double mu;             /* <X,Y>/K at the start of the iteration            */
  double d_obj, p_obj;   /* objectives at the START of the iteration (:564)  */
  double gap;            /* duality gap at the start of the iteration        */
  double err_P, err_p, err_d;   /* max-abs of P, p, d of this iterate        */
  double alpha_d, alpha_p, beta_c;
  double d_obj_new, p_obj_new, gap_new;  /* of the iterate after the step    */
  double phase_ms[17];   /* decomp, predictor, corrector, alpha, Xinv, R, residuals,
                            schur, cholS, LinvB, Q, cholQ, Z, rhs_x, solve, dX, dY */
} clrs_iter_info;

/* ---- life cycle ------------------------------------------------------- */
void clrs_default_options(clrs_options* opt);
int  clrs_create(const clrs_options* opt, clrs_handle** out);
int  clrs_set_option_num(clrs_handle* h, int which, const void* wire_num);
void clrs_destroy(clrs_handle* h);
const char* clrs_last_error(const clrs_handle* h);
/* size in bytes of one wire number for this handle: 16 + 8*ceil(prec/64) */
size_t clrs_wire_size(const clrs_handle* h);

/* ---- problem upload (a ClusteredLowRankSDP, src/interface.jl:807-819) -- */
/* b (N), constant, maximize */
int clrs_set_free(clrs_handle* h, int32_t N, const void* b, const void* constant, int32_t maximize);
/* cluster j (0-based): B_j is P_j x N row-major, c_j has P_j entries */
int clrs_add_cluster(clrs_handle* h, int32_t j, int32_t P_j, const void* B_j, const void* c_j);
/* PSD block l of cluster j: m x m subblocks of size delta; C is n x n, n = m*delta.
 * high_rank != 0: constraint matrices are dense n x n (m must be 1). */
int clrs_add_block(clrs_handle* h, int32_t j, int32_t l, int32_t m, int32_t delta,
                   int32_t high_rank, const void* C);
/* dense constraint matrix A[j][l][1,1][p]; p is the 0-based COMPACT row of the
 * constraint inside cluster j (already mapped through cs_map, src/solver.jl:156-167) */
int clrs_add_dense_term(clrs_handle* h, int32_t j, int32_t l, int32_t p, const void* A);
/* the same dense-block constraint matrix given by its nnz nonzero entries A[rows[t]][cols[t]] = vals[t] (each position at most once;
 * mirror != 0 also sets the transposed position, for input that lists one triangle like the SDPA sparse format read by
 * src/SDPAtoCLRS.jl:3-31).  Only the triplets cross PCIe; the dense matrix is assembled on the device. */
int clrs_add_sparse_term(clrs_handle* h, int32_t j, int32_t l, int32_t p, int32_t nnz,
                         const int32_t* rows, const int32_t* cols, const void* vals, int32_t mirror);
/* low-rank constraint matrix A[j][l][r,s][p] = sum_k lambda[k] vs[k] ws[k]^T
 * (r, s 0-based subblock indices; vs/ws are rank x delta row-major) */
int clrs_add_lowrank_term(clrs_handle* h, int32_t j, int32_t l, int32_t r, int32_t s,
                          int32_t p, int32_t rank, const void* lambda,
                          const void* vs, const void* ws);
/* Build the deduplicated pairing bases and pointer tables, allocate and fill
 * device memory, initialise x=0, y=0, X=omega_p I, Y=omega_d I and compute the
 * initial objectives, residuals and errors. */
int clrs_finalize(clrs_handle* h);

/* ---- state ------------------------------------------------------------ */
/* x: sum_j P_j numbers; y: N numbers; X, Y: for every block of the SDP in (j,l) order an
 * n x n row-major matrix, concatenated.  NULL pointers are skipped.  A handle that shares the
 * SDP with other ranks (clrs_comm_init) reads / writes only the clusters it owns: the caller
 * passes the full arrays to every rank and merges what the ranks return. */
int clrs_set_state(clrs_handle* h, const void* x, const void* X, const void* y, const void* Y);
int clrs_get_state(clrs_handle* h, void* x, void* X, void* y, void* Y);
/* total count of numbers in X (= in Y): sum over blocks of n^2 */
int64_t clrs_state_matrix_count(const clrs_handle* h);

/* ---- the hot path ----------------------------------------------------- */
/* One predictor-corrector IPM iteration (src/solver.jl:362-592).  Returns
 * CLRS_OK and fills *info; when info->stop != CLRS_CONTINUE no step was taken
 * for stop codes 4/5, and the loop-top `terminate` test fired for 1/2/3. */
int clrs_iterate(clrs_handle* h, clrs_iter_info* info);
/* final objectives and duality gap as full-precision wire numbers (:626-628) */
int clrs_get_objectives(clrs_handle* h, void* d_obj, void* p_obj, void* gap);

/* ---- multi-GPU (clusters / blocks sharded over ranks) -- */
/* nccl_unique_id: the 128-byte ncclUniqueId created on rank 0 and broadcast
 * by the host program.  Must be called before clrs_finalize. */
int clrs_comm_init(clrs_handle* h, int32_t rank, int32_t nranks, const void* nccl_unique_id);
int clrs_comm_unique_id(void* out128);
/* rank that owns cluster j after clrs_finalize (LPT on P_j^3 + sum n^3, src/threadinginfo.jl:88,97) */
int clrs_cluster_owner(clrs_handle* h, int32_t j);
/* A cluster that alone outweighs a rank's fair share and has several PSD blocks is SPLIT (the three-point bound
 * is one cluster of 49-61 blocks): its blocks are spread over the ranks (same LPT, weights n^3), each rank adds its blocks' part of
 * S_j and of <A_*, .>, the parts are summed with one all-reduce, and S_j, its factor and the solves are replicated.  Returns the
 * rank that holds block l of cluster j (= clrs_cluster_owner(j) for a cluster that is not split). */
int clrs_block_owner(clrs_handle* h, int32_t j, int32_t l);
/* the partitioner itself (host only, no GPU needed): owner[j] for J clusters of the given weights */
int clrs_partition_clusters(int32_t J, const double* weight, int32_t nranks, int32_t* owner);
/* the whole plan of a sharded solve (host only): p3[j] = P_j^3, nblocks[j] blocks with weights block_weight[] (concatenated;
 * 15 n^3, or (2 * #A_p + 15) n^3 for a dense block), column_split[j] != 0 for clusters on the column-split path (may be NULL);
 * split_mode -1 by weight (a cluster heavier than 1.25 x the fair share with >= 2 blocks), 0 never, 1 always.
 * Out: cluster_owner[J] (the lead of a split cluster), split[J], block_owner[] (concatenated like block_weight). */
int clrs_plan_shards(int32_t J, const double* p3, const int32_t* nblocks, const double* block_weight, const int32_t* column_split,
                     int32_t nranks, int32_t split_mode, int32_t* cluster_owner, int32_t* split, int32_t* block_owner);

/* ---- standalone kernels of the path (parity tests / microbenchmarks) ---- */
/* C = A * B (M x K times K x N) in multi-limb arithmetic on the device,
 * through the same split -> int8 GEMM -> recombine pipeline the solver uses.
 * path: 0 auto, 1 CUDA-core int8 (dp4a), 2 tcgen05.  Host wire buffers. */
int clrs_mp_gemm(clrs_handle* h, int32_t M, int32_t N, int32_t K,
                 const void* A, const void* B, void* C, int32_t path, double* device_ms);
/* lower Cholesky factor of an n x n matrix; returns CLRS_ERR_CHOL_X on a
 * non-positive pivot (src/tools.jl:75-107) */
int clrs_mp_cholesky(clrs_handle* h, int32_t n, const void* A, void* L);
/* column-pivoted QR of an m x n matrix of wire numbers (modified Gram-Schmidt on the device, pivot = the largest remaining column
 * norm, first index on ties): R (min(m,n) x n, row-major, caller-allocated) in the pivoted column order with a non-negative
 * diagonal, perm[k] = original index of pivoted column k.  The numerical core of preprocess! (src/pre_postprocessing.jl:36:
 * `qr(mpsd, ColumnNorm())` in BigFloat on the host); |R[i][i]| < tol marks the linearly dependent constraints. */
int clrs_mp_qr_pivot(clrs_handle* h, int32_t m, int32_t n, const void* A, void* R, int32_t* perm);
/* device self-test of the warp-cooperative arithmetic against the single-thread routines; returns the
 * number of mismatching samples (0 = pass), -1 on a CUDA error */
int clrs_debug_selftest(clrs_handle* h);
/* debug / parity access to intermediates of the last iteration.  what:
 * "S" (cluster j, l ignored), "Xinv","R","P","dX","dY","X","Y" (block j,l),
 * "Q","d","p","dx","dy","x","y","LinvB" (cluster j).  Returns count written.
 * "sparse?" writes nothing and returns 1 when block (j,l) takes the sparsity shortcut of clrs_options.sparse_schur. */
int64_t clrs_debug_get(clrs_handle* h, const char* what, int32_t j, int32_t l, void* out, int64_t capacity);

/* ---- measurement hooks (bench.py) --------------------------------------- */
/* enable != 0: time every multi-precision GEMM launch with CUDA events on the
 * library's stream and count kernel launches; resets the counters. */
void clrs_profile(clrs_handle* h, int32_t enable);
/* out10 = { ms, mp_flops (2*M*N*K), launches } for three classes of GEMM launches — CUDA-core
 * int8 path, tcgen05 with fewer than 1e6 outputs, tcgen05 with at least 1e6 outputs — then
 * the number of kernel launches since clrs_profile. */
void clrs_profile_get(clrs_handle* h, double* out10);
/* Kernel benchmark of C = A*B (M x K times K x N) on device-generated operands.
 * out3 = { split ms, ms per product (kernel + recombine), ms per tcgen05 kernel launch alone }. */
int clrs_bench_gemm(clrs_handle* h, int32_t M, int32_t N, int32_t K, int32_t reps, int32_t path, double* out3);
/* device time of the last clrs_iterate (CUDA events on the library's stream), ms; 0 when the call did not run an iteration
 * (loop-top terminate fired, or a factorisation failed) */
double clrs_last_iteration_ms(clrs_handle* h);
/* From its second iteration on a handle replays the iteration as a CUDA graph (the loop body is pure stream work: all
 * scalar decisions are taken on the device).  enable = 0 returns to eager launches (parity tests; the environment
 * variable CLRS_GRAPH=0 does the same for every handle).  Default: enabled. */
void clrs_use_graph(clrs_handle* h, int32_t enable);

#ifdef __cplusplus
}
#endif
#endif /* CLRS_B200_H */
