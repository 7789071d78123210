/*
 * fea_gpu.h -- C-ABI of the B200 (sm_100a) finite-strain hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int return codes
 * (0 = ok), no C++/torch types.  Each entry point names the reference
 * function(s) of zbw2577/fea-large `solver-large/` it replaces.  The host C
 * layer under fea-large_b200/host/ keeps the reference's own public names
 * (solve(), solver_create_stiffness(), ...) and forwards to these.
 *
 * There is NO CPU fallback behind any of these calls: without a CUDA device
 * fea_gpu_create() fails with FEA_GPU_ERR_CUDA.
 *
 * Numbering: `nodes`/`conn`/prescribed node ids are the caller's GLOBAL
 * 0-based ids (sexp_loader.c:232, exporter.py:480).  With nranks > 1 every
 * rank passes the same global mesh; the library partitions node ranges
 * (matrix rows) and elements itself and talks NCCL.
 */
#ifndef FEA_GPU_H
#define FEA_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fea_gpu_ctx *fea_gpu_handle;
typedef struct fea_plan *fea_plan_handle;

enum {
  FEA_GPU_OK = 0,
  FEA_GPU_ERR_ARG = 1,      /* bad argument (null, size, unknown enum)       */
  FEA_GPU_ERR_CUDA = 2,     /* CUDA runtime failure / no device              */
  FEA_GPU_ERR_NCCL = 3,     /* NCCL failure                                  */
  FEA_GPU_ERR_MESH = 4,     /* connectivity out of range, no elements, ...   */
  FEA_GPU_ERR_NOT_CONVERGED = 5, /* fea_gpu_solve hit max_iter (u is still set) */
  FEA_GPU_ERR_STALLED = 6   /* fea_gpu_solve ended on its stall / divergence guard before reaching
                               the tolerance; u holds the best checkpointed iterate */
};

/* model_type, reference fea_model.h:37-40 */
enum { FEA_MODEL_A5 = 0, FEA_MODEL_COMPRESSIBLE_NEOHOOKEAN = 1 };

/* fea_gpu_solve() flags */
enum {
  FEA_SOLVE_X0_ZERO = 0,      /* start from u = 0                                        */
  FEA_SOLVE_X0_RHS = 1,       /* start from u = R, as the reference call site does
                                 (fea_solver.c:251-256 passes x0 = b)                    */
  FEA_SOLVE_ABS_TOL = 2,      /* stop on ||r||_2 <= tol instead of ||r||_2 <= tol*||b||_2 */
  FEA_SOLVE_ACCEPT_STALL = 4  /* a solve that ends on the stall / divergence guard returns FEA_GPU_OK
                                 (the caller reads `relres`) instead of FEA_GPU_ERR_STALLED */
};

/* ---- lifetime --------------------------------------------------------- */

/*
 * Replaces fea_solver_alloc (fea_solver.c:387-456) + solver_create_element_database
 * (:556) + fea_model_init (fea_model.c:7): uploads the mesh, builds the sparsity
 * pattern (full 3x3 blocks for every node pair sharing an element, explicit
 * zeros kept, columns ascending -- SURVEY 8c) and the element->nonzero gather map.
 *
 *   X0          [n_nodes][3] reference coordinates (nodes0_p); current = X0 at start (:400)
 *   conn        [n_elems][10] TETRAHEDRA10 connectivity, reference node order (:1287-1300)
 *   n_gauss     4 or 5 (gauss_nodes4/5_tetr10, :32-54)
 *   presc_node/type/vals  prescribed_bnd_node array (fea_solver.h:142-157); type is the
 *               presc_boundary_type bitmask 1=x 2=y 4=z (:74-83), vals [n_presc][3]
 *   rank,nranks,nccl_unique_id  nranks==1 -> id may be NULL.  Otherwise 128 bytes of a
 *               ncclUniqueId created by rank 0 and distributed by the caller.
 *   device      CUDA device ordinal
 */
int fea_gpu_create(fea_gpu_handle *out, int32_t n_nodes, int32_t n_elems,
                   const double *X0, const int32_t *conn, int32_t model_type,
                   double lambda, double mu, int32_t n_gauss, int32_t n_presc,
                   const int32_t *presc_node, const int32_t *presc_type,
                   const double *presc_vals, int32_t rank, int32_t nranks,
                   const void *nccl_unique_id, int32_t device);
/*
 * The same for ONE process driving n_gpus GPUs of the box (the reference's boundary is one process:
 * do_main -> solve, fea_solver.c:100-242).  The library creates one context and one host thread per GPU
 * (devices[i], or 0..n_gpus-1 when devices is NULL) and its own NCCL communicator; the handle that comes
 * back stands for all ranks and every entry point below accepts it: phase calls run on all ranks at
 * once, read-backs return the assembled global arrays, scalar results are the all-reduced ones.
 * fea_gpu_get_csr is per rank and not available on such a handle.  fea_gpu_counts [13] = ranks.
 */
int fea_gpu_create_multi(fea_gpu_handle *out, int32_t n_nodes, int32_t n_elems,
                         const double *X0, const int32_t *conn, int32_t model_type,
                         double lambda, double mu, int32_t n_gauss, int32_t n_presc,
                         const int32_t *presc_node, const int32_t *presc_type,
                         const double *presc_vals, int32_t n_gpus, const int32_t *devices);
/* fea_solver_free (fea_solver.c:459-501) */
int fea_gpu_destroy(fea_gpu_handle h);
/* ncclGetUniqueId for the caller to broadcast (128 bytes) */
int fea_gpu_nccl_unique_id(void *out128);
const char *fea_gpu_last_error(void);

/* ---- node coordinates -------------------------------------------------- */

/* nodes_p <- x  (host [n_nodes][3], global ids); each rank takes its owned+ghost part */
int fea_gpu_set_nodes(fea_gpu_handle h, const double *x);
/* x <- nodes_p; with nranks>1 a collective (all-gather), every rank gets all nodes */
int fea_gpu_get_nodes(fea_gpu_handle h, double *x);
/* solver_update_nodes_with_bc(self, lambda) (fea_solver.c:1281, :1205-1242, :1259) */
int fea_gpu_apply_increment(fea_gpu_handle h, double lambda);
/* solver_update_nodes_with_solution(self, global_solution_vct) (:1270-1279) */
int fea_gpu_update_nodes(fea_gpu_handle h);
/* nodes_p += eta * global_solution_vct: the step of a line search along the Newton direction
 * (solver-prototype/cartesian3d/large/cartesian3d_large.m:85-119 evaluates R at nodes1 + eta X) */
int fea_gpu_update_nodes_scaled(fea_gpu_handle h, double eta);
/* keep / restore nodes_p on the device: the `nodes1 = nodes` of the prototype's line search (:65) and the
 * roll-back of a load increment that inverted elements (host/fea_solver.c: solve) */
int fea_gpu_save_nodes(fea_gpu_handle h);
int fea_gpu_restore_nodes(fea_gpu_handle h);
/* nodes_p <- nodes_p + alpha (nodes_p - saved), saved <- the old nodes_p.  With the nodes of the previous
 * increment saved and equal increments, alpha = 1 moves the prescribed nodes by exactly one more increment
 * (what solver_update_nodes_with_bc(self, 1) does, fea_solver.c:168) and gives the interior a secant guess
 * instead of leaving it behind: same equilibrium, fewer Newton iterations (tools/load_sequence.py). */
int fea_gpu_extrapolate_nodes(fea_gpu_handle h, double alpha);

/* ---- element phase ----------------------------------------------------- */

/* solver_create_current_shape_gradients (:831) + solver_create_stresses (:843):
 * J, detJ, grad N, F (through F^-1, the CURRENT_SHAPE_GRADIENTS branch :1131-1152),
 * Cauchy stress (fea_model.c:26/79) for every (element, Gauss point) */
int fea_gpu_update_state(fea_gpu_handle h);
/* solver_create_stiffness (:873): K_e constitutive (:887) + initial stress (:986),
 * deterministic gather into the global matrix (replaces sp_matrix_element_add :966,:1055) */
int fea_gpu_assemble_stiffness(fea_gpu_handle h);
/* solver_create_residual_forces (:863, :1072-1114): global_forces_vct = -int sigma grad N */
int fea_gpu_assemble_residual(fea_gpu_handle h);
/* state + stiffness + residual from the current nodes in ONE element pass.
 * flags: FEA_ASSEMBLE_STIFFNESS also rebuilds K (0 = modified-Newton iteration, R only);
 * FEA_ASSEMBLE_FUSE_BC folds solver_apply_prescribed_bc(self, 0) (:1200, :1244-1257) into the two
 * gathers -- bitwise what fea_gpu_apply_bc(h, 0) gives afterwards, one sweep over K less */
#define FEA_ASSEMBLE_STIFFNESS 1
#define FEA_ASSEMBLE_FUSE_BC 2
int fea_gpu_assemble_all(fea_gpu_handle h, int32_t flags);
/* solver_apply_prescribed_bc(self, lambda) (:1200, :1244-1257 + sp_matrix_cross_cancellation) */
int fea_gpu_apply_bc(fea_gpu_handle h, double lambda);
/* keep / restore the assembled matrix: sp_matrix_copy at :179 and :194-195 (modified Newton) */
int fea_gpu_save_stiffness(fea_gpu_handle h);
int fea_gpu_restore_stiffness(fea_gpu_handle h);

/* ---- linear solve and Newton bookkeeping ------------------------------- */

/* solver_solve_slae (:300-321): Jacobi-preconditioned CG on K u = R.
 * Stops when ||r|| <= tol ||b|| (or tol, FEA_SOLVE_ABS_TOL).  If ||r|| stops improving
 * (rounding floor) or diverges -- K of the reference's "analytical" models is singular, so a
 * noise-level right-hand side is inconsistent -- the solve ends with the checkpointed
 * near-minimum-residual iterate in u and returns FEA_GPU_ERR_STALLED (FEA_GPU_OK under
 * FEA_SOLVE_ACCEPT_STALL); `relres` tells what was reached.  iters/relres may be NULL.
 * Returns FEA_GPU_ERR_NOT_CONVERGED at max_iter.
 * "pcg_variant" = 1 (fea_gpu_set_param) runs the single-reduction (Chronopoulos-Gear) form of the same
 * iteration -- one all-reduce of four doubles per iteration instead of two small ones; it is not the
 * default because its step-length recurrence can break down where classic PCG converges. */
int fea_gpu_solve(fea_gpu_handle h, double tol, int32_t max_iter, int32_t flags,
                  int32_t *iters, double *relres);
/* cdot(global_forces_vct, global_solution_vct, n) at :208 */
int fea_gpu_dot_R_u(fea_gpu_handle h, double *out);
/* y = K x with host vectors in global numbering (tests / diagnostics) */
int fea_gpu_spmv(fea_gpu_handle h, const double *x, double *y);

/* ---- read-back (host arrays, global numbering) ------------------------- */

/* graddefs / stresses [n_elems][n_gauss][3][3] (fea_solver.h:262-269).  With
 * nranks>1 only elements owned by this rank are written. */
int fea_gpu_get_state(fea_gpu_handle h, double *graddefs, double *stresses);
/* the same for a list of n elements (global ids): graddefs / stresses [n][n_gauss][3][3]; found[k]
 * (may be NULL) tells whether element k is local to this rank, rows of others are left untouched */
int fea_gpu_get_state_elems(fea_gpu_handle h, int32_t n, const int32_t *elems, double *graddefs,
                            double *stresses, int32_t *found);
int fea_gpu_get_forces(fea_gpu_handle h, double *R);      /* global_forces_vct   */
int fea_gpu_set_forces(fea_gpu_handle h, const double *R);
int fea_gpu_get_solution(fea_gpu_handle h, double *u);    /* global_solution_vct */
/* scalar CSR of the rows this rank owns: int32 indices, columns ascending, explicit
 * zeros kept.  Call with NULL arrays to get sizes: n_rows, nnz.  `rows` (may be NULL)
 * receives the global DOF id of each returned row. */
int fea_gpu_get_csr(fea_gpu_handle h, int64_t *n_rows, int64_t *nnz, int32_t *rows,
                    int32_t *rowptr, int32_t *colidx, double *vals);
/* dense K_e [30][30] (row 3a+i, column 3b+j) of one element, global id, as the last element pass with
 * stiffness staged it: constitutive + initial-stress part (fea_solver.c:887-1068).  The element must be
 * local to this rank (FEA_GPU_ERR_ARG otherwise).  Diagnostic / parity checks. */
int fea_gpu_get_element_matrix(fea_gpu_handle h, int32_t element, double *ke900);
/* elements whose |J| or det F was <= 0 (or J singular) in the last element pass */
int fea_gpu_bad_points(fea_gpu_handle h, int64_t *count);

/* ---- host-buffer (end-to-end) path --------------------------------------- */

/* page-locked host memory for the arrays the host layer hands to the calls below
 * (nodes_p, global_forces_vct, ...), so copies are direct DMA */
int fea_gpu_host_alloc(void **out, uint64_t bytes);
int fea_gpu_host_free(void *p);
/* One assembly pass driven from HOST arrays, as the reference's phase sequence
 * solver_create_current_shape_gradients + _stresses + _stiffness + _residual_forces +
 * solver_apply_prescribed_bc(0) (fea_solver.c:171-203) would be with host-resident nodes:
 * copies x [n_nodes][3] host->device, runs the element pass, gathers, cancels BC rows,
 * copies global_forces_vct (this rank's rows; all rows when nranks == 1) device->host.
 * h2d_bytes / d2h_bytes (may be NULL) report what crossed the bus. */
int fea_gpu_step_from_host(fea_gpu_handle h, const double *x, int32_t with_stiffness,
                           double *R, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* ---- introspection / measurement --------------------------------------- */

/* out[0]=owned nodes, [1]=local nodes (owned+ghost), [2]=local elements,
 * [3]=block nonzeros (3x3), [4]=gather contributions, [5]=neighbour ranks,
 * [6]=halo nodes sent, [7]=halo nodes received, [8]=global nodes, [9]=global elements,
 * [10]=SELL block slots (incl. padding), [11]=SELL slices, [12]=1 if the nine-lane gather can run
 * on this pattern (fea_gpu_counts only) */
int fea_gpu_counts(fea_gpu_handle h, int64_t out[16]);
/* kernels launched by this library in this process (all handles) */
int64_t fea_gpu_launch_count(void);
/* device-side stopwatch on the library's stream (CUDA events) */
int fea_gpu_timer_start(fea_gpu_handle h);
int fea_gpu_timer_stop(fea_gpu_handle h, double *ms);
int fea_gpu_sync(fea_gpu_handle h);
/* per-phase device time, ms, averaged over the calls of each phase since the previous
 * fea_gpu_phase_ms (CUDA events around every launch, up to the last 32 calls; reading synchronises
 * the stream and restarts the averages; [14] = element-pass calls averaged):
 * [0]=element kernel, [1]=matrix gather, [2]=residual gather, [3]=bc,
 * [4]=pcg total, [5]=average in-solve spmv launch, [6]=halo, [8]=spmv launches timed,
 * [9]=pcg iterations, [10]=pcg exit (0 max_iter, 1 tolerance, 2 stall/divergence guard),
 * [11]=best relative residual seen, [12]=relative residual of the last iterate, [13]=stall count */
int fea_gpu_phase_ms(fea_gpu_handle h, double out[16]);
/* repeated SpMV on device vectors for roofline measurement: average ms per SpMV */
int fea_gpu_bench_spmv(fea_gpu_handle h, int32_t reps, double *ms_per_spmv);
/* measured machine peaks on this device: FP64 FMA TFLOP/s, copy GB/s (read+write) */
int fea_gpu_measure_peaks(int32_t device, double *dfma_tflops, double *copy_gbs);
/* FP64 tensor-core peak of this device: mma.sync.m8n8k4.f64 issue rate in TFLOP/s (north_star: DMMA is
 * used for the element contraction only if it beats the FMA pipe -- this is the measurement) */
int fea_gpu_measure_dmma(int32_t device, double *dmma_tflops);
/* average device ms of the two collectives of a PCG iteration, each timed alone (`reps` back to back on
 * the context's stream): the halo exchange of one [local nodes][3] vector and the all-reduce of the
 * four iteration sums.  0 when nranks == 1.  Either pointer may be NULL.  Collective. */
int fea_gpu_bench_comm(fea_gpu_handle h, int32_t reps, double *halo_ms, double *allreduce_ms);
/* tuning knobs (INTEGRATION.md has the table with the measurements behind every default):
 * "gather_mode" (1 = pull gather, one lane per block slot, the default; 9 = nine lanes per staged block, flat-staging
 * builds only; 2 = direct assembly: the element kernel writes destination-ordered cells, gather_cells_kernel adds
 * them with coalesced loads), "gather_sym" (1 = the pull gather sums the upper block triangle only and stores every
 * block into its mirror slot too, the default; 0 = every slot sums its own list), "chunk_tiles" (> 0: assembly in
 * chunks of that many 32-element tiles captured into one CUDA graph, staging read back from L2; default 0),
 * "chunk_overlap" (chunked assembly: gather beside the next chunk's elements; default 1), "gather_threads",
 * "gather_split", "elem_ratio" -- all bitwise neutral for K;
 * "pcg_variant" (0 = classic two-reduction PCG, the default; 1 = single-reduction),
 * "pcg_overlap" (1 = halo exchange beside the interior SpMV slices; default 0),
 * "precond" (0 = Jacobi, the default; 1 = Chebyshev-accelerated Jacobi z = p_d(D^-1 A) D^-1 r -- what the
 * host layer selects for a task file's PCG_ILU: several times fewer iterations and reductions for about
 * the same number of matrix products), "cheb_degree" (default 4), "cheb_ratio" (the polynomial targets
 * [lmax / ratio, lmax] of D^-1 A, lmax from a power iteration; default 100), "pcg_batch" (iterations
 * queued between host convergence checks), "pcg_stall" (iterations without a new best
 * ||r|| before PCG declares the rounding floor; 0 = automatic, max(500, 50 n^(1/3))) */
int fea_gpu_set_param(fea_gpu_handle h, const char *name, double value);
/* overwrite >= `bytes` of scratch so L2 holds none of the caller's data */
int fea_gpu_flush_l2(fea_gpu_handle h);

/* ---- host-only planning (no CUDA calls; testable on a CPU box) --------- */

/* partition + symbolic phase exactly as fea_gpu_create runs it */
int fea_plan_create(fea_plan_handle *out, int32_t n_nodes, int32_t n_elems,
                    const double *X0, const int32_t *conn, int32_t rank, int32_t nranks);
int fea_plan_destroy(fea_plan_handle p);
/* same slots as fea_gpu_counts */
int fea_plan_counts(fea_plan_handle p, int64_t out[16]);
/* any pointer may be NULL.
 *   local_node_gid  [local nodes]      global id of each local node (owned first, Morton order)
 *   local_elem_gid  [local elements]   (Morton order of the centroids)
 *   browptr         [owned nodes + 1]  block-row pointers
 *   bcol            [nnzb]             local column node ids, ascending
 *   cptr            [nnzb + 1], csrc [contributions]  gather map (ascending global element id);
 *                   csrc = elem*55 + code(a,b) | (1<<31 if the stored block is transposed);
 *                   code = 11*min(a,9-a) + pos orders the a<=b blocks as the element kernel
 *                   stages them (fea_plan.hpp: ke_code)
 *   nbr_rank [nbr], send_ptr [nbr+1], send_nodes [sent] (local ids), recv_ptr [nbr+1]
 *                   (ghost offset ranges, in local numbering minus owned count) */
int fea_plan_arrays(fea_plan_handle p, int32_t *local_node_gid, int32_t *local_elem_gid,
                    int32_t *browptr, int32_t *bcol, int32_t *cptr, uint32_t *csrc,
                    int32_t *nbr_rank, int32_t *send_ptr, int32_t *send_nodes,
                    int32_t *recv_ptr);
int fea_plan_node_owner(fea_plan_handle p, int32_t *owner /* [n_nodes] */);
/* the SELL-32 device layout of the same pattern (any pointer may be NULL):
 *   slice_ptr [slices + 1], sell_row [32 * slices] (-1 = padding lane), sbcol [slots],
 *   scptr [slots + 1], scsrc [contributions], sdiag [owned nodes] */
int fea_plan_sell_arrays(fea_plan_handle p, int32_t *slice_ptr, int32_t *sell_row, int32_t *sbcol,
                         int32_t *scptr, uint32_t *scsrc, int32_t *sdiag);

/* the cell layout of the direct (push) assembly (fea_plan.cpp "cell layout"; any pointer may be NULL; sizes from
 * fea_plan_counts: [10] slots, [12] cells, [13] columns with upper slots):
 *   cmeta [slots] contributions | rank << 11 of every upper slot (0 = lower triangle / padding),
 *   ccell [slots / 32 + 1] first cell of each 32-slot column, cmirror [slots] value index of the mirror slot or -1,
 *   edest [55][elements rounded up to 32] cell | (1<<31: store transposed) of every staged block, 0xffffffff = none,
 *   col_ready [slots / 32] last local element contributing (-1 = no upper slot), col_order [.] ascending col_ready */
int fea_plan_cell_arrays(fea_plan_handle p, uint16_t *cmeta, int32_t *ccell, int32_t *cmirror, uint32_t *edest,
                         int32_t *col_ready, int32_t *col_order);

/* Kuhn (Freudenthal) 6-tet block, 10-node tets in the reference node order:
 * nx*ny*nz cubes on [0,lx]x[y0,y0+ly]x[0,lz]; nodes (2nx+1)(2ny+1)(2nz+1), tets 6*nx*ny*nz.
 * Call with NULL arrays for sizes.  bc_style 0 = "analytical" (face y=y0: type 2 value 0
 * plus one corner type 7; face y=y0+ly: type 2 value dy -- K keeps a free rotation about y, as in
 * the reference's *_analytical.sexp), 1 = clamped (type 7 on both), 2 = style 0 plus a second bottom
 * corner held in z (no rigid mode left; the homogeneous uniaxial state is still the solution). */
int fea_mesh_block(int32_t nx, int32_t ny, int32_t nz, double lx, double ly, double lz,
                   double y0, int32_t bc_style, double dy, int64_t *n_nodes,
                   int64_t *n_elems, int64_t *n_presc, double *nodes, int32_t *conn,
                   int32_t *presc_node, int32_t *presc_type, double *presc_vals);

/* Thick-walled hollow cylinder about the z axis (BASELINE configs[1], the Lame problem of
 * exact-solutions/lame): nr x nt x nz cells in (r, theta, z), each cut into 6 Kuhn tets with
 * straight edges (mid-side nodes at edge midpoints, as every shipped mesh).  Nodes
 * (2nr+1)(2nt)(2nz+1), tets 6 nr nt nz.  Boundary conditions, per load increment: inner wall
 * r = r_in moves radially by `delta` (x and y prescribed), outer wall r = r_out is held in x and
 * y, both end faces are held in z (plane strain).  Call with NULL arrays for sizes. */
int fea_mesh_cylinder(int32_t nr, int32_t nt, int32_t nz, double r_in, double r_out, double length,
                      double delta, int64_t *n_nodes, int64_t *n_elems, int64_t *n_presc,
                      double *nodes, int32_t *conn, int32_t *presc_node, int32_t *presc_type,
                      double *presc_vals);

#ifdef __cplusplus
}
#endif
#endif /* FEA_GPU_H */
