/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Plain-C CPU restatement of the reference hot path (solver-large/), one
 * function per reference function, same loop order and operation order so
 * results agree with the reference-compiled objects (oracle/_ref) to the
 * last bit where the arithmetic is the reference's own.  Pinned against
 * oracle/_ref and the golden vectors by tests/test_oracle_*.py.
 *
 * Parity status: element arithmetic, BC bookkeeping and the Newton driver are
 * pinned against the reference's compiled code; the sparse container and the
 * linear solver restate libspmatrix's *call-site semantics* only (source
 * absent from /root/reference) -> that part is "parity unpinned".
 */
#ifndef ORACLE_FEA_H
#define ORACLE_FEA_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MODEL_A5 = 0, ORC_MODEL_NEOHOOKEAN = 1 };

/* ---- tables: fea_solver.c:32-54, 1287-1373 ---- */
void orc_gauss_table(int count, double *out /* [count][4] = {w,r,s,t} */);
void orc_shape_functions(double r, double s, double t, double *N /*[10]*/,
                         double *dN /*[3][10]*/);

/* ---- dense 3x3: dense_matrix.c:25-110 ---- */
double orc_det3(const double *m);
int orc_inv3(double *m, double *det);
void orc_mul3(const double *A, const double *B, double *R);      /* A  B  */
void orc_mul3_tn(const double *A, const double *B, double *R);   /* A' B  */
void orc_mul3_nt(const double *A, const double *B, double *R);   /* A  B' */

/* ---- per (element, Gauss point) ---- */
/* fea_solver.c:656-722; returns 0 when det J == 0 exactly */
int orc_shape_gradients(const double *dN /*[3][10]*/, const double *xe /*[10][3]*/,
                        double *g /*[3][10]*/, double *detJ);
/* fea_solver.c:1131-1152 (CURRENT_SHAPE_GRADIENTS branch) */
void orc_graddef(const double *g /*[3][10]*/, const double *X0e /*[10][3]*/,
                 double *F /*[3][3]*/);
/* fea_model.c:26-107 */
void orc_stress(int model, double lambda, double mu, const double *F, double *S);
/* fea_model.c:110-148 */
void orc_ctensor(int model, double lambda, double mu, const double *F,
                 double *c /*[3][3][3][3]*/);

/* ---- solver object ---- */
typedef struct orc_solver orc_solver;

orc_solver *orc_create(int n_nodes, const double *nodes, int n_elems,
                       const int *conn, int n_presc, const int *presc_node,
                       const int *presc_type, const double *presc_vals,
                       int model, double lambda, double mu, int gauss_count);
void orc_destroy(orc_solver *s);

void orc_set_nodes(orc_solver *s, const double *x);
void orc_get_nodes(const orc_solver *s, double *x);
void orc_apply_increment(orc_solver *s, double lambda);  /* fea_solver.c:1281 */
void orc_update_state(orc_solver *s);                     /* :831 + :843 */
void orc_get_state(const orc_solver *s, double *F, double *S);
void orc_get_gradients(const orc_solver *s, double *g, double *detJ);

/* K_e of one element, part 0 = both, 1 = constitutive (:887), 2 = initial stress (:986) */
void orc_element_matrix(const orc_solver *s, int element, double *ke /*[30][30]*/, int part);

void orc_assemble_stiffness(orc_solver *s);               /* :873 */
void orc_assemble_residual(orc_solver *s);                /* :863 */
void orc_apply_bc(orc_solver *s, double lambda);          /* :1200-1257 */
long orc_nnz(const orc_solver *s);
void orc_get_csr(const orc_solver *s, int *rowptr, int *colidx, double *vals);
void orc_get_forces(const orc_solver *s, double *R);
void orc_set_forces(orc_solver *s, const double *R);
void orc_get_solution(const orc_solver *s, double *u);
/* Jacobi-PCG stand-in for libspmatrix's solve; returns iterations */
int orc_solve_slae(orc_solver *s, double rel_tol, int max_iter);
void orc_update_with_solution(orc_solver *s);             /* :1270 */
double orc_dot_forces_solution(const orc_solver *s);      /* :208 */

/* the Newton driver, fea_solver.c:130-242.  Returns the number of completed
 * load increments; `trace_u` (may be NULL) receives u of every Newton
 * iteration back to back, `trace_tol` the <R,u> values, `n_trace` their count. */
int orc_newton_solve(orc_solver *s, int load_increments, double desired_tol,
                     int modified_newton, int max_newton, double lin_tol,
                     int lin_max_iter, double *trace_u, double *trace_tol,
                     int trace_cap, int *n_trace);

/* timing helper for bench.py's cpu_baseline leg: state + K + R of the first
 * `n_elems_sample` elements, repeated `reps` times; returns seconds */
double orc_time_assembly(orc_solver *s, int reps);

#ifdef __cplusplus
}
#endif
#endif
