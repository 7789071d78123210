/*
 * ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Host stand-in for the three libraries the reference links but does not
 * vendor (libspmatrix, liblogger, libsexp), plus a small C harness so the
 * reference's OWN compiled objects (fea_solver.c, fea_model.c, dense_matrix.c,
 * tests.c -- compiled from /root/reference by oracle/Makefile into
 * oracle/_ref/) can be driven from ctypes.
 *
 * What is reference-compiled: every element-level number (shape gradients, F,
 * Cauchy stress, tangent, all K_e/R_e contributions as they arrive at
 * sp_matrix_element_add), the BC bookkeeping, and the Newton driver solve().
 * What is a stand-in (libspmatrix source absent => "parity unpinned"): the
 * sparse container and the linear solvers below.  The solvers are a tight
 * Jacobi-PCG so that K u = R is solved to ~1e-15 relative residual whichever
 * solver type the model file names.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline
 * legs may load the resulting library.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "fea_solver.h" /* the reference's own header, found via -I */
#include "logger.h"
#include "sexp_loader.h"
#include "sp_iter.h"
#include "sp_utils.h"
#include "tests.h"

/* ------------------------------------------------------------------ */
/* liblogger stand-in: capture "Tolerance <X,R>" values, optional echo  */

static int g_log_echo = 0;
static double *g_tol_trace = NULL;
static int g_tol_count = 0, g_tol_cap = 0;

void logger_init_with_params(logger_parameters *p) { (void)p; }
void logger_fini(void) {}

void ref_log_capture(int level, const char *fmt, ...) {
  va_list ap;
  if (strncmp(fmt, "Tolerance", 9) == 0) {
    double v;
    va_start(ap, fmt);
    v = va_arg(ap, double);
    va_end(ap);
    if (g_tol_count == g_tol_cap) {
      g_tol_cap = g_tol_cap ? 2 * g_tol_cap : 256;
      g_tol_trace = (double *)realloc(g_tol_trace, sizeof(double) * g_tol_cap);
    }
    g_tol_trace[g_tol_count++] = v;
  }
  if (g_log_echo || level == 2) {
    va_start(ap, fmt);
    vfprintf(level == 2 ? stderr : stdout, fmt, ap);
    va_end(ap);
    fputc('\n', level == 2 ? stderr : stdout);
  }
}

/* libsexp is absent: the reference loader is not compiled; the symbol the
 * reference's initial_data_load() references is satisfied by a refusal. */
BOOL sexp_data_load(char *filename, fea_task **task,
                    fea_solution_params **fea_params, nodes_array **nodes,
                    elements_array **elements,
                    presc_bnd_array **presc_boundary) {
  (void)filename; (void)task; (void)fea_params; (void)nodes; (void)elements;
  (void)presc_boundary;
  return FALSE;
}

/* ------------------------------------------------------------------ */
/* sp_utils stand-ins                                                   */

const char *sp_parse_file_extension(const char *filename) {
  const char *dot = strrchr(filename, '.');
  return dot ? dot + 1 : NULL;
}
const char *sp_parse_file_basename(const char *filename, char *out) {
  const char *dot = strrchr(filename, '.');
  size_t n = dot ? (size_t)(dot - filename) : strlen(filename);
  memcpy(out, filename, n);
  out[n] = 0;
  return out;
}
int sp_istrcmp(const char *a, const char *b) { return strcasecmp(a, b); }

/* ------------------------------------------------------------------ */
/* sparse container stand-in: per-column sorted arrays                  */

static long g_add_calls = 0;
static int g_scatter_mode = 0; /* 0 = store, 1 = no-op (arithmetic-only timing) */
/* element-matrix capture (ref_element_matrix) */
static double *g_cap_ke = NULL;
static const int *g_cap_conn = NULL;

void sp_matrix_init(sp_matrix_ptr m, int rows, int cols, int bandwidth,
                    sparse_storage_type type) {
  int i;
  if (bandwidth < 4) bandwidth = 4;
  m->rows_count = rows;
  m->cols_count = cols;
  m->ordered = 1;
  m->storage_type = type;
  m->storage = (indexed_array *)malloc(sizeof(indexed_array) * (size_t)cols);
  for (i = 0; i < cols; ++i) {
    m->storage[i].width = 0;
    m->storage[i].last_index = -1;
    m->storage[i].indexes = NULL;
    m->storage[i].values = NULL;
  }
  (void)bandwidth;
}

sp_matrix_ptr sp_matrix_free(sp_matrix_ptr m) {
  int i;
  if (m && m->storage) {
    for (i = 0; i < m->cols_count; ++i) {
      free(m->storage[i].indexes);
      free(m->storage[i].values);
    }
    free(m->storage);
    m->storage = NULL;
  }
  return NULL;
}

void sp_matrix_clear(sp_matrix_ptr m) {
  int i;
  for (i = 0; i < m->cols_count; ++i)
    if (m->storage[i].last_index >= 0)
      memset(m->storage[i].values, 0,
             sizeof(double) * (size_t)(m->storage[i].last_index + 1));
}

void sp_matrix_copy(sp_matrix_ptr src, sp_matrix_ptr dst) {
  int i;
  dst->rows_count = src->rows_count;
  dst->cols_count = src->cols_count;
  dst->ordered = src->ordered;
  dst->storage_type = src->storage_type;
  dst->storage =
      (indexed_array *)malloc(sizeof(indexed_array) * (size_t)src->cols_count);
  for (i = 0; i < src->cols_count; ++i) {
    int n = src->storage[i].last_index + 1;
    dst->storage[i].width = n;
    dst->storage[i].last_index = n - 1;
    dst->storage[i].indexes = n ? (int *)malloc(sizeof(int) * (size_t)n) : NULL;
    dst->storage[i].values =
        n ? (double *)malloc(sizeof(double) * (size_t)n) : NULL;
    if (n) {
      memcpy(dst->storage[i].indexes, src->storage[i].indexes,
             sizeof(int) * (size_t)n);
      memcpy(dst->storage[i].values, src->storage[i].values,
             sizeof(double) * (size_t)n);
    }
  }
}

/* position of `key` in the sorted array, or -(insertion point)-1 */
static int ia_find(const indexed_array *a, int key) {
  int lo = 0, hi = a->last_index;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    int v = a->indexes[mid];
    if (v == key) return mid;
    if (v < key) lo = mid + 1; else hi = mid - 1;
  }
  return -lo - 1;
}

double sp_matrix_element_add(sp_matrix_ptr m, int i, int j, double value) {
  indexed_array *col;
  int pos;
  ++g_add_calls;
  if (g_cap_ke) { /* capture into a dense 30x30 element matrix */
    int a, b, li = -1, lj = -1;
    for (a = 0; a < 10; ++a) {
      if (g_cap_conn[a] == i / 3) li = 3 * a + i % 3;
      if (g_cap_conn[a] == j / 3) lj = 3 * a + j % 3;
    }
    (void)b;
    g_cap_ke[li * 30 + lj] += value;
    return value;
  }
  if (g_scatter_mode == 1) return value;
  col = &m->storage[j]; /* CCS: storage indexed by column, holds row ids */
  pos = ia_find(col, i);
  if (pos < 0) {
    int ins = -pos - 1, n = col->last_index + 1;
    if (n == col->width) {
      col->width = col->width ? 2 * col->width : 64;
      col->indexes = (int *)realloc(col->indexes, sizeof(int) * (size_t)col->width);
      col->values =
          (double *)realloc(col->values, sizeof(double) * (size_t)col->width);
    }
    memmove(col->indexes + ins + 1, col->indexes + ins,
            sizeof(int) * (size_t)(n - ins));
    memmove(col->values + ins + 1, col->values + ins,
            sizeof(double) * (size_t)(n - ins));
    col->indexes[ins] = i;
    col->values[ins] = 0.0;
    col->last_index = n;
    pos = ins;
  }
  col->values[pos] += value;
  return col->values[pos];
}

/* zero row+column `index`, keep and return the diagonal (fea_solver.c:1255) */
double sp_matrix_cross_cancellation(sp_matrix_ptr m, int index) {
  indexed_array *col = &m->storage[index];
  double diag = 0.0;
  int k;
  for (k = 0; k <= col->last_index; ++k) {
    int r = col->indexes[k];
    if (r == index) {
      diag = col->values[k];
    } else {
      int p = ia_find(&m->storage[r], index); /* entry (index, r) */
      if (p >= 0) m->storage[r].values[p] = 0.0;
      col->values[k] = 0.0;
    }
  }
  return diag;
}

/* CSR built as the exact transpose of the CCS storage */
void sp_matrix_yale_init(sp_matrix_yale_ptr y, sp_matrix_ptr m) {
  int n = m->rows_count, j, k;
  long nnz = 0;
  int *fill;
  for (j = 0; j < m->cols_count; ++j) nnz += m->storage[j].last_index + 1;
  y->storage_type = CRS;
  y->rows_count = n;
  y->cols_count = m->cols_count;
  y->nonzeros = (int)nnz;
  y->offsets = (int *)calloc((size_t)n + 1, sizeof(int));
  y->indexes = (int *)malloc(sizeof(int) * (size_t)nnz);
  y->values = (double *)malloc(sizeof(double) * (size_t)nnz);
  for (j = 0; j < m->cols_count; ++j)
    for (k = 0; k <= m->storage[j].last_index; ++k)
      y->offsets[m->storage[j].indexes[k] + 1]++;
  for (j = 0; j < n; ++j) y->offsets[j + 1] += y->offsets[j];
  fill = (int *)malloc(sizeof(int) * (size_t)n);
  memcpy(fill, y->offsets, sizeof(int) * (size_t)n);
  for (j = 0; j < m->cols_count; ++j) /* columns ascending => sorted rows */
    for (k = 0; k <= m->storage[j].last_index; ++k) {
      int r = m->storage[j].indexes[k];
      y->indexes[fill[r]] = j;
      y->values[fill[r]] = m->storage[j].values[k];
      fill[r]++;
    }
  free(fill);
}

void sp_matrix_yale_free(sp_matrix_yale_ptr y) {
  free(y->offsets); free(y->indexes); free(y->values);
  y->offsets = NULL; y->indexes = NULL; y->values = NULL;
}

void sp_matrix_create_ilu(sp_matrix_ptr m, sp_matrix_skyline_ilu_ptr ilu) {
  (void)m; ilu->dummy = 0;
}
void sp_matrix_skyline_ilu_free(sp_matrix_skyline_ilu_ptr ilu) { (void)ilu; }

/* ------------------------------------------------------------------ */
/* linear solver stand-in + Newton trace                                */

typedef struct {
  int n;
  double *rhs;
  double *sol;
} solve_record;
static solve_record *g_trace = NULL;
static int g_trace_count = 0, g_trace_cap = 0, g_trace_on = 0;
static int g_last_iters = 0;
static double g_last_relres = 0.0;

static void trace_push(int n, const double *b, const double *x) {
  solve_record *r;
  if (!g_trace_on) return;
  if (g_trace_count == g_trace_cap) {
    g_trace_cap = g_trace_cap ? 2 * g_trace_cap : 64;
    g_trace = (solve_record *)realloc(g_trace, sizeof(solve_record) * g_trace_cap);
  }
  r = &g_trace[g_trace_count++];
  r->n = n;
  r->rhs = (double *)malloc(sizeof(double) * (size_t)n);
  r->sol = (double *)malloc(sizeof(double) * (size_t)n);
  memcpy(r->rhs, b, sizeof(double) * (size_t)n);
  memcpy(r->sol, x, sizeof(double) * (size_t)n);
}

static void csr_mv(const sp_matrix_yale *A, const double *x, double *y) {
  int i, k;
  for (i = 0; i < A->rows_count; ++i) {
    double s = 0.0;
    for (k = A->offsets[i]; k < A->offsets[i + 1]; ++k)
      s += A->values[k] * x[A->indexes[k]];
    y[i] = s;
  }
}

/* Jacobi-PCG, x0 = 0, relative residual, with stagnation guard */
static void standin_pcg(const sp_matrix_yale *A, const double *b, double *x,
                        double rel_tol, int max_iter) {
  int n = A->rows_count, i, k, it = 0;
  double *r = (double *)malloc(sizeof(double) * (size_t)n * 5);
  double *z = r + n, *p = z + n, *q = p + n, *dinv = q + n;
  double bb = 0.0, rz = 0.0, rr = 0.0, best = 1e300;
  int stall = 0;
  for (i = 0; i < n; ++i) {
    double d = 1.0;
    for (k = A->offsets[i]; k < A->offsets[i + 1]; ++k)
      if (A->indexes[k] == i) d = A->values[k];
    dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += b[i] * b[i];
    rz += r[i] * z[i];
  }
  rr = bb;
  if (bb > 0.0) {
    for (it = 0; it < max_iter; ++it) {
      double pq = 0.0, alpha, rz_new = 0.0, beta;
      if (sqrt(rr) <= rel_tol * sqrt(bb)) break;
      csr_mv(A, p, q);
      for (i = 0; i < n; ++i) pq += p[i] * q[i];
      alpha = rz / pq;
      rr = 0.0;
      for (i = 0; i < n; ++i) {
        x[i] += alpha * p[i];
        r[i] -= alpha * q[i];
        z[i] = dinv[i] * r[i];
        rz_new += r[i] * z[i];
        rr += r[i] * r[i];
      }
      beta = rz_new / rz;
      rz = rz_new;
      for (i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
      if (rr < best * 0.999) { best = rr; stall = 0; }
      else if (++stall > 200) break; /* at the rounding floor */
    }
  }
  g_last_iters = it;
  g_last_relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  free(r);
}

void sp_matrix_yale_solve_cg(sp_matrix_yale_ptr m, double *b, double *x0,
                             int *max_iter, double *tolerance, double *x) {
  int n = m->rows_count;
  double *rhs = (double *)malloc(sizeof(double) * (size_t)n);
  (void)x0; /* the reference passes x0 == b (fea_solver.c:252-253) */
  memcpy(rhs, b, sizeof(double) * (size_t)n);
  standin_pcg(m, rhs, x, 1e-15, 200000);
  *max_iter = g_last_iters;
  *tolerance = g_last_relres;
  trace_push(n, rhs, x);
  free(rhs);
}

void sp_matrix_yale_solve_pcg_ilu(sp_matrix_yale_ptr m,
                                  sp_matrix_skyline_ilu_ptr ilu, double *b,
                                  double *x0, int *max_iter, double *tolerance,
                                  double *x) {
  (void)ilu;
  sp_matrix_yale_solve_cg(m, b, x0, max_iter, tolerance, x);
}

int sp_matrix_yale_chol_symbolic(sp_matrix_yale_ptr m, sp_chol_symbolic_ptr s) {
  (void)m; s->valid = 1; return 1;
}
int sp_matrix_yale_chol_symbolic_solve(sp_matrix_yale_ptr m,
                                       sp_chol_symbolic_ptr s, double *b,
                                       double *x) {
  (void)s;
  standin_pcg(m, b, x, 1e-15, 200000);
  trace_push(m->rows_count, b, x);
  return 1;
}

/* ------------------------------------------------------------------ */
/* harness: flat-array API over the reference's structs                 */

typedef struct {
  fea_solver_ptr solver;
} ref_ctx;

static void build_inputs(int n_nodes, const double *nodes, int n_elems,
                         const int *conn, int n_presc, const int *presc_node,
                         const int *presc_type, const double *presc_vals,
                         int model, double lambda, double mu, int gauss_count,
                         int load_increments, double desired_tol,
                         int modified_newton, int max_newton, int solver_type,
                         double solver_tol, int solver_max_iter,
                         fea_task_ptr *task, fea_solution_params_ptr *params,
                         nodes_array_ptr *na, elements_array_ptr *ea,
                         presc_bnd_array_ptr *pa) {
  int i, j;
  *task = fea_task_alloc();
  *params = fea_solution_params_alloc();
  *na = nodes_array_alloc();
  *ea = elements_array_alloc();
  *pa = presc_bnd_array_alloc();
  (*task)->model.model = model ? MODEL_COMPRESSIBLE_NEOHOOKEAN : MODEL_A5;
  (*task)->model.parameters_count = 2;
  (*task)->model.parameters[0] = lambda;
  (*task)->model.parameters[1] = mu;
  (*task)->load_increments_count = load_increments;
  (*task)->desired_tolerance = desired_tol;
  (*task)->modified_newton = modified_newton ? TRUE : FALSE;
  (*task)->max_newton_count = max_newton;
  (*task)->solver_type = (slae_solver_type)solver_type;
  (*task)->solver_tolerance = solver_tol;
  (*task)->solver_max_iter = solver_max_iter;
  (*params)->gauss_nodes_count = gauss_count;
  (*params)->nodes_per_element = 10;
  (*na)->nodes_count = n_nodes;
  (*na)->nodes = (real **)malloc(sizeof(real *) * (size_t)n_nodes);
  for (i = 0; i < n_nodes; ++i) {
    (*na)->nodes[i] = (real *)malloc(sizeof(real) * MAX_DOF);
    for (j = 0; j < 3; ++j) (*na)->nodes[i][j] = nodes[3 * i + j];
  }
  (*ea)->elements_count = n_elems;
  (*ea)->elements = (int **)malloc(sizeof(int *) * (size_t)n_elems);
  for (i = 0; i < n_elems; ++i) {
    (*ea)->elements[i] = (int *)malloc(sizeof(int) * 10);
    for (j = 0; j < 10; ++j) (*ea)->elements[i][j] = conn[10 * i + j];
  }
  (*pa)->prescribed_nodes_count = n_presc;
  (*pa)->prescribed_nodes = n_presc
      ? (prescribed_bnd_node *)malloc(sizeof(prescribed_bnd_node) * (size_t)n_presc)
      : NULL;
  for (i = 0; i < n_presc; ++i) {
    (*pa)->prescribed_nodes[i].node_number = presc_node[i];
    (*pa)->prescribed_nodes[i].type = (presc_boundary_type)presc_type[i];
    for (j = 0; j < 3; ++j)
      (*pa)->prescribed_nodes[i].values[j] = presc_vals[3 * i + j];
  }
}

ref_ctx *ref_create(int n_nodes, const double *nodes, int n_elems,
                    const int *conn, int n_presc, const int *presc_node,
                    const int *presc_type, const double *presc_vals, int model,
                    double lambda, double mu, int gauss_count) {
  fea_task_ptr task; fea_solution_params_ptr params; nodes_array_ptr na;
  elements_array_ptr ea; presc_bnd_array_ptr pa;
  ref_ctx *c = (ref_ctx *)calloc(1, sizeof(ref_ctx));
  build_inputs(n_nodes, nodes, n_elems, conn, n_presc, presc_node, presc_type,
               presc_vals, model, lambda, mu, gauss_count, 1, 1e-8, 1, 1,
               CHOLESKY, 1e-14, 20000, &task, &params, &na, &ea, &pa);
  c->solver = fea_solver_alloc(task, params, na, ea, pa);
  solver_create_element_database(c->solver);
  return c;
}

void ref_destroy(ref_ctx *c) {
  if (!c) return;
  if (c->solver->symb_chol) free(c->solver->symb_chol);
  c->solver->current_load_step = 0; /* no stored load steps in harness mode */
  fea_solver_free(c->solver);
  free(c);
}

void ref_set_nodes(ref_ctx *c, const double *x) {
  int i, j;
  for (i = 0; i < c->solver->nodes_p->nodes_count; ++i)
    for (j = 0; j < 3; ++j) c->solver->nodes_p->nodes[i][j] = x[3 * i + j];
}
void ref_get_nodes(ref_ctx *c, double *x) {
  int i, j;
  for (i = 0; i < c->solver->nodes_p->nodes_count; ++i)
    for (j = 0; j < 3; ++j) x[3 * i + j] = c->solver->nodes_p->nodes[i][j];
}
void ref_apply_increment(ref_ctx *c, double lambda) {
  solver_update_nodes_with_bc(c->solver, lambda);
}
void ref_update_state(ref_ctx *c) {
  solver_create_current_shape_gradients(c->solver);
  solver_create_stresses(c->solver);
}
void ref_shape_gradients_only(ref_ctx *c) {
  solver_create_current_shape_gradients(c->solver);
}
void ref_stresses_only(ref_ctx *c) { solver_create_stresses(c->solver); }

void ref_get_state(ref_ctx *c, double *F, double *S) {
  int e, g, ne = c->solver->elements_p->elements_count;
  int ng = c->solver->fea_params_p->gauss_nodes_count;
  for (e = 0; e < ne; ++e)
    for (g = 0; g < ng; ++g) {
      memcpy(F + ((size_t)e * ng + g) * 9, c->solver->graddefs[e][g].components,
             sizeof(double) * 9);
      memcpy(S + ((size_t)e * ng + g) * 9, c->solver->stresses[e][g].components,
             sizeof(double) * 9);
    }
}

/* grads[e][g][3][10], detJ[e][g]; a missing (singular) entry is NaN-filled */
void ref_get_gradients(ref_ctx *c, double *grads, double *detJ) {
  int e, g, i, k, ne = c->solver->elements_p->elements_count;
  int ng = c->solver->fea_params_p->gauss_nodes_count;
  for (e = 0; e < ne; ++e)
    for (g = 0; g < ng; ++g) {
      shape_gradients_ptr s = c->solver->shape_gradients[e][g];
      double *dst = grads + ((size_t)e * ng + g) * 30;
      for (i = 0; i < 3; ++i)
        for (k = 0; k < 10; ++k) dst[i * 10 + k] = s ? s->grads[i][k] : NAN;
      detJ[(size_t)e * ng + g] = s ? s->detJ : NAN;
    }
}

void ref_set_scatter_mode(int mode) { g_scatter_mode = mode; }
long ref_add_calls(void) { return g_add_calls; }

void ref_stiffness(ref_ctx *c) { solver_create_stiffness(c->solver); }
void ref_residual(ref_ctx *c) { solver_create_residual_forces(c->solver); }
void ref_apply_bc(ref_ctx *c, double lambda) {
  solver_apply_prescribed_bc(c->solver, lambda);
}
void ref_get_forces(ref_ctx *c, double *R) {
  memcpy(R, c->solver->global_forces_vct,
         sizeof(double) * (size_t)c->solver->global_mtx.rows_count);
}
void ref_set_forces(ref_ctx *c, const double *R) {
  memcpy(c->solver->global_forces_vct, R,
         sizeof(double) * (size_t)c->solver->global_mtx.rows_count);
}
void ref_get_solution(ref_ctx *c, double *u) {
  memcpy(u, c->solver->global_solution_vct,
         sizeof(double) * (size_t)c->solver->global_mtx.rows_count);
}
int ref_solve_slae(ref_ctx *c) {
  solver_solve_slae(c->solver);
  return g_last_iters;
}
void ref_update_with_solution(ref_ctx *c) {
  solver_update_nodes_with_solution(c->solver, c->solver->global_solution_vct);
}

long ref_nnz(ref_ctx *c) {
  long nnz = 0; int j;
  for (j = 0; j < c->solver->global_mtx.cols_count; ++j)
    nnz += c->solver->global_mtx.storage[j].last_index + 1;
  return nnz;
}
/* sorted CSR (row-major) of the stand-in container */
void ref_get_csr(ref_ctx *c, int *rowptr, int *colidx, double *vals) {
  sp_matrix_yale y;
  sp_matrix_yale_init(&y, &c->solver->global_mtx);
  memcpy(rowptr, y.offsets, sizeof(int) * ((size_t)y.rows_count + 1));
  memcpy(colidx, y.indexes, sizeof(int) * (size_t)y.nonzeros);
  memcpy(vals, y.values, sizeof(double) * (size_t)y.nonzeros);
  sp_matrix_yale_free(&y);
}

/* dense 30x30 K_e (constitutive + initial stress) of one element, captured at
 * the reference's sp_matrix_element_add call sites (fea_solver.c:966,1055) */
void ref_element_matrix(ref_ctx *c, int element, double *ke, int part) {
  memset(ke, 0, sizeof(double) * 900);
  g_cap_ke = ke;
  g_cap_conn = c->solver->elements_p->elements[element];
  if (part == 0 || part == 1) solver_local_constitutive_part(c->solver, element);
  if (part == 0 || part == 2) solver_local_initial_stess_part(c->solver, element);
  g_cap_ke = NULL;
  g_cap_conn = NULL;
}

/* material hooks on a bare F (fea_model.c:26,79,110,129) */
void ref_model_eval(int model, double lambda, double mu, const double *F,
                    double *sigma, double *ctens) {
  fea_model m;
  real Fm[3][3], S[3][3], C[3][3][3][3];
  m.model = model ? MODEL_COMPRESSIBLE_NEOHOOKEAN : MODEL_A5;
  m.parameters_count = 2;
  m.parameters[0] = lambda;
  m.parameters[1] = mu;
  fea_model_init(&m, m.model);
  memcpy(Fm, F, sizeof(Fm));
  m.stress(&m, Fm, S);
  m.ctensor(&m, Fm, C);
  memcpy(sigma, S, sizeof(S));
  memcpy(ctens, C, sizeof(C));
}

/* dense 3x3 helpers (dense_matrix.c) for the tests.c known-answer vectors */
void ref_matmul(int which, const double *A, const double *B, double *R) {
  real a[3][3], b[3][3], r[3][3];
  memcpy(a, A, sizeof(a)); memcpy(b, B, sizeof(b));
  if (which == 0) matrix_mul3x3(a, b, r);
  else if (which == 1) matrix_transpose_mul3x3(a, b, r);
  else matrix_transpose2_mul3x3(a, b, r);
  memcpy(R, r, sizeof(r));
}
int ref_inv3x3(double *M, double *det) {
  real m[3][3]; int ok;
  memcpy(m, M, sizeof(m));
  ok = inv3x3(m, det);
  memcpy(M, m, sizeof(m));
  return ok;
}
int ref_do_tests(void) { return do_tests(); }

void ref_gauss_table(int count, double *out /* [count][4] */) {
  extern real gauss_nodes4_tetr10[4][4];
  extern real gauss_nodes5_tetr10[5][4];
  memcpy(out, count == 4 ? (void *)gauss_nodes4_tetr10 : (void *)gauss_nodes5_tetr10,
         sizeof(double) * 4 * (size_t)count);
}
void ref_shape_tables(double r, double s, double t, double *N, double *dN) {
  int a, d;
  for (a = 0; a < 10; ++a) {
    N[a] = tetrahedra10_isoform(a, r, s, t);
    for (d = 0; d < 3; ++d) dN[d * 10 + a] = tetrahedra10_disoform(a, d, r, s, t);
  }
}

/* ---- the reference's full Newton driver solve() (fea_solver.c:130-242) ---- */

void ref_trace_reset(void) {
  int i;
  for (i = 0; i < g_trace_count; ++i) { free(g_trace[i].rhs); free(g_trace[i].sol); }
  g_trace_count = 0;
  g_tol_count = 0;
}
int ref_trace_count(void) { return g_trace_count; }
void ref_trace_get(int k, double *rhs, double *sol) {
  memcpy(rhs, g_trace[k].rhs, sizeof(double) * (size_t)g_trace[k].n);
  memcpy(sol, g_trace[k].sol, sizeof(double) * (size_t)g_trace[k].n);
}
int ref_tolerance_count(void) { return g_tol_count; }
void ref_tolerance_get(double *out) {
  memcpy(out, g_tol_trace, sizeof(double) * (size_t)g_tol_count);
}
void ref_set_log_echo(int on) { g_log_echo = on; }

/* Runs solve() on fresh copies (solve() takes ownership and frees them) and
 * leaves the Gmsh export at `msh_path`.  Returns the number of linear solves. */
int ref_run_solve(int n_nodes, const double *nodes, int n_elems, const int *conn,
                  int n_presc, const int *presc_node, const int *presc_type,
                  const double *presc_vals, int model, double lambda, double mu,
                  int gauss_count, int load_increments, double desired_tol,
                  int modified_newton, int max_newton, int solver_type,
                  double solver_tol, int solver_max_iter, const char *msh_path) {
  fea_task_ptr task; fea_solution_params_ptr params; nodes_array_ptr na;
  elements_array_ptr ea; presc_bnd_array_ptr pa;
  build_inputs(n_nodes, nodes, n_elems, conn, n_presc, presc_node, presc_type,
               presc_vals, model, lambda, mu, gauss_count, load_increments,
               desired_tol, modified_newton, max_newton, solver_type,
               solver_tol, solver_max_iter, &task, &params, &na, &ea, &pa);
  task->export_file = strdup(msh_path);
  ref_trace_reset();
  g_trace_on = 1;
  solve(task, params, na, ea, pa);
  g_trace_on = 0;
  return g_trace_count;
}
