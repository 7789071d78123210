/* Host driver of the B200 build: the reference's solver-large/fea_solver.c call structure
 * (main -> do_main -> solve, the phase functions, the Gmsh exporter) with every numerical
 * phase forwarded to the CUDA library through include/fea_gpu.h.  No element or matrix
 * arithmetic happens on the host; if the GPU library fails the process stops with the
 * reference's error() convention.
 */
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "fea_solver.h"
#include "sexp_loader.h"
#include "tests.h"

/* ---------------------------------------------------------------------------------------
 * logging: the reference writes through liblogger (S-expression file + stdout echo,
 * fea_solver.c:70-94); same message texts here so logs can be diffed, plain stdout/file */

static FILE *g_log_file = NULL;

static void log_line(const char *level, const char *fmt, ...) {
  va_list ap;
  char buf[512];
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  printf("%s\n", buf);
  if (g_log_file) {
    fprintf(g_log_file, "(%s \"%s\")\n", level, buf);
    fflush(g_log_file);
  }
}
#define LOG(...) log_line("log", __VA_ARGS__)
#define LOGINFO(...) log_line("info", __VA_ARGS__)
#define LOGERROR(...) log_line("error", __VA_ARGS__)

void error(char *msg) {
  LOGERROR("feasolve error encountered: %s", msg);
  exit(EXIT_FAILURE);
}

/* GPU call that must succeed: fatal otherwise (there is nothing to fall back to) */
static void gpu_must(int rc, const char *what) {
  if (rc != FEA_GPU_OK) {
    char buf[400];
    snprintf(buf, sizeof(buf), "%s failed (%d): %s", what, rc, fea_gpu_last_error());
    error(buf);
  }
}

/* ---------------------------------------------------------------------------------------
 * process level */

int parse_cmdargs(int argc, char **argv, char **filename) {
  if (argc < 2) {
    printf("Usage: fea_solve input_data.sexp\n");
    return 1;
  }
  *filename = argv[1];
  return 0;
}

int main(int argc, char **argv) {
  char *filename = NULL;
  char logname[300];
  int result;
  if (!do_tests()) {                     /* fea_solver.c:76-80 */
    fprintf(stderr, "Error! Tests failed!\n");
    return 1;
  }
  if (parse_cmdargs(argc, argv, &filename)) return 1;
  snprintf(logname, sizeof(logname), "%s.log", argv[0]);
  g_log_file = fopen(logname, "w");
  result = do_main(filename);
  if (g_log_file) fclose(g_log_file);
  g_log_file = NULL;
  return result;
}

BOOL initial_data_load(char *filename, fea_task_ptr *task, fea_solution_params_ptr *fea_params,
                       nodes_array_ptr *nodes, elements_array_ptr *elements,
                       presc_bnd_array_ptr *presc_boundary) {
  const char *dot = strrchr(filename, '.');
  BOOL ok = FALSE;
  if (dot && strcasecmp(dot + 1, "sexp") == 0)   /* format by extension, fea_solver.c:1672-1680 */
    ok = sexp_data_load(filename, task, fea_params, nodes, elements, presc_boundary);
  if (ok && *task) {
    /* <basename>.msh beside the input; sized properly (the reference's malloc(strlen) at
     * :1683 only works because ".sexp" is longer than ".msh") */
    const size_t stem = (size_t)(dot - filename);
    char *out = (char *)malloc(stem + 5);
    memcpy(out, filename, stem);
    memcpy(out + stem, ".msh", 5);
    (*task)->export_file = out;
  }
  return ok;
}

int do_main(char *filename) {
  fea_task_ptr task = NULL;
  fea_solution_params_ptr fea_params = NULL;
  nodes_array_ptr nodes = NULL;
  elements_array_ptr elements = NULL;
  presc_bnd_array_ptr presc = NULL;
  if (!initial_data_load(filename, &task, &fea_params, &nodes, &elements, &presc)) {
    LOGERROR("Error. Unable to load %s.", filename);
    return 1;
  }
  LOG("Initial data loaded");
  solve(task, fea_params, nodes, elements, presc);
  return 0;
}

/* ---------------------------------------------------------------------------------------
 * the load-increment / Newton driver, fea_solver.c:130-242.
 *
 * Same control flow and stop rule; the element passes the reference makes one after another
 * on the same coordinates (stresses at the end of an iteration :217-218, residual :185 and,
 * for full Newton, stiffness :200 at the start of the next) are one fused device pass here. */

/* Golden-section search of the step length along the Newton direction, as the reference's prototype does
 * when :line-search :max is positive (solver-prototype/cartesian3d/large/cartesian3d_large.m:85-119):
 * minimise f(eta) = |eta <u, R(x + eta u)>| on [0.5, 1]; if both probes are worse than the plain Newton
 * value |<u, R(x)>| the full step is kept.  The shipped solver-large parses the key and never reads it
 * (fea_solver.h:105, sexp_loader.c:153-167); all shipped models say 0, which leaves solve() as it was. */
static real solver_line_search(fea_solver_ptr solver, real tolerance, int max_iter) {
  const real tau = (sqrt(5.0) - 1.0) / 2.0;
  real a = 0.5, b = 1.0, eta = 1.0;
  int it;
  for (it = 0; it < max_iter; ++it) {
    const real x1 = b - tau * (b - a), x2 = a + tau * (b - a);
    real f1, f2;
    gpu_must(fea_gpu_restore_nodes(solver->gpu), "fea_gpu_restore_nodes");
    gpu_must(fea_gpu_update_nodes_scaled(solver->gpu, x1), "fea_gpu_update_nodes_scaled");
    gpu_must(fea_gpu_assemble_all(solver->gpu, FEA_ASSEMBLE_FUSE_BC), "fea_gpu_assemble_all");   /* R only, BC rows zero */
    gpu_must(fea_gpu_dot_R_u(solver->gpu, &f1), "fea_gpu_dot_R_u");
    f1 = fabs(x1 * f1);
    gpu_must(fea_gpu_restore_nodes(solver->gpu), "fea_gpu_restore_nodes");
    gpu_must(fea_gpu_update_nodes_scaled(solver->gpu, x2), "fea_gpu_update_nodes_scaled");
    gpu_must(fea_gpu_assemble_all(solver->gpu, FEA_ASSEMBLE_FUSE_BC), "fea_gpu_assemble_all");
    gpu_must(fea_gpu_dot_R_u(solver->gpu, &f2), "fea_gpu_dot_R_u");
    f2 = fabs(x2 * f2);
    LOG("Line search: f(%f) = %e, f(%f) = %e", x1, f1, x2, f2);
    if (f1 > f2) a = x1; else b = x2;
    if (fabs(tolerance) < f1 && fabs(tolerance) < f2) {
      eta = 1.0;
      break;
    }
    eta = (x1 + x2) / 2.0;
  }
  gpu_must(fea_gpu_restore_nodes(solver->gpu), "fea_gpu_restore_nodes");
  return eta;
}

void solve(fea_task_ptr task, fea_solution_params_ptr fea_params, nodes_array_ptr nodes,
           elements_array_ptr elements, presc_bnd_array_ptr presc_boundary) {
  fea_solver_ptr solver = fea_solver_alloc(task, fea_params, nodes, elements, presc_boundary);
  int it = 0;
  real tolerance;
  const char *keep = getenv("FEA_KEEP_STEPS");
  /* FEA_KEEP_STEPS=0: no host snapshot per increment (9.5 GB each at 50 M DOF); only the last state is pulled
   * for the exporter.  Default: every increment, as the reference (fea_solver.c:233, :605-636). */
  const BOOL keep_steps = !(keep && atoi(keep) == 0);
  /* FEA_PREDICTOR=1: from the second increment on, start from x_k + (x_k - x_{k-1}) instead of moving the boundary
   * nodes alone (:168).  The prescribed nodes still move by exactly one increment (they moved by one in the previous
   * one), the interior gets a secant guess: the same equilibrium in ~2 instead of ~5 Newton iterations
   * (profiles/r2_large_runs.md).  Off by default: the reference's iterates are the default. */
  const char *pred = getenv("FEA_PREDICTOR");
  const BOOL predictor = pred && atoi(pred) == 1 && task->linesearch_max == 0;
  BOOL prev_whole = FALSE;       /* the previous increment went in one part: the saved nodes are its start */
  LOG("Create elements database");
  solver_create_element_database(solver);
  LOG("Create an array of shape functions gradients in initial configuration");
  solver_create_initial_shape_gradients(solver);

  for (; solver->current_load_step < task->load_increments_count; ++solver->current_load_step) {
    /* The reference applies the whole increment at once (lambda = 1, :168).  If that inverts elements
     * (det J <= 0 or det F <= 0 right after the boundary nodes moved: increments larger than an element)
     * its log(J) turns NaN; here the increment is rolled back and applied in halves of what remains,
     * each part equilibrated by the same Newton loop.  solver_update_nodes_with_bc and
     * solver_apply_prescribed_bc already take the load fraction (fea_solver.c:573, :610). */
    real remaining = 1.0, part = 1.0;
    BOOL failed = FALSE;
    while (remaining > 0.0 && !failed) {
      int64_t bad = 0;
      BOOL predicted = FALSE;
      if (part > remaining) part = remaining;
      it = 0;
      if (predictor && prev_whole && part == 1.0) {
        gpu_must(fea_gpu_extrapolate_nodes(solver->gpu, 1.0), "fea_gpu_extrapolate_nodes");   /* saved <- x_k */
        predicted = TRUE;
      } else {
        gpu_must(fea_gpu_save_nodes(solver->gpu), "fea_gpu_save_nodes");
        solver_update_nodes_with_bc(solver, part);               /* full value every increment, :168 */
      }
      /* :171-179: gradients, stresses, K, keep K for modified Newton; the residual of the first
       * iteration (:185) comes out of the same element pass */
      gpu_must(fea_gpu_assemble_all(solver->gpu, 1), "fea_gpu_assemble_all");
      gpu_must(fea_gpu_bad_points(solver->gpu, &bad), "fea_gpu_bad_points");
      if (bad > 0 && predicted) {                                /* the guess folded elements: plain boundary move */
        LOGERROR("Load increment %d: %ld Gauss points inverted by the predictor, moving the boundary alone",
                 solver->current_load_step + 1, (long)bad);
        gpu_must(fea_gpu_restore_nodes(solver->gpu), "fea_gpu_restore_nodes");
        prev_whole = FALSE;
        continue;
      }
      if (bad > 0 && part > 1.0 / 64.0) {
        LOGERROR("Load increment %d: %ld Gauss points inverted by a load fraction of %g, halving it",
                 solver->current_load_step + 1, (long)bad, part);
        gpu_must(fea_gpu_restore_nodes(solver->gpu), "fea_gpu_restore_nodes");
        part *= 0.5;
        continue;
      }
      gpu_must(fea_gpu_save_stiffness(solver->gpu), "fea_gpu_save_stiffness");
      do {
        real eta = 1.0;
        it++;
        if (it > 1) {
          /* state of the updated nodes + residual (+ K for full Newton) in one pass */
          gpu_must(fea_gpu_assemble_all(solver->gpu, task->modified_newton ? 0 : 1), "fea_gpu_assemble_all");
          if (task->modified_newton) gpu_must(fea_gpu_restore_stiffness(solver->gpu), "fea_gpu_restore_stiffness");
        }
        solver_apply_prescribed_bc(solver, 0);                   /* :203 */
        solver_solve_slae(solver);                               /* :205 */
        gpu_must(fea_gpu_dot_R_u(solver->gpu, &tolerance), "fea_gpu_dot_R_u");   /* :208 */
        LOG("Tolerance <X,R> = %e", tolerance);
        if (task->linesearch_max > 0) {
          gpu_must(fea_gpu_save_nodes(solver->gpu), "fea_gpu_save_nodes");
          eta = solver_line_search(solver, tolerance, task->linesearch_max);
          LOG("Line search: eta = %f", eta);
        }
        LOG("Newton iteration %d finished", it);
        if (eta == 1.0)
          solver_update_nodes_with_solution(solver, NULL);       /* :216 */
        else
          gpu_must(fea_gpu_update_nodes_scaled(solver->gpu, eta), "fea_gpu_update_nodes_scaled");
      } while (fabs(tolerance) > task->desired_tolerance && it < task->max_newton_count);
      if (it == task->max_newton_count) failed = TRUE;           /* the reference's test is on the count alone, :225 */
      prev_whole = part == 1.0;
      remaining -= part;
      if (remaining > 0.0 && !failed)
        LOG("Load increment %d: fraction %g done, %g to go", solver->current_load_step + 1, part, remaining);
    }
    /* the reference recomputes gradients and stresses after the last update (:217-218) */
    solver_create_stresses(solver);
    LOG("Load increment %d finished", solver->current_load_step + 1);
    if (failed) {                                                /* :225-231 */
      solver->current_load_step--;
      LOGERROR("Unable to finish load step in %d Newton iterations,exit", task->max_newton_count);
      break;
    }
    if (keep_steps || solver->current_load_step + 1 == task->load_increments_count)
      solver_load_step_init(solver, &solver->load_steps_p[solver->current_load_step], solver->current_load_step);
  }
  LOG("Exporting data...");
  solver->export_function(solver, task->export_file);
  fea_solver_free(solver);
}

/* ---------------------------------------------------------------------------------------
 * allocators */

fea_task_ptr fea_task_alloc(void) {   /* defaults of fea_solver.c:1509-1529 */
  fea_task_ptr t = (fea_task_ptr)calloc(1, sizeof(fea_task));
  t->desired_tolerance = 1e-8;
  t->dof = 3;
  t->ele_type = TETRAHEDRA10;
  t->type = CARTESIAN3D;
  t->modified_newton = TRUE;
  t->model.model = MODEL_A5;
  t->model.parameters_count = 2;
  t->model.parameters[0] = 100;
  t->model.parameters[1] = 100;
  /* the reference leaves these three uninitialised without an slae-solver form;
   * defined defaults here */
  t->solver_type = CG;
  t->solver_tolerance = MAX_ITERATIVE_TOLERANCE;
  t->solver_max_iter = MAX_ITERATIVE_ITERATIONS;
  return t;
}
fea_task_ptr fea_task_free(fea_task_ptr t) {
  if (t) {
    free((void *)t->export_file);
    free(t);
  }
  return NULL;
}
fea_solution_params_ptr fea_solution_params_alloc(void) {
  fea_solution_params_ptr p = (fea_solution_params_ptr)malloc(sizeof(fea_solution_params));
  p->gauss_nodes_count = 5;
  p->nodes_per_element = 10;
  return p;
}
fea_solution_params_ptr fea_solution_params_free(fea_solution_params_ptr p) {
  free(p);
  return NULL;
}

nodes_array_ptr nodes_array_alloc(void) { return (nodes_array_ptr)calloc(1, sizeof(nodes_array)); }

BOOL nodes_array_reserve(nodes_array_ptr a, int count) {
  int i;
  real *flat;
  if (!a || count <= 0) return FALSE;
  flat = (real *)calloc((size_t)count * MAX_DOF, sizeof(real));
  a->nodes = (real **)malloc(sizeof(real *) * (size_t)count);
  if (!flat || !a->nodes) return FALSE;
  for (i = 0; i < count; ++i) a->nodes[i] = flat + (size_t)i * MAX_DOF;
  a->nodes_count = count;
  return TRUE;
}
nodes_array_ptr nodes_array_copy_alloc(nodes_array_ptr src) {
  nodes_array_ptr c = nodes_array_alloc();
  if (src && src->nodes_count && src->nodes && nodes_array_reserve(c, src->nodes_count))
    memcpy(c->nodes[0], src->nodes[0], sizeof(real) * MAX_DOF * (size_t)src->nodes_count);
  return c;
}
nodes_array_ptr nodes_array_free(nodes_array_ptr a) {
  if (a) {
    if (a->nodes) free(a->nodes[0]);
    free(a->nodes);
    free(a);
  }
  return NULL;
}

elements_array_ptr elements_array_alloc(void) { return (elements_array_ptr)calloc(1, sizeof(elements_array)); }

BOOL elements_array_reserve(elements_array_ptr a, int count, int npe) {
  int i;
  int *flat;
  if (!a || count <= 0 || npe <= 0) return FALSE;
  flat = (int *)calloc((size_t)count * (size_t)npe, sizeof(int));
  a->elements = (int **)malloc(sizeof(int *) * (size_t)count);
  if (!flat || !a->elements) return FALSE;
  for (i = 0; i < count; ++i) a->elements[i] = flat + (size_t)i * npe;
  a->elements_count = count;
  return TRUE;
}
elements_array_ptr elements_array_free(elements_array_ptr a) {
  if (a) {
    if (a->elements) free(a->elements[0]);
    free(a->elements);
    free(a);
  }
  return NULL;
}

presc_bnd_array_ptr presc_bnd_array_alloc(void) { return (presc_bnd_array_ptr)calloc(1, sizeof(presc_bnd_array)); }
presc_bnd_array_ptr presc_bnd_array_free(presc_bnd_array_ptr p) {
  if (p) {
    free(p->prescribed_nodes);
    free(p);
  }
  return NULL;
}

static tensor **tensor_table_alloc(int elnum, int gauss) {
  tensor *flat = (tensor *)calloc((size_t)elnum * (size_t)gauss, sizeof(tensor));   /* zeros, :434-435 */
  tensor **rows = (tensor **)malloc(sizeof(tensor *) * (size_t)elnum);
  int e;
  for (e = 0; e < elnum; ++e) rows[e] = flat + (size_t)e * gauss;
  return rows;
}
static void tensor_table_free(tensor **rows) {
  if (rows) {
    free(rows[0]);
    free(rows);
  }
}

/* ---------------------------------------------------------------------------------------
 * solver object */

void solver_create_element_params(fea_solver_ptr self) {
  if (self->task_p->ele_type != TETRAHEDRA10) error("Error: unknown element type");            /* :600 */
  if (self->fea_params_p->gauss_nodes_count != 4 && self->fea_params_p->gauss_nodes_count != 5)
    error("solver_create_element_params_tetrahedra10: gauss nodes");                           /* :1503 */
  if (self->fea_params_p->nodes_per_element != 10) error("TETRAHEDRA10 needs :nodes-count 10");
  self->export_function = solver_export_tetrahedra10_gmsh;
}

fea_solver_ptr fea_solver_alloc(fea_task_ptr task, fea_solution_params_ptr fea_params, nodes_array_ptr nodes,
                                elements_array_ptr elements, presc_bnd_array_ptr presc) {
  fea_solver_ptr s = (fea_solver_ptr)calloc(1, sizeof(fea_solver));
  const int elnum = elements->elements_count, ng = fea_params->gauss_nodes_count, np = presc->prescribed_nodes_count;
  int *pnode = (int *)malloc(sizeof(int) * (size_t)(np + 1)), *ptype = (int *)malloc(sizeof(int) * (size_t)(np + 1));
  real *pval = (real *)malloc(sizeof(real) * 3 * (size_t)(np + 1));
  const char *dev = getenv("FEA_GPU_DEVICE"), *cnt = getenv("FEA_GPU_COUNT");
  const int n_gpus = cnt && atoi(cnt) > 1 ? atoi(cnt) : 1;   /* FEA_GPU_COUNT=8: this process drives the whole box */
  int i, d, rc;
  s->task_p = task;
  s->fea_params_p = fea_params;
  s->nodes0_p = nodes;
  s->nodes_p = nodes_array_copy_alloc(nodes);                                                   /* :400 */
  s->elements_p = elements;
  s->presc_boundary_p = presc;
  solver_create_element_params(s);
  fea_model_init(&task->model, task->model.model);
  s->graddefs = tensor_table_alloc(elnum, ng);
  s->stresses = tensor_table_alloc(elnum, ng);
  s->current_load_step = 0;
  s->load_steps_p = (load_step_ptr)calloc((size_t)(task->load_increments_count > 0 ? task->load_increments_count : 1),
                                          sizeof(load_step));
  s->global_mtx.rows_count = nodes->nodes_count * task->dof;
  s->global_forces_vct = (real *)calloc((size_t)s->global_mtx.rows_count, sizeof(real));
  s->global_solution_vct = (real *)calloc((size_t)s->global_mtx.rows_count, sizeof(real));
  for (i = 0; i < np; ++i) {
    pnode[i] = presc->prescribed_nodes[i].node_number;
    ptype[i] = (int)presc->prescribed_nodes[i].type;
    for (d = 0; d < 3; ++d) pval[3 * i + d] = presc->prescribed_nodes[i].values[d];
  }
  if (n_gpus > 1)
    rc = fea_gpu_create_multi(&s->gpu, nodes->nodes_count, elnum, nodes->nodes[0], elements->elements[0],
                              task->model.model == MODEL_A5 ? FEA_MODEL_A5 : FEA_MODEL_COMPRESSIBLE_NEOHOOKEAN,
                              task->model.parameters[0], task->model.parameters[1], ng, np, pnode, ptype, pval,
                              n_gpus, NULL);
  else
    rc = fea_gpu_create(&s->gpu, nodes->nodes_count, elnum, nodes->nodes[0], elements->elements[0],
                        task->model.model == MODEL_A5 ? FEA_MODEL_A5 : FEA_MODEL_COMPRESSIBLE_NEOHOOKEAN,
                        task->model.parameters[0], task->model.parameters[1], ng, np, pnode, ptype, pval,
                        0, 1, NULL, dev ? atoi(dev) : 0);
  free(pnode);
  free(ptype);
  free(pval);
  gpu_must(rc, "fea_gpu_create");
  return s;
}

fea_solver_ptr fea_solver_free(fea_solver_ptr s) {
  int i;
  if (!s) return NULL;
  for (i = 0; i < s->current_load_step; ++i) solver_load_step_free(s, &s->load_steps_p[i]);   /* :485-486 */
  free(s->load_steps_p);
  tensor_table_free(s->graddefs);
  tensor_table_free(s->stresses);
  fea_gpu_destroy(s->gpu);
  fea_task_free(s->task_p);
  fea_solution_params_free(s->fea_params_p);
  nodes_array_free(s->nodes0_p);
  nodes_array_free(s->nodes_p);
  elements_array_free(s->elements_p);
  presc_bnd_array_free(s->presc_boundary_p);
  free(s->global_forces_vct);
  free(s->global_solution_vct);
  free(s);
  return NULL;
}

/* Gauss tables and shape-function derivatives are constants of the device kernels */
void solver_create_element_database(fea_solver_ptr self) { (void)self; }
void solver_free_element_database(fea_solver_ptr self) { (void)self; }
/* only needed without CURRENT_SHAPE_GRADIENTS, which the reference build defines (Makefile:12) */
void solver_create_initial_shape_gradients(fea_solver_ptr self) { (void)self; }
/* current-configuration gradients are recomputed inside every element pass, never stored */
void solver_create_current_shape_gradients(fea_solver_ptr self) { (void)self; }

void solver_create_stresses(fea_solver_ptr self) { gpu_must(fea_gpu_update_state(self->gpu), "fea_gpu_update_state"); }
void solver_create_residual_forces(fea_solver_ptr self) {
  gpu_must(fea_gpu_assemble_residual(self->gpu), "fea_gpu_assemble_residual");
}
void solver_create_stiffness(fea_solver_ptr self) {
  gpu_must(fea_gpu_assemble_stiffness(self->gpu), "fea_gpu_assemble_stiffness");
}
void solver_apply_prescribed_bc(fea_solver_ptr self, real lambda) {
  gpu_must(fea_gpu_apply_bc(self->gpu, lambda), "fea_gpu_apply_bc");
}
void solver_update_nodes_with_bc(fea_solver_ptr self, real lambda) {
  gpu_must(fea_gpu_apply_increment(self->gpu, lambda), "fea_gpu_apply_increment");
}
/* x == NULL or x == global_solution_vct: add the device solution; any other vector is
 * uploaded first (the reference takes an arbitrary x, fea_solver.c:1270) */
void solver_update_nodes_with_solution(fea_solver_ptr self, real *x) {
  if (x && x != self->global_solution_vct) {
    int i;
    solver_pull_state(self, FALSE);
    for (i = 0; i < self->global_mtx.rows_count; ++i) self->nodes_p->nodes[0][i] += x[i];
    solver_push_nodes(self);
    return;
  }
  gpu_must(fea_gpu_update_nodes(self->gpu), "fea_gpu_update_nodes");
}

/* solver_solve_slae (fea_solver.c:300-321): every solver type of the task file maps onto the
 * Jacobi-PCG; CHOLESKY (a direct solve in the reference) asks for the tightest tolerance */
BOOL solver_solve_slae(fea_solver_ptr solver) {
  const fea_task *t = solver->task_p;
  const real tol = t->solver_type == CHOLESKY ? MAX_ITERATIVE_TOLERANCE : t->solver_tolerance;
  const int max_iter = t->solver_type == CHOLESKY ? 10 * MAX_ITERATIVE_ITERATIONS : t->solver_max_iter;
  int32_t iters = 0;
  int rc;
  LOGINFO("Starting to solve SLAE");
  /* PCG_ILU (fea_solver.c:260-280) asks for a stronger preconditioner than plain CG: Chebyshev-accelerated
   * Jacobi on the device; CG and CHOLESKY run the north star's Jacobi-PCG */
  gpu_must(fea_gpu_set_param(solver->gpu, "precond", t->solver_type == PCG_ILU ? 1 : 0), "fea_gpu_set_param");
  rc = fea_gpu_solve(solver->gpu, tol, max_iter, FEA_SOLVE_X0_ZERO, &iters, &solver->last_linear_residual);
  solver->last_linear_iterations = iters;
  if (rc == FEA_GPU_ERR_STALLED)      /* u holds the best checkpointed iterate; the Newton loop goes on with it */
    LOGERROR("SLAE solver stopped on its stall/divergence guard after %d iterations, relative residual %e (asked %e)",
             (int)iters, solver->last_linear_residual, tol);
  else if (rc == FEA_GPU_ERR_NOT_CONVERGED)
    LOGERROR("SLAE solver reached %d iterations, relative residual %e (asked %e)", max_iter,
             solver->last_linear_residual, tol);
  else if (rc != FEA_GPU_OK)
    gpu_must(rc, "fea_gpu_solve");
  LOGINFO("SLAE solved: %d PCG iterations, relative residual %e", (int)iters, solver->last_linear_residual);
  return rc == FEA_GPU_OK;
}

void solver_push_nodes(fea_solver_ptr self) {
  gpu_must(fea_gpu_set_nodes(self->gpu, self->nodes_p->nodes[0]), "fea_gpu_set_nodes");
}
void solver_pull_state(fea_solver_ptr self, BOOL with_tensors) {
  gpu_must(fea_gpu_get_nodes(self->gpu, self->nodes_p->nodes[0]), "fea_gpu_get_nodes");
  gpu_must(fea_gpu_get_forces(self->gpu, self->global_forces_vct), "fea_gpu_get_forces");
  gpu_must(fea_gpu_get_solution(self->gpu, self->global_solution_vct), "fea_gpu_get_solution");
  if (with_tensors)
    gpu_must(fea_gpu_get_state(self->gpu, &self->graddefs[0][0].components[0][0],
                               &self->stresses[0][0].components[0][0]), "fea_gpu_get_state");
}

/* snapshot of a converged increment for the exporter (fea_solver.c:605-636) */
void solver_load_step_init(fea_solver_ptr self, load_step_ptr step, int step_number) {
  const int elnum = self->elements_p->elements_count, ng = self->fea_params_p->gauss_nodes_count;
  if (!step) return;
  solver_pull_state(self, TRUE);
  step->step_number = step_number;
  step->nodes_p = nodes_array_copy_alloc(self->nodes_p);
  step->graddefs = tensor_table_alloc(elnum, ng);
  step->stresses = tensor_table_alloc(elnum, ng);
  memcpy(step->graddefs[0], self->graddefs[0], sizeof(tensor) * (size_t)elnum * ng);
  memcpy(step->stresses[0], self->stresses[0], sizeof(tensor) * (size_t)elnum * ng);
}
void solver_load_step_free(fea_solver_ptr self, load_step_ptr step) {
  (void)self;
  if (!step) return;
  tensor_table_free(step->graddefs);
  tensor_table_free(step->stresses);
  nodes_array_free(step->nodes_p);
  step->graddefs = step->stresses = NULL;
  step->nodes_p = NULL;
}

/* ---------------------------------------------------------------------------------------
 * tet10 shape functions (reference node order, fea_solver.c:1287-1373); host tools only */

real tetrahedra10_isoform(int i, real r, real s, real t) {
  const real L[4] = {1 - r - s - t, r, s, t};   /* barycentric coordinates of vertices 0..3 */
  static const int edge[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
  if (i >= 0 && i < 4) return (2 * L[i] - 1) * L[i];
  if (i >= 4 && i < 10) return 4 * L[edge[i - 4][0]] * L[edge[i - 4][1]];
  error("tetrahedra10_isoform: wrong index");
  return 0;
}
real tetrahedra10_disoform(int shape, int dof, real r, real s, real t) {
  const real L[4] = {1 - r - s - t, r, s, t};
  real dL[4] = {-1, 0, 0, 0};                    /* d L_k / d xi_dof */
  static const int edge[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
  if (dof < 0 || dof > 2) error("tetrahedra10_disoform: wrong dof");
  dL[dof + 1] = 1;
  if (shape >= 0 && shape < 4) return (4 * L[shape] - 1) * dL[shape];
  if (shape >= 4 && shape < 10) {
    const int a = edge[shape - 4][0], b = edge[shape - 4][1];
    return 4 * (dL[a] * L[b] + L[a] * dL[b]);
  }
  error("tetrahedra10_disoform: wrong index");
  return 0;
}

/* ---------------------------------------------------------------------------------------
 * Gmsh 2.0 ASCII export: byte-compatible with solver_export_tetrahedra10_gmsh
 * (fea_solver.c:1375-1488): initial nodes, elements of type 11 with local nodes 8 and 9
 * swapped, then per stored increment the displacements and the stress of Gauss point 0. */

void solver_export_tetrahedra10_gmsh(fea_solver_ptr solver, const char *filename) {
  static const int gmsh_order[10] = {0, 1, 2, 3, 4, 5, 6, 7, 9, 8};
  const int nn = solver->nodes_p->nodes_count, ne = solver->elements_p->elements_count;
  FILE *f = fopen(filename, "w+");
  int i, j, k, load;
  if (!f) return;
  fprintf(f, "$MeshFormat\n2.0 0 8\n$EndMeshFormat\n");
  fprintf(f, "$Nodes\n%d\n", nn);
  for (i = 0; i < nn; ++i) {
    const real *X = solver->nodes0_p->nodes[i];
    fprintf(f, "%d %f %f %f\n", i + 1, X[0], X[1], X[2]);
  }
  fprintf(f, "$EndNodes\n$Elements\n%d\n", ne);
  for (i = 0; i < ne; ++i) {
    fprintf(f, "%d 11 3 1 1 1 ", i + 1);
    for (j = 0; j < 10; ++j) fprintf(f, "%d ", solver->elements_p->elements[i][gmsh_order[j]] + 1);
    fprintf(f, "\n");
  }
  fprintf(f, "$EndElements\n");
  for (load = 0; load <= solver->current_load_step; ++load) {
    const load_step *st = load ? &solver->load_steps_p[load - 1] : NULL;
    if (st && !st->nodes_p) continue;                          /* FEA_KEEP_STEPS=0: increment not kept */
    fprintf(f, "$NodeData\n1\n\"Displacements\"\n1\n%f\n3\n%d\n3\n%d\n", load * 0.83333333, load, nn);
    for (i = 0; i < nn; ++i) {
      real u[3] = {0.0, 0.0, 0.0};
      if (st)
        for (k = 0; k < 3; ++k) u[k] = st->nodes_p->nodes[i][k] - solver->nodes0_p->nodes[i][k];
      fprintf(f, "%d %f %f %f\n", i + 1, u[0], u[1], u[2]);
    }
    fprintf(f, "$EndNodeData\n");
    fprintf(f, "$ElementData\n1\n\"Stress tensor\"\n1\n%f\n3\n%d\n9\n%d\n", load * 0.83333333, load, ne);
    for (i = 0; i < ne; ++i) {
      fprintf(f, "%d ", i + 1);
      for (j = 0; j < MAX_DOF; ++j)
        for (k = 0; k < MAX_DOF; ++k) fprintf(f, "%f ", st ? st->stresses[i][0].components[j][k] : 0.0);
      fprintf(f, "\n");
    }
    fprintf(f, "$EndElementData\n");
  }
  fclose(f);
}
