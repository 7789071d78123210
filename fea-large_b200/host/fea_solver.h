/* Host layer of the B200 build: the reference's public solver API (solver-large/fea_solver.h)
 * over the CUDA C-ABI of include/fea_gpu.h.
 *
 * Same type and function names as the reference so that its callers (do_main, tools that
 * drive the phases one by one) compile against this header unchanged.  What differs, and why:
 *   - nodes_array / elements_array keep the `nodes[i][d]` / `elements[e][a]` access syntax but
 *     the rows point into ONE contiguous block (row 0 is its start): the reference's one malloc
 *     per node (sexp_loader.c:181) cannot survive 17 M nodes and the device wants flat arrays.
 *   - fea_solver embeds no libspmatrix types (fea_solver.h:276-277 of the reference); the
 *     global matrix lives on the GPU behind `gpu`, `global_mtx.rows_count` is kept because
 *     solve() reads it (fea_solver.c:210).
 *   - shape gradients are never materialised on the host (the reference mallocs 4 blocks per
 *     element and Gauss point, :700-707); graddefs / stresses are pulled from the device when a
 *     load step is stored.
 */
#ifndef FEA_B200_FEA_SOLVER_H
#define FEA_B200_FEA_SOLVER_H

#include <stdio.h>
#include "defines.h"
#include "dense_matrix.h"
#include "fea_model.h"
#include "fea_gpu.h"

#define MAX_ITERATIVE_TOLERANCE 1e-14   /* defaults of the iterative solvers (fea_solver.h:18-20) */
#define MAX_ITERATIVE_ITERATIONS 20000

typedef struct fea_solver_tag *fea_solver_ptr;
typedef void (*export_solution_t)(fea_solver_ptr, const char *filename);
typedef void (*apply_bc_t)(fea_solver_ptr self, int index, real arg);

typedef enum { CARTESIAN3D } task_type;
typedef enum { CG, PCG_ILU, CHOLESKY } slae_solver_type;
typedef enum { TETRAHEDRA10 } element_type;
typedef enum {
  FREE = 0, PRESCRIBEDX = 1, PRESCRIBEDY = 2, PRESCRIBEDXY = 3, PRESCRIBEDZ = 4,
  PRESCRIBEDXZ = 5, PRESCRIBEDYZ = 6, PRESCRIBEDXYZ = 7
} presc_boundary_type;

typedef struct {
  task_type type;
  fea_model model;
  slae_solver_type solver_type;
  real solver_tolerance;
  int solver_max_iter;
  int dof;
  element_type ele_type;
  int load_increments_count;
  real desired_tolerance;
  int max_newton_count;
  int linesearch_max;      /* parsed, unused -- as in the reference */
  int arclength_max;       /* parsed, unused -- as in the reference */
  BOOL modified_newton;
  const char *export_file;
} fea_task;
typedef fea_task *fea_task_ptr;

typedef struct {
  int nodes_per_element;
  int gauss_nodes_count;
} fea_solution_params;
typedef fea_solution_params *fea_solution_params_ptr;

typedef struct {
  int nodes_count;
  real **nodes;            /* nodes[i][d]; nodes[0] is a contiguous [nodes_count][3] block */
} nodes_array;
typedef nodes_array *nodes_array_ptr;

typedef struct {
  int elements_count;
  int **elements;          /* elements[e][a]; elements[0] is a contiguous [count][10] block */
} elements_array;
typedef elements_array *elements_array_ptr;

typedef struct {
  int node_number;
  real values[MAX_DOF];
  presc_boundary_type type;
} prescribed_bnd_node;
typedef prescribed_bnd_node *prescribed_bnd_node_ptr;

typedef struct {
  int prescribed_nodes_count;
  prescribed_bnd_node *prescribed_nodes;
} presc_bnd_array;
typedef presc_bnd_array *presc_bnd_array_ptr;

/* one stored load increment (fea_solver.h:212-224): what the exporter reads */
typedef struct {
  int step_number;
  nodes_array_ptr nodes_p;
  tensor **graddefs;       /* [elements][gauss]; row 0 is a contiguous block */
  tensor **stresses;
} load_step;
typedef load_step *load_step_ptr;

typedef struct { int rows_count; } fea_global_matrix;   /* the matrix itself is on the device */

typedef struct fea_solver_tag {
  export_solution_t export_function;
  fea_task_ptr task_p;
  fea_solution_params_ptr fea_params_p;
  nodes_array_ptr nodes0_p;
  nodes_array_ptr nodes_p;            /* host mirror of the current nodes, see solver_pull_state */
  elements_array_ptr elements_p;
  presc_bnd_array_ptr presc_boundary_p;
  tensor **graddefs;                  /* host mirrors, refreshed by solver_pull_state */
  tensor **stresses;
  int current_load_step;
  load_step_ptr load_steps_p;
  fea_global_matrix global_mtx;
  real *global_forces_vct;            /* host mirrors of R and u */
  real *global_solution_vct;
  fea_gpu_handle gpu;                 /* device side of everything above */
  int last_linear_iterations;
  real last_linear_residual;
} fea_solver;

/* ---- process level (fea_solver.c:57-128, 324) ---------------------------------------- */
void error(char *msg);
int parse_cmdargs(int argc, char **argv, char **filename);
int do_main(char *filename);
BOOL initial_data_load(char *filename, fea_task_ptr *task, fea_solution_params_ptr *fea_params,
                       nodes_array_ptr *nodes, elements_array_ptr *elements,
                       presc_bnd_array_ptr *presc_boundary);
/* takes ownership of its five arguments, exports `task->export_file`, frees everything */
void solve(fea_task_ptr task, fea_solution_params_ptr fea_params, nodes_array_ptr nodes,
           elements_array_ptr elements, presc_bnd_array_ptr presc_boundary);

/* ---- allocators (fea_solver.c:1509-1659) ---------------------------------------------- */
fea_task_ptr fea_task_alloc(void);
fea_task_ptr fea_task_free(fea_task_ptr task);
fea_solution_params_ptr fea_solution_params_alloc(void);
fea_solution_params_ptr fea_solution_params_free(fea_solution_params_ptr ptr);
nodes_array_ptr nodes_array_alloc(void);
nodes_array_ptr nodes_array_copy_alloc(nodes_array_ptr nodes);
nodes_array_ptr nodes_array_free(nodes_array_ptr nodes);
elements_array_ptr elements_array_alloc(void);
elements_array_ptr elements_array_free(elements_array_ptr elements);
presc_bnd_array_ptr presc_bnd_array_alloc(void);
presc_bnd_array_ptr presc_bnd_array_free(presc_bnd_array_ptr presc);
/* flat storage behind the row-pointer views (new; used by the loader) */
BOOL nodes_array_reserve(nodes_array_ptr nodes, int count);
BOOL elements_array_reserve(elements_array_ptr elements, int count, int nodes_per_element);

/* ---- solver object and phases (fea_solver.c:387-501, 556-571, 787-883, 1200-1284) ----- */
fea_solver_ptr fea_solver_alloc(fea_task_ptr task, fea_solution_params_ptr fea_params,
                                nodes_array_ptr nodes, elements_array_ptr elements,
                                presc_bnd_array_ptr presc);
fea_solver_ptr fea_solver_free(fea_solver_ptr solver);
void solver_create_element_params(fea_solver_ptr self);
void solver_create_element_database(fea_solver_ptr self);
void solver_free_element_database(fea_solver_ptr self);
void solver_create_initial_shape_gradients(fea_solver_ptr self);
void solver_create_current_shape_gradients(fea_solver_ptr self);
void solver_create_stresses(fea_solver_ptr self);
void solver_create_residual_forces(fea_solver_ptr self);
void solver_create_stiffness(fea_solver_ptr self);
void solver_apply_prescribed_bc(fea_solver_ptr self, real lambda);
void solver_update_nodes_with_bc(fea_solver_ptr self, real lambda);
void solver_update_nodes_with_solution(fea_solver_ptr self, real *x);
BOOL solver_solve_slae(fea_solver_ptr solver);
void solver_load_step_init(fea_solver_ptr self, load_step_ptr step, int step_number);
void solver_load_step_free(fea_solver_ptr self, load_step_ptr step);
void solver_export_tetrahedra10_gmsh(fea_solver_ptr solver, const char *filename);
/* device <-> host mirrors (new): call after touching nodes_p on the host / before reading
 * nodes_p, graddefs, stresses, global_forces_vct, global_solution_vct on the host */
void solver_push_nodes(fea_solver_ptr self);
void solver_pull_state(fea_solver_ptr self, BOOL with_tensors);

/* tet10 shape functions, kept for host-side tools (fea_solver.c:1287-1373) */
real tetrahedra10_isoform(int i, real r, real s, real t);
real tetrahedra10_disoform(int shape, int dof, real r, real s, real t);

#endif
