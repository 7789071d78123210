/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Declaration-level stand-in for the un-vendored libspmatrix
 * (github.com/fourier/libspmatrix, no version pinned by the reference:
 * solver-large/Makefile:8,14-15).  Only what the reference's fea_solver.c
 * touches is declared; the semantics are derived from its call sites
 * (fea_solver.c:179,194,195,223,251,268-278,288-294,304,319,448,496,877,966,
 * 1055,1250-1255).  The implementation lives in ref_standin.c.
 *
 * Parity note: libspmatrix source is absent, so storage order and the linear
 * solvers are "parity unpinned" (see DESIGN.md); the element arithmetic that
 * flows through sp_matrix_element_add is the reference's own compiled code.
 */
#ifndef ORACLE_STUB_SP_MATRIX_H
#define ORACLE_STUB_SP_MATRIX_H

typedef enum { CRS = 0, CCS = 1 } sparse_storage_type;

/* one compressed column (CCS) or row (CRS): fields used at fea_solver.c:1250-1252 */
typedef struct {
  int width;        /* allocated entries */
  int last_index;   /* index of the last stored entry, -1 when empty */
  int *indexes;     /* kept sorted ascending by the stand-in */
  double *values;
} indexed_array;

typedef struct {
  int rows_count;
  int cols_count;
  indexed_array *storage;
  int ordered;
  sparse_storage_type storage_type;
} sp_matrix;
typedef sp_matrix *sp_matrix_ptr;

/* compressed ("Yale") form produced per linear solve, fea_solver.c:304 */
typedef struct {
  sparse_storage_type storage_type;
  int rows_count;
  int cols_count;
  int nonzeros;
  int *offsets;
  int *indexes;
  double *values;
} sp_matrix_yale;
typedef sp_matrix_yale *sp_matrix_yale_ptr;

typedef struct { int dummy; } sp_matrix_skyline_ilu;
typedef sp_matrix_skyline_ilu *sp_matrix_skyline_ilu_ptr;

void sp_matrix_init(sp_matrix_ptr m, int rows, int cols, int bandwidth,
                    sparse_storage_type type);
sp_matrix_ptr sp_matrix_free(sp_matrix_ptr m);
void sp_matrix_clear(sp_matrix_ptr m);
void sp_matrix_copy(sp_matrix_ptr src, sp_matrix_ptr dst);
double sp_matrix_element_add(sp_matrix_ptr m, int i, int j, double value);
double sp_matrix_cross_cancellation(sp_matrix_ptr m, int index);

void sp_matrix_yale_init(sp_matrix_yale_ptr y, sp_matrix_ptr m);
void sp_matrix_yale_free(sp_matrix_yale_ptr y);

void sp_matrix_create_ilu(sp_matrix_ptr m, sp_matrix_skyline_ilu_ptr ilu);
void sp_matrix_skyline_ilu_free(sp_matrix_skyline_ilu_ptr ilu);

#endif
