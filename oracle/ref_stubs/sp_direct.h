/* ORACLE ONLY: stand-in declarations for libspmatrix's direct solvers
 * (used at reference fea_solver.c:285-295).  See sp_matrix.h header note. */
#ifndef ORACLE_STUB_SP_DIRECT_H
#define ORACLE_STUB_SP_DIRECT_H
#include "sp_matrix.h"
typedef struct { int valid; } sp_chol_symbolic;
typedef sp_chol_symbolic *sp_chol_symbolic_ptr;
int sp_matrix_yale_chol_symbolic(sp_matrix_yale_ptr m, sp_chol_symbolic_ptr s);
int sp_matrix_yale_chol_symbolic_solve(sp_matrix_yale_ptr m,
                                       sp_chol_symbolic_ptr s,
                                       double *b, double *x);
#endif
