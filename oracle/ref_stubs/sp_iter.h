/* ORACLE ONLY: stand-in declarations for libspmatrix's iterative solvers
 * (used at reference fea_solver.c:251-256, 270-276). */
#ifndef ORACLE_STUB_SP_ITER_H
#define ORACLE_STUB_SP_ITER_H
#include "sp_matrix.h"
void sp_matrix_yale_solve_cg(sp_matrix_yale_ptr m, double *b, double *x0,
                             int *max_iter, double *tolerance, double *x);
void sp_matrix_yale_solve_pcg_ilu(sp_matrix_yale_ptr m,
                                  sp_matrix_skyline_ilu_ptr ilu,
                                  double *b, double *x0,
                                  int *max_iter, double *tolerance, double *x);
#endif
