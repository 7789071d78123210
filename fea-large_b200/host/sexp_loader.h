#ifndef FEA_B200_SEXP_LOADER_H
#define FEA_B200_SEXP_LOADER_H
#include "fea_solver.h"
/* reads a task file in the reference's S-expression format (sexp_loader.c:275) */
BOOL sexp_data_load(char *filename, fea_task **task, fea_solution_params **fea_params,
                    nodes_array **nodes, elements_array **elements,
                    presc_bnd_array **presc_boundary);
#endif
