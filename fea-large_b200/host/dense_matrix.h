/* 3x3 helpers of the host layer: the reference's dense_matrix.h API.  Only the start-up
 * self test (tests.c) and host-side post-processing use them; all element arithmetic of
 * the solver runs on the GPU. */
#ifndef FEA_B200_DENSE_MATRIX_H
#define FEA_B200_DENSE_MATRIX_H
#include "defines.h"

typedef struct tensor_tag {
  real components[MAX_DOF][MAX_DOF];
} tensor;
typedef tensor *tensor_ptr;

real vector_norm(real *vector, int size);
real cdot(real *vector1, real *vector2, int size);
real det3x3(real (*m)[3]);
BOOL inv3x3(real (*m)[3], real *det);                                   /* in place */
void matrix_mul3x3(real (*A)[3], real (*B)[3], real (*R)[3]);            /* R = A B   */
void matrix_transpose_mul3x3(real (*A)[3], real (*B)[3], real (*R)[3]);  /* R = A' B  */
void matrix_transpose2_mul3x3(real (*A)[3], real (*B)[3], real (*R)[3]); /* R = A B'  */

#endif
