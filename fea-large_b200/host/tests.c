/* Start-up known-answer test of the host 3x3 helpers: the reference's only automated
 * test (solver-large/tests.c:12-51); main() refuses to run if it fails. */
#include <stdio.h>
#include "dense_matrix.h"
#include "tests.h"

static BOOL same3x3(real (*got)[3], const real (*want)[3]) {
  int i, j;
  BOOL ok = TRUE;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) ok &= EQUAL(got[i][j], want[i][j]);
  return ok;
}

BOOL do_tests(void) {
  /* the reference's vectors, tests.c:17-22 */
  real A[3][3] = {{1, 2, 0}, {2, 0, 3}, {0, 2, 3}};
  real B[3][3] = {{0, 2, 1}, {1, 1, 1}, {3, 2, -1}};
  static const real AB[3][3] = {{2, 4, 3}, {9, 10, -1}, {11, 8, -1}};
  static const real AtB[3][3] = {{2, 4, 3}, {6, 8, 0}, {12, 9, 0}};
  static const real ABt[3][3] = {{4, 3, 7}, {3, 5, 3}, {7, 5, 1}};
  real R[3][3], M[3][3] = {{2, 0, 0}, {0, 4, 0}, {0, 0, 8}}, det = 0;
  BOOL ok = TRUE;
  matrix_mul3x3(A, B, R);
  ok &= same3x3(R, AB);
  matrix_transpose_mul3x3(A, B, R);
  ok &= same3x3(R, AtB);
  matrix_transpose2_mul3x3(A, B, R);
  ok &= same3x3(R, ABt);
  ok &= EQUAL(det3x3(A), -18.0);
  ok &= inv3x3(M, &det) && EQUAL(det, 64.0) && EQUAL(M[1][1], 0.25);
  printf("test_matrix result: *%s*\n", ok ? "pass" : "fail");
  return ok;
}
