/* Host 3x3 helpers (reference API: solver-large/dense_matrix.c:6-110). */
#include "dense_matrix.h"

real vector_norm(real *v, int n) {
  real acc = 0.0;
  while (n-- > 0) { acc += (*v) * (*v); ++v; }
  return sqrt(acc);
}

real cdot(real *a, real *b, int n) {
  real acc = 0.0;
  int k;
  for (k = 0; k < n; ++k) acc += a[k] * b[k];
  return acc;
}

/* 2x2 minor of m that deletes row r and column c, rows/cols taken cyclically */
static real minor2(real (*m)[3], int r, int c) {
  const int r1 = (r + 1) % 3, r2 = (r + 2) % 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
  return m[r1][c1] * m[r2][c2] - m[r1][c2] * m[r2][c1];
}

real det3x3(real (*m)[3]) {
  /* cyclic minors carry the cofactor sign already */
  return m[0][0] * minor2(m, 0, 0) + m[0][1] * minor2(m, 0, 1) + m[0][2] * minor2(m, 0, 2);
}

BOOL inv3x3(real (*m)[3], real *det) {
  real adj[3][3];
  int r, c;
  *det = det3x3(m);
  if (EQUAL(*det, 0.0)) return FALSE;
  for (r = 0; r < 3; ++r)
    for (c = 0; c < 3; ++c) adj[c][r] = minor2(m, r, c) / (*det); /* transpose of the cofactors */
  for (r = 0; r < 3; ++r)
    for (c = 0; c < 3; ++c) m[r][c] = adj[r][c];
  return TRUE;
}

/* R[i][j] = sum_k A(i,k) B(k,j) with optional transposes of A or B */
static void mul3(real (*A)[3], int ta, real (*B)[3], int tb, real (*R)[3]) {
  int i, j, k;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      real acc = 0.0;
      for (k = 0; k < 3; ++k) acc += (ta ? A[k][i] : A[i][k]) * (tb ? B[j][k] : B[k][j]);
      R[i][j] = acc;
    }
}

void matrix_mul3x3(real (*A)[3], real (*B)[3], real (*R)[3]) { mul3(A, 0, B, 0, R); }
void matrix_transpose_mul3x3(real (*A)[3], real (*B)[3], real (*R)[3]) { mul3(A, 1, B, 0, R); }
void matrix_transpose2_mul3x3(real (*A)[3], real (*B)[3], real (*R)[3]) { mul3(A, 0, B, 1, R); }
