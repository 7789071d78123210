/* Basic types shared by the host layer.  Same public names as the reference's
 * solver-large/defines.h (real, BOOL, TRUE/FALSE, MAX_DOF, EQUAL, DELTA) so host code
 * written against the reference headers compiles unchanged; the device path is FP64
 * only, so `real` is always double here (the reference's -DSINGLE switch is not offered). */
#ifndef FEA_B200_DEFINES_H
#define FEA_B200_DEFINES_H

#include <float.h>
#include <math.h>

typedef double real;
typedef int BOOL;
#define TRUE 1
#define FALSE 0

enum { MAX_DOF = 3, MAX_MATERIAL_PARAMETERS = 10 };

#define REAL_EPSILON DBL_EPSILON
/* relative equality to one epsilon; with y == 0 it is an exact-zero test (defines.h:50) */
#define EQUAL(x, y) (fabs((x) - (y)) <= fmax(fabs(x), fabs(y)) * REAL_EPSILON ? TRUE : FALSE)
#define DELTA(i, j) ((i) == (j) ? 1 : 0)

#endif
