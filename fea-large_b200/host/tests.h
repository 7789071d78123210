#ifndef FEA_B200_TESTS_H
#define FEA_B200_TESTS_H
#include "defines.h"
/* start-up self test, run by main() before anything else (reference tests.c:53, fea_solver.c:76) */
BOOL do_tests(void);
#endif
