/* ORACLE ONLY: the reference's sexp_loader.c is NOT compiled into the oracle
 * (libsexp is absent); meshes reach the oracle as flat arrays. */
#ifndef ORACLE_STUB_LIBSEXP_H
#define ORACLE_STUB_LIBSEXP_H
#endif
