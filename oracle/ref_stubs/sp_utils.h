/* ORACLE ONLY: string helpers used at reference fea_solver.c:1672-1684. */
#ifndef ORACLE_STUB_SP_UTILS_H
#define ORACLE_STUB_SP_UTILS_H
const char *sp_parse_file_extension(const char *filename);
const char *sp_parse_file_basename(const char *filename, char *out);
int sp_istrcmp(const char *a, const char *b);
#endif
