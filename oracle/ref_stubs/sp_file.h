/* ORACLE ONLY: the reference includes this header but uses nothing from it. */
#ifndef ORACLE_STUB_SP_FILE_H
#define ORACLE_STUB_SP_FILE_H
#endif
