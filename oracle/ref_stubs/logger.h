/* ORACLE ONLY: stand-in for the un-vendored liblogger (reference
 * fea_solver.c:17,57-61,70-94,212-238).  Messages go to a capture hook in
 * ref_standin.c so tests can read the "Tolerance <X,R>" trace. */
#ifndef ORACLE_STUB_LOGGER_H
#define ORACLE_STUB_LOGGER_H
typedef enum { LOG_LEVEL_ALL = 0 } log_level_t;
typedef enum { LOG_FORMAT_SEXP = 0 } log_format_t;
typedef struct {
  log_level_t log_level;
  log_format_t log_format;
  const char *log_file_path;
  int log_rotate_count;
  int use_stdout;
} logger_parameters;
void logger_init_with_params(logger_parameters *p);
void logger_fini(void);
void ref_log_capture(int level, const char *fmt, ...);
#define LOG(...)      ref_log_capture(0, __VA_ARGS__)
#define LOGINFO(...)  ref_log_capture(1, __VA_ARGS__)
#define LOGERROR(...) ref_log_capture(2, __VA_ARGS__)
#endif
