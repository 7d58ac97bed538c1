/* Stand-in for R's <R_ext/Rdynload.h> (registration types only). */
#ifndef MSIM_RSHIM_RDYNLOAD_H
#define MSIM_RSHIM_RDYNLOAD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
#ifdef __cplusplus
}
#endif
#endif
