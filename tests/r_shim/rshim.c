/* A miniature of the R runtime for testing glue/msim_glue.c without R (test infrastructure only).
 *
 * Implements the R API subset of include/Rinternals.h on malloc'ed nodes (no garbage collector: objects live until
 * shim_reset()), and a driver the Python tests use through ctypes to do what R's evaluator would: build argument
 * objects, invoke a .Call / .C entry point by symbol with the argument count R would check, catch Rf_error, and
 * read the result back. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>

struct attr {
  SEXP name, value;
  struct attr *next;
};
struct SEXPREC {
  int type;
  R_xlen_t n;
  void *data; /* double[], int[], SEXP[] or char[] */
  struct attr *attrs;
  struct SEXPREC *all_next;
};

static struct SEXPREC nil_node = {NILSXP, 0, NULL, NULL, NULL}, unbound_node = {SYMSXP, 0, NULL, NULL, NULL},
                      global_node = {ENVSXP, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_node, R_UnboundValue = &unbound_node, R_GlobalEnv = &global_node;
SEXP R_NamesSymbol, R_ClassSymbol, R_RowNamesSymbol;
int R_NaInt = INT32_MIN;

static SEXP all_nodes = NULL;
static int protect_depth = 0;
static jmp_buf error_jmp;
static int error_armed = 0;
static char error_msg[512], stderr_msg[512];

static SEXP node(int type, R_xlen_t n, size_t elt) {
  SEXP x = (SEXP)calloc(1, sizeof *x);
  x->type = type;
  x->n = n;
  x->data = calloc((size_t)(n > 0 ? n : 1), elt);
  x->all_next = all_nodes;
  all_nodes = x;
  return x;
}

void Rf_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_msg, sizeof error_msg, fmt, ap);
  va_end(ap);
  if (!error_armed) {
    fprintf(stderr, "Rf_error outside a shim call: %s\n", error_msg);
    abort();
  }
  longjmp(error_jmp, 1);
}
void REprintf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(stderr_msg, sizeof stderr_msg, fmt, ap);
  va_end(ap);
}

int TYPEOF(SEXP x) { return x->type; }
R_xlen_t Rf_xlength(SEXP x) { return x->n; }
int Rf_length(SEXP x) { return (int)x->n; }

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
  switch (type) {
    case REALSXP: return node(REALSXP, n, sizeof(double));
    case INTSXP:
    case LGLSXP: return node((int)type, n, sizeof(int));
    case VECSXP:
    case STRSXP: {
      SEXP x = node((int)type, n, sizeof(SEXP));
      for (R_xlen_t i = 0; i < n; i++) ((SEXP *)x->data)[i] = R_NilValue;
      return x;
    }
    default: Rf_error("shim: allocVector of type %u not supported", type);
  }
}
static void want(SEXP x, int type, const char *who) {
  if (x->type != type) Rf_error("shim: %s() applied to an object of type %d", who, x->type);
}
static void in_range(SEXP x, R_xlen_t i, const char *who) {
  if (i < 0 || i >= x->n) Rf_error("shim: %s index %ld out of bounds (length %ld)", who, (long)i, (long)x->n);
}
double *REAL(SEXP x) { want(x, REALSXP, "REAL"); return (double *)x->data; }
int *INTEGER(SEXP x) {
  if (x->type != INTSXP && x->type != LGLSXP) Rf_error("shim: INTEGER() applied to an object of type %d", x->type);
  return (int *)x->data;
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { want(x, VECSXP, "VECTOR_ELT"); in_range(x, i, "VECTOR_ELT"); return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { want(x, VECSXP, "SET_VECTOR_ELT"); in_range(x, i, "SET_VECTOR_ELT"); return ((SEXP *)x->data)[i] = v; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { want(x, STRSXP, "STRING_ELT"); in_range(x, i, "STRING_ELT"); return ((SEXP *)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { want(x, STRSXP, "SET_STRING_ELT"); want(v, CHARSXP, "SET_STRING_ELT value"); in_range(x, i, "SET_STRING_ELT"); ((SEXP *)x->data)[i] = v; }
SEXP Rf_mkChar(const char *s) {
  SEXP x = node(CHARSXP, (R_xlen_t)strlen(s), 1);
  free(x->data);
  x->data = strdup(s);
  return x;
}
SEXP Rf_mkString(const char *s) {
  SEXP x = Rf_allocVector(STRSXP, 1);
  SET_STRING_ELT(x, 0, Rf_mkChar(s));
  return x;
}
const char *R_CHAR(SEXP x) { want(x, CHARSXP, "CHAR"); return (const char *)x->data; }

/* symbols are interned; variables of the global environment hang off it as attributes */
static SEXP symbols = NULL;
SEXP Rf_install(const char *name) {
  for (SEXP s = symbols; s; s = (SEXP)s->attrs)
    if (strcmp((const char *)s->data, name) == 0) return s;
  SEXP s = (SEXP)calloc(1, sizeof *s);
  s->type = SYMSXP;
  s->data = strdup(name);
  s->attrs = (struct attr *)symbols; /* intern list link (symbols carry no attributes here) */
  symbols = s;
  return s;
}
static SEXP attr_get(SEXP x, SEXP name) {
  for (struct attr *a = x->attrs; a; a = a->next)
    if (a->name == name) return a->value;
  return NULL;
}
static void attr_set(SEXP x, SEXP name, SEXP value) {
  for (struct attr *a = x->attrs; a; a = a->next)
    if (a->name == name) {
      a->value = value;
      return;
    }
  struct attr *a = (struct attr *)calloc(1, sizeof *a);
  a->name = name;
  a->value = value;
  a->next = x->attrs;
  x->attrs = a;
}
SEXP Rf_findVar(SEXP symbol, SEXP env) {
  if (env != R_GlobalEnv) Rf_error("shim: only the global environment exists");
  SEXP v = attr_get(env, symbol);
  return v ? v : R_UnboundValue;
}
void Rf_defineVar(SEXP symbol, SEXP value, SEXP env) {
  if (env != R_GlobalEnv) Rf_error("shim: only the global environment exists");
  attr_set(env, symbol, value);
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value) {
  if (x == R_NilValue) Rf_error("shim: attempt to set an attribute on NULL");
  if (name == R_NamesSymbol && (value->type != STRSXP || value->n != x->n)) Rf_error("shim: 'names' must be a character vector of the object's length");
  attr_set(x, name, value);
  return value;
}
SEXP Rf_getAttrib(SEXP x, SEXP name) {
  SEXP v = attr_get(x, name);
  return v ? v : R_NilValue;
}
int Rf_asInteger(SEXP x) {
  if (x->n < 1) return R_NaInt;
  if (x->type == INTSXP || x->type == LGLSXP) return ((int *)x->data)[0];
  if (x->type == REALSXP) return (int)((double *)x->data)[0];
  return R_NaInt;
}
double Rf_asReal(SEXP x) {
  if (x->n < 1) return 0.0 / 0.0;
  if (x->type == REALSXP) return ((double *)x->data)[0];
  if (x->type == INTSXP || x->type == LGLSXP) return (double)((int *)x->data)[0];
  return 0.0 / 0.0;
}
SEXP Rf_protect(SEXP x) { protect_depth++; return x; }
void Rf_unprotect(int n) {
  protect_depth -= n;
  if (protect_depth < 0) {
    fprintf(stderr, "shim: protect stack underflow\n");
    abort();
  }
}

/* ---------------- the driver (called from Python) ---------------- */

static void shim_boot(void) {
  if (R_NamesSymbol) return;
  R_NamesSymbol = Rf_install("names");
  R_ClassSymbol = Rf_install("class");
  R_RowNamesSymbol = Rf_install("row.names");
}

void shim_reset(void) {
  shim_boot();
  while (all_nodes) {
    SEXP x = all_nodes;
    all_nodes = x->all_next;
    for (struct attr *a = x->attrs; a;) {
      struct attr *nx = a->next;
      free(a);
      a = nx;
    }
    free(x->data);
    free(x);
  }
  for (struct attr *a = R_GlobalEnv->attrs; a;) {
    struct attr *nx = a->next;
    free(a);
    a = nx;
  }
  R_GlobalEnv->attrs = NULL;
  protect_depth = 0;
  error_msg[0] = stderr_msg[0] = 0;
}

SEXP shim_int(const int *v, int64_t n) {
  shim_boot();
  SEXP x = Rf_allocVector(INTSXP, (R_xlen_t)n);
  if (n > 0) memcpy(x->data, v, sizeof(int) * (size_t)n);
  return x;
}
SEXP shim_real(const double *v, int64_t n) {
  shim_boot();
  SEXP x = Rf_allocVector(REALSXP, (R_xlen_t)n);
  if (n > 0) memcpy(x->data, v, sizeof(double) * (size_t)n);
  return x;
}
/* list(name1 = elt1, ...) */
SEXP shim_list(int n, const char *const *names, SEXP *elts) {
  shim_boot();
  SEXP x = Rf_allocVector(VECSXP, n), nm = Rf_allocVector(STRSXP, n);
  for (int i = 0; i < n; i++) {
    SET_VECTOR_ELT(x, i, elts[i]);
    SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  }
  attr_set(x, R_NamesSymbol, nm);
  return x;
}
void shim_set_global(const char *name, SEXP value) {
  shim_boot();
  attr_set(R_GlobalEnv, Rf_install(name), value);
}
SEXP shim_get_global(const char *name) {
  shim_boot();
  SEXP v = attr_get(R_GlobalEnv, Rf_install(name));
  return v;
}

int shim_typeof(SEXP x) { return x->type; }
int64_t shim_length(SEXP x) { return (int64_t)x->n; }
void *shim_data(SEXP x) { return x->data; }
SEXP shim_elt(SEXP x, int64_t i) { return ((SEXP *)x->data)[i]; }
const char *shim_string(SEXP x, int64_t i) { return (const char *)((SEXP *)x->data)[i]->data; }
SEXP shim_attr(SEXP x, const char *name) { return attr_get(x, Rf_install(name)); }
const char *shim_last_error(void) { return error_msg; }
const char *shim_last_stderr(void) { return stderr_msg; }
int shim_protect_depth(void) { return protect_depth; }

/* .Call(symbol, args...): NULL (with shim_last_error) if the callee raised an R error */
SEXP shim_dot_call(void *fn, int nargs, SEXP *a) {
  shim_boot();
  SEXP out = NULL;
  error_msg[0] = 0;
  const int depth = protect_depth;
  error_armed = 1;
  if (setjmp(error_jmp) == 0) {
    switch (nargs) {
      case 0: out = ((SEXP(*)(void))fn)(); break;
      case 1: out = ((SEXP(*)(SEXP))fn)(a[0]); break;
      case 2: out = ((SEXP(*)(SEXP, SEXP))fn)(a[0], a[1]); break;
      case 3: out = ((SEXP(*)(SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2]); break;
      default: snprintf(error_msg, sizeof error_msg, "shim: too many arguments");
    }
  } else {
    out = NULL;
    protect_depth = depth; /* R unwinds the protect stack on error */
  }
  error_armed = 0;
  return out;
}
/* .C(symbol, args...): arguments are pointers to the vectors' data; 0 on success, 1 on an R error */
int shim_dot_c(void *fn, int nargs, void **a) {
  shim_boot();
  int rc = 0;
  error_msg[0] = 0;
  error_armed = 1;
  if (setjmp(error_jmp) == 0) {
    switch (nargs) {
      case 0: ((void (*)(void))fn)(); break;
      case 1: ((void (*)(void *))fn)(a[0]); break;
      case 2: ((void (*)(void *, void *))fn)(a[0], a[1]); break;
      default: rc = 1;
    }
  } else {
    rc = 1;
  }
  error_armed = 0;
  return rc;
}
