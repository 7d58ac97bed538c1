/* R glue for libmsim_b200: the file a maintainer adds to the reference package's src/ IN PLACE OF the model and
 * engine sources (person-r.cc, calibperson-r.cc, illness-death.cpp, simple-example.cc, simple-example2.cc, ssim.cc,
 * RngStream.cpp and the extern "C" block of microsimulation.cc).  It defines exactly the symbols src/init.c:11-24
 * declares and registers (init.c itself, the R wrappers and the tests stay as they are) and forwards each to the C ABI
 * of include/msim.h.  Plain R API, no Rcpp.  Link with -lmsim_b200.
 *
 *   .Call  callPersonSimulation(inseed, parms)      src/person-r.cc:195-226
 *          callSimplePerson(parms)                  src/simple-example.cc:72-83
 *          callSimplePerson2(parms)                 src/simple-example2.cc:62-83
 *          callIllnessDeath(parms)                  src/illness-death.cpp:73-94
 *          callCalibrationSimulation(parms)         src/calibperson-r.cc:176-200
 *   .C     r_create_current_stream, r_remove_current_stream, r_set_user_random_seed, r_get_user_random_seed,
 *          r_next_rng_substream, r_rng_advance_substream          src/microsimulation.cc:33-75
 *   hook   user_unif_rand (resolved by name under RNGkind("user")) src/microsimulation.cc:77-85
 *
 * The GPU is chosen with the environment variable MSIM_DEVICE (default 0) at the first call.
 * R is not in the build image: tests/test_r_glue.py compiles this file against a small stand-in for the R API
 * (tests/r_shim) and drives the entry points the way R's .Call/.C would. */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "msim.h"

static void check(int rc) {
  if (rc != MSIM_OK) Rf_error("microsimulation (GPU engine): %s", msim_last_error());
}

static void ensure_device(void) {
  static int ready = 0;
  if (ready) return;
  const char *dev = getenv("MSIM_DEVICE");
  check(msim_init(dev ? atoi(dev) : 0));
  ready = 1;
}

/* parms[["name"]] */
static SEXP list_get(SEXP list, const char *name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (TYPEOF(list) == VECSXP && names != R_NilValue)
    for (R_xlen_t i = 0; i < Rf_xlength(list); i++)
      if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  Rf_error("microsimulation: parms has no element '%s'", name);
  return R_NilValue;
}

/* msim_table -> named list of vectors.  Row logs are std::map<std::string, std::vector<double>> in the reference, so
 * every column is numeric (integer_keys = 0).  Report tables come from wrap_map (microsimulation.h:1800-1841): state
 * and event keys are short and counts int, which Rcpp wraps as integer vectors (integer_keys = 1). */
static int is_integer_column(const char *name) {
  return strcmp(name, "Key") == 0 || strcmp(name, "event") == 0 || strcmp(name, "number") == 0;
}
static SEXP table_to_list(const msim_table *t, int integer_keys) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, t->ncol));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, t->ncol));
  for (int j = 0; j < t->ncol; j++) {
    const double *src = t->data + (size_t)j * (size_t)t->nrow;
    SEXP col;
    if (integer_keys && is_integer_column(t->colnames[j])) {
      col = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)t->nrow));
      for (int64_t i = 0; i < t->nrow; i++) INTEGER(col)[i] = (int)src[i];
    } else {
      col = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)t->nrow));
      if (t->nrow > 0) memcpy(REAL(col), src, sizeof(double) * (size_t)t->nrow);
    }
    SET_VECTOR_ELT(out, j, col);
    SET_STRING_ELT(names, j, Rf_mkChar(t->colnames[j]));
    UNPROTECT(1);
  }
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(2);
  return out;
}

/* a report table as a data.frame (Rcpp::DataFrame::create): class attribute and compact row names c(NA, -nrow) */
static SEXP table_to_df(const msim_table *t) {
  SEXP out = PROTECT(table_to_list(t, 1));
  SEXP rn = PROTECT(Rf_allocVector(INTSXP, 2));
  INTEGER(rn)[0] = NA_INTEGER;
  INTEGER(rn)[1] = -(int)t->nrow;
  Rf_setAttrib(out, R_RowNamesSymbol, rn);
  Rf_setAttrib(out, R_ClassSymbol, Rf_mkString("data.frame"));
  UNPROTECT(2);
  return out;
}

/* EventReport::wrap() (microsimulation.h:974-986): list(pt=, ut=, events=, prev=), or list() when no event was reported */
static SEXP report_to_list(msim_report *r) {
  static const char *const names[4] = {"pt", "ut", "events", "prev"};
  if (r->events.nrow == 0) {
    msim_report_free(r);
    return Rf_allocVector(VECSXP, 0);
  }
  const msim_table *tabs[4] = {&r->pt, &r->ut, &r->events, &r->prev};
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 4));
  for (int k = 0; k < 4; k++) {
    SET_VECTOR_ELT(out, k, table_to_df(tabs[k]));
    SET_STRING_ELT(nm, k, Rf_mkChar(names[k]));
  }
  Rf_setAttrib(out, R_NamesSymbol, nm);
  msim_report_free(r);
  UNPROTECT(2);
  return out;
}

/* R's Mersenne-Twister state travels through .Random.seed (INTSXP: kind, position, 624 words), which is what
 * Rcpp::RNGScope's GetRNGstate()/PutRNGstate() read and write around the model (illness-death.cpp:75) */
static SEXP mt_seed_var(void) {
  SEXP s = Rf_findVar(Rf_install(".Random.seed"), R_GlobalEnv);
  if (s == R_UnboundValue || TYPEOF(s) != INTSXP || Rf_xlength(s) != 626 || INTEGER(s)[0] % 100 != 3)
    Rf_error("microsimulation: expected RNGkind(\"Mersenne-Twister\") with a seed set (set.seed) before this call");
  return s;
}
static void mt_in(void) { check(msim_mt_set_state((const uint32_t *)(INTEGER(mt_seed_var()) + 1))); }
/* PutRNGstate(): a FRESH .Random.seed is installed (the old vector may be shared, e.g. after s <- .Random.seed, and
 * must not change under its other owners).  The state is fetched by the caller first, so that nothing of the library's
 * is allocated when an error is raised. */
static void mt_store(const uint32_t st[625]) {
  SEXP old = mt_seed_var();
  SEXP s = PROTECT(Rf_allocVector(INTSXP, 626));
  INTEGER(s)[0] = INTEGER(old)[0];
  memcpy(INTEGER(s) + 1, st, 625 * sizeof(uint32_t));
  Rf_defineVar(Rf_install(".Random.seed"), s, R_GlobalEnv);
  UNPROTECT(1);
}

SEXP callPersonSimulation(SEXP inseed, SEXP parms) {
  if (TYPEOF(inseed) != INTSXP || Rf_xlength(inseed) < 6) Rf_error("microsimulation: inseed must be six integers");
  const int n = Rf_asInteger(list_get(parms, "n"));
  int32_t seed[6];
  for (int i = 0; i < 6; i++) seed[i] = INTEGER(inseed)[i];
  ensure_device();
  msim_table t;
  check(msim_callPersonSimulation(seed, n, &t));
  SEXP out = PROTECT(table_to_list(&t, 0));
  msim_table_free(&t);
  UNPROTECT(1);
  return out;
}

SEXP callSimplePerson(SEXP parms) {
  const int n = Rf_asInteger(list_get(parms, "n"));
  ensure_device();
  mt_in();
  msim_table t;
  uint32_t st[625];
  check(msim_callSimplePerson(n, &t));
  const int rc = msim_mt_get_state(st);
  if (rc) msim_table_free(&t); /* before the R error unwinds past it */
  check(rc);
  mt_store(st);
  SEXP out = PROTECT(table_to_list(&t, 0));
  msim_table_free(&t);
  UNPROTECT(1);
  return out;
}

SEXP callSimplePerson2(SEXP parms) {
  const int n = Rf_asInteger(list_get(parms, "n"));
  ensure_device();
  mt_in();
  msim_report r;
  uint32_t st[625];
  check(msim_callSimplePerson2(n, &r));
  const int rc = msim_mt_get_state(st);
  if (rc) msim_report_free(&r);
  check(rc);
  mt_store(st);
  return report_to_list(&r);
}

SEXP callIllnessDeath(SEXP parms) {
  const int n = Rf_asInteger(list_get(parms, "n"));
  const double cure = Rf_asReal(list_get(parms, "cure")), zsd = Rf_asReal(list_get(parms, "zsd"));
  ensure_device();
  mt_in();
  msim_report r;
  uint32_t st[625];
  check(msim_callIllnessDeath(n, cure, zsd, &r));
  const int rc = msim_mt_get_state(st);
  if (rc) msim_report_free(&r);
  check(rc);
  mt_store(st);
  return report_to_list(&r);
}

/* Rcpp::wrap of CalibPerson::report, a std::map: keys in alphabetical order, stages only if they were ever counted */
SEXP callCalibrationSimulation(SEXP parms) {
  static const char *const key[5] = {"Clinical", "DiseaseFree", "PreClinical", "Precursor", "TimeAtRisk"};
  static const int column[4] = {3, 0, 2, 1}; /* msim_calibration.occupancy columns: DiseaseFree, Precursor, PreClinical, Clinical */
  SEXP runpar = list_get(parms, "runpar");
  if (TYPEOF(runpar) != REALSXP || Rf_xlength(runpar) < 6) Rf_error("microsimulation: runpar must be six doubles");
  const int n_persons = Rf_asInteger(list_get(parms, "n"));
  ensure_device();
  msim_calibration c;
  check(msim_callCalibrationSimulation(n_persons, REAL(runpar), &c));
  int n = 1;
  for (int k = 0; k < 4; k++) n += c.present[column[k]] != 0;
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, n));
  int at = 0;
  for (int k = 0; k < 4; k++) {
    if (!c.present[column[k]]) continue;
    SEXP v = PROTECT(Rf_allocVector(REALSXP, 10));
    memcpy(REAL(v), c.occupancy + 10 * column[k], 10 * sizeof(double));
    SET_VECTOR_ELT(out, at, v);
    SET_STRING_ELT(nm, at, Rf_mkChar(key[k]));
    UNPROTECT(1);
    at++;
  }
  SEXP tar = PROTECT(Rf_allocVector(REALSXP, c.n_time_at_risk));
  if (c.n_time_at_risk > 0) memcpy(REAL(tar), c.time_at_risk, sizeof(double) * (size_t)c.n_time_at_risk);
  SET_VECTOR_ELT(out, at, tar);
  SET_STRING_ELT(nm, at, Rf_mkChar(key[4]));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(3);
  return out;
}

/* ---- .C entries: same names and argument lists as the reference ---- */
void r_create_current_stream(void) { msim_r_create_current_stream(); }
void r_remove_current_stream(void) { msim_r_remove_current_stream(); }
void r_set_user_random_seed(double *seed) { check(msim_r_set_user_random_seed(seed)); }
void r_get_user_random_seed(double *seed) { check(msim_r_get_user_random_seed(seed)); }
void r_next_rng_substream(void) { check(msim_r_next_rng_substream()); }
void r_rng_advance_substream(double *seed, int *n) { check(msim_r_rng_advance_substream(seed, n)); }

/* R looks this one up by name when RNGkind("user") is selected */
double *user_unif_rand(void) {
  double *u = msim_user_unif_rand();
  if (!u) REprintf("user_unif_rand(): No stream created yet!");
  return u;
}
