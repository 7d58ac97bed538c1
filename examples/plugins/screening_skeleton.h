/* screening_skeleton.h — inputs of the screening-model SKELETON (examples/plugins/screening_skeleton.cu, and its
 * CPU restatement in oracle/).
 *
 * The reference's screening model (callFhcrc) is not in the reference tree — it lives in mclements/prostata; what the
 * tree keeps of it is the call signature and enum lists (test/cluster_mic.R:11-37, 74-83), the PSA natural-history
 * sketch (test/test_microsimulation.R:1949-1990) and the data tables (data/fhcrcData.rda).  The skeleton strings
 * those pieces together so that the device building blocks — Rpexp, Table, NumericInterpolate, gsm, EventReport — run
 * TOGETHER in one kernel; it is not the FHCRC model and nothing in the reference pins its results (parity unpinned:
 * the plugin is checked against the oracle's restatement of the same skeleton).
 *
 *   onset          age_o = 35 + sqrt(-2 log(U) / g0); grade = U < 0.006 (age_o - 35)            (:1959-1962)
 *   PSA            log psa(t) = min(log 20, beta0 + beta1 (t - 35) + beta2 max(0, t - age_o) + eps), betas drawn with
 *                  rnorm / rnormPos                                                              (:1963-1967)
 *   other death    Rpexp over one birth cohort of all_cause_mortality
 *   clinical dx    onset + gsm::randU(U, 0, grade)  (a generalised survival model for the sojourn time)
 *   screening      every `screen` years from 50 to 70: PSA >= threshold and U < compliance * biopsy_frequency[PSA, age]
 *                  (a Table) -> biopsy; cancer present and U < biopsy_sensitivity[cohort + age] (a step function)
 *                  -> screen diagnosis
 *   after dx       time to cancer death by inverting a survival curve (NumericInterpolate::invert_decreasing); none
 *                  if U falls below the curve's last value
 */
#ifndef SCREENING_SKELETON_H
#define SCREENING_SKELETON_H
#include "../../include/msim.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { SCR_Healthy, SCR_Localised, SCR_Diagnosed, SCR_N_STATE };
enum { SCR_toLocalised, SCR_toClinicalDiagnosis, SCR_toCancerDeath, SCR_toOtherDeath, SCR_toScreen, SCR_toBiopsy,
       SCR_toScreenDiagnosis, SCR_N_EVENT };

typedef struct screening_inputs {
  double g0, mubeta0, sebeta0, mubeta1, sebeta1, mubeta2[2], sebeta2[2], tau2;
  double psa_threshold, compliance, cohort;
  int32_t screen;          /* 0: no screening; k > 0: a screen every k years from 50 to 70 */
  int32_t n_mort;          /* all_cause_mortality of the cohort: hazards mort_rate[i] from age mort_age[i] on */
  const double *mort_age, *mort_rate;
  int32_t n_biopsy;        /* biopsy_frequency as rows (PSA from, age from, probability): Table<double, double, double> */
  const double *biopsy_rows;
  int32_t n_sens;          /* biopsy_sensitivity: (Year, Sensitivity), a step function */
  const double *sens_year, *sens_value;
  int32_t n_surv;          /* survival after diagnosis: (Time, Survival), decreasing */
  const double *surv_time, *surv_value;
  const msim_gsm *gsm;     /* sojourn time from onset to clinical diagnosis, covariate pattern = grade */
} screening_inputs;

/* per-person outputs (plugin values / oracle values): age at death, age at diagnosis (0 if never), 1 if the diagnosis
 * came from a screen, grade */
#define SCR_N_VALUES 4

#ifdef __cplusplus
}
#endif
#endif
