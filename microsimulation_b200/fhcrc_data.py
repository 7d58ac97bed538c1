"""The ten look-up tables of `data/fhcrcData.rda` (man/Data.Rd: "Old data used in the prostata model") as inputs of the
device look-ups: `Table<keys..., Outcome>` (msim_table_lookup; inst/include/rcpp_table.h:136-324) and
`NumericInterpolate` (msim_interpolate; rcpp_table.h:52-111).  The reference itself never reads the file from C++ (only
the external screening model does, through Rcpp data frames: test/cluster_mic.R:11-37), so the key / outcome split
below follows the column names.

    tables = load("/path/to/fhcrcData.rda")                  # rdata.read_rda, no R needed
    spec = TABLES["prtx"]                                      # keys ("Age", "DxY", "G"), outcomes ("CM", "RP", "RT")
    rows = table_rows(tables["prtx"], spec["keys"], "RP")      # n_rows x (n_keys + 1), keys first: msim_table_lookup's `rows`
"""
import numpy as np

from . import rdata

# table -> key columns (in the order a Table<...> would take them) and outcome columns
TABLES = {
    "all_cause_mortality": {"keys": ("BirthCohort", "Age"), "outcomes": ("AnnualMortality", "Survival")},
    "biopsy_frequency": {"keys": ("PSA.beg",), "outcomes": ("X55.59", "X60.64", "X65.69", "X70.74", "X75.79")},
    "biopsy_sensitivity": {"keys": ("Year",), "outcomes": ("Sensitivity",)},
    "dre": {"keys": ("psa.low",), "outcomes": ("sensitivity", "specificity")},
    "prob_grade7": {"keys": ("slope",), "outcomes": ("Pr.grade.7.",)},
    "pradt": {"keys": ("Tx", "Age", "DxY", "Grade"), "outcomes": ("NoADT", "ADT")},
    "prtx": {"keys": ("Age", "DxY", "G"), "outcomes": ("CM", "RP", "RT")},
    "survival_dist": {"keys": ("Grade", "Time"), "outcomes": ("Survival",)},
    "survival_local": {"keys": ("AgeLow", "Grade", "Time"), "outcomes": ("Survival",)},
    "prob_earlystage": {"keys": ("BeginAge", "BeginPSA", "Grade"), "outcomes": ("Prob",)},
}


def load(path):
    """The `fhcrcData` list of an .rda file: {table name: {column name: numpy array}}."""
    d = rdata.read_rda(path)
    if "fhcrcData" not in d:
        raise ValueError("%s holds %s, not fhcrcData" % (path, sorted(d)))
    return d["fhcrcData"]


def table_rows(table, keys, outcome):
    """Rows of a data frame as msim_table_lookup takes them: n_rows x (len(keys) + 1) doubles, keys first."""
    cols = [np.asarray(table[k], dtype=np.float64) for k in keys] + [np.asarray(table[outcome], dtype=np.float64)]
    return np.ascontiguousarray(np.stack(cols, axis=1))


def lookup(table, keys, outcome, query):
    """Table<keys..., Outcome>(query) on the GPU: per axis the largest key <= the argument (the smallest if none)."""
    from . import lookups
    return lookups.Table(table_rows(table, keys, outcome))(np.asarray(query, dtype=np.float64))


def mortality_hazards(tables, birth_cohort):
    """(ages, annual mortality rates) of one birth cohort of all_cause_mortality: the piecewise-constant hazards an
    Rpexp for other-cause death is built from (msim_rpexp)."""
    t = tables["all_cause_mortality"]
    sel = np.asarray(t["BirthCohort"]) == birth_cohort
    order = np.argsort(np.asarray(t["Age"])[sel])
    return np.asarray(t["Age"], dtype=np.float64)[sel][order], np.asarray(t["AnnualMortality"], dtype=np.float64)[sel][order]
