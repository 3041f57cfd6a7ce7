"""The oracle against the reference's own shipped golden vectors for the multi-asset path
(SURVEY.md 8c): cached XLL outputs of AutocallPricer.xlsx (price + 25 AAD risks of a 3-asset
autocallable) and testDLM.xlsx (27 moments of a 3-asset displaced model with dividends, 10^6 Sobol
paths).  The fixture tests/golden/golden_xlsx.json was extracted by tests/golden/make_golden_xlsx.py.
CPU only: this pins oracle/_ref (the reference compiled with g++), which the GPU tests then use."""
import json
import os

import numpy as np

from conftest import rel_err

X = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_xlsx.json")))


def named(labels, names):
    """The C drivers name the assets a0, a1, ...; the spreadsheets use tickers."""
    out = []
    for l in labels:
        for i, nm in enumerate(names):
            l = l.replace(f"a{i}", nm)
        out.append(l)
    return out


def put_dlm(api, g, id_):
    api.put_displaced(g["spots"], g["atms"], g["skews"], g["disc_rate"], g["repo_spreads"], g["div_dates"],
                      np.array(g["divs"]), np.array(g["correl"]), g["lam"], id_)


def test_reference_reproduces_autocall_pricer_xlsx(ref):
    g = X["autocall_pricer"]
    put_dlm(ref, g, "dlm_x")
    ref.put_autocall(g["spots"], g["maturity"], g["periods"], g["ko"], g["strike"], g["cpn"], g["smooth"], "auto_x")
    assert ref.labels("auto_x")[0] == g["payoff_label"]
    assert named(ref.labels("dlm_x", params=True), ["uber", "lyft", "luckin"]) == g["risk_labels"]
    price = ref.value("dlm_x", "auto_x", g["n_paths"])[0]
    assert abs(price / g["price"] - 1) < 1e-14
    pv, rv, risks = ref.aad_risk_one("dlm_x", "auto_x", g["n_paths"])
    assert abs(rv / g["risk_value"] - 1) < 1e-14
    assert np.max(np.abs(risks - np.array(g["risks"]))) < 1e-12       # thread summation order: ~3e-14 observed


def test_reference_reproduces_test_dlm_xlsx(ref):
    g = X["test_dlm"]
    put_dlm(ref, g, "dlm_t")
    ref.put_multistats(3, g["fix_dates"], g["fwd_dates"], "stats_t")
    got = ref.value("dlm_t", "stats_t", g["n_paths"])
    assert got.size == 27
    assert rel_err(got, g["values"]) < 1e-13
