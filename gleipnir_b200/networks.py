"""Mass-action networks of the reference's shipped PySB models (hand-transcribed: PySB and
BioNetGen are not installable here, so network generation from rules cannot run).

  dimerization   examples/pysb_dimerization_model/dimerization_model.py   A + A <> AA
  model_0        jupyter_notebooks/grouped_reaction_models/model_0.py     A + B <> AB
"""
from .odegen import MassActionNetwork


def dimerization():
    # Rule('ReversibleBinding', A(d=None) + A(d=None) | A(d=1) % A(d=1), kf, kr); A_0 = 20*V, V = 100
    return MassActionNetwork(
        "dimerization", species=["A", "AA"], parameters={"kf": 0.001, "kr": 1.0, "A_0": 2000.0},
        reactions=[("kf", ["A", "A"], ["AA"]), ("kr", ["AA"], ["A", "A"])],
        initials={"A": "A_0"}, observables={"A_free": {"A": 1}, "A_dimer": {"AA": 1}})


def model_0():
    # A(B=None) + B(A=None) | A(B=1) % B(A=1), kf = 1.5e-5, kr = 1.2e-4; A_0 = 2e5, B_0 = 5e4
    return MassActionNetwork(
        "model_0", species=["A", "B", "AB"],
        parameters={"kf": 1.5e-05, "kr": 0.00012, "A_0": 200000.0, "B_0": 50000.0},
        reactions=[("kf", ["A", "B"], ["AB"]), ("kr", ["AB"], ["A", "B"])],
        initials={"A": "A_0", "B": "B_0"},
        observables={"A_obs": {"A": 1, "AB": 1}, "B_obs": {"B": 1, "AB": 1}, "AB_complex": {"AB": 1}})


def jarm_like(n_sites=2):
    """SYNTHETIC stand-in for the JARM network (examples/pysb_jarm/run_jarm_multinest.py imports
    `jnk3_no_ask1.model`, which is not in the reference tree -- SURVEY 6.3 item 3).  Same shape
    of problem: a two-kinase (MKK4, MKK7) dual-site phosphorylation of JNK3 with an arrestin
    scaffold that binds the kinases and the substrate, all steps mass action with
    bind/unbind/catalyse constants spanning many decades.  16 species, 24 reactions,
    24 rate constants (18 of them sampled in the tests) + 4 initial amounts."""
    sp = ["J", "pJ_T", "pJ_Y", "ppJ", "K4", "K7", "Arr",
          "K4_J", "K7_J", "K4_pJT", "K7_pJY", "Arr_J", "Arr_K4", "Arr_K7", "Arr_K4_J", "Arr_K7_J"]
    par = {}
    rx = []

    def bind(a, b, c, name, kf, kr):
        par["kf_" + name], par["kr_" + name] = kf, kr
        rx.append(("kf_" + name, [a, b], [c]))
        rx.append(("kr_" + name, [c], [a, b]))

    def cat(c, prods, name, kc):
        par["kc_" + name] = kc
        rx.append(("kc_" + name, [c], prods))

    bind("K4", "J", "K4_J", "k4j", 2e-3, 0.5)
    cat("K4_J", ["K4", "pJ_Y"], "k4j", 0.8)                 # MKK4 prefers Tyr
    bind("K7", "J", "K7_J", "k7j", 1e-3, 0.3)
    cat("K7_J", ["K7", "pJ_T"], "k7j", 1.2)                 # MKK7 prefers Thr
    bind("K4", "pJ_T", "K4_pJT", "k4pt", 3e-3, 0.4)
    cat("K4_pJT", ["K4", "ppJ"], "k4pt", 0.6)
    bind("K7", "pJ_Y", "K7_pJY", "k7py", 2.5e-3, 0.2)
    cat("K7_pJY", ["K7", "ppJ"], "k7py", 1.0)
    bind("Arr", "J", "Arr_J", "aj", 5e-4, 0.05)
    bind("Arr", "K4", "Arr_K4", "ak4", 8e-4, 0.08)
    bind("Arr", "K7", "Arr_K7", "ak7", 6e-4, 0.06)
    bind("Arr_K4", "J", "Arr_K4_J", "ak4j", 4e-3, 0.1)
    cat("Arr_K4_J", ["Arr_K4", "pJ_Y"], "ak4j", 2.0)        # scaffold accelerates catalysis
    bind("Arr_K7", "J", "Arr_K7_J", "ak7j", 3e-3, 0.1)
    cat("Arr_K7_J", ["Arr_K7", "pJ_T"], "ak7j", 2.5)
    par.update({"J_0": 600.0, "K4_0": 50.0, "K7_0": 50.0, "Arr_0": 200.0})
    obs = {"pTyr_jnk3": {"pJ_Y": 1, "ppJ": 1, "K7_pJY": 1},
           "pThr_jnk3": {"pJ_T": 1, "ppJ": 1, "K4_pJT": 1},
           "all_jnk3": {s: 1 for s in sp if "J" in s or "pJ" in s}}
    return MassActionNetwork("jarm_like", sp, par, rx,
                             {"J": "J_0", "K4": "K4_0", "K7": "K7_0", "Arr": "Arr_0"}, obs)


SHIPPED = {"dimerization": dimerization, "model_0": model_0, "jarm_like": jarm_like}
