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


# ---- C5: the JARM-shaped network -------------------------------------------------------------
JARM_N_RATE = 44          # rate constants p[0..43]; p[44..49] are initial amounts
JARM_CALIBRATED = [1, 5, 9, 11, 15, 17, 19, 23, 25, 27, 31, 35, 36, 37, 38, 39, 41, 43]   # run_jarm_multinest.py:51
JARM_KCAT = [36, 37]                                                                       # :59
JARM_ARRESTIN = [44]                                                                       # :56
JARM_JNK3_INITIALS = [47, 48, 49]                                                          # :58
JARM_JNK3_TOTAL = 0.6                                                                      # :57
JARM_NOARRESTIN_JNK3 = [0.592841488, 0.0, 0.007158512]                                     # :111,123
# evaluate_cycles (run_jarm_multinest.py:72-90): box = product of (p[a]/p[b]) or 1/(p[a]/p[b])
JARM_CYCLES = [
    [(21, 20, 0), (23, 22, 0), (1, 0, 1), (5, 4, 1)],
    [(21, 20, 0), (25, 24, 0), (3, 2, 1), (27, 26, 1)],
    [(13, 12, 0), (23, 22, 0), (1, 0, 1), (15, 14, 1)],
    [(7, 6, 0), (25, 24, 0), (3, 2, 1), (11, 10, 1)],
]


def jarm():
    """SYNTHETIC network of the JARM problem's size and shape.  The reference's C5 script imports
    `jnk3_no_ask1.model`, `equilibration_function` and `data/exp_data_3min.csv`
    (examples/pysb_jarm/run_jarm_multinest.py:12-14,24), none of which is in the reference tree
    (SURVEY 6.3 item 3), so the network below is written to match every index the script uses:
    50 parameters = 44 rate constants in (kf, kr) pairs + 6 initial amounts; p[36], p[37] the two
    catalytic constants zeroed during pre-equilibration; p[44] the arrestin amount; p[47..49] the
    three JNK3 initial amounts; and the four thermodynamic cycles of `evaluate_cycles` (:72-90) are
    genuine cycles of THIS network (arrestin-kinase, arrestin-JNK3 and the two routes to each
    ternary complex), so the box products equal 1 for detailed-balance-consistent constants.
    JNK3 carries two sites (Thr phosphorylated by MKK7, Tyr by MKK4); arrestin scaffolds either
    kinase, JNK3, or both kinases.  34 species, so the cooperative solver's 32 lanes hold a 32-row
    core and two leaf species (odegen.py)."""
    states = ["uu", "pT", "pY", "pp"]
    sp = ["Arr", "K4", "K7"] + ["J" + s for s in states] + ["Arr_K4", "Arr_K7"] + ["Arr_J" + s for s in states]
    sp += ["Arr_K4_J" + s for s in states] + ["Arr_K7_J" + s for s in states]
    sp += ["K4_J" + s for s in states] + ["K7_J" + s for s in states]
    sp += ["Arr_K4_K7"] + ["Arr_K4_K7_J" + s for s in states]
    names = [None] * 50
    vals = [0.0] * 50

    def pair(i, name, kf, kr):
        names[i], names[i + 1] = "kf_" + name, "kr_" + name
        vals[i], vals[i + 1] = kf, kr

    pair(0, "Arr_K4", 10.0, 400.0)
    pair(2, "Arr_K7", 10.0, 100.0)
    pair(4, "ArrK4_Juu", 10.0, 10.0)
    pair(6, "Arr_JpY", 10.0, 20.0)
    pair(8, "Arr_Jpp", 10.0, 100.0)
    pair(10, "ArrK7_JpY", 10.0, 10.0)
    pair(12, "Arr_JpT", 10.0, 16.0)
    pair(14, "ArrK4_JpT", 10.0, 8.0)
    pair(16, "K4_Juu", 10.0, 60.0)
    pair(18, "K7_Juu", 10.0, 90.0)
    pair(20, "Arr_Juu", 10.0, 20.0)
    pair(22, "ArrJ_K4", 10.0, 200.0)
    pair(24, "ArrJ_K7", 10.0, 50.0)
    pair(26, "ArrK7_Juu", 10.0, 10.0)
    pair(28, "K4_JpT", 10.0, 50.0)
    pair(30, "K7_JpY", 10.0, 60.0)
    pair(32, "K4_prod", 10.0, 200.0)
    pair(34, "K7_prod", 10.0, 200.0)
    names[36], vals[36] = "kcat_pTyr", 0.5
    names[37], vals[37] = "kcat_pThr", 0.8
    pair(38, "ArrK4_prod", 5.0, 100.0)
    pair(40, "ArrK7_prod", 10.0, 100.0)
    pair(42, "ArrK4K7_J", 10.0, 15.0)
    for i, (n, v) in enumerate([("Arr_0", 5.0), ("K4_0", 0.05), ("K7_0", 0.05), ("Juu_0", 0.493), ("JpT_0", 0.0),
                                ("JpY_0", 0.107)]):
        names[44 + i], vals[44 + i] = n, v
    par = dict(zip(names, vals))
    rx = []

    def bind(a, b, c, i):
        rx.append((names[i], [a, b], [c]))
        rx.append((names[i + 1], [c], [a, b]))

    def cat(a, b, i):
        rx.append((names[i], [a], [b]))

    bind("Arr", "K4", "Arr_K4", 0)
    bind("Arr", "K7", "Arr_K7", 2)
    bind("Arr_K7", "K4", "Arr_K4_K7", 0)
    bind("Arr_K4", "K7", "Arr_K4_K7", 2)
    bind("Arr", "Juu", "Arr_Juu", 20)
    bind("Arr", "JpT", "Arr_JpT", 12)
    bind("Arr", "JpY", "Arr_JpY", 6)
    bind("Arr", "Jpp", "Arr_Jpp", 8)
    bind("Arr_K4", "Juu", "Arr_K4_Juu", 4)
    bind("Arr_K4", "JpT", "Arr_K4_JpT", 14)
    bind("Arr_K4", "JpY", "Arr_K4_JpY", 38)
    bind("Arr_K4", "Jpp", "Arr_K4_Jpp", 38)
    bind("Arr_K7", "Juu", "Arr_K7_Juu", 26)
    bind("Arr_K7", "JpY", "Arr_K7_JpY", 10)
    bind("Arr_K7", "JpT", "Arr_K7_JpT", 40)
    bind("Arr_K7", "Jpp", "Arr_K7_Jpp", 40)
    for s in states:
        bind("Arr_J" + s, "K4", "Arr_K4_J" + s, 22)
        bind("Arr_J" + s, "K7", "Arr_K7_J" + s, 24)
        bind("Arr_K4_K7", "J" + s, "Arr_K4_K7_J" + s, 42)
    bind("K4", "Juu", "K4_Juu", 16)
    bind("K4", "JpT", "K4_JpT", 28)
    bind("K4", "JpY", "K4_JpY", 32)
    bind("K4", "Jpp", "K4_Jpp", 32)
    bind("K7", "Juu", "K7_Juu", 18)
    bind("K7", "JpY", "K7_JpY", 30)
    bind("K7", "JpT", "K7_JpT", 34)
    bind("K7", "Jpp", "K7_Jpp", 34)
    for pre in ("Arr_K4_J", "K4_J", "Arr_K4_K7_J"):          # MKK4 phosphorylates Tyr
        cat(pre + "uu", pre + "pY", 36)
        cat(pre + "pT", pre + "pp", 36)
    for pre in ("Arr_K7_J", "K7_J", "Arr_K4_K7_J"):          # MKK7 phosphorylates Thr
        cat(pre + "uu", pre + "pT", 37)
        cat(pre + "pY", pre + "pp", 37)
    obs = {"pTyr_jnk3": {s: 1 for s in sp if s.endswith("JpY") or s.endswith("Jpp")},
           "pThr_jnk3": {s: 1 for s in sp if s.endswith("JpT") or s.endswith("Jpp")},
           "all_jnk3": {s: 1 for s in sp if "J" in s}}
    return MassActionNetwork("jarm", sp, par, rx,
                             {"Arr": "Arr_0", "K4": "K4_0", "K7": "K7_0", "Juu": "Juu_0", "JpT": "JpT_0",
                              "JpY": "JpY_0"}, obs)


SHIPPED["jarm"] = jarm
