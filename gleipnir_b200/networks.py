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


SHIPPED = {"dimerization": dimerization, "model_0": model_0}
