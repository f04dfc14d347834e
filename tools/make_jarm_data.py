"""Synthetic stand-in for `examples/pysb_jarm/data/exp_data_3min.csv`, which the reference's JARM
scripts read (run_jarm_multinest.py:24) but the reference tree does not contain.

Same columns.  Values: the nominal parameters of the synthetic JARM-shaped network
(gleipnir_b200/networks.py::jarm) simulated by the SciPy oracle through the script's own protocol
(pre-equilibration without catalysis, then 180 s with / without arrestin), observables divided by the
total JNK3 amount (0.6), multiplied by 1 + 0.05*N(0,1) noise from numpy default_rng(2024); standard
deviations 0.1*value + 0.01.  The first four no-arrestin pTyr points are scored by a half-normal
(:38-39), which has no density below its location: those four averages are set to 0.9 x the simulated
value so that the nominal parameters stay inside the support.

    python tools/make_jarm_data.py        (writes gleipnir_b200/data/jarm_exp_data_3min.csv)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from gleipnir_b200 import networks          # noqa: E402
from oracle import ode                      # noqa: E402

EXP_TIMES = [0, 1, 2, 4, 6, 8, 10, 15, 20, 30, 45, 60, 90, 120, 150, 180]


def main():
    net = networks.jarm()
    d = net.description()
    ts = np.linspace(0.0, 180.0, 181)
    proto = dict(sampled_idx=networks.JARM_CALIBRATED, log10=True, tspan=ts,
                 conditions=[[], [(networks.JARM_ARRESTIN[0], 0.0)] + list(zip(networks.JARM_JNK3_INITIALS,
                                                                                   networks.JARM_NOARRESTIN_JNK3))],
                 preeq=dict(t_end=100.0, n_points=100, zero_idx=networks.JARM_KCAT, eq_tol=None), cycles=None,
                 rows=[dict(cond=c, obs=0, scale=1.0, kind="norm", data=np.zeros(181), sd=np.ones(181),
                            mask=np.isin(np.arange(181), EXP_TIMES).astype(float)) for c in (0, 1)])
    Y = np.log10(d["param_values"][networks.JARM_CALIBRATED])
    _, sims = ode.protocol_loglike(d, proto, Y, detail=True)
    rng = np.random.default_rng(2024)
    cols = {"Time (secs)": np.array(EXP_TIMES, dtype=float)}
    for cond, tag in ((0, "arrestin"), (1, "noarrestin")):
        for o, name in ((0, "pTyr"), (1, "pThr")):
            clean = sims[cond][EXP_TIMES, o] / networks.JARM_JNK3_TOTAL
            avg = clean * (1.0 + 0.05 * rng.standard_normal(len(EXP_TIMES)))
            if tag == "noarrestin" and name == "pTyr":
                avg[:4] = 0.9 * clean[:4]
            cols["%s_%s_avg" % (name, tag)] = avg
            cols["%s_%s_std" % (name, tag)] = 0.1 * clean + 0.01
    out = os.path.join(os.path.dirname(HERE), "gleipnir_b200", "data", "jarm_exp_data_3min.csv")
    names = list(cols)
    with open(out, "w") as f:
        f.write(",".join(names) + "\n")
        for i in range(len(EXP_TIMES)):
            f.write(",".join(repr(float(cols[n][i])) for n in names) + "\n")
    print(out)


if __name__ == "__main__":
    main()
