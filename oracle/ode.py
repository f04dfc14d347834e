"""TEST INFRASTRUCTURE ONLY (oracle) -- mass-action ODE log-likelihoods on the CPU.

Restates the PySB likelihood bodies of ``gleipnir/pysb_utilities/nestedsample_it.py:284-294``
(sum of normal log-pdfs), ``:306-316`` (-MSE), ``:329-339`` (-SSE) and the script form
``examples/pysb_dimerization_model/run_NS_dimerization_model.py:56-65``:
    params[rate_mask] = 10**Y -> simulate -> reduce over observables x masked time points; NaN -> -inf.

Third-party arithmetic that is ABSENT here: ``pysb.simulator.ScipyOdeSimulator`` + BioNetGen
(version unpinned anywhere in the reference).  Its published algorithm is: generate
the mass-action network (rate = k * prod(reactants), 1/m! for m identical reactants), integrate
dy/dt = S v(y) with scipy (VODE/BDF rtol 1e-6 atol 1e-12 by default, LSODA when asked).  This
stand-in integrates the same equations with scipy's LSODA at rtol = atol = 1e-12, and is pinned
by the one golden trajectory the reference ships: ``jupyter_notebooks/model_0_AB_complex_data.npy``
(fixture tests/golden/ref_ode.npz; reproduced to < 1e-6 relative, the simulator's own tolerance).
"""
import numpy as np
from scipy.integrate import odeint


def rhs_jac(desc):
    ns = len(desc["species"])
    reactions = desc["reactions"]

    def rates(y, p):
        return np.array([sf * p[k] * np.prod([y[s] for s in reac]) for k, reac, prod, sf in reactions])

    def f(y, t, p):
        v = rates(y, p)
        out = np.zeros(ns)
        for i, (k, reac, prod, sf) in enumerate(reactions):
            for s in reac:
                out[s] -= v[i]
            for s in prod:
                out[s] += v[i]
        return out

    def jac(y, t, p):
        J = np.zeros((ns, ns))
        for k, reac, prod, sf in reactions:
            for q in set(reac):
                others = list(reac)
                others.remove(q)
                dv = sf * reac.count(q) * p[k] * np.prod([y[s] for s in others])
                for s in reac:
                    J[s, q] -= dv
                for s in prod:
                    J[s, q] += dv
        return J
    return f, jac


def simulate(desc, params, tspan, rtol=1e-12, atol=1e-12):
    """Species trajectories [nt, ns] and observables [nt, nobs]."""
    f, jac = rhs_jac(desc)
    ns = len(desc["species"])
    y0 = np.zeros(ns)
    for s, (kind, v) in desc["initials"].items():
        y0[s] = params[v] if kind == "p" else v
    y = odeint(f, y0, tspan, args=(params,), Dfun=jac, rtol=rtol, atol=atol, mxstep=500000)
    obs = np.stack([sum(c * y[:, s] for s, c in o) for o in desc["observables"]], axis=1)
    return y, obs


def loglike(desc, sampled_idx, Y, tspan, rows, kind="snlpdf", log10=True):
    """rows: list of (observable index, data[nt], sd[nt], mask[nt] in {0,1}) on the full tspan grid."""
    params = np.array(desc["param_values"], dtype=np.float64)
    params[sampled_idx] = 10.0 ** np.asarray(Y) if log10 else np.asarray(Y)
    _, obs = simulate(desc, params, tspan)
    logl = 0.0
    for o, data, sd, mask in rows:
        m = np.asarray(mask) != 0
        sim = obs[m, o]
        if kind == "snlpdf":
            z = (sim - data[m]) / sd[m]
            logl += np.sum(-0.5 * z * z - np.log(sd[m]) - 0.5 * np.log(2 * np.pi))
        elif kind == "mse":
            logl -= np.mean((data[m] - sim) ** 2)
        else:
            logl -= np.sum((data[m] - sim) ** 2)
    return -np.inf if np.isnan(logl) else float(logl)
