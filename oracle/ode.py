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
    """dy/dt = S v(y) and its analytic Jacobian, numpy-vectorised over reactions (the form PySB's
    lambdified right-hand side takes): v_r = sf_r p[k_r] prod(reactants)."""
    ns = len(desc["species"])
    reactions = desc["reactions"]
    nr = len(reactions)
    if any(not 1 <= len(r[1]) <= 2 for r in reactions):
        return _rhs_jac_loops(desc)
    ka = np.array([r[0] for r in reactions], dtype=int)
    sf = np.array([r[3] for r in reactions], dtype=float)
    a = np.array([r[1][0] for r in reactions], dtype=int)
    bi = np.array([r[1][1] if len(r[1]) > 1 else -1 for r in reactions], dtype=int)
    has_b = bi >= 0
    b = np.where(has_b, bi, 0)
    same = has_b & (a == bi)
    S = np.zeros((ns, nr))
    for i, (k, reac, prod, _) in enumerate(reactions):
        for s in reac:
            S[s, i] -= 1.0
        for s in prod:
            S[s, i] += 1.0
    Ea = np.zeros((nr, ns)); Ea[np.arange(nr), a] = 1.0            # scatter d v / d y_a
    Eb = np.zeros((nr, ns)); Eb[np.arange(nr)[has_b & ~same], bi[has_b & ~same]] = 1.0

    def f(y, t, p):
        yb = np.where(has_b, y[b], 1.0)
        return S @ (sf * p[ka] * y[a] * yb)

    def jac(y, t, p):
        c = sf * p[ka]
        da = np.where(has_b, np.where(same, 2.0 * c * y[a], c * y[b]), c)
        db = np.where(has_b & ~same, c * y[a], 0.0)
        return S @ (da[:, None] * Ea + db[:, None] * Eb)
    return f, jac


def _rhs_jac_loops(desc):
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


# ---- multi-condition protocol (the JARM likelihood) ----------------------------------------------
def protocol_loglike(desc, proto, Y, rtol=1e-12, atol=1e-12, detail=False):
    """Restates ``examples/pysb_jarm/run_jarm_multinest.py:92-143`` for an explicit network:

      :94      params[mask] = 10**Y
      :72-90   cycle products ("thermoboxes"), :101-103 sum of norm(loc, scale).logpdf, NaN -> -inf
      :105-119 pre-equilibration: per condition the parameter overrides (:110-111) with the
               catalytic constants zeroed (:114), integrated on linspace(0, t_end, n) -- the state
               at t_end is taken as the equilibrium (the reference's external
               `equilibration_function.pre_equilibration` is absent from the tree; it returns the
               state once every species' finite-difference derivative is below 1e-6, else raises;
               here an optional `eq_tol` on max|dy/dt| at t_end plays that role); exception -> -inf
      :121-125 main solve per condition from the equilibrated state, catalysis restored
      :126-134 rows: norm / halfnorm logpdf of observable[mask] / scale
      :137-143 sum, NaN -> -inf

    proto: dict(sampled_idx, log10, tspan, conditions=[[(param index, value), ...], ...],
    preeq=None | dict(t_end, n_points, zero_idx, eq_tol), cycles=None | dict(loc, scale, factors),
    rows=[dict(cond, obs, scale, kind, data[nt], sd[nt], mask[nt])])."""
    base = np.array(desc["param_values"], dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    base[proto["sampled_idx"]] = 10.0 ** Y if proto.get("log10", True) else Y
    f, jac = rhs_jac(desc)
    ns = len(desc["species"])
    tspan = np.asarray(proto["tspan"], dtype=np.float64)
    total = 0.0
    cyc = proto.get("cycles")
    with np.errstate(all="ignore"):
        if cyc:
            boxes = []
            for factors in cyc["factors"]:
                b = 1.0
                for a, bb, inv in factors:
                    t = base[a] / base[bb]
                    b = b * ((1.0 / t) if inv else t)
                boxes.append(b)
            z = (np.array(boxes) - cyc["loc"]) / cyc["scale"]
            lp = np.sum(-0.5 * z * z - np.log(cyc["scale"]) - 0.5 * np.log(2 * np.pi))
            total += -np.inf if np.isnan(lp) else lp
    sims = []
    for c, overrides in enumerate(proto["conditions"]):
        p = base.copy()
        for i, v in overrides:
            p[i] = v
        y0 = np.zeros(ns)
        for s, (kind, v) in desc["initials"].items():
            y0[s] = p[v] if kind == "p" else v
        pre = proto.get("preeq")
        if pre:
            pe = p.copy()
            pe[pre["zero_idx"]] = 0.0
            te = np.linspace(0.0, pre["t_end"], pre["n_points"])
            ye, info = odeint(f, y0, te, args=(pe,), Dfun=jac, rtol=rtol, atol=atol, mxstep=500000, full_output=True)
            if info["message"] != "Integration successful." or not np.all(np.isfinite(ye[-1])):
                return -np.inf
            y0 = ye[-1]
            tol = pre.get("eq_tol")
            if tol is not None and np.max(np.abs(f(y0, 0.0, pe))) > tol:
                return -np.inf
        rows = [r for r in proto["rows"] if r["cond"] == c]
        last = max([int(np.nonzero(np.asarray(r["mask"]))[0].max()) for r in rows] + [0])
        ts = tspan[:last + 1]
        if len(ts) > 1:
            y, info = odeint(f, y0, ts, args=(p,), Dfun=jac, rtol=rtol, atol=atol, mxstep=500000, full_output=True)
            if info["message"] != "Integration successful.":
                return -np.inf
        else:
            y = y0[None, :]
        obs = np.stack([sum(cf * y[:, s] for s, cf in o) for o in desc["observables"]], axis=1)
        sims.append(obs)
        with np.errstate(all="ignore"):
            for r in rows:
                m = np.asarray(r["mask"])[:last + 1] != 0
                x = obs[m, r["obs"]] / r["scale"]
                mu, sd = np.asarray(r["data"])[:last + 1][m], np.asarray(r["sd"])[:last + 1][m]
                z = (x - mu) / sd
                if r["kind"] == "halfnorm":               # scipy halfnorm(loc, scale).logpdf
                    lp = np.where(z >= 0.0, 0.5 * np.log(2.0 / np.pi) - np.log(sd) - 0.5 * z * z, -np.inf)
                else:
                    lp = -0.5 * z * z - np.log(sd) - 0.5 * np.log(2 * np.pi)
                total += np.sum(lp)
    total = -np.inf if np.isnan(total) else float(total)
    return (total, sims) if detail else total
