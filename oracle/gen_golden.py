"""TEST INFRASTRUCTURE ONLY -- records golden vectors from the REAL reference.

Run in the build container only (needs /root/reference, which does not travel):
    python oracle/gen_golden.py
Writes tests/golden/ref_runs.npz, ref_replay.npz, ref_priors.npz (and, with `stats` as argument,
ref_stats.npz: ~10 CPU-minutes on 8 cores).  The committed
.npz files are what tests read; this script is committed so they can be regenerated.

Recorded:
  ref_runs     full ``NestedSampling.run()`` outputs of the unmodified reference (pandas
               shim only, see oracle/ref_shim.py) for seeded configs, incl. the SURVEY 6.2
               anchor (seed 0, tests/test_nested_sampling.py config).
  ref_replay   outer-loop replay cases: initial live points (seeded), a pre-generated
               stream of replacement (point, logL) fed through a stub sampler, and the
               reference's dead table / Z / H / information -- with ties, -inf and NaN.
  ref_priors   scipy.stats pdf/logpdf/ppf tables + SampledParameter known answers.
  ref_stats    10-seed ln Z of the unmodified reference on config C1 (d = 2 and 5 Gaussians, N = 500,
               NumberOfIterations(5000), sampler(10, tc=1): SURVEY 8d) with its reported errors -- the
               statistical-parity fixture north_star asks for ("within its reported +-sigma over 10 seeds").
"""
import os
import sys

import numpy as np
from scipy.stats import loguniform, norm, uniform

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _gauss(d):
    const = -0.5 * np.log(2 * np.pi)
    return lambda x: -0.5 * np.sum(x ** 2) + d * const


def _egg(x):
    return (2.0 + (np.cos(x)).prod()) ** 5


def full_runs(m):
    SP = m["sampled_parameter"].SampledParameter
    Sampler = m["samplers"].MetropolisComponentWiseHardNSRejection
    NOI = m["stopping_criterion"].NumberOfIterations
    RPM = m["stopping_criterion"].RemainingPriorMass
    NS = m["nested_sampling"].NestedSampling
    cases = {
        # name: (seed, d, prior ctor, loglike, N, sampler kwargs, stop)
        "anchor_gauss5_N20": (0, 5, lambda: uniform(loc=-5.0, scale=10.0), _gauss(5), 20,
                              dict(iterations=10, tuning_cycles=1), NOI(120)),
        "gauss2_N50_tc0": (3, 2, lambda: uniform(loc=-5.0, scale=10.0), _gauss(2), 50,
                           dict(iterations=5, tuning_cycles=0), NOI(200)),
        "egg_N40_rpm": (7, 2, lambda: uniform(loc=0.0, scale=10.0 * np.pi), _egg, 40,
                        dict(iterations=8, burn_in=2, tuning_cycles=2), RPM(1e-2)),
        "gauss3_normprior_normprop": (11, 3, lambda: norm(0.5, 2.0), _gauss(3), 30,
                                      dict(iterations=6, tuning_cycles=1, proposal="normal"), NOI(90)),
    }
    out = {}
    for name, (seed, d, mk, ll, N, skw, stop) in cases.items():
        np.random.seed(seed)
        sps = [SP(name=i, prior=mk()) for i in range(d)]
        ns = NS(sps, ll, N, sampler=Sampler(**skw), stopping_criterion=stop)
        logz, err = ref_shim.run_reference(ns, m["nested_sampling"])
        out[name + "/logZ"] = float(logz)
        out[name + "/logZ_err"] = float(err)
        out[name + "/evidence"] = float(ns.evidence)
        out[name + "/evidence_error"] = float(ns.evidence_error)
        out[name + "/information"] = float(ns.information)
        out[name + "/dead"] = ns.dead_points.to_numpy(dtype=float)
        out[name + "/widths"] = np.asarray(ns.sampler._widths, dtype=float)
        post = ns.posteriors()
        keys = list(post.keys())
        out[name + "/post_marginal0"] = post[keys[0]][0]
        out[name + "/post_edges0"] = post[keys[0]][1]
        out[name + "/max_loglikelihood"] = float(ns.max_loglikelihood())
        out[name + "/akaike_ic"] = float(ns.akaike_ic())
        out[name + "/bayesian_ic5"] = float(ns.bayesian_ic(5))
        out[name + "/deviance_ic"] = float(ns.deviance_ic())
        out[name + "/best_fit_likelihood"] = np.asarray(ns.best_fit_likelihood(), dtype=float)
        print(name, logz, err, ns.dead_points.shape)
    np.savez_compressed(os.path.join(OUT, "ref_runs.npz"), **out)


class _StubSampler(object):
    """Feeds a pre-generated stream through the reference's sampler protocol
    (samplers.py:56): __call__(sampled_parameters, loglikelihood, start, boundary)."""

    def __init__(self, points, logls):
        self.points, self.logls, self.i = points, logls, 0

    def __call__(self, sampled_parameters, loglikelihood, start_param_vec, ns_boundary, **kw):
        p, l = self.points[self.i], self.logls[self.i]
        self.i += 1
        return p.copy(), l


def replay_cases(m):
    SP = m["sampled_parameter"].SampledParameter
    NOI = m["stopping_criterion"].NumberOfIterations
    NS = m["nested_sampling"].NestedSampling
    out = {}
    rng = np.random.default_rng(2024)
    specs = {
        # name: (N, d, n_iter, mode)
        "plain": (64, 3, 300, "plain"),
        "ties": (48, 2, 200, "ties"),          # quantised logL -> exact ties resolved by lowest index
        "neginf": (32, 2, 120, "neginf"),      # -inf logL entries (ODE-failure semantics)
        "nan": (16, 2, 12, "nan"),             # NaN-first argmin semantics
        "N1000": (1000, 2, 1500, "plain"),
    }
    for name, (N, d, n_iter, mode) in specs.items():
        np.random.seed(100 + N)
        sps = [SP(name="p%d" % i, prior=uniform(loc=-5.0, scale=10.0)) for i in range(d)]
        init_table = {}

        if mode == "ties":
            ll0 = lambda x: float(np.floor(-0.5 * np.sum(x ** 2) * 2.0) / 2.0)
        elif mode == "neginf":
            ll0 = lambda x: (-np.inf if x[0] < -2.5 else -0.5 * float(np.sum(x ** 2)))
        elif mode == "nan":
            ll0 = lambda x: (np.nan if x[0] > 4.0 else -0.5 * float(np.sum(x ** 2)))
        else:
            ll0 = lambda x: -0.5 * float(np.sum(x ** 2))

        pts = rng.uniform(-5, 5, size=(n_iter, d))
        # replacement logLs: drift upward like a real run, plus the mode's special values
        base = np.sort(rng.uniform(-25.0, 0.0, size=n_iter))
        if mode == "ties":
            base = np.floor(base * 2.0) / 2.0
        if mode == "neginf":
            base[rng.integers(0, n_iter, size=10)] = -np.inf
        if mode == "nan":
            base[5] = np.nan
        stub = _StubSampler(pts, base)
        ns = NS(sps, ll0, N, sampler=stub, stopping_criterion=NOI(n_iter))
        ref_shim.run_reference(ns, m["nested_sampling"])
        dead = ns.dead_points.to_numpy(dtype=float)
        # the initial live set = what rvs produced; recover it by re-seeding
        np.random.seed(100 + N)
        _ = [SP(name="p%d" % i, prior=uniform(loc=-5.0, scale=10.0)) for i in range(d)]
        live0 = np.empty((N, d))
        for i in range(N):
            for j in range(d):
                live0[i, j] = sps[j].rvs(1)[0]
        out[name + "/live0"] = live0
        out[name + "/logl0"] = np.array([ll0(r) for r in live0])
        out[name + "/repl_x"] = pts[: stub.i]
        out[name + "/repl_l"] = base[: stub.i]
        out[name + "/dead"] = dead
        out[name + "/evidence"] = float(ns.evidence)
        out[name + "/H"] = float(ns._H)
        out[name + "/information"] = float(ns.information)
        out[name + "/logZ"] = float(ns.log_evidence)
        out[name + "/logZ_err"] = float(ns.log_evidence_error)
        out[name + "/n_iter"] = n_iter
        print("replay", name, dead.shape, ns.log_evidence, ns.information)
    np.savez_compressed(os.path.join(OUT, "ref_replay.npz"), **out)


def prior_tables(m):
    SP = m["sampled_parameter"].SampledParameter
    out = {}
    xs = np.linspace(-7.0, 7.0, 57)
    us = np.concatenate(([1e-300, 1e-16, 1e-9], np.linspace(0.001, 0.999, 41), [1 - 1e-9, 1 - 1e-16]))
    fam = {
        "uniform": (uniform(loc=-5.0, scale=10.0), xs),
        "norm": (norm(0.5, 2.0), xs),
        "loguniform": (loguniform(1e-3, 1e2), np.concatenate(([5e-4, 1e-3], np.logspace(-3, 2, 41), [1e2, 2e2]))),
    }
    with np.errstate(all="ignore"):
        for name, (dist, x) in fam.items():
            sp = SP(name, dist)
            out[name + "/x"] = x
            out[name + "/prior"] = np.array([sp.prior(v) for v in x])
            out[name + "/logprior"] = np.array([sp.logprior(v) for v in x])
            out[name + "/u"] = us
            out[name + "/invcdf"] = np.array([sp.invcdf(v) for v in us])
    sp = SP("sample", norm(0.0, 1.0))      # tests/test_sampled_parameter.py:23-36
    out["known/logprior_0.5"] = float(sp.logprior(0.5))
    out["known/prior_0.5"] = float(sp.prior(0.5))
    out["known/invcdf_0.5"] = float(sp.invcdf(0.5))
    np.savez_compressed(os.path.join(OUT, "ref_priors.npz"), **out)


def ode_fixtures():
    """The reference's ODE data files, copied verbatim as fixtures (they are data, not code):
    model_0_AB_complex_data.npy is the one genuine ScipyOdeSimulator trajectory in the tree
    (examples/HypSelector/grouped_reactions/run_grouped_reactions_nsclassic.py:10-11,27: model_0 at
    default parameters, tspan = linspace(0, 5, 20), last 10 points); the dimerization pair is the
    likelihood input of examples/pysb_dimerization_model/run_NS_dimerization_model.py:22-24."""
    ref = ref_shim.REFERENCE_ROOT
    out = {
        "model_0/AB_complex_last10": np.load(os.path.join(ref, "jupyter_notebooks", "model_0_AB_complex_data.npy")),
        "model_0/tspan": np.linspace(0, 5, 20),
        "dimerization/data": np.load(os.path.join(ref, "examples", "pysb_dimerization_model",
                                                  "dimerization_model_dimer_data.npy")),
        "dimerization/sd": np.load(os.path.join(ref, "examples", "pysb_dimerization_model",
                                                "dimerization_model_dimer_sd.npy")),
        "dimerization/tspan": np.linspace(0, 1, num=51),
    }
    np.savez_compressed(os.path.join(OUT, "ref_ode.npz"), **out)
    print("ode fixtures", {k: v.shape for k, v in out.items()})


def _stat_one(args):
    d, seed = args
    m = ref_shim.load_reference()
    SP = m["sampled_parameter"].SampledParameter
    Sampler = m["samplers"].MetropolisComponentWiseHardNSRejection
    NOI = m["stopping_criterion"].NumberOfIterations
    NS = m["nested_sampling"].NestedSampling
    np.random.seed(seed)
    sps = [SP(name=i, prior=uniform(loc=-5.0, scale=10.0)) for i in range(d)]
    ns = NS(sps, _gauss(d), 500, sampler=Sampler(iterations=10, tuning_cycles=1), stopping_criterion=NOI(5000))
    logz, err = ref_shim.run_reference(ns, m["nested_sampling"])
    return d, seed, float(logz), float(err), float(ns.information)


def stat_cases():
    import multiprocessing as mp
    jobs = [(d, seed) for d in (2, 5) for seed in range(10)]
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(_stat_one, jobs, chunksize=1)
    out = {}
    for d in (2, 5):
        rows = sorted(r for r in res if r[0] == d)
        out["gauss%d/seeds" % d] = np.array([r[1] for r in rows])
        out["gauss%d/logZ" % d] = np.array([r[2] for r in rows])
        out["gauss%d/logZ_err" % d] = np.array([r[3] for r in rows])
        out["gauss%d/information" % d] = np.array([r[4] for r in rows])
        print("C1 d=%d: ln Z = %.4f +- %.4f (seed scatter), reported sigma %.4f, truth %.4f"
              % (d, out["gauss%d/logZ" % d].mean(), out["gauss%d/logZ" % d].std(ddof=1),
                 out["gauss%d/logZ_err" % d].mean(), -d * np.log(10.0)))
    np.savez_compressed(os.path.join(OUT, "ref_stats.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "stats" in sys.argv[1:]:
        stat_cases()
        sys.exit(0)
    mods = ref_shim.load_reference()
    full_runs(mods)
    replay_cases(mods)
    prior_tables(mods)
    ode_fixtures()
