"""CPU side of the JARM path (no GPU): the elimination plan and index tables of the cooperative
solver executed with numpy, the oracle's restatement of the multi-condition likelihood protocol
(examples/pysb_jarm/run_jarm_multinest.py:72-143) pinned to scipy.stats and to a literal
transcription of the script's cycle products, and the problem builder."""
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.stats import halfnorm, norm

from gleipnir_b200 import jarm, networks
from oracle import ode


@pytest.mark.parametrize("name", ["jarm", "jarm_like", "dimerization", "model_0"])
def test_elimination_plan_solves_like_numpy(name):
    """odeplan.CoopPlan.emulate_solve runs the generated LU / substitution schedule with numpy lane vectors."""
    net = networks.SHIPPED[name]()
    d = net.description()
    f, jac = ode.rhs_jac(d)
    pl = net.plan()
    assert pl.nc + pl.nx == net.ns and pl.nc <= 32
    rng = np.random.default_rng(3)
    for trial in range(4):
        y = rng.uniform(0.01, 2.0, net.ns)
        p = d["param_values"] * 10.0 ** rng.uniform(-1, 1, net.np_)
        J = jac(y, 0.0, p)
        diag = 10.0 ** rng.uniform(0, 4)
        b = rng.normal(size=net.ns)
        x = pl.emulate_solve(J, diag, b)
        want = np.linalg.solve(diag * np.eye(net.ns) - J, b)
        np.testing.assert_allclose(x, want, rtol=1e-9, atol=1e-12)


def test_jarm_plan_shape():
    pl = networks.jarm().plan()
    assert (pl.ns, pl.nc, pl.nx, pl.gl) == (34, 32, 2, 32)
    # leaves are mutually independent and are eliminated without fill outside their neighbourhoods
    for l in range(pl.nx):
        assert len(pl.leaf_rows[l]) <= 4 and len(pl.leaf_cols[l]) <= 4
    assert pl.nnz_u < 200 and pl.nnz_l < 200                    # sparse: a dense 32 x 32 factor has 496 each


@pytest.mark.parametrize("name", ["jarm", "jarm_like", "dimerization", "model_0"])
def test_index_tables_reproduce_rhs_and_jacobian(name):
    """The signed value tables + index lists the device reads (MassActionNetwork.coop_tables) give f and J."""
    net = networks.SHIPPED[name]()
    d = net.description()
    f, jac = ode.rhs_jac(d)
    rng = np.random.default_rng(0)
    y = rng.uniform(0.01, 2.0, net.ns)
    p = d["param_values"] * rng.uniform(0.5, 2, net.np_)
    fe, Je = net.emulate_rhs_jac(y, p)
    np.testing.assert_allclose(fe, f(y, 0, p), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Je, jac(y, 0, p), rtol=1e-13, atol=0)
    T = net.coop_tables()
    assert T["ld"] >= net.ns                                      # stride of the leaf rows kept in shared memory
    # the column lists cover every structural entry of the core rows once: depth 1 off the diagonal for these networks
    assert len(T["mcols"]) == net.ns and len(T["mtab"]) == sum(dp for _, dp in T["mcols"]) * T["gl"]


def test_vectorised_oracle_rhs_equals_loop_form():
    d = networks.jarm().description()
    f, jac = ode.rhs_jac(d)
    f2, jac2 = ode._rhs_jac_loops(d)
    rng = np.random.default_rng(5)
    y, p = rng.uniform(0, 1, 34), d["param_values"] * rng.uniform(0.1, 10, 50)
    np.testing.assert_allclose(f(y, 0, p), f2(y, 0, p), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(jac(y, 0, p), jac2(y, 0, p), rtol=1e-12, atol=1e-14)


def test_jarm_network_matches_the_script_indices():
    """Everything run_jarm_multinest.py addresses by index exists with that meaning."""
    net = networks.jarm()
    assert net.np_ == 50 and net.ns == 34 and len(net.reactions) == 84
    assert networks.JARM_CALIBRATED == [1, 5, 9, 11, 15, 17, 19, 23, 25, 27, 31, 35, 36, 37, 38, 39, 41, 43]   # :51
    assert [net.param_names[i] for i in networks.JARM_KCAT] == ["kcat_pTyr", "kcat_pThr"]                     # :59
    assert net.param_names[44] == "Arr_0" and net.initials["Arr"] == "Arr_0"                                 # :56
    assert [net.param_names[i] for i in networks.JARM_JNK3_INITIALS] == ["Juu_0", "JpT_0", "JpY_0"]            # :58
    assert all(net.param_names[i].startswith("kf_") and net.param_names[i + 1].startswith("kr_") for i in range(0, 36, 2))
    # total JNK3 of the arrestin condition = 0.6 (:57) and of the no-arrestin condition (:111)
    v = net.param_values
    assert abs(v[47] + v[48] + v[49] - 0.6) < 1e-12 and abs(sum(networks.JARM_NOARRESTIN_JNK3) - 0.6) < 1e-12
    # mass conservation: d/dt of total JNK3, MKK4, MKK7, arrestin is zero for any state
    d = net.description()
    f, _ = ode.rhs_jac(d)
    rng = np.random.default_rng(1)
    y = rng.uniform(0, 1, net.ns)
    dy = f(y, 0.0, v)
    for tag in ("J", "K4", "K7", "Arr"):
        idx = [i for i, s in enumerate(net.species) if tag in s.replace("Arr", "A" if tag != "Arr" else "Arr").split("_") or
               any(part.startswith(tag) for part in s.split("_"))]
        assert abs(dy[idx].sum()) < 1e-9 * np.abs(dy).max(), tag


def _evaluate_cycles_literal(pars1):
    """run_jarm_multinest.py:72-90, transcribed."""
    box1 = (pars1[21] / pars1[20]) * (pars1[23] / pars1[22]) * (1 / (pars1[1] / pars1[0])) * (1 / (pars1[5] / pars1[4]))
    box2 = (pars1[21] / pars1[20]) * (pars1[25] / pars1[24]) * (1 / (pars1[3] / pars1[2])) * (1 / (pars1[27] / pars1[26]))
    box3 = (pars1[13] / pars1[12]) * (pars1[23] / pars1[22]) * (1 / (pars1[1] / pars1[0])) * (1 / (pars1[15] / pars1[14]))
    box4 = (pars1[7] / pars1[6]) * (pars1[25] / pars1[24]) * (1 / (pars1[3] / pars1[2])) * (1 / (pars1[11] / pars1[10]))
    return np.array([box1, box2, box3, box4])


def test_cycle_terms_follow_the_script():
    """The protocol's cycle term == like_thermobox.logpdf(evaluate_cycles(pars)).sum() (:44, 101)."""
    like = jarm.jarm_loglikelihood()
    proto = dict(like.oracle_protocol())
    desc = like.network.description()
    proto["rows"], proto["preeq"] = [], None                    # no data terms, no solves: cycles only
    rng = np.random.default_rng(2)
    for trial in range(5):
        Y = np.log10(desc["param_values"][networks.JARM_CALIBRATED]) + rng.uniform(-1, 1, 18)
        pars = np.array(desc["param_values"])
        pars[networks.JARM_CALIBRATED] = 10.0 ** Y
        want = norm(loc=1, scale=1e-0).logpdf(_evaluate_cycles_literal(pars)).sum()
        got = ode.protocol_loglike(desc, proto, Y)
        assert np.isclose(got, want, rtol=1e-12)
    # the synthetic network's nominal constants satisfy detailed balance on the four cycles
    np.testing.assert_allclose(_evaluate_cycles_literal(desc["param_values"]), 1.0, rtol=1e-12)


def test_protocol_data_terms_follow_scipy():
    """Rows of the protocol == the script's norm / halfnorm logpdf sums (:126-134) on the oracle's own trajectories."""
    like = jarm.jarm_loglikelihood()
    proto, desc = like.oracle_protocol(), like.network.description()
    exp = jarm.read_exp_data()
    Y = np.log10(desc["param_values"][networks.JARM_CALIBRATED])
    total, sims = ode.protocol_loglike(desc, proto, Y, detail=True)
    mask = np.array([i for i, t in enumerate(proto["tspan"]) if t in exp["Time (secs)"]])
    s0, s1 = sims[0], sims[1]
    want = np.sum(norm(loc=exp["pTyr_arrestin_avg"], scale=exp["pTyr_arrestin_std"]).logpdf(s0[mask, 0] / 0.6))
    want += np.sum(norm(loc=exp["pThr_arrestin_avg"], scale=exp["pThr_arrestin_std"]).logpdf(s0[mask, 1] / 0.6))
    want += np.sum(norm(loc=exp["pTyr_noarrestin_avg"][4:], scale=exp["pTyr_noarrestin_std"][4:]).logpdf(s1[mask, 0][4:] / 0.6))
    want += np.sum(halfnorm(loc=exp["pTyr_noarrestin_avg"][:4], scale=exp["pTyr_noarrestin_std"][:4]).logpdf(s1[mask, 0][:4] / 0.6))
    want += np.sum(norm(loc=exp["pThr_noarrestin_avg"], scale=exp["pThr_noarrestin_std"]).logpdf(s1[mask, 1] / 0.6))
    want += norm(loc=1, scale=1.0).logpdf(_evaluate_cycles_literal(desc["param_values"])).sum()
    assert np.isfinite(total) and np.isclose(total, want, rtol=1e-12)
    # below the half-normal's location the density is zero: -inf, as scipy gives it
    lowered = dict(proto)
    lowered["rows"] = [dict(r) for r in proto["rows"]]
    lowered["rows"][3]["data"] = lowered["rows"][3]["data"] + 1.0
    assert ode.protocol_loglike(desc, lowered, Y) == -np.inf


def test_pre_equilibration_semantics():
    """Catalysis is off during pre-equilibration (:114) and on afterwards; the no-arrestin condition starts from
    its own initial amounts (:110-111); eq_tol plays the role of the external pre_equilibration's failure."""
    like = jarm.jarm_loglikelihood()
    proto, desc = like.oracle_protocol(), like.network.description()
    Y = np.log10(desc["param_values"][networks.JARM_CALIBRATED])
    _, sims = ode.protocol_loglike(desc, proto, Y, detail=True)
    # t = 0 of the main solve is the equilibrated state: nothing is phosphorylated beyond the initial pY pool
    assert abs(sims[0][0, 1]) < 1e-12 and abs(sims[1][0, 1]) < 1e-12          # no pThr before catalysis starts
    assert abs(sims[0][0, 2] - 0.6) < 1e-9 and abs(sims[1][0, 2] - 0.6) < 1e-9
    assert abs(sims[1][0, 0] - 0.007158512) < 1e-9                            # no-arrestin initial pTyr pool
    assert sims[0][-1, 1] > 0.1                                                # catalysis is back on in the main solve
    strict = dict(proto)
    strict["preeq"] = dict(proto["preeq"], eq_tol=1e-30)
    assert ode.protocol_loglike(desc, strict, Y) == -np.inf
    loose = dict(proto)
    loose["preeq"] = dict(proto["preeq"], eq_tol=1e-3)
    assert np.isfinite(ode.protocol_loglike(desc, loose, Y))


def test_problem_builder_mirrors_the_script():
    sps, like = jarm.jarm_problem()
    assert len(sps) == 18 and like.ndim == 18                                   # :63
    lo, width = np.log10(5e-8), np.log10(1.9e3) - np.log10(5e-8)
    for i, sp in enumerate(sps):
        loc, scale = sp.prior_dist.args if sp.prior_dist.args else (sp.prior_dist.kwds["loc"], sp.prior_dist.kwds["scale"])
        if i == 0:
            assert np.isclose(loc, np.log10(120)) and np.isclose(scale, 1.0)    # :68
        elif i == 6:
            assert np.isclose(loc, np.log10(28)) and np.isclose(scale, 1.0)     # :69
        else:
            assert np.isclose(loc, lo) and np.isclose(scale, width)
    assert len(like.tspan) == 181 and like.tspan[-1] == 180.0                   # :26
    assert [r["kind"] for r in like.rows] == ["norm", "norm", "norm", "halfnorm", "norm"]
    assert [int(r["mask"].sum()) for r in like.rows] == [16, 16, 12, 4, 16]     # :37-40 split at index 4
    assert all(r["scale"] == 0.6 for r in like.rows)
    b = like._blob
    assert (b[0], b[1], b[2], b[7], b[8], b[9], b[10], b[12], b[13]) == (18, 50, 181, 5, 2, 1, 100.0, 2, 4)
    # the nominal parameters sit inside the prior box and inside the half-normal's support
    nominal = np.log10(like.network.param_values[networks.JARM_CALIBRATED])
    for y, (loc, scale) in zip(nominal, jarm.jarm_prior_table()):
        assert loc <= y <= loc + scale
    assert np.isfinite(ode.protocol_loglike(like.network.description(), like.oracle_protocol(), nominal))


def test_data_file_is_what_the_generator_writes(tmp_path):
    """tools/make_jarm_data.py is the committed provenance of data/jarm_exp_data_3min.csv."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    before = open(jarm.DATA_CSV).read()
    subprocess.run(["python", os.path.join(here, "tools", "make_jarm_data.py")], check=True, capture_output=True)
    assert open(jarm.DATA_CSV).read() == before


def test_cooperative_plugin_exports_counters():
    from gleipnir_b200 import odegen
    so = odegen.compile_plugin(networks.jarm())
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    for sym in ("gl_plugin_walk_launch", "gl_plugin_eval_launch", "gl_plugin_last_error", "gl_plugin_info",
                "gl_plugin_counters", "gl_plugin_trace"):
        assert re.search(r" T %s\b" % sym, out), sym
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "DFMA" in sass and "SHFL.IDX" in sass


def test_generated_sources_do_not_depend_on_the_hash_seed():
    """The plugin file name is a hash of the generated source: it must be the same in every process, or the prebuilt
    plugins shipped to the GPU box are ignored and rebuilt there."""
    import sys
    code = ("from gleipnir_b200 import networks; "
            "print([mk().source_hash() for mk in networks.SHIPPED.values()])")
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = set()
    for seed in ("0", "1", "77"):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        outs.add(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True, cwd=here,
                                env=env).stdout.strip())
    assert len(outs) == 1, outs


def _random_network(rng, ns, nr, name):
    """Random uni-/bimolecular mass-action network; duplicate reactant pairs on purpose, so that Jacobian entries
    off the diagonal get several terms (the shipped networks only have single-term ones)."""
    from gleipnir_b200.odegen import MassActionNetwork
    sp = ["S%d" % i for i in range(ns)]
    par, rx = {}, []
    pairs = []
    for r in range(nr):
        par["k%d" % r] = float(10.0 ** rng.uniform(-2, 1))
        if pairs and rng.random() < 0.3:
            reac = list(pairs[rng.integers(len(pairs))])            # same reactants, other products
        elif rng.random() < 0.5:
            reac = [sp[rng.integers(ns)]]
        else:
            reac = [sp[rng.integers(ns)], sp[rng.integers(ns)]]      # may be a homodimerisation
        pairs.append(tuple(reac))
        prod = [sp[i] for i in rng.choice(ns, size=rng.integers(1, 3), replace=False)]
        rx.append(("k%d" % r, reac, prod))
    par["S0_0"] = 3.0
    return MassActionNetwork(name, species=sp, parameters=par, reactions=rx, initials={"S0": "S0_0", "S1": 1.5},
                             observables={"o": {sp[0]: 1, sp[-1]: 2}})


@pytest.mark.parametrize("ns,nr,seed", [(6, 9, 0), (12, 30, 1), (16, 40, 2), (17, 30, 3), (28, 70, 4), (32, 90, 5),
                                         (33, 60, 6), (35, 70, 7)])
def test_generated_tables_and_plan_on_random_networks(ns, nr, seed):
    """coop_tables (column lists of M, diagonal lists, leaf rows) and the elimination plan on random networks:
    f and J equal the oracle's, (diag*I - J) x = b solved with the plan equals numpy.linalg.solve, and the generated
    CUDA source mentions every structural column.  Networks above 32 species need independent leaf species; a random
    network may not have them, which the plan must say loudly."""
    rng = np.random.default_rng(seed)
    net = _random_network(rng, ns, nr, "rnd%d" % seed)
    d = net.description()
    f, jac = ode.rhs_jac(d)
    y = rng.uniform(0.05, 2.0, ns)
    p = d["param_values"] * rng.uniform(0.5, 2.0, net.np_)
    try:
        pl = net.plan()
    except ValueError as exc:
        assert ns > 32 and "leaf" in str(exc)
        return
    fe, Je = net.emulate_rhs_jac(y, p)
    np.testing.assert_allclose(fe, f(y, 0, p), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Je, jac(y, 0, p), rtol=1e-12, atol=1e-13)
    T = net.coop_tables()
    assert pl.gl == (16 if pl.nc <= 16 else 32) and T["gl"] == pl.gl
    assert max(dp for _, dp in T["mcols"]) >= 1 and len(T["dtab"]) == T["ed"] * pl.gl
    J = jac(y, 0, p)
    diag = 1.0 / (0.25 * 0.05)
    b = rng.normal(size=ns)
    x = pl.emulate_solve(J, diag, b)
    assert x is not None
    np.testing.assert_allclose(x, np.linalg.solve(diag * np.eye(ns) - J, b), rtol=1e-9, atol=1e-12)
    src = net.plugin_source()
    assert "load_m" in src and src.count("M[%d] = " % (pl.nc - 1)) >= 1
