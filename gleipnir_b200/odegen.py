"""Code generation for mass-action ODE networks (PySB/BioNetGen-style reaction networks).

``MassActionNetwork`` is the explicit network description a PySB model exports after network
generation: species, parameters, reactions (rate constant x product of reactant amounts, with
BioNetGen's 1/m! statistical factor for m identical reactants), initial conditions and
observables (integer-weighted species sums).  From it this module generates a CUDA translation
unit with straight-line ``rhs`` / analytic ``jac`` code and compiles it (nvcc, sm_100a) into a
plugin library that instantiates the walker and batch-evaluation kernel templates of
``csrc/`` for the network's functor (``csrc/ode_like.cuh``: per-parameter-set Rosenbrock
integrator with the data misfit fused in).

Reference: the likelihood closures over ``ScipyOdeSimulator`` in
``gleipnir/pysb_utilities/nestedsample_it.py:273-339`` and
``examples/pysb_dimerization_model/run_NS_dimerization_model.py:56-65``.
"""
import ctypes
import hashlib
import math
import os
import subprocess

import numpy as np

from . import _lib
from .likelihoods import DeviceLogLikelihood

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
PLUGIN_DIR = os.path.join(HERE, "_plugins")

LIKE_TYPES = {"snlpdf": 0, "mse": 1, "sse": 2}


class MassActionNetwork(object):
    """species: list of names; parameters: {name: nominal value} (ordered);
    reactions: list of (rate_parameter_name, [reactant species...], [product species...]) --
    a species listed twice reacts with itself; optional 4th element overrides the statistical
    factor (default 1/prod(m_s!) over reactant multiplicities, the BioNetGen convention);
    initials: {species: parameter name or number}; observables: {name: {species: coefficient}}."""

    def __init__(self, name, species, parameters, reactions, initials, observables):
        self.name = name
        self.species = list(species)
        self.param_names = list(parameters.keys())
        self.param_values = np.array([float(parameters[k]) for k in self.param_names])
        self.reactions = []
        for r in reactions:
            rate, reac, prod = r[0], list(r[1]), list(r[2])
            if len(r) > 3:
                sf = float(r[3])
            else:
                sf = 1.0
                for sp in set(reac):
                    sf /= math.factorial(reac.count(sp))
            self.reactions.append((rate, reac, prod, sf))
        self.initials = dict(initials)
        self.observables = {k: dict(v) for k, v in observables.items()}
        self.obs_names = list(self.observables.keys())

    # ---- numbers the generated code needs ------------------------------------------------
    @property
    def ns(self):
        return len(self.species)

    @property
    def np_(self):
        return len(self.param_names)

    def _sp(self, name):
        return self.species.index(name)

    def _pp(self, name):
        return self.param_names.index(name)

    def description(self):
        """Plain-data description (what the oracle consumes; no code)."""
        return dict(species=list(self.species), param_names=list(self.param_names),
                    param_values=self.param_values.copy(),
                    reactions=[(self._pp(k), [self._sp(s) for s in re], [self._sp(s) for s in pr], sf)
                               for k, re, pr, sf in self.reactions],
                    initials={self._sp(s): (("p", self._pp(v)) if isinstance(v, str) else ("c", float(v)))
                              for s, v in self.initials.items()},
                    observables=[[(self._sp(s), float(c)) for s, c in self.observables[o].items()]
                                 for o in self.obs_names])

    # ---- CUDA source -----------------------------------------------------------------------
    def cuda_struct(self):
        ns = self.ns
        lines = ["struct Net {",
                 "    static constexpr int NS = %d, NP = %d, NOBS = %d;" % (ns, self.np_, len(self.obs_names))]
        # rate expressions
        rates = []
        for k, reac, prod, sf in self.reactions:
            terms = ["p[%d]" % self._pp(k)] + ["y[%d]" % self._sp(s) for s in reac]
            expr = " * ".join(terms)
            if sf != 1.0:
                expr = "%s * %s" % (repr(sf), expr)
            rates.append(expr)
        lines.append("    static __device__ __forceinline__ void rhs(const double *y, const double *p, double *f) {")
        for i, e in enumerate(rates):
            lines.append("        const double v%d = %s;" % (i, e))
        for s in range(ns):
            terms = []
            for i, (k, reac, prod, sf) in enumerate(self.reactions):
                net = prod.count(self.species[s]) - reac.count(self.species[s])
                if net:
                    terms.append(("+ " if net > 0 else "- ") + ("%d.0 * " % abs(net) if abs(net) != 1 else "") + "v%d" % i)
            lines.append("        f[%d] = %s;" % (s, " ".join(terms).lstrip("+ ") if terms else "0.0"))
        lines.append("    }")
        # Jacobian: d f_s / d y_q
        lines.append("    static __device__ __forceinline__ void jac(const double *y, const double *p, double *J) {")
        dv = {}
        for i, (k, reac, prod, sf) in enumerate(self.reactions):
            for q in set(reac):
                m = reac.count(q)
                others = list(reac)
                others.remove(q)                      # d/dy_q of y_q^m = m*y_q^(m-1)
                terms = ["p[%d]" % self._pp(k)] + ["y[%d]" % self._sp(s) for s in others]
                coef = sf * m
                expr = " * ".join(terms)
                if coef != 1.0:
                    expr = "%s * %s" % (repr(coef), expr)
                dv[(i, self._sp(q))] = expr
        for (i, q), e in dv.items():
            lines.append("        const double d%d_%d = %s;" % (i, q, e))
        for s in range(ns):
            for q in range(ns):
                terms = []
                for i, (k, reac, prod, sf) in enumerate(self.reactions):
                    net = prod.count(self.species[s]) - reac.count(self.species[s])
                    if net and (i, q) in dv:
                        terms.append(("+ " if net > 0 else "- ") + ("%d.0 * " % abs(net) if abs(net) != 1 else "") + "d%d_%d" % (i, q))
                lines.append("        J[%d] = %s;" % (s * ns + q, " ".join(terms).lstrip("+ ") if terms else "0.0"))
        lines.append("    }")
        lines.append("    static __device__ __forceinline__ void initials(const double *p, double *y0) {")
        for s in range(ns):
            v = self.initials.get(self.species[s], 0.0)
            lines.append("        y0[%d] = %s;" % (s, "p[%d]" % self._pp(v) if isinstance(v, str) else repr(float(v))))
        lines.append("    }")
        lines.append("    static __device__ __forceinline__ double observable(int o, const double *y) {")
        lines.append("        switch (o) {")
        for o, name in enumerate(self.obs_names):
            terms = ["%s * y[%d]" % (repr(float(c)), self._sp(s)) if float(c) != 1.0 else "y[%d]" % self._sp(s)
                     for s, c in self.observables[name].items()]
            lines.append("            case %d: return %s;" % (o, " + ".join(terms) if terms else "0.0"))
        lines.append("            default: return 0.0;")
        lines.append("        }")
        lines.append("    }")
        lines.append("};")
        return "\n".join(lines)

    COOP_MIN_SPECIES = 5           # NS <= 4: one system per thread; above: one system per 16/32 lanes

    @property
    def cooperative(self):
        return self.ns >= self.COOP_MIN_SPECIES

    def cuda_tables(self):
        """Table form of the network for the cooperative solver (csrc/ode_coop.cuh)."""
        ns = self.ns
        if ns > 32:
            raise ValueError("networks with more than 32 species are not supported yet")
        for k, reac, prod, sf in self.reactions:
            if not 1 <= len(reac) <= 2:
                raise ValueError("the cooperative solver handles uni- and bimolecular mass-action reactions")
        gl = 16 if ns <= 16 else 32
        rx_a = [self._sp(r[1][0]) for r in self.reactions]
        rx_b = [self._sp(r[1][1]) if len(r[1]) > 1 else -1 for r in self.reactions]
        rx_k = [self._pp(r[0]) for r in self.reactions]
        rx_sf = [r[3] for r in self.reactions]
        row_ptr, row_rx, row_coef = [0], [], []
        for s_ in self.species:
            for i, (k, reac, prod, sf) in enumerate(self.reactions):
                net = prod.count(s_) - reac.count(s_)
                if net:
                    row_rx.append(i)
                    row_coef.append(net)
            row_ptr.append(len(row_rx))
        init_kind, init_val, init_par = [], [], []
        for s_ in self.species:
            v = self.initials.get(s_, 0.0)
            init_kind.append(1 if isinstance(v, str) else 0)
            init_par.append(self._pp(v) if isinstance(v, str) else 0)
            init_val.append(0.0 if isinstance(v, str) else float(v))
        obs_ptr, obs_sp, obs_coef = [0], [], []
        for name in self.obs_names:
            for s_, c in self.observables[name].items():
                obs_sp.append(self._sp(s_))
                obs_coef.append(float(c))
            obs_ptr.append(len(obs_sp))

        def arr(ctype, name, vals):
            body = ", ".join(repr(v) if ctype == "double" else str(v) for v in vals) or "0"
            return "__device__ const %s T_%s[] = {%s};" % (ctype, name, body)

        tables = [arr("int", "rx_a", rx_a), arr("int", "rx_b", rx_b), arr("int", "rx_k", rx_k),
                  arr("double", "rx_sf", rx_sf), arr("int", "row_ptr", row_ptr), arr("int", "row_rx", row_rx),
                  arr("int", "row_coef", row_coef), arr("int", "init_kind", init_kind),
                  arr("double", "init_val", init_val), arr("int", "init_par", init_par),
                  arr("int", "obs_ptr", obs_ptr), arr("int", "obs_sp", obs_sp), arr("double", "obs_coef", obs_coef)]
        acc = []
        for t, name in (("int", "rx_a"), ("int", "rx_b"), ("int", "rx_k"), ("double", "rx_sf"), ("int", "row_ptr"),
                        ("int", "row_rx"), ("int", "row_coef"), ("int", "init_kind"), ("double", "init_val"),
                        ("int", "init_par"), ("int", "obs_ptr"), ("int", "obs_sp"), ("double", "obs_coef")):
            acc.append("    static __device__ __forceinline__ %s %s(int i) { return T_%s[i]; }" % (t, name, name))
        struct = ["struct Net {",
                  "    static constexpr int NS = %d, NP = %d, NR = %d, NOBS = %d, GL = %d;"
                  % (ns, self.np_, len(self.reactions), len(self.obs_names), gl)] + acc + ["};"]
        return "\n".join(tables + struct)

    def plugin_source(self):
        if self.cooperative:
            return COOP_PLUGIN_TEMPLATE % dict(name=self.name, struct=self.cuda_tables())
        return PLUGIN_TEMPLATE % dict(name=self.name, struct=self.cuda_struct())

    def source_hash(self):
        h = hashlib.sha256(self.plugin_source().encode())
        for fn in ("ode_like.cuh", "ode_coop.cuh", "walker.cuh", "loglike.cuh", "likelihoods.cuh", "engine.cuh", "common.cuh",
                   "priors.cuh", "philox.cuh", "select.cuh"):
            with open(os.path.join(CSRC, fn), "rb") as f:
                h.update(f.read())
        return h.hexdigest()[:16]


PLUGIN_TEMPLATE = r'''// GENERATED by gleipnir_b200/odegen.py for network "%(name)s" -- do not edit.
#include "walker.cuh"
#include "loglike.cuh"
#include "ode_like.cuh"

namespace {
%(struct)s
using Like = gl::LikeOde<Net>;
}  // namespace

extern "C" {
GL_API int gl_plugin_walk_launch(const void *P, const void *M, void *stream) {
    gl::WalkLauncher L{*(const gl::WalkParams *)P, *(const gl::ModelDev *)M, (cudaStream_t)stream};
    return L.template operator()<Like>();
}
GL_API int gl_plugin_eval_launch(const void *M, const double *x, long long n, double *out, void *stream) {
    return gl::loglike_rows_launch<Like>(*(const gl::ModelDev *)M, x, n, out, (cudaStream_t)stream, 64);
}
GL_API const char *gl_plugin_last_error(void) { return gl::error_buffer(); }
GL_API void gl_plugin_info(int *ns, int *np, int *nobs) {
    *ns = Net::NS; *np = Net::NP; *nobs = Net::NOBS;
}
}
'''


COOP_PLUGIN_TEMPLATE = r'''// GENERATED by gleipnir_b200/odegen.py for network "%(name)s" (cooperative solver) -- do not edit.
#include "ode_coop.cuh"

namespace {
%(struct)s
}  // namespace

extern "C" {
GL_API int gl_plugin_walk_launch(const void *P, const void *M, void *stream) {
    return gl::walk_coop_launch<Net>(*(const gl::WalkParams *)P, *(const gl::ModelDev *)M, (cudaStream_t)stream);
}
GL_API int gl_plugin_eval_launch(const void *M, const double *x, long long n, double *out, void *stream) {
    return gl::ode_coop_eval_launch<Net>(*(const gl::ModelDev *)M, x, n, out, (cudaStream_t)stream);
}
GL_API const char *gl_plugin_last_error(void) { return gl::error_buffer(); }
GL_API void gl_plugin_info(int *ns, int *np, int *nobs) {
    *ns = Net::NS; *np = Net::NP; *nobs = Net::NOBS;
}
}
'''


def compile_plugin(network, force=False):
    """nvcc the generated source into gleipnir_b200/_plugins/ode_<name>_<hash>.so (cached)."""
    from .build import NVCC_FLAGS, _nvcc
    os.makedirs(PLUGIN_DIR, exist_ok=True)
    tag = "%s_%s" % (network.name, network.source_hash())
    so = os.path.join(PLUGIN_DIR, "ode_%s.so" % tag)
    if os.path.exists(so) and not force:
        return so
    cu = os.path.join(PLUGIN_DIR, "ode_%s.cu" % tag)
    with open(cu, "w") as f:
        f.write(network.plugin_source())
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-shared", cu, "-o", so, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for ODE network %s:\n%s\n%s" % (network.name, r.stdout, r.stderr))
    return so


_loaded = {}


def load_plugin(network):
    so = compile_plugin(network)
    if so not in _loaded:
        lib = ctypes.CDLL(so)
        lib.gl_plugin_last_error.restype = ctypes.c_char_p
        _loaded[so] = lib
    return _loaded[so]


def pack_blob(network, sampled, tspan, rows, like_type, rtol, atol, max_steps):
    """Data block read by LikeOde (layout: csrc/ode_like.cuh::OdeBlob).
    sampled: list of (parameter name, is_log10); rows: list of (observable index, data[nt], sd[nt], mask[nt])."""
    nt, d, nrows = len(tspan), len(sampled), len(rows)
    header = np.zeros(32)
    nominal = network.param_values.astype(np.float64)
    pmap = np.array([network._pp(n) for n, _ in sampled], dtype=np.float64)
    islog = np.array([1.0 if lg else 0.0 for _, lg in sampled])
    obsidx = np.array([r[0] for r in rows], dtype=np.float64)
    data = np.concatenate([np.asarray(r[1], dtype=np.float64) for r in rows]) if rows else np.zeros(0)
    sd = np.concatenate([np.asarray(r[2], dtype=np.float64) for r in rows]) if rows else np.zeros(0)
    mask = np.concatenate([np.asarray(r[3], dtype=np.float64) for r in rows]) if rows else np.zeros(0)
    parts = [nominal, pmap, islog, np.asarray(tspan, dtype=np.float64), obsidx, data, sd, mask]
    off = 32
    offs = []
    for a in parts:
        offs.append(off)
        off += len(a)
    header[0:8] = [d, len(nominal), nt, like_type, max_steps, rtol, atol, nrows]
    header[16:24] = offs
    return np.ascontiguousarray(np.concatenate([header] + parts))


class OdeNetworkLogLikelihood(DeviceLogLikelihood):
    """Log-likelihood of time-course data under a mass-action ODE network.

    network            MassActionNetwork
    sampled            names of the sampled parameters, in sampled-vector order; the engine's
                       vector holds log10 of each (``params[mask] = 10**Y``) unless
                       ``log10=False``
    tspan              output time points (tspan[0] is the initial time)
    observable_data    {observable: (data, sd, mask)} as NestedSampleIt takes it
                       (nestedsample_it.py:224-247): ``mask`` = indices into tspan (None = all),
                       data/sd given at those points
    kind               'snlpdf' (sum of normal log-pdfs), 'mse' or 'sse'
    """

    def __init__(self, network, sampled, tspan, observable_data, kind="snlpdf", log10=True, rtol=1e-8,
                 atol=1e-12, max_steps=100000):
        super().__init__()
        self.network = network
        self.sampled = [(n, bool(log10)) for n in sampled]
        self.ndim = len(self.sampled)
        self.tspan = np.asarray(tspan, dtype=np.float64)
        self.kind = kind
        self.rtol, self.atol, self.max_steps = float(rtol), float(atol), int(max_steps)
        nt = len(self.tspan)
        self.rows = []
        for obs, (dat, sd, mask) in observable_data.items():
            idx = np.arange(nt) if mask is None else np.asarray(mask, dtype=int)
            full_d, full_s, full_m = np.zeros(nt), np.ones(nt), np.zeros(nt)
            full_d[idx] = np.asarray(dat, dtype=np.float64)
            if kind == "mse":
                full_s[idx] = 1.0 / len(idx)               # the mse kernel multiplies by 1/count
            elif sd is not None:
                full_s[idx] = np.asarray(sd, dtype=np.float64)
            full_m[idx] = 1.0
            self.rows.append((network.obs_names.index(obs), full_d, full_s, full_m))
        self._blob = pack_blob(network, self.sampled, self.tspan, self.rows, LIKE_TYPES[kind], self.rtol, self.atol,
                               self.max_steps)
        self._plugin = None

    def _create(self, lib):
        self._plugin = load_plugin(self.network)
        walk = ctypes.cast(self._plugin.gl_plugin_walk_launch, ctypes.c_void_p)
        ev = ctypes.cast(self._plugin.gl_plugin_eval_launch, ctypes.c_void_p)
        h = ctypes.c_void_p()
        _lib.check(lib.gl_model_create_external(self.ndim, walk, ev, self._blob.ctypes.data_as(ctypes.c_void_p),
                                                self._blob.nbytes, ctypes.byref(h)))
        return h
