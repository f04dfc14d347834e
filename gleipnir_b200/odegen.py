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
                for sp in sorted(set(reac)):
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
            for q in sorted(set(reac)):            # sorted: the generated text (and its hash) must not depend on PYTHONHASHSEED
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

    def coop_min_blocks(self):
        """Resident 128-thread blocks per SM the cooperative kernels are compiled for: 4 (128 registers per thread,
        16 warps per SM).  The matrix row is gathered straight into registers, so a block of four JARM-size systems
        needs ~40 KB of shared memory and the register file is the only limit."""
        env = os.environ.get("GLEIPNIR_COOP_MIN_BLOCKS")
        if env:
            return int(env)
        return 4

    def plan(self):
        """Elimination plan of the cooperative solver (odeplan.CoopPlan), cached."""
        if getattr(self, "_plan", None) is None:
            from .odeplan import CoopPlan
            self._plan = CoopPlan(self)
        return self._plan

    def coop_tables(self):
        """Index tables of the cooperative solver's right-hand side and Jacobian (csrc/ode_coop.cuh), species in
        the plan's internal numbering.  Everything a lane needs is a list of positions in two small
        shared-memory value tables the group refreshes once per state:

          sv[2r], sv[2r+1]            = v_r, -v_r                      (reaction rates, both signs)
          sd[4r .. 4r+3]              = dv_r/dy_a, -dv_r/dy_a, dv_r/dy_b, -dv_r/dy_b
          sd[DG + i]                  = diag - J_ii, written by lane i from its own list dtab[.][i]

        so that  f_i = sum of sv[ftab[e][i]]  and the register row of  M = diag*I - J  is gathered COLUMN BY COLUMN:
        M[q] of lane i = sum over e < depth_q of sd[mtab[(off_q + e)][i]]  (the opposite-sign slots, so no negation;
        the diagonal entry points at the lane's own slot DG + i; lanes without an entry in a column point at the
        zero slot) -- loads with addresses known up front and adds, no per-entry decoding of the network, and no
        Jacobian rows in shared memory for the core species.  Only the few LEAF rows are stored (jtab2: position in
        sd | column << 16 | "first term of this sum" << 24), because every lane reads them during the elimination."""
        pl = self.plan()
        gl, nc, nx, ns = pl.gl, pl.nc, pl.nx, self.ns
        inv = pl.inv
        isp = lambda name: inv[self._sp(name)]
        nr = len(self.reactions)
        nrm = (nr + gl - 1) // gl
        nrp = nrm * gl
        ONE = ns                                             # y[NS] = 1.0: second "reactant" of a unimolecular step
        ZV, ZD = 2 * nrp, 4 * nrp                            # slots holding 0.0 (list padding)
        rtab = []
        for r in range(nrp):
            if r < nr:
                reac = self.reactions[r][1]
                a, b = isp(reac[0]), (isp(reac[1]) if len(reac) > 1 else ONE)
            else:
                a = b = ONE
            rtab.append(a | (b << 16))
        f_lists = [[] for _ in range(ns)]                    # per internal species: sv indices
        j_lists = [dict() for _ in range(ns)]                # per internal species: {column: [sd indices]}
        for r, (k, reac, prod, sf) in enumerate(self.reactions):
            for name in sorted(set(reac) | set(prod)):
                net = prod.count(name) - reac.count(name)
                if not net:
                    continue
                i = isp(name)
                f_lists[i] += [2 * r + (0 if net > 0 else 1)] * abs(net)
                sgn = 0 if net > 0 else 1
                a = isp(reac[0])
                j_lists[i].setdefault(a, []).extend([4 * r + sgn] * abs(net))
                if len(reac) > 1:
                    b = isp(reac[1])
                    j_lists[i].setdefault(b, []).extend([4 * r + 2 + sgn] * abs(net))

        def ftable(species, width):
            depth = max([len(f_lists[s]) for s in species] + [1])
            tab = [ZV] * (depth * width)
            for lane, s in enumerate(species):
                for e, idx in enumerate(f_lists[s]):
                    tab[e * width + lane] = idx
            return depth, tab

        def jtable(species, width):
            flat = []
            for s in species:
                ent = []
                for col in sorted(j_lists[s]):
                    terms = j_lists[s][col]
                    for t, idx in enumerate(terms):          # bit 24: first term of this (row, column) sum
                        ent.append(idx | (col << 16) | ((1 << 24) if t == 0 else 0))
                flat.append(ent)
            depth = max([len(e) for e in flat] + [1])
            tab = [ZD] * (depth * width)                     # lanes without a row: add 0.0 into column 0 of a dump row
            for lane, ent in enumerate(flat):
                for e in range(depth):                       # padding: add 0.0 to the last sum and store it again
                    tab[e * width + lane] = ent[e] if e < len(ent) else (ZD | (((ent[-1] >> 16) & 0xff) << 16) if ent else ZD)
            return depth, tab

        nxa = max(nx, 1)
        ef, ftab = ftable(list(range(nc)), gl)
        ef2, ftab2 = ftable(list(range(nc, ns)), nxa)
        ej2, jtab2 = jtable(list(range(nc, ns)), nxa)
        # M = diag*I - J gathered by column: the sign of every term is flipped (index ^ 1)
        DG = 4 * nrp + 2                                     # per-lane slots holding diag - J_ii
        ed = max([len(j_lists[i].get(i, [])) for i in range(nc)] + [1])
        dtab = [ZD] * (ed * gl)
        for i in range(nc):
            for e, idx in enumerate(j_lists[i].get(i, [])):
                dtab[e * gl + i] = idx ^ 1
        mcols, mtab = [], []                                 # per column of [core | leaf]: (offset, depth)
        for q in range(ns):
            per_lane = [([DG + i] if i == q else [idx ^ 1 for idx in j_lists[i].get(q, [])]) if i < nc else []
                        for i in range(gl)]
            depth = max(len(t) for t in per_lane)
            mcols.append((len(mtab) // gl, depth))
            for e in range(depth):
                mtab += [t[e] if e < len(t) else ZD for t in per_lane]
        if not mtab:
            mtab = [ZD] * gl
        ld = ns + (ns & 1)
        return dict(nrm=nrm, nrp=nrp, rtab=rtab, ef=ef, ftab=ftab, ef2=ef2, ftab2=ftab2, ej2=ej2, jtab2=jtab2,
                    ed=ed, dtab=dtab, mtab=mtab, mcols=mcols, dg=DG, ld=ld, gl=gl, nc=nc, nx=nx, nxa=nxa, ns=ns)

    def emulate_rhs_jac(self, y, p):
        """f and J (ORIGINAL numbering) computed the way the device does, from coop_tables(); CPU tests compare
        it with the oracle's right-hand side."""
        T = self.coop_tables()
        pl = self.plan()
        ns, nc, gl, nxa = T["ns"], T["nc"], T["gl"], T["nxa"]
        yi = np.ones(ns + 1)
        yi[:ns] = np.asarray(y, dtype=np.float64)[pl.perm]
        c = np.zeros(T["nrp"])
        for r, (k, reac, prod, sf) in enumerate(self.reactions):
            c[r] = sf * p[self._pp(k)]
        sv = np.zeros(2 * T["nrp"] + 2)
        sd = np.zeros(4 * T["nrp"] + 2 + gl)
        for r in range(T["nrp"]):
            a, b = T["rtab"][r] & 0xffff, T["rtab"][r] >> 16
            cy = c[r] * yi[b]
            v, da, db = cy * yi[a], cy, c[r] * yi[a]
            sv[2 * r], sv[2 * r + 1] = v, -v
            sd[4 * r:4 * r + 4] = [da, -da, db, -db]
        f = np.zeros(ns)
        J = np.zeros((ns, ns))
        for width, ef, ftab, base in ((gl, T["ef"], T["ftab"], 0), (nxa, T["ef2"], T["ftab2"], nc)):
            for lane in range(width):
                if base + lane >= ns or (base == 0 and lane >= nc):
                    continue
                f[base + lane] = sum(sv[ftab[e * width + lane]] for e in range(ef))
        for lane in range(min(nxa, ns - nc)):                  # leaf rows, stored as rows
            acc = 0.0
            for e in range(T["ej2"]):
                w = T["jtab2"][e * nxa + lane]
                acc = sd[w & 0xffff] + (0.0 if w >> 24 else acc)
                J[nc + lane, (w >> 16) & 0xff] = acc
        diag = 0.0                                              # the device adds 1/(gamma h) here
        for lane in range(gl):                                  # every lane's own diagonal slot
            sd[T["dg"] + lane] = diag + sum(sd[T["dtab"][e * gl + lane]] for e in range(T["ed"]))
        for q, (off, depth) in enumerate(T["mcols"]):           # the register row, column by column: M = -J
            for lane in range(nc):
                J[lane, q] = -sum(sd[T["mtab"][(off + e) * gl + lane]] for e in range(depth))
        perm = np.array(pl.perm)
        fo = np.zeros(ns)
        fo[perm] = f
        Jo = np.zeros((ns, ns))
        Jo[np.ix_(perm, perm)] = J
        return fo, Jo

    def cuda_tables(self):
        """Table form of the network for the cooperative solver (csrc/ode_coop.cuh), species in the
        plan's internal numbering, plus the generated elimination / substitution code."""
        for k, reac, prod, sf in self.reactions:
            if not 1 <= len(reac) <= 2:
                raise ValueError("the cooperative solver handles uni- and bimolecular mass-action reactions")
        pl = self.plan()
        T = self.coop_tables()
        ns = self.ns
        if ns + 1 >= 256 or 4 * T["nrp"] + 1 >= 65536:
            raise ValueError("network too large for the cooperative solver's 8/16-bit index tables")
        inv = pl.inv                                        # original species index -> internal
        order = [self.species[s] for s in pl.perm]          # internal -> species name
        isp = lambda name: inv[self._sp(name)]
        rx_k = [self._pp(r[0]) for r in self.reactions]
        rx_sf = [r[3] for r in self.reactions]
        init_kind, init_val, init_par = [], [], []
        for s_ in order:
            v = self.initials.get(s_, 0.0)
            init_kind.append(1 if isinstance(v, str) else 0)
            init_par.append(self._pp(v) if isinstance(v, str) else 0)
            init_val.append(0.0 if isinstance(v, str) else float(v))
        obs_ptr, obs_sp, obs_coef = [0], [], []
        for name in self.obs_names:
            for s_, c in self.observables[name].items():
                obs_sp.append(isp(s_))
                obs_coef.append(float(c))
            obs_ptr.append(len(obs_sp))

        def arr(ctype, name, vals):
            body = ", ".join(repr(v) if ctype == "double" else str(v) for v in vals) or "0"
            return "__device__ const %s T_%s[] = {%s};" % (ctype, name, body)

        spec = (("int", "rx_k", rx_k), ("double", "rx_sf", rx_sf), ("int", "init_kind", init_kind),
                ("double", "init_val", init_val), ("int", "init_par", init_par), ("int", "obs_ptr", obs_ptr),
                ("int", "obs_sp", obs_sp), ("double", "obs_coef", obs_coef),
                ("unsigned int", "rtab", T["rtab"]), ("unsigned short", "ftab", T["ftab"]),
                ("unsigned short", "ftab2", T["ftab2"]), ("unsigned short", "dtab", T["dtab"]),
                ("unsigned short", "mtab", T["mtab"]), ("unsigned int", "jtab2", T["jtab2"]))
        tables = [arr(t, n, v) for t, n, v in spec]
        acc = ["    static __device__ __forceinline__ %s %s(int i) { return T_%s[i]; }" % (t, n, n) for t, n, v in spec]
        struct = ["// species, internal order: " + " ".join(order),
                  "struct Net {",
                  "    static constexpr int NS = %d, NC = %d, NX = %d, NXA = %d, NP = %d, NR = %d, NOBS = %d, GL = %d, "
                  "LD = %d;" % (ns, pl.nc, pl.nx, T["nxa"], self.np_, len(self.reactions), len(self.obs_names), pl.gl,
                                T["ld"]),
                  "    static constexpr int NRM = %d, NRP = %d, EF = %d, EF2 = %d, ED = %d, EM = %d, EJ2 = %d, XS = %d;"
                  % (T["nrm"], T["nrp"], T["ef"], T["ef2"], T["ed"], len(T["mtab"]) // pl.gl, T["ej2"], pl.xs_stride())]
        struct += acc
        lm = []
        for q, (off, depth) in enumerate(T["mcols"]):
            dst = "M[%d]" % q if q < pl.nc else "A[%d]" % (q - pl.nc)
            terms = ["sd[mt[%d * GL + lane]]" % (off + e) for e in range(depth)]
            lm.append("        %s = %s;" % (dst, " + ".join(terms) if terms else "0.0"))
        for l in range(pl.nx, T["nxa"]):
            lm.append("        A[%d] = 0.0;" % l)
        struct += ["    // this lane's row of M = diag*I - J and its entries A in the leaf columns, gathered column by column",
                   "    // from the signed derivative table sd (mt = the column index lists in shared memory)",
                   "    static __device__ __forceinline__ void load_m(double (&M)[NC], double (&A)[NXA], const double *sd, "
                   "const unsigned short *mt, int lane) {",
                   "\n".join(lm), "    }"]
        struct += ["    // M = diag*I - J factored in place: lane i holds core row i (L below the diagonal, U on and",
                   "    // above it), A the lane's multipliers of the leaf columns; J = the LEAF rows of the Jacobian in shared memory",
                   "    static __device__ __forceinline__ bool lu(double (&M)[NC], double (&A)[NXA], const double *J, "
                   "double diag, int lane,",
                   "                                              unsigned gmask, int gbase, double &inv_d, double &inv_d2, "
                   "double *xs) {",
                   pl.cuda_lu(), "    }",
                   "    // U -> D^-1 U; ltm = bit mask of the lanes above this one",
                   "    static __device__ __forceinline__ void scale_u(double (&M)[NC], double inv_d, unsigned ltm) {",
                   pl.cuda_scale_u(), "    }",
                   "    static __device__ __forceinline__ void solve(const double (&M)[NC], const double (&A)[NXA], "
                   "const double *J, double inv_d,",
                   "                                                 double inv_d2, double &b, double &b2, int lane, "
                   "unsigned gmask, int gbase, unsigned gtm, unsigned ltm, double *xs) {",
                   pl.cuda_solve(), "    }", "};"]
        return "\n".join(tables + struct)

    def plugin_source(self, coop=None):
        if self.cooperative if coop is None else coop:
            return COOP_PLUGIN_TEMPLATE % dict(name=self.name, struct=self.cuda_tables(), min_blocks=self.coop_min_blocks())
        return PLUGIN_TEMPLATE % dict(name=self.name, struct=self.cuda_struct())

    def source_hash(self, coop=None):
        h = hashlib.sha256(self.plugin_source(coop).encode())
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
#define GL_COOP_MIN_BLOCKS %(min_blocks)d
#include "ode_coop.cuh"

namespace {
%(struct)s
}  // namespace

extern "C" {
GL_API int gl_plugin_walk_launch(const void *P, const void *M, void *stream) {
    const gl::ModelDev &md = *(const gl::ModelDev *)M;          // s0 = blob word 8: number of conditions
    return gl::walk_coop_launch<Net>(*(const gl::WalkParams *)P, md, md.s0 >= 1.0 ? (int)md.s0 : 1, (cudaStream_t)stream);
}
GL_API int gl_plugin_eval_launch(const void *M, const double *x, long long n, double *out, void *stream) {
    return gl::ode_coop_eval_launch<Net>(*(const gl::ModelDev *)M, x, n, out, (cudaStream_t)stream);
}
GL_API const char *gl_plugin_last_error(void) { return gl::error_buffer(); }
GL_API void gl_plugin_info(int *ns, int *np, int *nobs) {
    *ns = Net::NS; *np = Net::NP; *nobs = Net::NOBS;
}
GL_API int gl_plugin_counters(unsigned long long *out3, int reset) { return gl::coop_counters<Net>(out3, reset); }
GL_API long long gl_plugin_trace(int enable, unsigned long long *out, long long max_records) {
    return gl::coop_trace<Net>(enable, out, max_records);
}
}
'''


def compile_plugin(network, force=False, coop=None):
    """nvcc the generated source into gleipnir_b200/_plugins/ode_<name>_<hash>.so (cached).
    coop=True forces the cooperative (lane-group) solver, which the protocol likelihood needs."""
    from .build import NVCC_FLAGS, _nvcc
    os.makedirs(PLUGIN_DIR, exist_ok=True)
    tag = "%s_%s" % (network.name, network.source_hash(coop))
    so = os.path.join(PLUGIN_DIR, "ode_%s.so" % tag)
    if os.path.exists(so) and not force:
        return so
    cu = os.path.join(PLUGIN_DIR, "ode_%s.cu" % tag)
    with open(cu, "w") as f:
        f.write(network.plugin_source(coop))
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-shared", cu, "-o", so, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for ODE network %s:\n%s\n%s" % (network.name, r.stdout, r.stderr))
    return so


_loaded = {}


def load_plugin(network, coop=None):
    so = compile_plugin(network, coop=coop)
    if so not in _loaded:
        lib = ctypes.CDLL(so)
        lib.gl_plugin_last_error.restype = ctypes.c_char_p
        _loaded[so] = lib
    return _loaded[so]


def pack_protocol_blob(network, sampled, tspan, conditions, rows, like_type, rtol, atol, max_steps, preeq=None,
                       cycles=None):
    """Data block read by the cooperative solver (layout: csrc/ode_coop.cuh::ProtoBlob).

    sampled     list of (parameter name, is_log10)
    conditions  list (one per condition) of [(parameter index, value), ...] overrides
    rows        list of dicts(cond, obs (index), scale, kind ('norm'|'halfnorm'), data[nt], sd[nt], mask[nt])
    preeq       None | dict(t_end, zero_idx=[parameter indices], eq_tol=None|float)
    cycles      None | dict(loc, scale, factors=[[(a, b, inverted), ...], ...])"""
    nt, d = len(tspan), len(sampled)
    ncond = len(conditions)
    rows = sorted(rows, key=lambda r: r["cond"])             # stable: rows of a condition stay in the given order
    nrows = len(rows)
    header = np.zeros(48)
    nominal = network.param_values.astype(np.float64)
    pmap = np.array([network._pp(n) for n, _ in sampled], dtype=np.float64)
    islog = np.array([1.0 if lg else 0.0 for _, lg in sampled])
    meta = np.zeros(4 * nrows)
    data, sd, mask, cterm = (np.zeros(nrows * nt) for _ in range(4))
    sd[:] = 1.0
    rowptr = np.zeros(ncond + 1)
    stops, stopptr = [], [0]
    for i, r in enumerate(rows):
        kind = {"norm": 0, "halfnorm": 1}[r.get("kind", "norm")]
        meta[4 * i:4 * i + 4] = [r["cond"], r["obs"], r.get("scale", 1.0), kind]
        data[i * nt:(i + 1) * nt] = r["data"]
        sd[i * nt:(i + 1) * nt] = r["sd"]
        mask[i * nt:(i + 1) * nt] = r["mask"]
        with np.errstate(divide="ignore", invalid="ignore"):
            lsd = np.log(np.asarray(r["sd"], dtype=np.float64))
        # scipy: norm.logpdf = -z^2/2 - log(sd) - log(2 pi)/2; halfnorm.logpdf = -z^2/2 + log(2/pi)/2 - log(sd), z >= 0
        cterm[i * nt:(i + 1) * nt] = (0.5 * np.log(2.0 / np.pi) - lsd) if kind == 1 else (-lsd - 0.5 * np.log(2.0 * np.pi))
        rowptr[r["cond"] + 1:] += 1
    for c in range(ncond):
        m = np.zeros(nt, dtype=bool)
        for r in rows:
            if r["cond"] == c:
                m |= np.asarray(r["mask"]) != 0
        stops += [float(i) for i in np.nonzero(m)[0]]
        stopptr.append(len(stops))
    overptr, overidx, overval = [0], [], []
    for ov in conditions:
        for i, v in ov:
            overidx.append(float(i))
            overval.append(float(v))
        overptr.append(len(overidx))
    zeroidx = [float(i) for i in (preeq["zero_idx"] if preeq else [])]
    cycptr, cycfac = [0], []
    if cycles:
        for factors in cycles["factors"]:
            for a, b, inv in factors:
                cycfac += [float(a), float(b), 1.0 if inv else 0.0]
            cycptr.append(len(cycfac) // 3)
    parts = [nominal, pmap, islog, np.asarray(tspan, dtype=np.float64), meta, data, sd, mask, cterm, rowptr,
             np.array(overptr, dtype=np.float64), np.array(overidx), np.array(overval), np.array(zeroidx),
             np.array(cycptr, dtype=np.float64), np.array(cycfac), np.array(stopptr, dtype=np.float64), np.array(stops)]
    off = 48
    offs = []
    for a in parts:
        offs.append(off)
        off += len(a)
    eq_tol = preeq.get("eq_tol") if preeq else None
    header[0:16] = [d, len(nominal), nt, like_type, max_steps, rtol, atol, nrows, ncond, 1.0 if preeq else 0.0,
                    preeq["t_end"] if preeq else 0.0, -1.0 if eq_tol is None else eq_tol, len(zeroidx),
                    len(cycptr) - 1, cycles["loc"] if cycles else 0.0, cycles["scale"] if cycles else 1.0]
    header[16:16 + len(offs)] = offs
    return np.ascontiguousarray(np.concatenate([header] + [np.asarray(a, dtype=np.float64) for a in parts]))


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
        self._coop = network.cooperative
        if self._coop:                 # one condition, no pre-equilibration, no cycle terms
            prows = [dict(cond=0, obs=o, scale=1.0, kind="norm", data=fd, sd=fs, mask=fm) for o, fd, fs, fm in self.rows]
            self._blob = pack_protocol_blob(network, self.sampled, self.tspan, [[]], prows, LIKE_TYPES[kind], self.rtol,
                                            self.atol, self.max_steps)
        else:
            self._blob = pack_blob(network, self.sampled, self.tspan, self.rows, LIKE_TYPES[kind], self.rtol, self.atol,
                                   self.max_steps)
        self._plugin = None

    def _create(self, lib):
        self._plugin = load_plugin(self.network, coop=self._coop)
        walk = ctypes.cast(self._plugin.gl_plugin_walk_launch, ctypes.c_void_p)
        ev = ctypes.cast(self._plugin.gl_plugin_eval_launch, ctypes.c_void_p)
        h = ctypes.c_void_p()
        _lib.check(lib.gl_model_create_external(self.ndim, walk, ev, self._blob.ctypes.data_as(ctypes.c_void_p),
                                                self._blob.nbytes, ctypes.byref(h)))
        return h

    def counters(self, reset=False):
        """Work counters of the cooperative solver since the last reset: likelihood evaluations actually
        integrated, ODE solves, Rosenbrock steps attempted (all likelihood objects of one network share them)."""
        if not self._coop:
            raise TypeError("work counters exist for the cooperative solver only")
        self.device_model()
        out = (ctypes.c_ulonglong * 3)()
        _lib.check(self._plugin.gl_plugin_counters(out, 1 if reset else 0))
        return dict(evals=int(out[0]), solves=int(out[1]), steps=int(out[2]))

    def trace(self, enable=None):
        """Walker-launch trace of the cooperative solver (diagnostics): ``trace(True)`` / ``trace(False)`` switch it on
        (clearing it) / off; ``trace()`` returns an int64 array [n, 4] of (phase << 32 | SM id, start ns, end ns, ODE
        steps) per walker block since it was switched on."""
        self.device_model()
        fn = self._plugin.gl_plugin_trace
        fn.restype = ctypes.c_longlong
        fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
        if enable is not None:
            if fn(1 if enable else 0, None, 0) < 0:
                raise RuntimeError("trace switch failed")
            return None
        buf = np.zeros((4096, 4), dtype=np.uint64)
        n = fn(0, buf.ctypes.data, 4096)
        if n < 0:
            raise RuntimeError("trace read failed")
        return buf[:n].astype(np.int64)


class OdeProtocolLogLikelihood(OdeNetworkLogLikelihood):
    """Multi-condition ODE likelihood: the structure of ``examples/pysb_jarm/run_jarm_multinest.py:92-143``.

    network      MassActionNetwork
    sampled      names of the sampled parameters (``params[mask] = 10**Y``, :94)
    tspan        output time points (:26)
    conditions   one dict {parameter name or index: value} of overrides per simulated condition (:110-111, 122-123)
    rows         data terms (:31-43, 126-134): dicts with ``cond``, ``obs`` (observable name), ``scale`` (the
                 simulated observable is divided by it, :57), ``kind`` 'norm' | 'halfnorm', ``data``/``sd`` at the
                 tspan indices ``mask``
    preeq        None, or dict(t_end=100.0, zero=[parameter names or indices], eq_tol=None): every condition is first
                 integrated to t_end with the listed constants zeroed and the main solve starts from that state
                 (:105-119); with eq_tol the state must satisfy max|dy/dt| <= eq_tol or the likelihood is -inf
                 (the role of the reference's external ``pre_equilibration``, which raises when no equilibrium is found)
    cycles       None, or dict(loc, scale, factors): thermodynamic-cycle products scored by norm(loc, scale) (:72-103)
    rtol, atol   local error tolerances of the RODAS4 integrator.  The defaults, 1e-6 / 1e-12, are the ones the
                 reference's own solver runs with (ScipyOdeSimulator -> scipy `ode`/VODE defaults); RODAS4 at these
                 settings agrees with a tight-tolerance LSODA solution to ~1e-9 relative in the log-likelihood
                 (4e-8 in the final state), so the north_star bound of 1e-6 holds with three orders to spare.
    """

    def __init__(self, network, sampled, tspan, conditions, rows, preeq=None, cycles=None, log10=True, rtol=1e-6,
                 atol=1e-12, max_steps=100000):
        DeviceLogLikelihood.__init__(self)
        self.network = network
        self.sampled = [(n, bool(log10)) for n in sampled]
        self.ndim = len(self.sampled)
        self.tspan = np.asarray(tspan, dtype=np.float64)
        self.kind = "snlpdf"
        self.rtol, self.atol, self.max_steps = float(rtol), float(atol), int(max_steps)
        nt = len(self.tspan)
        pidx = lambda k: network._pp(k) if isinstance(k, str) else int(k)
        self.conditions = [[(pidx(k), float(v)) for k, v in c.items()] for c in conditions]
        self.rows = []
        for r in rows:
            idx = np.arange(nt) if r.get("mask") is None else np.asarray(r["mask"], dtype=int)
            fd, fs, fm = np.zeros(nt), np.ones(nt), np.zeros(nt)
            fd[idx], fs[idx], fm[idx] = np.asarray(r["data"], dtype=np.float64), np.asarray(r["sd"], dtype=np.float64), 1.0
            obs = r["obs"]
            self.rows.append(dict(cond=int(r["cond"]), obs=network.obs_names.index(obs) if isinstance(obs, str) else int(obs),
                                  scale=float(r.get("scale", 1.0)), kind=r.get("kind", "norm"), data=fd, sd=fs, mask=fm))
        self.preeq = None
        if preeq:
            self.preeq = dict(t_end=float(preeq.get("t_end", 100.0)), zero_idx=[pidx(k) for k in preeq.get("zero", [])],
                              eq_tol=preeq.get("eq_tol"), n_points=int(preeq.get("n_points", 100)))
        self.cycles = None
        if cycles:
            self.cycles = dict(loc=float(cycles.get("loc", 1.0)), scale=float(cycles.get("scale", 1.0)),
                               factors=[[(pidx(a), pidx(b), bool(inv)) for a, b, inv in f] for f in cycles["factors"]])
        self._coop = True
        self._blob = pack_protocol_blob(network, self.sampled, self.tspan, self.conditions, self.rows, LIKE_TYPES["snlpdf"],
                                        self.rtol, self.atol, self.max_steps, self.preeq, self.cycles)
        self._plugin = None

    def oracle_protocol(self):
        """The same protocol as the plain-data dict ``oracle.ode.protocol_loglike`` consumes (tests, bench baseline)."""
        return dict(sampled_idx=[self.network._pp(n) for n, _ in self.sampled], log10=self.sampled[0][1],
                    tspan=self.tspan, conditions=self.conditions, preeq=self.preeq, cycles=self.cycles, rows=self.rows)
