"""Symbolic analysis of a mass-action network for the cooperative (warp-per-system) ODE solver.

The solver (csrc/ode_coop.cuh) factors ``M = I/(gamma h) - J`` once per Rosenbrock step and
back-substitutes four times.  One warp (or half-warp) owns one system: lane i holds ROW i of M in
registers.  Everything that depends only on the network's STRUCTURE is decided here, once, and
emitted as straight-line CUDA (no loops over zeros, no runtime indices into the register row):

* **species order** -- a minimum-degree elimination order on the Jacobian's sparsity pattern,
  so that the rows of U stay short: a lane-uniform elimination step costs one warp-wide FMA per
  structural nonzero of the pivot row, whether or not a given lane's multiplier is zero;
* **leaf species** -- a network with more species than lanes (NS > 32) keeps a 32-row *core* in
  the lanes and eliminates NS-32 mutually independent low-degree species first (block
  elimination): their rows stay in shared memory, are never updated (they are eliminated first
  and do not couple to each other), and cost a handful of FMAs each.  Lane l carries leaf l's
  vector entries in a second register slot;
* **symbolic LU** of the core with fill, giving the (pivot, column) update list;
* **level schedules** of the two triangular solves: columns whose unknowns do not depend on each
  other broadcast together, so the shuffle->FMA dependency chain is as long as the elimination
  DAG is deep, not as long as the matrix is wide.

``emulate_solve`` executes the very same plan with numpy lane vectors; the CPU tests compare it
with ``numpy.linalg.solve`` so a wrong pattern or schedule is caught without a GPU.

Reference for what is being integrated: the ScipyOdeSimulator runs behind
``gleipnir/pysb_utilities/nestedsample_it.py:287`` and ``examples/pysb_jarm/run_jarm_multinest.py:116,125``.
"""
import numpy as np

MAX_LANES = 32


def jacobian_pattern(network):
    """Boolean [ns, ns]: entry (s, q) is True when d f_s / d y_q is structurally nonzero."""
    ns = network.ns
    P = np.zeros((ns, ns), dtype=bool)
    for k, reac, prod, sf in network.reactions:
        rows = set()
        for s in sorted(set(reac) | set(prod)):
            if prod.count(s) - reac.count(s) != 0:
                rows.add(network._sp(s))
        for q in sorted(set(reac)):
            for s in rows:
                P[s, network._sp(q)] = True
    return P


class CoopPlan(object):
    """Elimination plan in INTERNAL species numbering: internal index i < nc is core row i (lane i),
    internal index nc + l is leaf l (second slot of lane l).  ``perm[i]`` = original species index."""

    def __init__(self, network, max_lanes=MAX_LANES):
        ns = network.ns
        P0 = jacobian_pattern(network) | np.eye(ns, dtype=bool)
        sym = P0 | P0.T
        nx = max(0, ns - max_lanes)
        leaves = self._pick_leaves(sym, nx)
        if len(leaves) < nx:
            raise ValueError("network %s has %d species: the cooperative solver needs %d mutually independent "
                             "leaf species next to its %d-row core and found %d"
                             % (network.name, ns, nx, max_lanes, len(leaves)))
        core = [s for s in range(ns) if s not in leaves]
        # pattern after leaf elimination (fill among each leaf's neighbours), in original numbering
        P = P0.copy()
        for e in leaves:
            rows = [i for i in core if P[i, e]]
            cols = [q for q in core if P[e, q]]
            for i in rows:
                P[i, cols] = True
        order = self._min_degree(P, core)
        self.perm = order + leaves
        self.ns, self.nc, self.nx = ns, len(order), len(leaves)
        self.gl = 16 if self.nc <= 16 else 32
        inv = {s: i for i, s in enumerate(self.perm)}
        self.inv = inv
        n = ns
        Q = np.zeros((n, n), dtype=bool)                       # original pattern, internal numbering
        for s in range(ns):
            for q in range(ns):
                if P0[s, q]:
                    Q[inv[s], inv[q]] = True
        self.pattern0 = Q.copy()
        nc = self.nc
        # ---- leaf elimination --------------------------------------------------------------------
        self.leaf_rows = []                                    # per leaf: core columns of its row
        self.leaf_cols = []                                    # per leaf: core rows of its column
        for l in range(self.nx):
            e = nc + l
            assert not Q[e, nc:][np.arange(self.nx) != l].any() and not Q[nc:, e][np.arange(self.nx) != l].any()
            cols = [q for q in range(nc) if Q[e, q]]
            rows = [i for i in range(nc) if Q[i, e]]
            self.leaf_rows.append(cols)
            self.leaf_cols.append(rows)
            for i in rows:
                Q[i, cols] = True
        # ---- symbolic LU of the core ----------------------------------------------------------------
        self.lu_ops = []                                       # (c, [q > c with U_cq != 0])
        for c in range(nc):
            cols = [q for q in range(c + 1, nc) if Q[c, q]]
            rows = [i for i in range(c + 1, nc) if Q[i, c]]
            self.lu_ops.append((c, cols))
            for i in rows:
                Q[i, cols] = True
        self.filled = Q
        Lp = np.tril(Q[:nc, :nc], -1)
        Up = np.triu(Q[:nc, :nc], 1)
        # ---- level schedules -------------------------------------------------------------------------
        lev_f = [0] * nc
        for c in range(nc):
            deps = [lev_f[q] for q in range(c) if Lp[c, q]]
            lev_f[c] = 1 + max(deps) if deps else 0
        lev_b = [0] * nc
        for c in range(nc - 1, -1, -1):
            deps = [lev_b[q] for q in range(c + 1, nc) if Up[c, q]]
            lev_b[c] = 1 + max(deps) if deps else 0
        # forward: a column only broadcasts if some row below uses it
        self.fwd_levels = [[c for c in range(nc) if lev_f[c] == lv and Lp[:, c].any()]
                           for lv in range(max(lev_f) + 1 if nc else 0)]
        self.fwd_levels = [lv for lv in self.fwd_levels if lv]
        # backward: every column is scaled at its level; it broadcasts only if some row above uses it
        self.bwd_levels = [([c for c in range(nc) if lev_b[c] == lv],
                            [c for c in range(nc) if lev_b[c] == lv and Up[:, c].any()])
                           for lv in range(max(lev_b) + 1 if nc else 0)]
        self.nnz_u = int(Up.sum())
        self.nnz_l = int(Lp.sum())

    @staticmethod
    def _pick_leaves(sym, nx):
        """nx mutually non-adjacent species of lowest degree (greedy)."""
        if nx == 0:
            return []
        deg = sym.sum(axis=1)
        chosen = []
        for s in np.argsort(deg, kind="stable"):
            if all(not sym[s, e] for e in chosen):
                chosen.append(int(s))
                if len(chosen) == nx:
                    break
        return chosen

    @staticmethod
    def _min_degree(P, core):
        """Greedy Markowitz order on the (unsymmetric) pattern restricted to `core`."""
        P = P.copy()
        left = list(core)
        order = []
        while left:
            best, cost = None, None
            for s in left:
                r = sum(1 for q in left if q != s and P[s, q])
                c = sum(1 for i in left if i != s and P[i, s])
                key = (r * c, r, s)
                if cost is None or key < cost:
                    best, cost = s, key
            order.append(best)
            left.remove(best)
            rows = [i for i in left if P[i, best]]
            cols = [q for q in left if P[best, q]]
            for i in rows:
                P[i, cols] = True
        return order

    # ---- the plan executed with numpy lane vectors (what the generated CUDA does) -----------------
    def emulate_solve(self, J, diag, rhs):
        """Solve (diag*I - J) x = rhs with the plan; J, rhs in ORIGINAL numbering.  Returns x (original
        numbering) or None when a pivot vanishes."""
        ns, nc, nx, gl = self.ns, self.nc, self.nx, self.gl
        perm = np.array(self.perm)
        Ji = np.asarray(J, dtype=np.float64)[np.ix_(perm, perm)]
        bi = np.asarray(rhs, dtype=np.float64)[perm]
        lane = np.arange(gl)
        M = np.zeros((gl, nc))                                  # M[lane, q]
        A = np.zeros((gl, max(nx, 1)))
        for i in range(gl):
            if i < nc:
                M[i] = -Ji[i, :nc]
                M[i, i] += diag
                for l in range(nx):
                    A[i, l] = -Ji[i, nc + l]
            elif i < nc:
                pass
        inv_d2 = np.zeros(gl)
        # leaf elimination
        for l in range(nx):
            e = nc + l
            de = diag - Ji[e, e]
            if not np.isfinite(de) or abs(de) <= 1e-300:
                return None
            ie = 1.0 / de
            A[:, l] *= ie
            for q in self.leaf_rows[l]:
                M[:, q] += A[:, l] * Ji[e, q]                   # M_iq -= l_ie * M_eq, M_eq = -J_eq
            inv_d2[l] = ie
        # core LU
        for c, cols in self.lu_ops:
            piv = M[c, c]
            if not np.isfinite(piv) or abs(piv) <= 1e-300:
                return None
            lm = M[:, c] * (1.0 / piv)
            below = lane > c
            M[below, c] = lm[below]
            for q in cols:
                pq = M[c, q]
                M[below, q] = M[below, q] - lm[below] * pq
        inv_d = np.array([1.0 / M[i, i] if i < nc else 1.0 for i in range(gl)])
        for i in range(nc):                                     # U -> D^-1 U (cuda_scale_u)
            M[i, i + 1:] *= inv_d[i]
        # solve
        b = np.zeros(gl)
        b[:nc] = bi[:nc]
        b2 = np.zeros(gl)
        b2[:nx] = bi[nc:]
        for l in range(nx):
            b = b - A[:, l] * b2[l]
        for level in self.fwd_levels:
            t = {c: b[c] for c in level}
            for c in level:
                m = lane > c
                b[m] = b[m] - M[m, c] * t[c]
        b = b * inv_d
        for scale, bcast in self.bwd_levels:
            t = {c: b[c] for c in bcast}
            for c in bcast:
                m = lane < c
                b[m] = b[m] - M[m, c] * t[c]
        for l in range(nx):
            e = nc + l
            for q in self.leaf_rows[l]:
                b2[l] += Ji[e, q] * b[q]
            b2[l] *= inv_d2[l]
        x = np.zeros(ns)
        x[perm[:nc]] = b[:nc]
        x[perm[nc:]] = b2[:nx]
        return x

    # ---- CUDA emission ---------------------------------------------------------------------------------
    def cuda_lu(self):
        """Body of Net::lu: factor in place.  Registers: M[NC] (this lane's core row), A[NXA] (its
        entries in the leaf columns); J = the LEAF rows of the Jacobian in shared memory (row l = leaf l); inv_d receives 1/U_ii of the lane's own row; returns the lane's verdict on
        the pivots.  Branch-free: a lane at or above the pivot row takes part with a zero multiplier."""
        nc, nx = self.nc, self.nx
        out = ["        bool ok = true;"]
        for l in range(nx):
            e = nc + l
            out.append("        {   // leaf %d (internal species %d)" % (l, e))
            out.append("            const double de = diag - J[%d * LD + %d];" % (l, e))
            out.append("            ok = ok & (fabs(de) > 1e-300);")
            out.append("            const double ie = gl::rcp_newton(de);")
            out.append("            A[%d] *= ie;" % l)
            for q in self.leaf_rows[l]:
                out.append("            M[%d] = fma(A[%d], J[%d * LD + %d], M[%d]);" % (q, l, l, q, q))
            out.append("            inv_d2 = lane == %d ? ie : inv_d2;" % l)
            out.append("        }")
        if self.LU_BROADCAST == "smem" and self.LU_LEVELS:
            # pivots that do not couple (no L or U entry between them) form a LEVEL of the elimination DAG: their
            # lanes publish pivot + row together, ONE sync, then every lane applies them all -- 15 store/sync/load
            # round trips instead of 32 on the JARM network, and the reciprocals of a level overlap
            stride = self.lu_level_stride()
            for li, level in enumerate(self.lu_levels()):
                out.append("        {   // level %d: pivots %s" % (li, " ".join(str(c) for c in level)))
                out.append("            double2 *X = reinterpret_cast<double2 *>(xs + %d);" % ((li & 1) * stride))
                offs, off = {}, 0
                for c in level:
                    cols = dict(self.lu_ops)[c]
                    vals = ["M[%d]" % c] + ["M[%d]" % q for q in cols]
                    if len(vals) & 1:
                        vals.append("0.0")
                    offs[c] = off
                    out.append("            if (lane == %d) {" % c)
                    for j in range(0, len(vals), 2):
                        out.append("                X[%d] = make_double2(%s, %s);" % (off + j // 2, vals[j], vals[j + 1]))
                    out.append("            }")
                    off += len(vals) // 2
                out.append("            __syncwarp(gmask);")
                for c in level:
                    cols = dict(self.lu_ops)[c]
                    nv = len(cols) + 1
                    for j in range((nv + 1) // 2):
                        out.append("            const double2 r%d_%d = X[%d];" % (c, j, offs[c] + j))
                for c in level:
                    cols = dict(self.lu_ops)[c]
                    bc = lambda z, c=c: "r%d_%d.%s" % (c, z // 2, "xy"[z & 1])
                    out.append("            const double piv%d = %s;" % (c, bc(0)))
                    out.append("            ok = ok & (fabs(piv%d) > 1e-300);" % c)
                    out.append("            const double ip%d = gl::rcp_newton(piv%d);" % (c, c))
                    out.append("            inv_d = lane == %d ? ip%d : inv_d;" % (c, c))
                for c in level:
                    cols = dict(self.lu_ops)[c]
                    bc = lambda z, c=c: "r%d_%d.%s" % (c, z // 2, "xy"[z & 1])
                    if cols or any(c in lv for lv in self.fwd_levels):
                        out.append("            const double lz%d = lane > %d ? M[%d] * ip%d : 0.0;" % (c, c, c, c))
                        out.append("            M[%d] = lane > %d ? lz%d : M[%d];" % (c, c, c, c))
                        for z, q in enumerate(cols):
                            out.append("            M[%d] = fma(-lz%d, %s, M[%d]);" % (q, c, bc(z + 1), q))
                out.append("        }")
            out.append("        return ok;")
            return "\n".join(out)
        for c, cols in self.lu_ops:
            out.append("        {   // pivot %d" % c)
            used = bool(cols or any(c in lv for lv in self.fwd_levels))
            if self.LU_BROADCAST == "smem":
                # the pivot lane publishes its pivot and the structural nonzeros of its row (128-bit stores), all
                # lanes read them back as broadcasts: ~k + 2 instructions per pivot where 64-bit shuffles cost
                # 2 SHFL + the moves that re-pair the halves per value.  Two buffers alternate, one sync per pivot.
                vals = ["M[%d]" % c] + ["M[%d]" % q for q in cols]
                if len(vals) & 1:
                    vals.append("0.0")
                out.append("            double2 *X = reinterpret_cast<double2 *>(xs + %d);" % ((c & 1) * self.lu_buf_stride()))
                out.append("            if (lane == %d) {" % c)
                for j in range(0, len(vals), 2):
                    out.append("                X[%d] = make_double2(%s, %s);" % (j // 2, vals[j], vals[j + 1]))
                out.append("            }")
                out.append("            __syncwarp(gmask);")
                for j in range(0, len(vals), 2):
                    out.append("            const double2 r%d = X[%d];" % (j // 2, j // 2))
                bc = lambda z: "r%d.%s" % (z // 2, "xy"[z & 1])
                out.append("            const double piv = %s;" % bc(0))
            else:
                bc = None
                out.append("            const double piv = gl::shfl_f64(gmask, M[%d], gbase + %d);" % (c, c))
            out.append("            ok = ok & (fabs(piv) > 1e-300);")
            out.append("            const double ip = gl::rcp_newton(piv);")
            out.append("            inv_d = lane == %d ? ip : inv_d;" % c)
            if used:
                if self.NEG_FACTORS:
                    out.append("            const double nlz = lane > %d ? -(M[%d] * ip) : 0.0;" % (c, c))
                    out.append("            M[%d] = lane > %d ? nlz : M[%d];" % (c, c, c))
                else:
                    out.append("            const double lz = lane > %d ? M[%d] * ip : 0.0;" % (c, c))
                    out.append("            M[%d] = lane > %d ? lz : M[%d];" % (c, c, c))
                for z, q in enumerate(cols):
                    src = bc(z + 1) if bc else "gl::shfl_f64(gmask, M[%d], gbase + %d)" % (q, c)
                    out.append("            M[%d] = fma(%s, %s, M[%d]);" % (q, "nlz" if self.NEG_FACTORS else "-lz", src, q))
            out.append("        }")
        out.append("        return ok;")
        return "\n".join(out)

    # levels of at most this many columns are broadcast by shuffles instead of shared memory.  Measured on the JARM
    # network (tools/jarm_check.py, steps/s): 0 -> 7.23e7, 2 -> 7.20e7, 4 -> 7.16e7, 32 (all shuffles) -> 6.23e7.
    SHUFFLE_LEVEL_WIDTH = int(__import__("os").environ.get("GLEIPNIR_SHFL_WIDTH", "0"))

    # how the elimination broadcasts a pivot row: "shfl" (64-bit shuffles) or "smem" (pivot lane stores, all load).
    # Measured on the JARM network (tools/jarm_check.py, 40 000 prior draws, ODE steps/s): shfl 9.2e7, smem 9.4e7;
    # substitution columns: select 9.4e7, mask_t 9.2e7.
    LU_BROADCAST = __import__("os").environ.get("GLEIPNIR_LU_BROADCAST", "smem")
    # how a substitution column is masked: "select" (FMA, then select the result) or "mask_t" (zero the broadcast
    # value on the lanes that do not take part: the selects leave the dependent chain of b)
    SOLVE_FORM = __import__("os").environ.get("GLEIPNIR_SOLVE_FORM", "select")

    # the factors are kept NEGATED in the register row (-L below the diagonal, -D^-1 U above it): the substitution
    # FMAs then need no negated operand
    NEG_FACTORS = __import__("os").environ.get("GLEIPNIR_NEG_FACTORS", "0") == "1"

    # group the pivots of the elimination by levels of its DAG (shared-memory broadcast only).  Measured on the JARM
    # network (15 levels instead of 32 pivots; tools/jarm_check.py, ODE steps/s): off 9.48e7, whole levels 8.45e7 (a
    # level of ten pivots keeps 25 register pairs of row values alive), at most 2 pivots per group 9.37e7, 3: 9.2e7
    # -- the elimination's store/sync/load round trips are not what a step waits for; left off.
    LU_LEVELS = __import__("os").environ.get("GLEIPNIR_LU_LEVELS", "0") == "1"
    LU_LEVEL_WIDTH = int(__import__("os").environ.get("GLEIPNIR_LU_LEVEL_WIDTH", "0"))

    def lu_levels(self):
        """Pivots grouped so that no two of a level couple: level(c) = 1 + max level of the earlier pivots c' with
        L[c][c'] != 0 or U[c'][c] != 0 (filled pattern)."""
        Q, nc = self.filled, self.nc
        lev = [0] * nc
        for c in range(nc):
            deps = [lev[q] for q in range(c) if Q[c, q] or Q[q, c]]
            lev[c] = 1 + max(deps) if deps else 0
        levels = [[c for c in range(nc) if lev[c] == l] for l in range(max(lev) + 1 if nc else 0)]
        w = self.LU_LEVEL_WIDTH                             # wide levels are cut: every pivot keeps its row in registers
        return [lv[i:i + w] for lv in levels for i in range(0, len(lv), w)] if w > 0 else levels

    def lu_level_stride(self):
        """Doubles per broadcast buffer of a level (pivot + row of every pivot of the level, each padded to a pair)."""
        ops = dict(self.lu_ops)
        return max([sum(len(ops[c]) + 1 + ((len(ops[c]) + 1) & 1) for c in level) for level in self.lu_levels()] + [2])

    def lu_buf_stride(self):
        """Doubles per broadcast buffer of the elimination (pivot + longest pivot row, padded to a pair)."""
        m = max([len(cols) + 1 for _, cols in self.lu_ops] + [2])
        return m + (m & 1)

    def xs_stride(self):
        """Doubles per broadcast buffer of the substitution: one slot per lane plus one per leaf."""
        return max(max(self.gl, self.nc) + max(self.nx, 0), self.lu_buf_stride(),
                   self.lu_level_stride() if self.LU_LEVELS else 0)

    def cuda_scale_u(self):
        """After the factorisation: U -> D^-1 U (unit upper triangular), so that the backward substitution needs
        one unconditional scaling of the right-hand side instead of a predicated one per level."""
        f = "ninv_d" if self.NEG_FACTORS else "inv_d"
        head = ["        const double ninv_d = -inv_d;"] if self.NEG_FACTORS else []
        return "\n".join(head + ["        if (ltm & %du) M[%d] *= %s;" % (1 << q if q < 32 else 0, q, f)
                                 for q in range(1, self.nc)])

    def cuda_solve(self):
        """Body of Net::solve: b (core entry of this lane), b2 (leaf entry of lanes < NX).

        The vector travels through shared memory: at every level of the schedule all lanes store their entry
        (two buffers, alternating, so one warp-level sync per level orders stores and loads), the columns of the
        level are read back as broadcasts (one LDS.64 each, where a 64-bit shuffle costs two SHFL and the moves
        that re-pair the halves) and applied as predicated FMAs.  Predicates are bit tests of two per-lane
        masks, gtm = lanes below this one, ltm = lanes above (one LOP3 producing the predicate; chains of
        `lane > c` comparisons make the compiler build branch ladders).  M holds L below the diagonal and
        D^-1 U above it."""
        nc, nx = self.nc, self.nx
        leaf0 = max(self.gl, self.nc)
        out = ["        int buf = 0;"]
        if nx:
            out.append("        if (lane < %d) xs[%d + lane] = b2;" % (nx, leaf0))
            out.append("        __syncwarp(gmask);")
            for l in range(nx):
                out.append("        b = fma(-A[%d], xs[%d], b);" % (l, leaf0 + l))
        def level_code(cols, mask):
            # narrow levels (the sequential tail of the elimination: the hub species form a clique) take the
            # broadcast from a shuffle -- ~25 cycles against ~55 for the store / sync / load round trip
            out.append("        {")
            if len(cols) <= self.SHUFFLE_LEVEL_WIDTH:
                for c in cols:
                    out.append("            const double t%d = gl::shfl_f64(gmask, b, gbase + %d);" % (c, c))
            else:
                out.append("            double *X = xs + buf * XS; buf ^= 1;")
                out.append("            X[lane] = b;")
                out.append("            __syncwarp(gmask);")
                for c in cols:
                    out.append("            const double t%d = X[%d];" % (c, c))
            for c in cols:
                m = "M[%d]" % c if self.NEG_FACTORS else "-M[%d]" % c
                if self.SOLVE_FORM == "mask_t":
                    out.append("            b = fma(%s, (%s & %du) ? t%d : 0.0, b);" % (m, mask, 1 << c, c))
                else:
                    out.append("            if (%s & %du) b = fma(%s, t%d, b);" % (mask, 1 << c, m, c))
            out.append("        }")
        for level in self.fwd_levels:
            level_code(level, "gtm")
        out.append("        b *= inv_d;")
        for scale, bcast in self.bwd_levels:
            if bcast:
                level_code(bcast, "ltm")
        if nx:
            out.append("        {")
            out.append("            double *X = xs + buf * XS; buf ^= 1;")
            out.append("            X[lane] = b;")
            out.append("            __syncwarp(gmask);")
            for l in range(nx):
                e = nc + l
                out.append("            if (lane == %d) {" % l)
                for q in self.leaf_rows[l]:
                    out.append("                b2 = fma(J[%d * LD + %d], X[%d], b2);" % (l, q, q))
                out.append("                b2 *= inv_d2;")
                out.append("            }")
            out.append("        }")
        out.append("        __syncwarp(gmask);")
        return "\n".join(out)

    def flops_dense(self, n_reactions):
        """SURVEY 8d's algorithmic count per Rosenbrock step: Jacobian 2r*order, LU 2/3 n^3,
        4 x (RHS 3r + solve 2 n^2)."""
        n = self.ns
        return 4.0 * n_reactions + 2.0 / 3.0 * n ** 3 + 4.0 * (3.0 * n_reactions + 2.0 * n * n)

    def flops_executed(self, n_reactions):
        """FMAs x 2 the generated code actually issues per step, counted per USEFUL lane entry
        (structural nonzeros), not per warp instruction."""
        lu = 2.0 * sum(len(self.leaf_rows[l]) * len(self.leaf_cols[l]) for l in range(self.nx))
        Q = self.filled
        for c, cols in self.lu_ops:
            rows = int(Q[c + 1:self.nc, c].sum())
            lu += 2.0 * rows * len(cols) + rows
        solve = 2.0 * (self.nnz_l + self.nnz_u + self.nc + sum(len(r) + len(c) for r, c in zip(self.leaf_rows, self.leaf_cols)))
        return 4.0 * n_reactions + lu + 4.0 * (3.0 * n_reactions + solve)
