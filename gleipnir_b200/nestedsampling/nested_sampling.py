"""``NestedSampling`` -- drop-in for ``gleipnir.nestedsampling.NestedSampling``
(``gleipnir/nestedsampling/nested_sampling.py:22-350``) whose ``run()`` executes on a B200.

Same constructor, ``run(verbose=False) -> (log_evidence, log_evidence_error)``, read-only
properties (``evidence``, ``evidence_error``, ``log_evidence``, ``log_evidence_error``,
``information``, ``dead_points``), ``posteriors``, ``max_loglikelihood``, ``deviance_ic``,
``best_fit_likelihood`` and the inherited ABC helpers.  Extra keyword arguments (all optional;
the call SHAPE stays the reference's, the default ``batch_k`` does change the sequencing):

    batch_k   deaths per device step.  1 reproduces upstream's one-death-per-iteration
              sequencing exactly; the default ``None`` picks ``max(1, N // 10)``, i.e. batched
              Nested Sampling: the k worst die together, each with weight X/(N+1), then
              X <- X (N-k+1)/(N+1)  (measured unbiased for k <= N/2).
              Stopping criteria count DEATHS (``_n_iterations``), as upstream counts iterations:
              ``RemainingPriorMass(c)`` stops after the first n with (N/(N+1))**n <= c -- upstream's
              literal predicate -- while the engine's own ln X after those n deaths differs from
              ln c by O(k/N) for k > 1.  User-defined criteria are polled once per batch, so a run
              can pass their threshold by up to k-1 deaths.
    seed      Philox key.  Default: drawn from numpy's global RNG, so ``np.random.seed``
              still makes a script reproducible.
    device    torch device (default: current CUDA device).
    sharded   None (default): shard the live set over the ranks of an initialised
              ``torch.distributed`` process group when there is more than one (one process per
              GPU; ``population_size`` is then the GLOBAL population, divisible by the world
              size); False: run the whole population on this process's GPU.

The whole run -- prior draws, log-likelihoods, k-worst selection, evidence / information
accumulation, constrained Metropolis walkers, stop test -- is device resident (see
``csrc/engine.cu``); the host only converts the stopping criterion into a death count and
copies the dead-point log back at the end.
"""
import warnings

import numpy as np
import pandas as pd

from ..engine import DeviceEngine
from ..likelihoods import require_device_loglikelihood
from ..nsbase import NestedSamplingBase
from ..sampled_parameter import prior_table
from .samplers import MetropolisComponentWiseHardNSRejection
from .stopping_criterion import NumberOfIterations


def _default_sampler():
    return MetropolisComponentWiseHardNSRejection(10, tuning_cycles=1)


class NestedSampling(NestedSamplingBase):

    def __init__(self, sampled_parameters, loglikelihood, population_size, sampler=None,
                 stopping_criterion=None, batch_k=None, seed=None, device=None, sharded=None,
                 initial_live_points=None):
        self.sampled_parameters = sampled_parameters
        self._sampled_parameters_dict = {sp.name: sp for sp in sampled_parameters}
        self.loglikelihood = loglikelihood
        # upstream's defaults (nested_sampling.py:51-52), but a fresh object per instance
        # instead of a shared mutable default argument
        self.sampler = sampler if sampler is not None else _default_sampler()
        self.population_size = population_size
        self.stopping_criterion = stopping_criterion if stopping_criterion is not None else NumberOfIterations(1000)
        self.batch_k = batch_k
        self.seed = seed
        self.device = device
        self.sharded = sharded
        # extension: start from these live points [population_size, d] (host array) instead of prior draws --
        # resuming a run, or a population the caller has conditioned; their log-likelihoods are evaluated on the device
        self.initial_live_points = initial_live_points

        self._alpha = population_size / (population_size + 1)
        self._evidence = 0.0
        self._evidence_error = 0.0
        self._logZ_err = 0.0
        self._log_evidence = 0.0
        self._information = 0.0
        self._H = 0.0
        self._n_iterations = 0
        self._n_evals = 0
        self._dead_points = list()
        self._dead_values = None                    # host array [rows, d+2]; the DataFrame is built on first access
        self._live_values = None
        self._names = [sp.name for sp in sampled_parameters]
        self._post_eval = False
        self._posteriors = None
        self._engine = None
        self._state = None

    @classmethod
    def from_dead_points(cls, sampled_parameters, loglikelihood, population_size, dead_points, log_evidence,
                         information=0.0, device=None):
        """A finished run rebuilt from its dead-point table (``dead_points`` of this class or of upstream Gleipnir:
        columns log_l, weight, parameters): every post-processing output -- ``posteriors``, ``max_loglikelihood``,
        ``akaike_ic``, ``bayesian_ic``, ``deviance_ic``, ``best_fit_*`` (nested_sampling.py:266-336,
        nsbase.py:93-175) -- then works on it without re-running."""
        ns = cls(sampled_parameters, loglikelihood, population_size, device=device)
        vals = dead_points.to_numpy(dtype=float) if hasattr(dead_points, "to_numpy") else np.asarray(dead_points, dtype=float)
        if vals.ndim != 2 or vals.shape[1] != len(sampled_parameters) + 2:
            raise ValueError("dead_points must be [rows, 2 + number of sampled parameters]")
        ns._dead_values = np.ascontiguousarray(vals, dtype=np.float64)
        ns._log_evidence = float(log_evidence)
        with np.errstate(over="ignore"):
            ns._evidence = float(np.exp(log_evidence))
        ns._information = float(information)
        ns._logZ_err = float(np.sqrt(max(ns._information, 0.0) / population_size))
        ns._evidence_error = float(np.exp(ns._logZ_err))
        ns._dead_points = None
        return ns

    # ------------------------------------------------------------------------------------
    def _resolve_batch_k(self):
        n = int(self.population_size)
        k = self.batch_k
        if k is None:
            k = max(1, n // 10)
        k = int(k)
        if not (1 <= k < n):
            raise ValueError("batch_k must satisfy 1 <= batch_k < population_size")
        return k

    def _death_count(self):
        crit = self.stopping_criterion
        if hasattr(crit, "death_count"):
            return int(crit.death_count(self))
        return None

    def _dist_world(self):
        if self.sharded is False:
            return None
        try:
            import torch.distributed as dist
        except ImportError:          # pragma: no cover
            return None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
        if self.sharded:
            raise RuntimeError("sharded=True needs an initialised torch.distributed group with world_size > 1")
        return None

    def _run_sharded(self, dist, loglike, table, n, k, seed, scfg, n_deaths, verbose, chunked=False):
        """One shard per rank, global threshold by all-gather (gleipnir_b200/sharded.py).  ``chunked``: a user-defined
        stopping criterion, evaluated by every rank on the (identical) run state after every batch; ``n_deaths`` is
        then the capacity of the dead-point log."""
        import torch
        from ..sharded import (gather_dead_table_device, gather_dead_tables, get_nccl_comm, run_sharded_native,
                               run_sharded_sliced, shard_bounds)
        comm = get_nccl_comm()
        rank = comm.rank
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        seed_t = torch.tensor([seed], dtype=torch.int64, device=dev)
        dist.broadcast(seed_t, 0)                       # every rank must use rank 0's Philox key
        off, n_loc = shard_bounds(n, comm.world, rank)
        if k >= n_loc:
            raise ValueError("batch_k (%d) must be smaller than the per-rank population (%d)" % (k, n_loc))
        eng = DeviceEngine(table, loglike, n_loc, k, n_deaths, seed=int(seed_t.item()),
                           device=self.device, row_offset=off, global_population=n, **scfg)
        self._engine = eng
        self._adopt_initial(eng, loglike, off, n_loc)
        if chunked:
            def should_stop(state):
                self._n_iterations = int(state.n_dead)
                self._evidence = float(np.exp(state.log_z))
                self._information = float(state.information)
                return bool(self._stopping_criterion())
            n_deaths = run_sharded_sliced(eng, comm, k, n_deaths, should_stop)
        else:
            run_sharded_native(eng, comm, n_deaths, k)
        st = eng.state()
        if k <= 8192:
            table_rows = gather_dead_table_device(eng, k, n_deaths)     # ordered on the device
        else:
            rows, idx, step = gather_dead_tables(eng)
            from ..engine import death_order
            table_rows = rows[death_order(rows[:, 0], idx, step)]
        evals = torch.tensor([st.n_evals], dtype=torch.int64, device=eng.device)
        dist.all_reduce(evals)
        live = torch.empty((n, eng.d), dtype=torch.float64, device=eng.device)
        dist.all_gather_into_tensor(live, eng.live)
        return st, table_rows, int(evals.item()), live.cpu().numpy()

    def run_replicas(self, seeds, max_concurrent=16):
        """This run repeated once per seed, all of them concurrently on the device (gleipnir_b200/replicas.py);
        returns the finished ``NestedSampling`` objects, one per seed -- e.g. the 10-seed scatter of ln Z that the
        reported ``log_evidence_error`` should be compared with."""
        import copy
        from ..replicas import run_many
        clones = [NestedSampling(self.sampled_parameters, self.loglikelihood, self.population_size,
                                 sampler=copy.copy(self.sampler), stopping_criterion=self.stopping_criterion,
                                 batch_k=self.batch_k, seed=int(s), device=self.device, sharded=False,
                                 initial_live_points=self.initial_live_points) for s in seeds]
        run_many(clones, max_concurrent=max_concurrent)
        return clones

    def _prepare(self, loglike):
        """Everything run() settles before touching the device: (loglike, prior table, N, k, seed, sampler config,
        closed-form death count or None, chunked, dead-log capacity)."""
        if loglike.ndim != len(self.sampled_parameters):
            raise ValueError("loglikelihood.ndim = %d but %d sampled parameters"
                             % (loglike.ndim, len(self.sampled_parameters)))
        table = prior_table(self.sampled_parameters)
        n = int(self.population_size)
        k = self._resolve_batch_k()
        seed = self.seed
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
        scfg = self.sampler.device_config()
        n_deaths = self._death_count()
        chunked = n_deaths is None
        if chunked:
            # user-defined criterion: evaluated on the host between device batches
            capacity = int(getattr(self.stopping_criterion, "max_iterations", 100 * n))
        else:
            capacity = n_deaths
        return loglike, table, n, k, seed, scfg, n_deaths, chunked, capacity

    def run(self, verbose=False):
        loglike = require_device_loglikelihood(self.loglikelihood)
        loglike, table, n, k, seed, scfg, n_deaths, chunked, capacity = self._prepare(loglike)
        dist = self._dist_world()
        if dist is not None:
            st, table_rows, n_evals, live = self._run_sharded(dist, loglike, table, n, k, seed, scfg,
                                                              capacity if chunked else n_deaths, verbose, chunked)
            return self._publish(st, table_rows, n_evals, live, n)
        if verbose:
            print("Generating the initial set of live points with population size {}...".format(n))
        eng = DeviceEngine(table, loglike, n, k, capacity, seed=seed, device=self.device, **scfg)
        self._engine = eng
        self._adopt_initial(eng, loglike, 0, n)
        if verbose:
            print("Evaluating the loglikelihood function for each live point...")
        eng.init()
        if not chunked:
            if verbose:
                # progress printing needs host visibility: advance in slices of ~10 batches
                done = 0
                while done < n_deaths:
                    done = min(n_deaths, done + 10 * k)
                    eng.advance(done)
                    st = eng.state()
                    print("Iteration: {} Evidence estimate: {} Remaining prior mass: {}".format(
                        st.n_dead, np.exp(st.log_z), np.exp(st.log_x)))
            else:
                eng.advance(n_deaths)
        else:
            done, stopped = 0, False
            while done < capacity:
                done = min(capacity, done + k)
                eng.advance(done)
                st = eng.state()
                self._n_iterations = int(st.n_dead)
                self._evidence = float(np.exp(st.log_z))
                self._information = float(st.information)
                if self._stopping_criterion():
                    stopped = True
                    break
            if not stopped:
                warnings.warn("the stopping criterion was still False after %d deaths, the capacity of the dead-point "
                              "log (stopping_criterion.max_iterations, default 100 x population): the run was "
                              "finalised there and is NOT converged by its own criterion" % capacity)
        eng.finalize()
        st = eng.state()
        if k <= 8192:
            table_rows = eng.dead_table_device()       # ordered on the device, one D2H copy
        else:
            table_rows, _idx, _step = eng.dead_table()
        return self._publish(st, table_rows, int(st.n_evals), eng.live.cpu().numpy(), n)

    def _adopt_initial(self, eng, loglike, offset, count):
        if self.initial_live_points is None:
            return
        import torch
        x = np.asarray(self.initial_live_points, dtype=np.float64)
        if x.shape != (int(self.population_size), eng.d):
            raise ValueError("initial_live_points must be [population_size, number of sampled parameters]")
        with torch.cuda.device(eng.device):
            xt = torch.as_tensor(np.ascontiguousarray(x[offset:offset + count])).to(eng.device)   # host -> device
            eng.adopt_population(xt, loglike.evaluate(xt))

    def _publish(self, st, table_rows, n_evals, live, n):
        self._state = st
        self._n_iterations = int(st.n_dead)
        self._n_evals = int(n_evals)
        self._log_evidence = float(st.log_z)
        with np.errstate(over="ignore"):
            self._evidence = float(np.exp(st.log_z))
        self._information = float(st.information)
        self._H = float(st.h_linear)
        self._logZ_err = float(np.sqrt(max(self._information, 0.0) / n))
        self._evidence_error = float(np.exp(self._logZ_err))
        # the dead-point log is on the host now; wrapping it in a DataFrame (upstream's
        # nested_sampling.py:211) costs milliseconds and is deferred to the first `dead_points` access
        self._dead_values = np.ascontiguousarray(table_rows, dtype=np.float64)
        self._live_values = live
        self._dead_points = None
        self._post_eval = False
        return self._log_evidence, self._logZ_err

    def _stopping_criterion(self):
        return self.stopping_criterion(self)

    # ---- read-only outputs (nested_sampling.py:219-264,338-344) ------------------------------
    @property
    def evidence(self):
        return self._evidence

    @evidence.setter
    def evidence(self, value):
        warnings.warn("evidence is not settable")

    @property
    def evidence_error(self):
        return self._evidence_error

    @evidence_error.setter
    def evidence_error(self, value):
        warnings.warn("evidence_error is not settable")

    @property
    def log_evidence(self):
        return self._log_evidence

    @log_evidence.setter
    def log_evidence(self, value):
        warnings.warn("log_evidence is not settable")

    @property
    def log_evidence_error(self):
        return self._logZ_err

    @log_evidence_error.setter
    def log_evidence_error(self, value):
        warnings.warn("log_evidence_error is not settable")

    @property
    def information(self):
        return self._information

    @information.setter
    def information(self, value):
        warnings.warn("information is not settable")

    @property
    def dead_points(self):
        if self._dead_points is None and self._dead_values is not None:
            self._dead_points = pd.DataFrame(self._dead_values, columns=['log_l', 'weight'] + self._names, copy=False)
        return self._dead_points

    @property
    def _live_points(self):
        if self._live_values is None:
            return None
        return pd.DataFrame(self._live_values, columns=self._names)

    @dead_points.setter
    def dead_points(self, value):
        warnings.warn("dead_points is not settable")

    # ---- post-processing (nested_sampling.py:266-336) ----------------------------------------
    def _normalised_weights(self):
        logl = self._dead_values[:, 0]
        w = self._dead_values[:, 1]
        with np.errstate(all="ignore"):
            nw = np.exp(np.log(w) + logl - self._log_evidence)     # = w*exp(logl)/Z without overflow
        return logl, np.where(np.isnan(nw), 0.0, nw)

    def posteriors(self, nbins=None):
        """{name: (marginal, bin_edges, bin_centers)}: weighted histograms of the dead points
        (Rice-rule bin count), computed on the device (csrc/posteriors.cu: numpy's own bin
        assignment rule, deterministic sums); cached after the first call, as upstream."""
        if not self._post_eval:
            from ..posteriors import device_posteriors
            vals = self._dead_values
            marg, edges, centers = device_posteriors(vals[:, 0], vals[:, 1], vals[:, 2:], self._log_evidence, nbins,
                                                     device=self.device)
            self._posteriors = dict()
            for j, name in enumerate(self._names):
                self._posteriors[name] = (marg[j], edges[j], centers[j])
            self._post_eval = True
        return self._posteriors

    def max_loglikelihood(self):
        return self._dead_values[:, 0].max()

    def deviance_ic(self):
        """DIC = p_D + D_bar with D = -2 logL averaged over the posterior weights."""
        logl, nw = self._normalised_weights()
        mask = nw > 0.0
        params = self._dead_values[:, 2:]
        d_bar = np.average(-2.0 * logl[mask], weights=nw[mask])
        theta_bar = np.average(params[mask], weights=nw[mask], axis=0)
        p_d = d_bar - (-2.0 * self.loglikelihood(theta_bar))
        return p_d + d_bar

    def best_fit_likelihood(self):
        vals = self._dead_values
        return vals[int(np.argmax(vals[:, 0]))][2:]
