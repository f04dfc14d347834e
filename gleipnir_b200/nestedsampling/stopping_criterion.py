"""Stopping criteria -- host mirror of ``gleipnir/nestedsampling/stopping_criterion.py``.

Same classes, attributes and ``criterion(nested_sampler) -> bool`` protocol.  Because the
B200 engine enqueues a whole run on the device, each built-in criterion additionally
reports, in closed form, the number of deaths ("iterations") at which its predicate first
turns true (``death_count``); the predicate itself is evaluated literally around that
point so the count is exactly what upstream's ``while not criterion(self)`` loop gives.
"""
import math


class NumberOfIterations(object):
    """Stop once ``_n_iterations >= n_iterations`` (stopping_criterion.py:23-32)."""

    def __init__(self, iterations):
        self.n_iterations = iterations

    def __call__(self, nested_sampler):
        return nested_sampler._n_iterations >= self.n_iterations

    def death_count(self, nested_sampler):
        return max(1, int(math.ceil(self.n_iterations)))


class RemainingPriorMass(object):
    """Stop once ``alpha**n <= cutoff`` (stopping_criterion.py:48-57), n = deaths so far, alpha = N/(N+1):
    upstream's literal predicate, evaluated in closed form.  With ``batch_k`` > 1 the engine's own estimate of the
    remaining prior mass after n deaths (X shrinks by (N-k+1)/(N+1) per batch) differs from alpha**n by O(k/N) in
    the logarithm; the criterion deliberately keeps upstream's definition in terms of the death count."""

    def __init__(self, cutoff):
        self.cutoff = cutoff

    def __call__(self, nested_sampler):
        return nested_sampler._alpha ** nested_sampler._n_iterations <= self.cutoff

    def death_count(self, nested_sampler):
        alpha, cutoff = nested_sampler._alpha, self.cutoff
        if not (0.0 < cutoff):
            raise ValueError("RemainingPriorMass cutoff must be > 0 to terminate")
        if cutoff >= 1.0:
            return 1
        n = max(1, int(math.ceil(math.log(cutoff) / math.log(alpha))))
        while n > 1 and alpha ** (n - 1) <= cutoff:      # literal predicate decides the boundary
            n -= 1
        while alpha ** n > cutoff:
            n += 1
        return n
