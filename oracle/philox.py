"""TEST INFRASTRUCTURE ONLY (oracle) -- Philox4x32-10 counter-based RNG in numpy.

The reference draws from numpy's global MT19937 (``sampled_parameter.py:49``,
``samplers.py:92-93,128-129``, ``nested_sampling.py:147``); stream parity with
a sharded device run is impossible, so the B200 path is keyed by a
counter-based generator instead and this file is its CPU restatement
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11 --
published algorithm, constants below; pinned by the Random123 known-answer
vectors in ``tests/test_oracle_philox.py``).

Counter layout shared with ``gleipnir_b200/csrc/philox.cuh``:
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (c0, c1, c2, tag)      tag selects the stream (GL_TAG_* below)
Each call yields four 32-bit words (w0..w3) -> two doubles
    u_a = u52(w0, w1), u_b = u52(w2, w3),   u52(a, b) = (2*m + 1) * 2**-53,
    m = (a >> 6) * 2**26 + (b >> 6)         (52 random bits, strictly in (0,1))
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

# stream tags (counter word 3)
GL_TAG_INIT = 0      # initial population: c0,c1 = global row lo/hi, c2 = component
GL_TAG_WIDTH = 1     # initial proposal widths: c0 = draw index (<100), c2 = component
GL_TAG_PICK = 2      # survivor pick: c0 = dead slot (global idx), c1 = step, c2 = attempt
GL_TAG_MOVE = 3      # trial move: c0 = dead slot, c1 = step, c2 = sweep*d + component
GL_TAG_USER = 4      # gl_prior_draw(): c0,c1 = row lo/hi, c2 = component


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable integer arrays.

    Returns four uint64 arrays holding 32-bit words.
    """
    c0, c1, c2, c3 = np.broadcast_arrays(
        np.asarray(c0, dtype=np.uint64) & _MASK32, np.asarray(c1, dtype=np.uint64) & _MASK32,
        np.asarray(c2, dtype=np.uint64) & _MASK32, np.asarray(c3, dtype=np.uint64) & _MASK32)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0          # < 2**64: no wrap
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK32
        hi1, lo1 = p1 >> _S32, p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u52(a, b):
    """Two 32-bit words -> double strictly inside (0, 1), exact in fp64."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    m = ((a >> np.uint64(6)) << np.uint64(26)) | (b >> np.uint64(6))
    return (2.0 * m.astype(np.float64) + 1.0) * (2.0 ** -53)


def uniforms(seed, c0, c1, c2, tag):
    """(u_a, u_b) for the given counters under ``seed`` (see module docstring)."""
    seed = int(seed)
    w0, w1, w2, w3 = philox4x32_10(c0, c1, c2, tag, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return u52(w0, w1), u52(w2, w3)
