"""GPU parity: prior / likelihood / selection kernels against the oracle and the
reference-derived golden vectors, all through the C ABI (ctypes)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import likelihoods as olk
from oracle import philox as px
from oracle import priors as pr
from oracle.engine import select_k_worst

pytestmark = pytest.mark.gpu

RTOL = 1e-12      # north_star: analytic logL / transforms within 1e-12 relative


def _lib():
    from gleipnir_b200 import _lib
    return _lib, _lib.load()


FAM = {"uniform": (pr.GL_PRIOR_UNIFORM, -5.0, 10.0), "norm": (pr.GL_PRIOR_NORM, 0.5, 2.0),
       "loguniform": (pr.GL_PRIOR_LOGUNIFORM, 1e-3, 1e2)}


@pytest.mark.parametrize("name", sorted(FAM))
def test_prior_maps_match_reference_golden(cuda, golden_dir, name):
    L, lib = _lib()
    g = np.load(os.path.join(golden_dir, "ref_priors.npz"))
    row = L.prior_array([FAM[name]])
    x, u = np.ascontiguousarray(g[name + "/x"]), np.ascontiguousarray(g[name + "/u"])
    out = np.empty_like(x)
    L.check(lib.gl_prior_pdf_host(row, L.as_double_ptr(x), L.as_double_ptr(out), len(x)))
    np.testing.assert_allclose(out, g[name + "/prior"], rtol=RTOL, atol=0)
    L.check(lib.gl_prior_logpdf_host(row, L.as_double_ptr(x), L.as_double_ptr(out), len(x)))
    np.testing.assert_allclose(out, g[name + "/logprior"], rtol=RTOL, atol=1e-14)
    outu = np.empty_like(u)
    L.check(lib.gl_prior_ppf_host(row, L.as_double_ptr(u), L.as_double_ptr(outu), len(u)))
    np.testing.assert_allclose(outu, g[name + "/invcdf"], rtol=RTOL, atol=1e-14)


def test_prior_known_answers_from_reference_tests(cuda):
    # reference tests/test_sampled_parameter.py:23-36
    from scipy.stats import norm
    from gleipnir_b200.sampled_parameter import SampledParameter
    sp = SampledParameter('sample', norm(0.0, 1.0))
    v = cuda.tensor([0.5], dtype=cuda.float64, device="cuda")
    assert np.isclose(sp.device_logprior(v).item(), -1.0439385332046727, rtol=RTOL)
    assert np.isclose(sp.device_prior(v).item(), 0.3520653267642995, rtol=RTOL)
    assert abs(sp.device_invcdf(v).item()) < 1e-15
    assert len(sp.device_rvs(10)) == 10


@pytest.mark.parametrize("n,d", [(1, 1), (1000, 3), (100003, 5), (2 ** 20 + 7, 2)])
def test_prior_draw_matches_oracle_philox(cuda, n, d):
    L, lib = _lib()
    table = [FAM["uniform"], FAM["norm"], FAM["loguniform"], FAM["uniform"], FAM["norm"]][:d]
    out = cuda.empty((n, d), dtype=cuda.float64, device="cuda")
    seed, first = 0x1234567887654321, 2 ** 32 - 5       # crosses the 32-bit row boundary
    L.check(lib.gl_prior_draw(L.prior_array(table), d, seed, px.GL_TAG_USER, first, n, out.data_ptr(), None))
    got = out.cpu().numpy()
    g = np.arange(first, first + n, dtype=np.uint64)
    for j, (kind, p0, p1) in enumerate(table):
        u, _ = px.uniforms(seed, g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), j, px.GL_TAG_USER)
        # loc + scale*u cancels near 0: error is absolute, a few ulp of max(|loc|, scale)
        np.testing.assert_allclose(got[:, j], pr.ppf(kind, p0, p1, u), rtol=RTOL, atol=4e-16 * max(abs(p0), abs(p1)))


def test_prior_transform_block(cuda):
    from scipy.stats import loguniform, norm, uniform
    from gleipnir_b200.sampled_parameter import SampledParameter, hypercube_transform
    sps = [SampledParameter('a', uniform(-5, 10)), SampledParameter('b', norm(0.5, 2.0)),
           SampledParameter('c', loguniform(1e-3, 1e2))]
    u = np.random.default_rng(0).uniform(1e-9, 1 - 1e-9, size=(4097, 3))
    got = hypercube_transform(sps, u).cpu().numpy()
    for j, sp in enumerate(sps):
        np.testing.assert_allclose(got[:, j], sp.invcdf(u[:, j]), rtol=RTOL)


def _models():
    from gleipnir_b200 import likelihoods as lk
    mu, P, ln = olk.make_c3_dense_gaussian(50)
    return [
        (lk.IIDGaussianLogLikelihood(5), olk.Model("iid_gaussian", 5), (-5, 5)),
        (lk.IIDGaussianLogLikelihood(3, mu=0.3, sigma=1.7), olk.Model("iid_gaussian", 3, mu=0.3, sigma=1.7), (-5, 5)),
        (lk.EggCartonLogLikelihood(2), olk.Model("eggcarton", 2), (0, 10 * np.pi)),
        (lk.Mixture1DLogLikelihood(), olk.Model("mixture1d", 1, positions=[0.1, 0.2, 0.5, 0.55, 0.9, 1.1], width=0.01), (0, 2)),
        (lk.DenseGaussianLogLikelihood(mu, P, ln), olk.Model("dense_gaussian", 50, mu=mu, P=P, log_norm=ln), (-5, 5)),
    ]


@pytest.mark.parametrize("which", range(5))
def test_loglike_matches_oracle(cuda, which):
    dev, ora, (lo, hi) = _models()[which]
    x = np.random.default_rng(which).uniform(lo, hi, size=(20011, dev.ndim))
    got = dev.evaluate(x)
    want = ora(x)
    # likelihoods in the fp64 denormal range (logL < -700) may flush to 0 -> -inf on the device;
    # everywhere else finiteness and values must agree
    fin = np.isfinite(want) & (want > -700.0)
    assert np.all(np.isfinite(got[fin]))
    assert np.all((got[~fin] == -np.inf) | (got[~fin] < -700.0))
    np.testing.assert_allclose(got[fin], want[fin], rtol=RTOL, atol=0 if which != 4 else 1e-10)
    # same kernel on device tensors, and the scalar host protocol of the reference
    gt = dev.evaluate(cuda.as_tensor(x, device="cuda")).cpu().numpy()
    assert np.array_equal(gt, got, equal_nan=True)
    assert dev(x[0]) == got[0]


def test_reference_example_likelihood_values(cuda):
    from gleipnir_b200 import likelihoods as lk
    # tests/test_nested_sampling.py:34-36,59-60: loglikelihood(5,5,5,5,5)
    ref = -0.5 * np.sum(np.array([5.] * 5) ** 2) + 5 * (-0.5 * np.log(2 * np.pi))
    assert np.isclose(lk.IIDGaussianLogLikelihood(5)(np.array([5.] * 5)), ref, rtol=RTOL)
    # examples/eggcarton/NS_eggcarton.py:25-27
    v = np.array([1.0, 2.5])
    assert np.isclose(lk.EggCartonLogLikelihood(2)(v), (2. + np.cos(v).prod()) ** 5, rtol=RTOL)
    # 1-D multimodal far from every mode underflows to -inf (NS_1d_multimodal.py:29-31)
    assert lk.Mixture1DLogLikelihood()(np.array([1.9])) == -np.inf


def _select(cuda, x, k):
    L, lib = _lib()
    t = cuda.as_tensor(x, device="cuda")
    idx = cuda.empty(k, dtype=cuda.int64, device="cuda")
    lstar = cuda.empty(1, dtype=cuda.float64, device="cuda")
    ws = cuda.empty(int(lib.gl_select_workspace_bytes(len(x))), dtype=cuda.uint8, device="cuda")
    L.check(lib.gl_select_kworst(t.data_ptr(), len(x), k, idx.data_ptr(), lstar.data_ptr(), ws.data_ptr(), None))
    return idx.cpu().numpy(), lstar.item()


@pytest.mark.parametrize("n,k", [(1, 1), (20, 1), (500, 50), (10000, 1000), (10000, 9999), (1 << 20, 100000),
                                 (1000003, 1)])
def test_select_matches_argmin_order(cuda, n, k):
    rng = np.random.default_rng(n + k)
    x = rng.normal(size=n) * 30.0
    got, lstar = _select(cuda, x, k)
    want = select_k_worst(x, k)
    assert np.array_equal(got, np.sort(want))          # device order is ascending index
    assert lstar == x[want[-1]]
    if k == 1:
        assert got[0] == np.argmin(x)


def test_select_ties_nan_neginf_negzero(cuda):
    rng = np.random.default_rng(9)
    x = np.round(rng.normal(size=5000), 1)              # heavy exact ties
    x[[7, 100, 4000]] = np.nan
    x[[3, 50]] = -np.inf
    x[[9]] = -0.0
    x[[2]] = 0.0
    for k in (1, 2, 3, 4, 5, 6, 100, 2500, 4999):
        got, lstar = _select(cuda, x, k)
        want = select_k_worst(x, k)
        assert np.array_equal(got, np.sort(want)), k
        assert (np.isnan(lstar) and np.isnan(x[want[-1]])) or lstar == x[want[-1]]
    y = np.full(3000, 1.25)                             # every value equal: lowest indices die
    got, _ = _select(cuda, y, 17)
    assert np.array_equal(got, np.arange(17))
    assert _select(cuda, x, 1)[0][0] == np.argmin(x) == 7   # np.argmin: first NaN wins


@pytest.mark.parametrize("d,n", [(50, 100003), (7, 1001), (8, 8), (128, 4099), (3, 1), (33, 1003), (12, 16), (12, 17),
                                 (64, 5000), (65, 3001), (1, 40), (50, 15)])
def test_dense_gaussian_dmma_matches_scalar_kernel_and_oracle(cuda, d, n):
    """FP64 tensor-core (DMMA) batch kernel vs the scalar DFMA kernel vs numpy, ragged sizes: the cp.async
    k-outer kernel (d <= 64, n >= 16; odd d, partial last tile), the n-outer kernel beyond, tiny batches."""
    from gleipnir_b200 import _lib as L
    from gleipnir_b200 import likelihoods as lk
    mu, P, ln = olk.make_c3_dense_gaussian(d, seed=d)
    mu = mu + np.linspace(-0.3, 0.4, d)
    like = lk.DenseGaussianLogLikelihood(mu, P, ln)
    x = np.random.default_rng(d + n).uniform(-5, 5, size=(n, d))
    tensor = like.evaluate(x)
    L.check(L.load().gl_model_set_tensor_path(like.device_model(), 0))
    scalar = like.evaluate(x)
    L.check(L.load().gl_model_set_tensor_path(like.device_model(), 1))
    want = olk.dense_gaussian(x, mu, P, ln)
    np.testing.assert_allclose(tensor, want, rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(scalar, want, rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(tensor, scalar, rtol=RTOL, atol=1e-10)


def test_dense_gaussian_dmma_unaligned_rows(cuda):
    """A batch that starts 8 bytes off a 16-byte boundary takes the kernel without 16-byte async copies."""
    import torch
    from gleipnir_b200 import _lib as L
    from gleipnir_b200 import likelihoods as lk
    d, n = 33, 515
    mu, P, ln = olk.make_c3_dense_gaussian(d, seed=3)
    like = lk.DenseGaussianLogLikelihood(mu, P, ln)
    x = np.random.default_rng(9).uniform(-5, 5, size=(n, d))
    flat = torch.empty(n * d + 1, dtype=torch.float64, device="cuda")
    flat[1:] = torch.as_tensor(x.ravel(), device="cuda")
    view = flat[1:].view(n, d)                         # data pointer = base + 8
    assert view.data_ptr() % 16 == 8
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    L.check(L.load().gl_loglike_eval(like.device_model(), view.data_ptr(), n, out.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream))
    np.testing.assert_allclose(out.cpu().numpy(), olk.dense_gaussian(x, mu, P, ln), rtol=RTOL, atol=1e-10)


def test_eggcarton_near_cosine_zeros_and_large_arguments(cuda):
    """The walker's own cos (immediate-coefficient polynomials after Cody-Waite reduction) must stay
    within 1e-12 of numpy also where cos ~ 0 (odd multiples of pi/2) and far outside the prior box."""
    from gleipnir_b200 import likelihoods as lk
    like = lk.EggCartonLogLikelihood(2)
    base = (np.arange(-40, 41) + 0.5) * np.pi
    x0 = np.concatenate([base, base + 1e-9, base - 3e-7, np.linspace(-1e4, 1e4, 1001), [1e9, -3e10, 0.0, -0.0]])
    x = np.stack([x0, np.full_like(x0, 0.3)], axis=1)
    got = like.evaluate(x)
    want = olk.eggcarton(x)
    np.testing.assert_allclose(got, want, rtol=RTOL)
    # d = 1 exposes cos itself: (2 + cos x)^5
    one = lk.EggCartonLogLikelihood(1)
    np.testing.assert_allclose(one.evaluate(x0[:, None]), (2.0 + np.cos(x0)) ** 5, rtol=RTOL)
