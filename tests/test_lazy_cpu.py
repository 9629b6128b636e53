"""Host logic of the deferred-vector backend (pygrank_b200/lazy.py) without a GPU: the expression trees the reference
drivers build (/root/reference/pygrank/algorithms/filters/adhoc.py:34-36,166-169, abstract_filters.py:126-136,225-256,
measures/supervised.py:93-130) are recognised as the fused shapes, everything else evaluates eagerly to the numpy
answer.  The fused runs themselves need the CUDA library: tests/test_plugin_dropin.py (GPU)."""
import numpy as np
import pytest
import torch

from pygrank_b200 import _capi as C
from pygrank_b200 import lazy
from pygrank_b200.graph import DeviceGraph
from pygrank_b200.lazy import LazyScalar, LazyVec


class _FakeGraph(DeviceGraph):
    """A DeviceGraph shell whose conv is a dense CPU product (matching only looks at the type and a few flags)."""

    def __init__(self, M):
        self.n = M.shape[0]
        self.shape = M.shape
        self.normalization = "symmetric"
        self.pathological = False
        self._M = torch.from_numpy(M)

    def conv(self, x):
        return x @ self._M


@pytest.fixture
def setup(monkeypatch):
    monkeypatch.setattr(lazy, "FUSE", False)      # no CUDA library here: matching is tested, values come from the eager path
    rng = np.random.default_rng(0)
    n = 50
    M = rng.random((n, n)) * (rng.random((n, n)) < 0.2)
    g = _FakeGraph(M)
    p = LazyVec.wrap(torch.from_numpy(rng.random(n)))
    x = LazyVec.wrap(torch.from_numpy(rng.random(n)))
    return g, M, p, x


def _conv(x, g):
    return LazyVec("conv", (x, g), x.n, x.dtype)


def test_pagerank_step_is_an_affine_step_with_quotient(setup):
    g, M, p, x = setup
    alpha = 0.85
    num = _conv(x, g) * alpha + p * (1 - alpha)                    # adhoc.py:36
    s = LazyScalar(("sum", num))                                   # backend.sum(ranks)
    assert (s == 0) is False                                       # safe_div's test is speculated, nothing evaluated
    assert num._val is None and s._value is None
    nxt = num / s                                                  # abstract_filters.py:134
    kind, operand, (quotient, f) = lazy._step_of(nxt)
    assert kind == "affine" and operand is x and quotient is True
    assert f.M is g and f.a_s == alpha and f.a_v is None
    assert lazy._pure_key(f.b) == ("scale", ("t", id(p._val)), ("num", 1 - alpha))
    # the same structure rebuilt by the next driver iteration has the same signature (loop-invariant terms by id)
    num2 = _conv(x, g) * alpha + p * (1 - alpha)
    f2 = lazy._as_affine(num2)
    assert lazy.AffineRun.signature(f, True) == lazy.AffineRun.signature(f2, True)
    # eager value == numpy
    want = (x._val.numpy() @ M) * alpha + p._val.numpy() * (1 - alpha)
    got = nxt.materialize().numpy()
    assert np.allclose(got, want / want.sum(), rtol=1e-14)


def test_absorbing_walks_step_has_a_vector_coefficient(setup):
    g, M, p, x = setup
    deg = LazyVec.wrap(torch.from_numpy(M.sum(1)))
    absorb = LazyVec.wrap(torch.full((g.n,), 0.17, dtype=torch.float64))
    ret = (_conv(x, g) * deg + p * absorb) / (absorb + deg)        # adhoc.py:167-168
    kind, operand, (quotient, f) = lazy._step_of(ret)
    assert kind == "affine" and quotient is False and operand is x
    a_v = lazy._pure_value(f.a_v, lazy._pure_key(f.a_v)).numpy()
    b = lazy._pure_value(f.b, lazy._pure_key(f.b)).numpy()
    d, a = M.sum(1), 0.17
    assert np.allclose(a_v, d / (a + d)) and np.allclose(b, p._val.numpy() * a / (a + d))
    want = ((x._val.numpy() @ M) * d + p._val.numpy() * a) / (a + d)
    assert np.allclose(ret.materialize().numpy(), want, rtol=1e-14)


def test_error_measures_are_recognised(setup):
    g, M, p, x = setup
    prev, cur = x, p
    mabs = LazyScalar(("sum", abs(prev - cur))) / 50               # supervised.py:101-106
    m = lazy._match_error(mabs.tree)
    assert m.mode == C.ERR_MABS and m.prev is prev and m.cur is cur and m.divisor == 50
    l1 = LazyScalar(("sum", abs(prev - cur)))
    assert lazy._match_error(l1.tree).mode == C.ERR_L1
    d1, d2 = prev - cur, prev - cur
    msq = LazyScalar(("sum", d1 * d2)) / 50                        # supervised.py:117-122
    assert lazy._match_error(msq.tree).mode == C.ERR_MSQ
    mx = LazyScalar(("max", abs(prev - cur)))
    assert lazy._match_error(mx.tree).mode == C.ERR_MAX
    # unmatched vectors fall back to the eager value, and comparisons give Python bools
    want = np.abs(x._val.numpy() - p._val.numpy()).sum() / 50
    assert (mabs <= want + 1e-12) is True and (mabs <= want - 1e-12) is False
    assert float(mabs) == pytest.approx(want, rel=1e-14)


def test_closed_form_chain_is_a_power_chain(setup):
    g, M, p, x = setup
    zeros = LazyVec.wrap(torch.zeros(g.n, dtype=torch.float64))
    r1 = zeros + p * 1.0                                           # abstract_filters.py:228, first term: no operator yet
    assert lazy._step_of(r1) is None
    p2 = _conv(p, g)
    r2 = r1 + p2 * 0.5
    kind, operand, (pw, coef) = lazy._step_of(r2)
    assert kind == "poly-res" and operand is r1 and pw is p2 and coef == 0.5
    assert lazy._power_chain(p2) == ("seed", p, g, 1)
    p3 = _conv(p2, g)
    assert lazy._power_chain(p3) == ("seed", p, g, 2)
    # eager evaluation of the chain == numpy
    want = p._val.numpy() + 0.5 * (p._val.numpy() @ M)
    assert np.allclose(r2.materialize().numpy(), want, rtol=1e-14)
    assert r2.op == "tensor"                                        # a materialised node forgets its expression


def test_laplacian_and_foreign_operators_are_not_fused(setup):
    g, M, p, x = setup
    g.normalization = "laplacian"
    assert lazy._as_affine(_conv(x, g) * 0.5 + p) is None
    assert lazy._power_chain(_conv(_conv(p, g), g)) is None
    g.normalization = "symmetric"
    dense = object()
    assert lazy._step_of(LazyVec("conv", (x, dense), x.n, x.dtype) * 0.5 + p) is None


def test_lazy_scalar_and_vector_arithmetic_matches_numpy(setup):
    g, M, p, x = setup
    xn, pn = x._val.numpy(), p._val.numpy()
    s = LazyScalar(("sum", x))
    assert float(s * 2 + 1) == pytest.approx(xn.sum() * 2 + 1)
    assert float(3 / s) == pytest.approx(3 / xn.sum())
    v = (x - p) * 2 + 1 - (p / x) ** 2 + abs(-x) / s
    want = (xn - pn) * 2 + 1 - (pn / xn) ** 2 + np.abs(-xn) / xn.sum()
    assert np.allclose(np.asarray(v), want, rtol=1e-13)
    assert len(v) == g.n and v.shape == (g.n,) and float(v[3]) == pytest.approx(want[3])
    v[3] = 7.0
    assert float(v[3]) == 7.0
    mask = x > p
    assert isinstance(mask, torch.Tensor) and int(mask.sum()) == int((xn > pn).sum())
    deep = x
    for _ in range(3000):                                           # chains as deep as max_iters do not recurse
        deep = deep * 1.0 + p * 0.0
    assert np.allclose(np.asarray(deep), xn)


def test_loop_invariant_cache_notices_in_place_writes(setup):
    """personalization * (1 - alpha) is evaluated once per run and memoised by tensor identity; writing into the
    personalization (GraphSignal.__setitem__, core/signals.py:92-93) must invalidate it."""
    g, M, p, x = setup
    expr = lambda: p * 0.15
    k = lazy._pure_key(expr())
    first = lazy._pure_value(expr(), k).clone()
    assert lazy._pure_value(expr(), k) is lazy._pure_value(expr(), k)            # memoised
    p[3] = 100.0
    second = lazy._pure_value(expr(), lazy._pure_key(expr()))
    assert float(second[3]) == pytest.approx(15.0) and float(first[3]) != float(second[3])


def test_scalar_division_by_zero_follows_numpy(setup):
    """RMabs divides by sum|previous| (measures/supervised.py:114), zero on the first test of a closed-form filter: the
    numpy backend gets inf and goes on; so must the deferred scalars."""
    g, M, p, x = setup
    zero = lazy.LazyScalar(("sum", p * 0.0))
    one = lazy.LazyScalar(("sum", p * 0.0 + 1.0))
    assert float(one / zero) == float("inf") and not ((one / zero) <= 1e-6)
    assert np.isnan(float(zero / zero)) and not ((zero / zero) <= 1e-6)
