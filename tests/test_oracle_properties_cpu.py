"""Property tests of the CPU oracle (hypothesis): the vectorised O(N) restatements against literal, loop-level
restatements of the reference formulas on random shapes incl. ties and collapsed tails."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import mip360_oracle as O

SET = dict(max_examples=40, deadline=None)


def _knots(B, N, seed, collapse):
    g = torch.Generator().manual_seed(seed)
    t = (torch.rand(B, N + 1, generator=g, dtype=torch.float64) * 0.5).cumsum(-1) + 0.1
    if collapse and N >= 3:
        t[0, N // 2:] = t[0, N // 2]  # zero-width tail
    return t, g


@settings(**SET)
@given(B=st.integers(1, 5), N=st.integers(1, 24), seed=st.integers(0, 10**6), collapse=st.booleans())
def test_distortion_linear_form_equals_double_sum(B, N, seed, collapse):
    s, g = _knots(B, N, seed, collapse)
    w = torch.rand(B, N, generator=g, dtype=torch.float64)
    # literal regularization.py:14-17
    ref = 0.0
    for i in range(N):
        for j in range(N):
            ref = ref + torch.sum(w[..., i] * w[..., j] * torch.abs((s[..., i] + s[..., i + 1]) / 2 - (s[..., j] + s[..., j + 1]) / 2))
    ref = ref + 1 / 3 * torch.sum(w ** 2 * (s[..., 1:] - s[..., :-1]))
    torch.testing.assert_close(O.loss_dist(s, w), ref, rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(O.loss_dist_quadratic(s, w), ref, rtol=1e-10, atol=1e-12)


@settings(**SET)
@given(B=st.integers(1, 5), N=st.integers(1, 20), seed=st.integers(0, 10**6), collapse=st.booleans(), ties=st.booleans())
def test_bounds_equal_literal_masked_sum(B, N, seed, collapse, ties):
    tf, g = _knots(B, N, seed, collapse)
    tc, _ = _knots(B, N, seed + 1, False)
    if ties and N >= 2:
        tc[-1, 1] = tf[-1, min(2, N)]
        tc[-1] = tc[-1].sort().values
    w = torch.rand(B, N, generator=g, dtype=torch.float64)
    # literal distillation.py:19-29 (the [B,N] mask selects over the whole batch)
    t0, t1, T0, T1 = tf[..., :-1], tf[..., 1:], tc[..., :-1], tc[..., 1:]
    ref = torch.zeros_like(w)
    for i in range(N):
        L, R = T0[..., i, None], T1[..., i, None]
        ref[..., i] = torch.sum(w[..., ~((t0 > R) | (t1 < L))], dim=-1)
    torch.testing.assert_close(O.bounds(tf, w, tc), ref, rtol=1e-12, atol=1e-14)
    # searchsorted form used by the CUDA kernels (App. B5)
    cum = torch.cat([torch.zeros(B, 1, dtype=w.dtype), w.cumsum(-1)], -1)
    lo = torch.searchsorted(t1.contiguous(), T0.contiguous(), right=False)
    hi = torch.searchsorted(t0.contiguous(), T1.contiguous(), right=True) - 1
    b = torch.where(hi >= lo, torch.gather(cum, 1, (hi + 1).clamp(0, N)) - torch.gather(cum, 1, lo.clamp(0, N)), torch.zeros_like(w))
    torch.testing.assert_close(b, O.bounds_per_ray(tf, w, tc), rtol=1e-10, atol=1e-12)


@settings(**SET)
@given(B=st.integers(1, 4), N=st.integers(1, 20), seed=st.integers(0, 10**6), randomized=st.booleans())
def test_inverse_cdf_equals_literal_mask_formulation(B, N, seed, randomized):
    bins, g = _knots(B, N, seed, True)
    bins = bins.float()
    w = torch.rand(B, N, generator=g) ** 3
    w[0, : N // 2] = 0  # flat CDF stretch -> ties between knots
    cdf = O.pdf_to_cdf(w)
    M = N + 1
    jitter = torch.empty(B, M).uniform_(0, 1 / M - torch.finfo(torch.float32).eps, generator=g)
    u = O.pdf_uniforms(B, M, randomized, jitter=jitter).contiguous()
    # literal ray.py:41-56
    mask = u[..., None, :] >= cdf[..., :, None]

    def find_interval(x):
        x0, _ = torch.max(torch.where(mask, x[..., None], x[..., :1, None]), -2)
        x1, _ = torch.min(torch.where(~mask, x[..., None], x[..., -1:, None]), -2)
        return x0, x1

    b0, b1 = find_interval(bins)
    c0, c1 = find_interval(cdf)
    tt = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    ref = b0 + tt * (b1 - b0)
    out, i0 = O.invert_cdf(bins, cdf, u)
    assert torch.equal(out, ref)
    assert (i0 >= 0).all() and (i0 <= N).all()
    assert (out[:, 1:] >= out[:, :-1]).all()


@settings(**SET)
@given(B=st.integers(1, 4), N=st.integers(1, 16), seed=st.integers(0, 10**6))
def test_contraction_jacobian_closed_form_matches_autograd(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g, dtype=torch.float64) * 3
    J = O.contract_jacobian(x)

    def contract_point(p):  # parameterization.py:23-29 on one 3-vector
        n = torch.linalg.vector_norm(p)
        return p if n <= 1 else (2 - 1 / n) * (p / n)

    for b in range(B):
        for n in range(N):
            ref = torch.autograd.functional.jacobian(contract_point, x[b, n])
            torch.testing.assert_close(J[b, n], ref, rtol=1e-9, atol=1e-12)


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: philox4x32 10)."""
    from oracle import philox
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(x) for x in philox.philox4x32_10(ctr, key)) == out
    u = philox.uniform(0x1234567890ABCDEF, 7, 3, (5, 9))
    assert u.dtype.name == "float32" and u.shape == (5, 9) and (u >= 0).all() and (u < 1).all()
    assert len(set(u.ravel().tolist())) == 45
    assert not (philox.uniform(0x1234567890ABCDEF, 8, 3, (5, 9)) == u).any()   # another call site
    assert not (philox.uniform(0x1234567890ABCDEF, 7, 4, (5, 9)) == u).any()   # another replay
