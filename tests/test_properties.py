"""Size-independent properties of the host logic and the oracle (hypothesis), the kind the full-size GPU tests lean on."""
import numpy as np
from hypothesis import given, settings, strategies as st

from quadrotor_manipulator_mppi_b200.sharded import decode_ordered, encode_ordered, shard_range

finite32 = st.floats(allow_nan=False, allow_infinity=False, width=32)


@given(finite32, finite32)
def test_ordered_key_preserves_float_order(a, b):
    """The int32 key the shards MIN-reduce orders exactly like the float costs (and round-trips)."""
    ka, kb = encode_ordered(a), encode_ordered(b)
    assert (np.float32(a) < np.float32(b)) == (ka < kb) or np.float32(a) == np.float32(b)
    assert np.float32(decode_ordered(ka)) == np.float32(a)
    assert -2 ** 31 <= ka < 2 ** 31


@given(st.integers(8, 1 << 26), st.integers(1, 8))
def test_shard_ranges_tile_the_samples(K, world):
    nxt, sizes = 0, []
    for r in range(world):
        off, n = shard_range(K, world, r)
        assert off == nxt and n >= 0
        nxt += n
        sizes.append(n)
    assert nxt == K and max(sizes) - min(sizes) <= 1


@settings(max_examples=25, deadline=None)
@given(st.lists(st.floats(-6.0, 6.0), min_size=7, max_size=7))
def test_oracle_fk_is_a_rigid_transform(q):
    from oracle import oracle as orc
    orc.build()
    T = orc.fk(np.asarray(q, np.float32)).astype(np.float64)
    R = T[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=5e-6) and abs(np.linalg.det(R) - 1.0) < 5e-6
    assert np.allclose(T[3], [0, 0, 0, 1]) and np.linalg.norm(T[:3, 3]) < 1.3         # arm reach of the j2s7s300


@settings(max_examples=25, deadline=None)
@given(st.integers(5, 64), st.sampled_from([5, 9]), st.floats(-2, 2), st.floats(-2, 2), st.floats(-0.05, 0.05))
def test_oracle_savgol_reproduces_quadratics_in_the_interior(T, window, a, b, c):
    """Order-2 smoothing taps leave polynomials of degree <= 2 unchanged away from the padded edges (svg_filter.py)."""
    from oracle import oracle as orc
    orc.build()
    if T <= window:
        return
    t = np.arange(T, dtype=np.float64)
    seq = np.stack([a + b * t * 0.1 + c * t * t] * 3, axis=1).astype(np.float32)
    out = orc.savgol(seq, window)
    h = window // 2
    assert np.allclose(out[h:T - h], seq[h:T - h], rtol=2e-5, atol=2e-4)
    assert abs(orc.savgol_taps(window).sum() - 1.0) < 1e-6


@settings(max_examples=25, deadline=None)
@given(st.lists(st.floats(0.0, 50.0, width=32), min_size=1, max_size=200), st.sampled_from([0.05, 0.1, 1.0]))
def test_oracle_weights_are_a_softmin(costs, lam):
    from oracle import oracle as orc
    orc.build()
    S = np.asarray(costs, np.float32) + np.float32(1000.0)
    w, rho, eta = orc.weights(S, lam)
    assert rho == S.min() and abs(w.sum() - 1.0) < 1e-5 and w[np.argmin(S)] == w.max()
    assert eta >= 1.0 - 1e-6                                                           # the minimum-cost sample has weight 1 before normalising
