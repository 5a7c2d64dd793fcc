"""The CPU twin of the in-kernel uniform generator (oracle/rng.py): vectorised numpy against the plain-integer
definition, range and a coarse uniformity check.  The device side is compared in tests/test_gpu_tc_policy.py."""
import numpy as np

from oracle.rng import uniform_scalar, uniforms


def test_vectorised_equals_definition():
    for seed in (0, 1, 1234, 2 ** 63 + 12345, 2 ** 64 - 1):
        for step in (0, 3, 131, 1256):
            u = uniforms(seed, 300, step)
            assert u.dtype == np.float32
            for r in (0, 1, 7, 255, 299):
                assert float(u[r]) == uniform_scalar(seed, r, step)


def test_range_and_uniformity():
    u = np.concatenate([uniforms(42, 65536, s) for s in range(8)])
    assert u.min() >= 0.0 and u.max() < 1.0
    hist = np.histogram(u, bins=64, range=(0.0, 1.0))[0]
    exp = u.size / 64
    assert np.abs(hist - exp).max() < 6 * np.sqrt(exp)
    # consecutive steps / rollouts are not correlated at lag 1
    a, b = uniforms(42, 65536, 0), uniforms(42, 65536, 1)
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.02
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.02
