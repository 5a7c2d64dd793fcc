"""CPU twin of the in-kernel counter-based uniform generator (TEST INFRASTRUCTURE ONLY).

``jampr_uniform(seed, rollout, step)`` of jampr_plus_b200/csrc/common.cuh: splitmix64 finaliser of
``seed + GOLDEN * (rollout * FNV_PRIME + step + 1)`` (all arithmetic modulo 2^64), top 24 bits scaled to [0, 1).
With it a sampling episode that was run with ``Policy.sampling_seed`` (no caller-supplied uniforms) can be replayed on
the CPU: ``select_inverse_cdf(logp, nbh, uniforms(seed, bs, step))`` must reproduce the kernel's actions
(tests/test_gpu_tc_policy.py::test_seeded_sampling_replays_on_cpu).  The reference samples with torch.multinomial
(policy.py:213-215), whose stream is not reproducible across devices: this generator replaces nothing in it."""
import numpy as np

_M = (1 << 64) - 1


def uniform_scalar(seed: int, rollout: int, step: int) -> float:
    """Plain Python integers: the definition."""
    z = (seed + 0x9E3779B97F4A7C15 * ((rollout * 0x100000001B3 + step + 1) & _M)) & _M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    z = z ^ (z >> 31)
    return float(np.float32(z >> 40) * np.float32(1.0 / 16777216.0))


def uniforms(seed: int, n_rollouts: int, step: int) -> np.ndarray:
    """(n_rollouts,) float32, vectorised (uint64 wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        r = np.arange(n_rollouts, dtype=np.uint64)
        z = np.uint64(seed & _M) + np.uint64(0x9E3779B97F4A7C15) * (r * np.uint64(0x100000001B3) + np.uint64(step + 1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
