"""Counter-based RNG specification (pure Python) -- TEST INFRASTRUCTURE ONLY.

The batched env samples jobs and task durations on the device with Philox4x32-10 (Salmon et al.,
"Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123 algorithm) instead of the
reference's sequential PCG64 `np.random.Generator`.  This file is the executable SPEC of how the
reference's sampler calls map to counters; `oracle/sim_oracle.c` and the CUDA kernels restate it.

    key  = (seed & 0xffffffff, seed >> 32)
    JOB stream   ctr = (job_idx,    0, 1, 0): w0 -> query (tpch.py:177 `integers(22)`),
                                              w1 -> size  (tpch.py:178 `choice(QUERY_SIZES)`),
                                              w2 -> inter-arrival (tpch.py:70 `exponential(mean)`)
    TASK stream  ctr = (launch_idx, 0, 2, 0): w0 -> `random()` of `_sample_executor_key` (tpch.py:225),
                                              w1 -> the one successful `choice(durations)` (tpch.py:211)
                 launch_idx = number of task launches so far in the episode (= duration-tape index)
    LIMIT stream ctr = (0, 0, 3, 0):          w0 -> StochasticTimeLimit's exponential
                                              (wrappers/stochastic_time_limit.py:17)

    bounded(w, n)  = (w * n) >> 32
    random(w)      = w * 2**-32                      (exact in f64)
    exponential(w) = mean * neglog((w + 1) * 2**-32) (neglog: fixed sequence of IEEE f64 +,-,*,/ so
                                                      CPU and GPU agree bit for bit)

`PhiloxNpRandom` is an `np_random` look-alike that `oracle/refrun.py` plugs into the UNMODIFIED
reference through gymnasium's seeding hook, so reference, oracle and kernels consume identical draws.
"""
from __future__ import annotations

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

STREAM_JOB, STREAM_TASK, STREAM_LIMIT = 1, 2, 3


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


LN2 = 0.6931471805599453
SQRT2 = 1.4142135623730951


def neglog_u32(w: int) -> float:
    """-ln((w+1) / 2**32) for w in [0, 2**32), from +,-,*,/ only (no libm)."""
    k = w + 1  # 1 .. 2**32
    e = k.bit_length() - 1  # floor(log2 k)
    m = float(k) / float(1 << e)  # exact, in [1, 2)
    if m > SQRT2:
        m = m * 0.5
        e += 1
    s = (m - 1.0) / (m + 1.0)
    z = s * s
    # 2*atanh(s) = 2 s (1 + z/3 + z^2/5 + ... + z^11/23), Horner from the highest term
    p = 1.0 / 23.0
    for d in (21.0, 19.0, 17.0, 15.0, 13.0, 11.0, 9.0, 7.0, 5.0, 3.0, 1.0):
        p = p * z + 1.0 / d
    lnm = (2.0 * s) * p
    return -(lnm + float(e - 32) * LN2)


def bounded(w: int, n: int) -> int:
    return (w * n) >> 32


class PhiloxNpRandom:
    """The subset of `np.random.Generator` the reference sampler uses, on Philox counters."""

    def __init__(self, seed):
        seed = int(seed or 0)
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.job_idx = 0
        self.launch_idx = 0
        self._jobw = None

    # -- job stream (tpch.py:176-178, :70) --
    def integers(self, n):
        self._jobw = philox4x32_10((self.job_idx, 0, STREAM_JOB, 0), self.key)
        return bounded(self._jobw[0], int(n))

    def exponential(self, scale):
        w = self._jobw[2]
        self.job_idx += 1
        return scale * neglog_u32(w)

    # -- task stream (tpch.py:208-229) --
    def random(self):
        w = philox4x32_10((self.launch_idx, 0, STREAM_TASK, 0), self.key)
        return w[0] * 2.0**-32

    def choice(self, a):
        if len(a) and isinstance(a[0], str):  # QUERY_SIZES
            return a[bounded(self._jobw[1], len(a))]
        if len(a) == 0:
            raise ValueError("'a' cannot be empty unless no samples are taken")
        w = philox4x32_10((self.launch_idx, 0, STREAM_TASK, 0), self.key)
        self.launch_idx += 1
        return a[bounded(w[1], len(a))]


def time_limit_draw(seed: int, mean: float) -> float:
    seed = int(seed or 0)
    w = philox4x32_10((0, 0, STREAM_LIMIT, 0), (seed & MASK, (seed >> 32) & MASK))
    return mean * neglog_u32(w[0])
