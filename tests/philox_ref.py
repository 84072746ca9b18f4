"""TEST INFRASTRUCTURE: numpy restatement of the device's counter-based RNG streams (sdc_core.h: Philox4x32-10 keyed by
the env seed, counter = (index, episode, stream, 0x5DCB200); RS_START = 0 -> start day / hour / weather roll,
RS_NOISE = 1 -> Box-Muller normals of the weather random walk).  Independent of the C++ source: written from the
published algorithm (Salmon et al., SC'11) and the stream layout documented in sdc_core.h."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c, k0, k1):
    """c: uint32 array [..., 4]; k0, k1: python ints.  Returns uint32 [..., 4]."""
    x, y, z, w = (c[..., i].astype(np.uint64) for i in range(4))
    for r in range(10):
        p0, p1 = M0 * x, M1 * z
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        kk0, kk1 = np.uint64((k0 + r * W0) & 0xFFFFFFFF), np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        x, y, z, w = hi1 ^ y ^ kk0, lo1, hi0 ^ w ^ kk1, lo0
    return np.stack([x, y, z, w], axis=-1).astype(np.uint32)


def env_random(seed, episode, stream, idx):
    idx = np.atleast_1d(np.asarray(idx, np.uint32))
    c = np.zeros(idx.shape + (4,), np.uint32)
    c[..., 0], c[..., 1], c[..., 2], c[..., 3] = idx, np.uint32(episode), np.uint32(stream), np.uint32(0x5DCB200)
    return philox4x32(c, int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)


def episode_start(seed, episode, day_lo, day_hi):
    """(day, hour, roll): random.randint(lo, hi), random.randint(0, 23), np.random.randint(0, 14) of the reference
    (sustaindc_env.py:454-455, utils/managers.py:601) drawn from stream RS_START."""
    r = env_random(seed, episode, 0, 0)[0]
    return int(day_lo + int(r[0]) % (day_hi - day_lo + 1)), int(r[1]) % 24, int(r[2]) % 14


def noise_increments(seed, episode, n=35040):
    """The n fp32 random-walk increments 0.02f * N(0,1) of one episode (stream RS_NOISE), as float64."""
    n4 = (n + 3) // 4
    r = env_random(seed, episode, 1, np.arange(n4, dtype=np.uint32))
    f32 = np.float32
    k = f32(2.3283064365386963e-10)
    u1 = ((r[:, 0] >> 8).astype(f32) + f32(0.5)) * f32(1.0 / 16777216.0)
    u2 = r[:, 1].astype(f32) * k
    u3 = ((r[:, 2] >> 8).astype(f32) + f32(0.5)) * f32(1.0 / 16777216.0)
    u4 = r[:, 3].astype(f32) * k
    ra = np.sqrt(f32(-2.0) * np.log(u1)).astype(f32)
    rb = np.sqrt(f32(-2.0) * np.log(u3)).astype(f32)
    a2, a4 = f32(6.283185307179586) * u2, f32(6.283185307179586) * u4
    z = np.stack([ra * np.cos(a2).astype(f32), ra * np.sin(a2).astype(f32), rb * np.cos(a4).astype(f32), rb * np.sin(a4).astype(f32)],
                 axis=1).astype(f32).reshape(-1)[:n]
    return (f32(0.02) * z).astype(np.float64)


class ReplayNpRng:
    """Stands in for `np.random` inside oracle.weather_reset: hands out the device's normals and day roll."""

    def __init__(self, increments, roll):
        self.steps, self.roll = increments / 0.02, roll

    def normal(self, loc=0, scale=1, size=None):
        assert size == len(self.steps)
        return self.steps

    def randint(self, lo, hi):
        assert (lo, hi) == (0, 14)
        return self.roll

    def random(self):
        return 0.0
