"""TEST INFRASTRUCTURE: numpy restatement of the device's counter-based RNG streams (sdc_core.h: Philox4x32-10 keyed by
the env seed, counter = (index, episode, stream, 0x5DCB200); RS_START = 0 -> start day / hour / weather roll,
RS_NOISE = 1 -> seeds of the per-segment PCG32 streams behind the Box-Muller normals of the weather random walk).  Independent of the C++ source: written from the
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


N_SEG, SEG = 256, 140          # sdc_core.h kNoiseSegs, kNoiseSeg


def noise_increments(seed, episode, n=35040):
    """The n fp32 random-walk increments 0.02f * N(0,1) of one episode, as float64.  Segment i (samples 140 i ...) draws
    from its own PCG32 stream (XSH-RR 64/32) seeded by the Philox block (i, episode, RS_NOISE); two normals per draw:
    radius from its top 16 bits, angle in [-pi, pi) from its low 16 bits read as int16 (Box-Muller)."""
    r = env_random(seed, episode, 1, np.arange(N_SEG, dtype=np.uint32)).astype(np.uint64)
    state = r[:, 0] | (r[:, 1] << np.uint64(32))
    inc = (r[:, 2] | (r[:, 3] << np.uint64(32))) | np.uint64(1)
    mult = np.uint64(6364136223846793005)
    draws = np.zeros((N_SEG, SEG // 2), np.uint32)
    with np.errstate(over="ignore"):
        for q in range(SEG // 2):
            old = state
            state = old * mult + inc
            xs = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
            rot = (old >> np.uint64(59)).astype(np.uint32)
            draws[:, q] = (xs >> rot) | (xs << ((np.uint32(32) - rot) & np.uint32(31)))
    f32 = np.float32
    a = draws
    u1 = ((a >> 16).astype(f32) + f32(0.5)) * f32(1.0 / 65536.0)
    th = (a & np.uint32(0xFFFF)).astype(np.uint16).view(np.int16).astype(f32) * f32(9.587379924285257e-5)
    rad = np.sqrt(f32(-2.0) * np.log(u1)).astype(f32)
    z = np.empty((N_SEG, SEG), f32)
    z[:, 0::2], z[:, 1::2] = rad * np.cos(th).astype(f32), rad * np.sin(th).astype(f32)
    return (f32(0.02) * z).astype(np.float64).reshape(-1)[:n]


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
