"""Independent Python restatement of the host RNG layer (simu_b200/host/rng.h):
MT19937 (init_genrand seeding), Boost's uniform_int bucket rejection, libstdc++'s
random_shuffle loop and Boost <= 1.55's cached Box-Muller normal.  Test-side only."""
import math

import numpy as np


class Mt19937:
    def __init__(self, seed: int):
        self.bg = np.random.MT19937()
        self.bg._legacy_seeding(int(seed) & 0xFFFFFFFF)

    def __call__(self) -> int:
        return int(self.bg.random_raw())


def uniform_int_below(eng, n: int) -> int:
    rng_ = n - 1
    if rng_ == 0:
        return 0
    brange = 0xFFFFFFFF
    if rng_ == brange:
        return eng()
    bucket = brange // (rng_ + 1)
    if brange % (rng_ + 1) == rng_:
        bucket += 1
    while True:
        r = eng() // bucket
        if r <= rng_:
            return r


def random_shuffle(v: list, eng) -> list:
    for i in range(1, len(v)):
        j = uniform_int_below(eng, i + 1)
        if i != j:
            v[i], v[j] = v[j], v[i]
    return v


class NormalDraws:
    def __init__(self, eng):
        self.eng, self.valid, self.r1, self.rho = eng, False, 0.0, 0.0

    def __call__(self) -> float:
        if not self.valid:
            self.r1 = self.eng() * (1.0 / 4294967296.0)
            r2 = self.eng() * (1.0 / 4294967296.0)
            self.rho = math.sqrt(-2.0 * math.log(1.0 - r2))
            self.valid = True
        else:
            self.valid = False
        ang = 2.0 * 3.14159265358979323846 * self.r1
        return self.rho * (math.cos(ang) if self.valid else math.sin(ang)) * 1.0 + 0.0
