"""Particle size distributions of `subsection particle type N` (source/dem/distributions.cc,
include/dem/distributions.h:404-450): uniform, normal and lognormal, number-weighted.

The reference samples with `std::mt19937(seed + rank)` and `std::normal_distribution<>` /
`std::lognormal_distribution<>`; to insert the same diameters the generators are restated here the
way libstdc++ implements them: generate_canonical<double, 53> from two 32-bit draws, Marsaglia's
polar method returning y*mult first and keeping x*mult for the next call."""
from __future__ import annotations

import math

import numpy as np


class Mt19937Canonical:
    """std::mt19937 + std::generate_canonical<double, 53>."""

    def __init__(self, seed: int):
        self._bg = np.random.MT19937()
        self._bg._legacy_seeding(int(seed) & 0xFFFFFFFF)  # init_genrand(seed), as std::mt19937(seed)

    def canonical(self) -> float:
        lo, hi = (int(v) for v in self._bg.random_raw(2))
        r = (lo + hi * 4294967296.0) / 18446744073709551616.0
        return math.nextafter(1.0, 0.0) if r >= 1.0 else r


class StdNormal:
    """std::normal_distribution<double> of libstdc++ (bits/random.tcc)."""

    def __init__(self, mean: float, stddev: float):
        self.mean, self.stddev = mean, stddev
        self._saved = None

    def __call__(self, gen: Mt19937Canonical) -> float:
        if self._saved is not None:
            ret, self._saved = self._saved, None
        else:
            while True:
                x = 2.0 * gen.canonical() - 1.0
                y = 2.0 * gen.canonical() - 1.0
                r2 = x * x + y * y
                if not (r2 > 1.0 or r2 == 0.0):
                    break
            mult = math.sqrt(-2 * math.log(r2) / r2)
            self._saved = x * mult
            ret = y * mult
        return ret * self.stddev + self.mean


class UniformDistribution:
    def __init__(self, diameter: float):
        self.diameter = diameter

    def sample(self, n: int):
        return np.full(n, self.diameter)

    def max_diameter(self) -> float:
        return self.diameter


class NormalDistribution:
    """NormalDistribution (distributions.cc:23-190), number-based weighting."""

    def __init__(self, average, std, seed, min_cutoff=-1.0, max_cutoff=-1.0):
        self.gen = Mt19937Canonical(seed)
        self.dist = StdNormal(average, std)
        self.min_cutoff = average - 2.5 * std if min_cutoff < 0 else min_cutoff
        self.max_cutoff = average + 2.5 * std if max_cutoff < 0 else max_cutoff
        if not (0.0 < self.min_cutoff < self.max_cutoff):
            raise ValueError("normal size distribution: cutoffs must satisfy 0 < min < max")

    def sample(self, n: int):
        out = []
        while len(out) < n:
            d = self.dist(self.gen)
            if self.min_cutoff < d < self.max_cutoff:
                out.append(d)
        return np.array(out)

    def max_diameter(self) -> float:
        return self.max_cutoff


class LogNormalDistribution:
    """LogNormalDistribution (distributions.cc:231-320), number-based weighting;
    std::lognormal_distribution is exp(normal(mu_ln, sigma_ln))."""

    def __init__(self, average, std, seed, min_cutoff=-1.0, max_cutoff=-1.0):
        self.gen = Mt19937Canonical(seed)
        self.sigma_ln = math.sqrt(math.log(1.0 + (std / average) ** 2))
        self.mu_ln = math.log(average) - 0.5 * self.sigma_ln**2
        self.dist = StdNormal(self.mu_ln, self.sigma_ln)
        self.min_cutoff = math.exp(self.mu_ln - 2.5 * self.sigma_ln) if min_cutoff < 0 else min_cutoff
        self.max_cutoff = math.exp(self.mu_ln + 2.5 * self.sigma_ln) if max_cutoff < 0 else max_cutoff

    def sample(self, n: int):
        out = []
        while len(out) < n:
            d = math.exp(self.dist(self.gen))
            if self.min_cutoff < d < self.max_cutoff:
                out.append(d)
        return np.array(out)

    def max_diameter(self) -> float:
        return self.max_cutoff


def make_distribution(ptype, rank: int = 0):
    """setup_distributions (distributions.h:404-450): seed + MPI rank."""
    kind = ptype.size_distribution_type
    if kind == "uniform":
        return UniformDistribution(ptype.diameter)
    if kind == "normal":
        return NormalDistribution(ptype.diameter, ptype.standard_deviation, ptype.seed + rank, ptype.min_cutoff, ptype.max_cutoff)
    if kind == "lognormal":
        return LogNormalDistribution(ptype.diameter, ptype.standard_deviation, ptype.seed + rank, ptype.min_cutoff, ptype.max_cutoff)
    raise ValueError(f"size distribution type {kind!r} is not mirrored (uniform, normal, lognormal are)")
