"""Explicit counter-based RNG state for the Philox kernels: (seed, offset), never hidden in the
native library.  The default generator is seeded lazily from NumPy's global RandomState so that
`np.random.seed(s)` (what the reference's worker_init_fn does) makes synthesis reproducible."""
import numpy as np


class PhiloxGenerator:
    def __init__(self, seed=None):
        self._seed = None if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0

    def manual_seed(self, seed: int):
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0
        return self

    @property
    def seed(self) -> int:
        if self._seed is None:
            hi, lo = np.random.randint(0, 2 ** 32, size=2, dtype=np.uint64)
            self._seed = (int(hi) << 32) | int(lo)
        return self._seed

    def next(self):
        """(seed, offset) for one launch; the offset advances by one per launch."""
        s, o = self.seed, self.offset
        self.offset += 1
        return s, o


default_generator = PhiloxGenerator()


def manual_seed(seed: int):
    return default_generator.manual_seed(seed)
