"""Random streams used by the oracle (test infrastructure — see oracle/__init__.py).

Two interchangeable back ends drive the *same* walk / sampling logic:

* ``MTStream``   – consumes numpy's global MT19937 (``np.random.choice``) and
  python's ``random.uniform`` exactly the way the reference does
  (anchor_patch_samplers.py:70,74,78,80,98-106), so that with the same seeds the
  oracle reproduces the reference's walks bit for bit (this is how the walk
  logic is pinned).
* ``PhiloxStream`` – counter-based Philox4x32-10 keyed by (seed, walk id, step);
  the CUDA kernels use the identical construction, so GPU walks are compared
  bit-exactly with the oracle.
"""
import random as _pyrandom

import numpy as np

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al. 2011). counter: 4 uint32, key: 2 uint32 -> 4 uint32."""
    c0, c1, c2, c3 = [int(c) & MASK32 for c in counter]
    k0, k1 = [int(k) & MASK32 for k in key]
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK32
        hi1, lo1 = p1 >> 32, p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + PHILOX_W0) & MASK32
        k1 = (k1 + PHILOX_W1) & MASK32
    return c0, c1, c2, c3


def u32_to_index(r, n):
    """Multiply-shift map of a uint32 onto [0, n): (r * n) >> 32 (same on the GPU)."""
    return (int(r) * int(n)) >> 32


def u32_to_unit_f32(r):
    """Top 24 bits -> float32 in [0,1) (same on the GPU)."""
    return np.float32(int(r) >> 8) * np.float32(1.0 / 16777216.0)


# stream tags (4th counter word); must match csrc/philox.cuh
TAG_WALK = 0x57414C4B   # 'WALK'
TAG_NEIGH = 0x4E454947  # 'NEIG'
TAG_POS = 0x504F5331    # 'POS1'
TAG_STRUC = 0x53545231  # 'STR1'
TAG_DROP = 0x44524F50   # 'DROP'


class PhiloxStream:
    """Per-item stream: draw(step) -> 4 uint32 from counter (item_lo, item_hi, step, tag)."""

    def __init__(self, seed, item, tag):
        self.key = (seed & MASK32, (seed >> 32) & MASK32)
        self.item = item
        self.tag = tag
        self.step = 0

    def draw(self):
        out = philox4x32_10((self.item & MASK32, (self.item >> 32) & MASK32, self.step, self.tag), self.key)
        self.step += 1
        return out

    # the walk logic only needs these three primitives ------------------------------------
    def choice(self, n):
        """uniform index in [0,n) (one Philox block per call; word 0)."""
        return u32_to_index(self.draw()[0], n)

    choice1 = choice

    def coin_and_choice(self):
        """one block: (unit float for the beta coin, raw uint32 for the subsequent choice)."""
        r = self.draw()
        return u32_to_unit_f32(r[1]), r[0]


class MTStream:
    """Reference-compatible stream on numpy's global RandomState + python ``random``."""

    def choice(self, n):          # np.random.choice(list)      (anchor_patch_samplers.py:70,74,98-106)
        return int(np.random.choice(n))

    def choice1(self, n):         # np.random.choice(list, 1)[0] (anchor_patch_samplers.py:78,80)
        return int(np.random.choice(n, 1)[0])

    def uniform(self):            # random.uniform(0, 1)          (anchor_patch_samplers.py:102)
        return _pyrandom.uniform(0, 1)
