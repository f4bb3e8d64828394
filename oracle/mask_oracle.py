"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's
patch-mask generator, bit-exact with numpy's legacy global RNG.

Follows  Pretraining/CM-UNet/cmae/models/backbones/UNet_encoder.py:106-139 (create_random_patch_mask)
and the numpy legacy `RandomState` algorithms it calls through `np.random.shuffle` (:124):
  * MT19937 seeded by `init_genrand(s)` (Knuth LCG 1812433253), 624-word state, twist (397, 0x9908b0df),
    tempering (11; 7,0x9d2c5680; 15,0xefc60000; 18);
  * `shuffle(x)` for a 1-d array: for i = n-1 .. 1: j = rk_interval(i); swap(x[i], x[j]);
  * `rk_interval(max)`: mask = smallest 2^k-1 >= max; draw u32 & mask until <= max (32-bit path).
Pinned against numpy itself in tests/test_oracle_mask.py and against the golden vectors minted from the
unmodified reference (tests/golden/masks.json).
"""
import numpy as np

N, M = 624, 397
UPPER, LOWER, MATRIX_A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF


class MT19937:
    """numpy legacy RandomState bit stream (32-bit outputs)."""

    def __init__(self, seed=None):
        self.mt = np.zeros(N, dtype=np.uint32)
        self.pos = N
        if seed is not None:
            self.seed(seed)

    def seed(self, s):
        mt = [0] * N
        mt[0] = int(s) & 0xFFFFFFFF
        for i in range(1, N):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt = np.array(mt, dtype=np.uint32)
        self.pos = N

    # --- interop with numpy (np.random.get_state()/set_state()) -------------------------------
    def set_state(self, key, pos):
        self.mt = np.array(key, dtype=np.uint32).copy()
        self.pos = int(pos)

    def get_state(self):
        return self.mt.copy(), self.pos

    def _regen(self):
        mt = [int(v) for v in self.mt]
        for kk in range(N - M):
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER)
            mt[kk] = mt[kk + M] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
        for kk in range(N - M, N - 1):
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER)
            mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
        y = (mt[N - 1] & UPPER) | (mt[0] & LOWER)
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
        self.mt = np.array(mt, dtype=np.uint32)
        self.pos = 0

    def next_u32(self):
        if self.pos >= N:
            self._regen()
        y = int(self.mt[self.pos])
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def interval(self, mx):
        if mx == 0:
            return 0
        mask = mx
        for sh in (1, 2, 4, 8, 16):
            mask |= mask >> sh
        while True:
            v = self.next_u32() & mask
            if v <= mx:
                return v

    def shuffle_arange(self, n):
        p = list(range(n))
        for i in range(n - 1, 0, -1):
            j = self.interval(i)
            p[i], p[j] = p[j], p[i]
        return p


def num_masked_patches(img_size, patch_size, mask_ratio):
    """K of UNet_encoder.py:119-138: patches are added while area + ps^2 <= int(ratio*S*S)."""
    target = int(mask_ratio * img_size * img_size)
    return min(target // (patch_size * patch_size), (img_size // patch_size) ** 2)


def patch_mask(rng, batch_size, img_size, patch_size=16, mask_ratio=0.65):
    """Returns (mask uint8 (B,S,S), perms int32 (B,P)); consumes B shuffles from `rng` even when
    mask_ratio == 0 (quirk Q2: the target encoder still shuffles, UNet_encoder.py:124 precedes :137)."""
    g = img_size // patch_size
    P = g * g
    K = num_masked_patches(img_size, patch_size, mask_ratio)
    mask = np.zeros((batch_size, img_size, img_size), dtype=np.uint8)
    perms = np.zeros((batch_size, P), dtype=np.int32)
    for b in range(batch_size):
        p = rng.shuffle_arange(P)
        perms[b] = p
        for idx in p[:K]:
            r = (idx // g) * patch_size
            c = (idx % g) * patch_size
            mask[b, r:r + patch_size, c:c + patch_size] = 1
    return mask, perms
