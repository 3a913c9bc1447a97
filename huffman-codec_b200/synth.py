"""Synthetic 8-bit grayscale workloads (BASELINE.json configs 3-5).

Deterministic numpy generators so that the GPU path, the CPU oracle and the reference
binary all see exactly the same bytes.  Pure host-side data generation; no codec logic.
"""
import numpy as np

CLASSES = ("walk", "smooth", "random", "const")


def image(kind, n, seed, m=None):
    """One m x n (rows x cols) u8 image of the named class (m defaults to n)."""
    m = n if m is None else m
    rng = np.random.default_rng(seed)
    if kind == "walk":
        a = (np.cumsum(rng.integers(-2, 3, (m, n)), axis=1) + 128) & 255
    elif kind == "smooth":
        a_, b_, c_ = (int(v) for v in rng.integers(1, 32, 3))
        y, x = np.mgrid[0:m, 0:n]
        a = ((a_ * x + b_ * y) // 16 + c_) & 255
    elif kind == "random":
        a = rng.integers(0, 256, (m, n))
    elif kind == "const":
        a = np.full((m, n), int(rng.integers(0, 256)))
    elif kind == "fib":
        # Fibonacci-skewed stream: symbol s repeated F(s) times (deep FGK tree), row-padded
        f = [1, 1]
        while len(f) < 24:
            f.append(f[-1] + f[-2])
        s = np.concatenate([np.full(f[i], i) for i in range(24)])
        rng.shuffle(s)
        a = np.resize(s, (m, n))
    elif kind == "longrun":
        # runs around the 258 reset: lengths 255..262 of alternating values
        out = []
        v = 0
        while sum(len(o) for o in out) < m * n:
            out.append(np.full(int(rng.integers(250, 530)), v))
            v = (v + int(rng.integers(1, 255))) & 255
        a = np.concatenate(out)[: m * n].reshape(m, n)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(a.astype(np.uint8))


def batch(count, n, seed0, classes=CLASSES, m=None):
    """count images of n x n; image i has seed seed0+i and class classes[i % len(classes)].

    Returns a (count, m*n) u8 array (one row per file) and the per-file class names."""
    m = n if m is None else m
    out = np.empty((count, m * n), np.uint8)
    names = []
    for i in range(count):
        k = classes[i % len(classes)]
        out[i] = image(k, n, seed0 + i, m).reshape(-1)
        names.append(k)
    return out, names


def c3(count=4096, n=512):
    """BASELINE config 3: batch of 4096 synthetic 512x512 images (seed 1234+i)."""
    return batch(count, n, 1234)


def c4(count=1024, n=4096):
    """BASELINE config 4: batch of 1024 synthetic 4096x4096 images (seed 2000+i)."""
    return batch(count, n, 2000)


def c5(count=16384, n=512):
    """BASELINE config 5: mixed-entropy batch in equal thirds random / smooth / const."""
    return batch(count, n, 3000, classes=("random", "smooth", "const"))
