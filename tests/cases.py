"""Seeded input builders shared by the CPU and GPU tests (numpy, fp32)."""
import numpy as np


def cloud(seed, b, n, kind="gauss", dup_frac=0.0):
    rng = np.random.RandomState(seed)
    if kind == "gauss":
        p = rng.randn(b, n, 3).astype(np.float32)
    elif kind == "uniform":
        p = rng.rand(b, n, 3).astype(np.float32) * np.array([80, 4, 70], dtype=np.float32)
    elif kind == "lattice":  # integer lattice: masses of EXACT distance ties
        p = rng.randint(0, 6, size=(b, n, 3)).astype(np.float32)
    elif kind == "identical":
        p = np.tile(rng.randn(b, 1, 3).astype(np.float32), (1, n, 1))
    else:
        raise ValueError(kind)
    nd = int(dup_frac * n)
    if nd:
        for s in range(b):
            src = rng.randint(0, n - nd, size=nd)
            p[s, n - nd:] = p[s, src]
    return np.ascontiguousarray(p)


def lidar(seed, b, n=16384):
    import torch  # noqa: F401
    from epnet_b200 import scenes
    return np.stack([scenes.lidar_scene(seed + i, n).numpy() for i in range(b)])
