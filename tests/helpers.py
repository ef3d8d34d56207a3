"""Helpers shared by the CPU and GPU test modules."""
import numpy as np


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def splitmix_src(n_dofs, constrained=None, salt=0):
    """Deterministic synthetic vector: 2*u01(splitmix64(i ^ GOLDEN)) - 1."""
    i = (np.arange(n_dofs, dtype=np.uint64) + np.uint64((salt * 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF)) ^ np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = i + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    v = 2.0 * ((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)) - 1.0
    if constrained is not None:
        v[constrained] = 0.0
    return v


def hierarchy_levels(kind, p, n_fine):
    """Level list (degree, cells) coarse -> fine.
    kind 'h': geometric coarsening down to one cell (reference geometric driver).
    kind 'hp': p -> ... -> 1 by halving the degree (4->2->1), then geometric levels (BASELINE config 2)."""
    levels = []
    if kind == "h":
        n = n_fine
        while True:
            levels.append((p, n))
            if n % 2 or n == 1:
                break
            n //= 2
        return levels[::-1]
    degs = [p]
    while degs[-1] > 1:
        degs.append(max(1, degs[-1] // 2))
    for d in degs:
        levels.append((d, n_fine))
    n = n_fine
    while n % 2 == 0 and n > 1:
        n //= 2
        levels.append((1, n))
    return levels[::-1]
