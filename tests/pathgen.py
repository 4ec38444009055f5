"""Seeded random path / paint generators shared by parity tests and bench.py (SURVEY.md §8(d) C2 recipe)."""
import math

import numpy as np

MOVE, LINE, QUAD, CUBIC, CLOSE = 0, 1, 2, 3, 4


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def u(self):  # [0, 1)
        return (self.next() >> 11) * (1.0 / (1 << 53))

    def uniform(self, a, b):
        return a + (b - a) * self.u()

    def randint(self, a, b):  # inclusive
        return a + int(self.u() * (b - a + 1))

    def log_uniform(self, a, b):
        return math.exp(self.uniform(math.log(a), math.log(b)))


def random_path(rng, cx, cy, radius, n_seg=None, kinds=(0.5, 0.3, 0.2)):
    """Closed path of n_seg segments: cubic / quad / line with the given probabilities; control points =
    centre + radius * U(-1,1)^2."""
    if n_seg is None:
        n_seg = rng.randint(3, 12)

    def pt():
        return (np.float32(cx + radius * rng.uniform(-1, 1)), np.float32(cy + radius * rng.uniform(-1, 1)))

    verbs, pts = [MOVE], [pt()]
    for _ in range(n_seg):
        k = rng.u()
        if k < kinds[0]:
            verbs.append(CUBIC); pts += [pt(), pt(), pt()]
        elif k < kinds[0] + kinds[1]:
            verbs.append(QUAD); pts += [pt(), pt()]
        else:
            verbs.append(LINE); pts.append(pt())
    verbs.append(CLOSE)
    return np.array(verbs, dtype=np.uint8), np.array(pts, dtype=np.float32)


def random_stops(rng, n=None, opaque=False):
    if n is None:
        n = rng.randint(2, 8)
    offs = sorted(rng.u() for _ in range(n))
    if rng.u() < 0.5:
        offs[0], offs[-1] = 0.0, 1.0
    stops = []
    for o in offs:
        a = 1.0 if opaque else rng.randint(32, 255) / 255.0
        stops.append([o, rng.randint(0, 255) / 255.0, rng.randint(0, 255) / 255.0, rng.randint(0, 255) / 255.0, a])
    return stops


def random_paint_spec(rng, cx, cy, radius, solid=0.5, linear=0.3):
    k = rng.u()
    if k < solid:
        return {"kind": "solid", "color": (rng.randint(0, 255) / 255.0, rng.randint(0, 255) / 255.0,
                                           rng.randint(0, 255) / 255.0, rng.randint(32, 255) / 255.0)}
    spread = ("pad", "reflect", "repeat")[rng.randint(0, 2)]
    if k < solid + linear:
        a = rng.uniform(0, 2 * math.pi)
        return {"kind": "linear", "x0": cx - radius * math.cos(a), "y0": cy - radius * math.sin(a),
                "x1": cx + radius * math.cos(a) * 0.7, "y1": cy + radius * math.sin(a) * 0.7,
                "stops": random_stops(rng), "spread": spread}
    f = rng.uniform(0, 0.8) * radius
    a = rng.uniform(0, 2 * math.pi)
    return {"kind": "radial", "x0": cx + f * math.cos(a), "y0": cy + f * math.sin(a), "r0": 0.0, "x1": cx, "y1": cy,
            "r1": radius, "stops": random_stops(rng), "spread": spread}
