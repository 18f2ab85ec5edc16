"""Generate near-minimax polynomial coefficients for sb_math.h (run once; output pasted into the header).

exp:  e^r   = 1 + r + r^2 * G(r),           r in [-ln2/2, ln2/2]
log:  log1p(r) = r - r^2/2 + r^3 * P(r),    r in [sqrt(.5)-1, sqrt(2)-1]
Chebyshev interpolation in float64 (within a hair of minimax), coefficients rounded to fp32.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as Pm
import mpmath as mp

def fit(f, a, b, deg):
    ch = C.Chebyshev.interpolate(f, deg, domain=[a, b])
    p = ch.convert(kind=Pm.Polynomial, domain=[-1, 1], window=[-1, 1])
    return p.coef

def G(r):
    r = np.asarray(r, dtype=np.float64)
    out = np.empty_like(r)
    for i, x in enumerate(r):
        x = mp.mpf(float(x))
        out[i] = float((mp.e**x - 1 - x) / x**2) if x != 0 else 0.5
    return out

def P(r):
    r = np.asarray(r, dtype=np.float64)
    out = np.empty_like(r)
    for i, x in enumerate(r):
        x = mp.mpf(float(x))
        out[i] = float((mp.log1p(x) - x + x * x / 2) / x**3) if x != 0 else 1.0 / 3
    return out

mp.mp.prec = 120
ln2 = float(np.log(2.0))
for deg in (4, 5):
    c = fit(G, -ln2 / 2, ln2 / 2, deg)
    print("exp G deg", deg, [np.float32(v).item().hex() for v in c], [float(np.float32(v)) for v in c])
for deg in (6, 7, 8):
    c = fit(P, np.sqrt(0.5) - 1, np.sqrt(2) - 1, deg)
    print("log P deg", deg, [np.float32(v).item().hex() for v in c], [float(np.float32(v)) for v in c])

def T(u):
    u = np.asarray(u, dtype=np.float64)
    out = np.empty_like(u)
    for i, x in enumerate(u):
        if x <= 0:
            out[i] = -1.0 / 3
        else:
            a = mp.sqrt(mp.mpf(float(x)))
            out[i] = float((mp.tanh(a) - a) / a**3)
    return out

for deg in (4, 5, 6):
    c = fit(T, 0.0, 0.55**2, deg)
    print("tanh T(a^2) deg", deg, [np.float32(v).item().hex() for v in c])
