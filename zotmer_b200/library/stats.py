# Log-domain statistics behind `zot jaccard -p`: log n!, log(a + b) from logs, log binomial coefficients
# (zotmer/library/stats.py:36-58, 77-92, 121-128) and the two routines the reference keeps in commands/jaccard.py:56-83 --
# the log of the regularised incomplete beta function as a series, and a beta quantile by bisection.  Every expression
# keeps the reference's operations in the reference's order: the printed values have to agree digit for digit.
import math


def factorial(n):
    r = 1
    for i in range(2, n + 1):
        r *= i
    return r


_small = [math.log(factorial(n)) for n in range(25)]


def logFac(n):
    if n < len(_small):
        return _small[n]
    return n * math.log(n) - n + math.log(n * (1 + 4 * n * (1 + 2 * n))) / 6.0 + math.log(math.pi) / 2.0


def logAdd(a, b):
    x = max(a, b)
    y = min(a, b)
    w = y - x
    return x + math.log1p(math.exp(w))


def logChoose(n, k):
    if k == 0 or k == n:
        return 0
    return logFac(n) - (logFac(n - k) + logFac(k))


def logBetaSeries(x, m, n):
    """log I_x(m, n) for integer m, n: n log(1 - x) + log sum_{j >= m} C(n + j - 1, j) x^j, the sum accumulated in the
    log domain until a term no longer changes it (commands/jaccard.py:56-70)"""
    logx = math.log(x)
    j = m
    coeff = logChoose(n + j - 1, j)
    total = coeff + j * logx
    while True:
        j += 1
        coeff += math.log((n + j - 1.0) / j)
        grown = logAdd(total, coeff + j * logx)
        if grown == total:
            return n * math.log1p(-x) + total
        total = grown


def betaQuantile(q, m, n):
    """the x with I_x(m, n) = q, by bisection on [1e-10, 1 - 1e-10] down to a bracket of 1e-7; returns the lower end
    (commands/jaccard.py:72-83)"""
    target = math.log(q)
    lo, hi = 1e-10, 1 - 1e-10
    while (hi - lo) > 1e-7:
        mid = (hi + lo) / 2.0
        if logBetaSeries(mid, m, n) < target:
            lo = mid
        else:
            hi = mid
    return lo
