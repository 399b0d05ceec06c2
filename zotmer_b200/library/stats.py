"""The two statistics helpers `zot jaccard -p` needs (mirrors zotmer/library/stats.py:36-58,77-92,121-128)."""
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
