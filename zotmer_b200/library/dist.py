"""
Qualitative set distances as functions of (a, b, c) = (|X n Y|, |X \\ Y|, |Y \\ X|)
(mirrors the `vec=False` branches of zotmer/library/dist.py:16-239).

The cardinalities come from the device (zb_pairs_abc replaces split(), dist.py:241-265); the
float formulas stay in host Python, written with the same operations in the same order as the
reference so the results agree to the last ulp.
"""
import math


def brayCurtis(a, b, c):
    return float(b + c) / float(2 * a + b + c)          # dist.py:40-41


def chord(a, b, c):
    return math.sqrt(2 * (1 - a / math.sqrt((a + b) * (a + c))))   # dist.py:67-68


hellinger = chord                                        # dist.py:93-94


def jaccard(a, b, c):
    return float(b + c) / float(a + b + c)               # dist.py:112-113


def kulczynski(a, b, c):
    a = float(a)
    b = float(b)
    c = float(c)
    return 1 - 0.5 * (a / (a + b) + a / (a + c))         # dist.py:168-172


def ochiai(a, b, c):
    return 1 - a / math.sqrt((a + b) * (a + c))          # dist.py:190-191


sorensen = brayCurtis                                    # dist.py:209-210


def whittaker(a, b, c):
    a = float(a)
    b = float(b)
    c = float(c)
    return 0.5 * (b / (a + b) + c / (a + c) + abs(a / (a + b) - a / (a + c)))   # dist.py:235-239
