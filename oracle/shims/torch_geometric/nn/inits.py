import math


def glorot(t):
    if t is not None:
        stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        t.data.uniform_(-stdv, stdv)


def zeros(t):
    if t is not None:
        t.data.fill_(0)
