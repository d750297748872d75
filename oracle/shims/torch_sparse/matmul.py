"""Names imported (and never used) by other_models.py:5."""


def spmm_max(*a, **k):
    raise NotImplementedError


spmm_mean = spmm_add = spmm_max
