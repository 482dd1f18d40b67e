from oracle.ref_ops import coalesce, spspmm  # noqa: F401


class SparseTensor:  # isinstance checks only (equivariant_scorenetwork.py:67)
    pass
