from oracle.ref_ops import scatter, scatter_mean, scatter_add  # noqa: F401
