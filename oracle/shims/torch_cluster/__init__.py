from oracle.ref_ops import radius_graph  # noqa: F401
