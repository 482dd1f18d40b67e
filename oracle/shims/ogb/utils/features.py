"""ogb 1.2.1 feature vocab sizes (restated; not verifiable offline)."""
from moleculesde_b200.data import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS


def get_atom_feature_dims():
    return list(ATOM_FEATURE_DIMS)


def get_bond_feature_dims():
    return list(BOND_FEATURE_DIMS)
