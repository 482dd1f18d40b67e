"""moleculesde_b200 -- B200-native (sm_100a) implementation of the MoleculeSDE hot path.

The CUDA library (`libmolsde_b200.so`, C ABI declared in `include/molsde_b200.h`) is
loaded lazily by `moleculesde_b200._abi`; every compute entry point raises if it is
missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
