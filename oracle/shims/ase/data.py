"""Only `atomic_masses` is read (schnet.py:47) and only as an unused float64 buffer
(dipole=False on the hot path); 119 placeholders keep the state_dict shape."""
import numpy as np

atomic_masses = np.arange(119, dtype=np.float64)
