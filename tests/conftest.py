import os
import sys

import pytest

os.environ.setdefault("MOLSDE_CHECK_ABI", "1")   # layout assertions on every pointer handed to the C ABI (off in production)
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    import torch
    path = os.path.join(REPO, "tests", "golden", "golden_pcqm8.pt")
    return torch.load(path, weights_only=False)


@pytest.fixture(scope="session")
def golden_batch(golden):
    """The seeded synthetic batch the golden file was generated from (extended graph from the
    oracle restatement; checked against the golden copy in test_oracle_golden)."""
    import torch
    from moleculesde_b200.data import Batch, synth_molecules
    from oracle.ref_ops import extend_graph_index
    mols = synth_molecules(golden["meta"]["num_mols"], golden["meta"]["data_seed"])
    for m in mols:
        m.extended_edge_index = extend_graph_index(m.edge_index, m.num_nodes)
    return mols, Batch.from_data_list(mols)


def sd_from_manifest(manifest, seed):
    import torch
    from oracle.weights import fill_state_dict
    sd = {}
    for k, (shape, dtype) in manifest.items():
        sd[k] = torch.zeros(shape, dtype=getattr(torch, dtype.replace("torch.", "")))
    filled = fill_state_dict(sd, seed)
    # buffers the fill keeps as-is must carry their constructor values
    if "distance_expansion.offset" in filled:
        filled["distance_expansion.offset"] = torch.linspace(0.0, 10.0, shape_of(manifest, "distance_expansion.offset"))
    for k in filled:
        if k.endswith("equi_modules.0.eps") or k.endswith("equi_modules.1.eps"):
            filled[k] = torch.zeros(1)
    return filled


def shape_of(manifest, key):
    return manifest[key][0][0]
