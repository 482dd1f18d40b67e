"""Import the UNMODIFIED reference sources over `oracle/shims` (build container only).

TEST INFRASTRUCTURE.  `/root/reference` does not exist on the GPU box; only
`tests/golden/make_golden.py` and the container-only tests call `load()`.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MOLSDE_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Geom3D"))


def load():
    """Put shims + reference on sys.path; returns a namespace of the hot-path symbols."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    repo = os.path.dirname(_HERE)
    for p in (REFERENCE_ROOT, os.path.join(_HERE, "shims"), repo):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    # `Geom3D/datasets/__init__.py` pulls rdkit/ogb featurisers (out of scope): register the
    # package bare so that `Geom3D.datasets.dataset_3D` (extend_graph) imports alone.
    import Geom3D  # noqa: F401
    if "Geom3D.datasets" not in sys.modules:
        pkg = types.ModuleType("Geom3D.datasets")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "Geom3D", "datasets")]
        sys.modules["Geom3D.datasets"] = pkg
    ns = types.SimpleNamespace()
    from Geom3D.models import SchNet, GNN
    from Geom3D.models.MoleculeSDE import SDEModel2Dto3D_02, SDEModel3Dto2D_node_adj_dense
    from Geom3D.models.MoleculeSDE.SDE_model_2D_to_3D import SDEModel2Dto3D_01
    from Geom3D.models.MoleculeSDE import SDE_sparse, SDE_dense
    from Geom3D.datasets.dataset_3D import extend_graph
    ns.SchNet, ns.GNN = SchNet, GNN
    ns.SDEModel2Dto3D_02 = SDEModel2Dto3D_02
    ns.SDEModel2Dto3D_01 = SDEModel2Dto3D_01
    ns.SDEModel3Dto2D_node_adj_dense = SDEModel3Dto2D_node_adj_dense
    ns.SDE_sparse, ns.SDE_dense = SDE_sparse, SDE_dense
    ns.extend_graph = extend_graph
    ns.examples = _load_examples()
    return ns


def _exec_slice(path: str, start: int, end: int, glb: dict) -> None:
    """Execute lines [start, end] (1-based, inclusive) of a reference script in `glb`.
    The inference scripts cannot be imported whole (broken imports and
    argparse at import time), so their sampler classes are executed from the file where it
    lies -- nothing is copied into this repo."""
    with open(path) as f:
        lines = f.readlines()
    src = "\n" * (start - 1) + "".join(lines[start - 1:end])
    exec(compile(src, path, "exec"), glb)


def _load_examples():
    import abc
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from tqdm import trange
    ex = types.SimpleNamespace()
    ex_dir = os.path.join(REFERENCE_ROOT, "examples")
    # examples/util.py: do_CL / dual_CL (lines 22-79)
    g = {"torch": torch, "nn": nn, "F": F}
    _exec_slice(os.path.join(ex_dir, "util.py"), 22, 79, g)
    ex.do_CL, ex.dual_CL = g["do_CL"], g["dual_CL"]
    # 2D->3D sampler: position_PC_generation + Predictor/Corrector classes (lines 92-212)
    from Geom3D.models.MoleculeSDE.SDE_sparse import VPSDE, VESDE, subVPSDE
    g = {"torch": torch, "abc": abc, "trange": trange, "VPSDE": VPSDE, "VESDE": VESDE,
         "subVPSDE": subVPSDE, "args": types.SimpleNamespace(device="cpu"), "print": lambda *a, **k: None}
    _exec_slice(os.path.join(ex_dir, "pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py"), 92, 212, g)
    ex.pc2d3d = types.SimpleNamespace(**{k: g[k] for k in (
        "position_PC_generation", "ReverseDiffusionPredictor", "LangevinCorrector")})
    ex.pc2d3d_globals = g
    return ex
