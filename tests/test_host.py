"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header declares
(no compute calls), state_dict layouts match the reference manifests, schedules and the tile
planner behave."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from moleculesde_b200 import build
    path = build.build()
    assert os.path.exists(path)
    return path


def test_library_exports_header_symbols(built_lib):
    header = open(os.path.join(REPO, "include", "molsde_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(molsde_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    L = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/molsde_b200.h but not exported"
    from moleculesde_b200 import _abi
    assert sorted(_abi.EXPORTS) == declared
    L.molsde_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.molsde_version()


def test_param_offsets_match_header():
    from moleculesde_b200 import _abi
    from moleculesde_b200 import sde_2d_to_3d as M
    hdr = open(os.path.join(REPO, "moleculesde_b200", "csrc", "sde2d3d_params.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(MOLSDE_\w+)\s+(\d+)", hdr)}
    assert defs["MOLSDE_P_TOTAL"] == M.P_TOTAL == defs["MOLSDE_P_BASIS0"] + 2 * defs["MOLSDE_P_BASIS_STRIDE"]
    assert defs["MOLSDE_P_GAT0"] == M.P_GAT0 == defs["MOLSDE_P_E0_END"] == M.P_E0_END and defs["MOLSDE_P_GAT_SZ"] == M.P_GAT_SZ
    assert defs["MOLSDE_P_BASIS0"] == M.P_BASIS0 == M.P_GAT0 + 4 * M.P_GAT_SZ
    assert defs["MOLSDE_P_BASIS_SZ"] == M.P_BASIS_SZ <= defs["MOLSDE_P_BASIS_STRIDE"] == M.P_BASIS_STRIDE
    for k, v in M._G.items():
        assert defs["MOLSDE_G_" + k] == v and v % 4 == 0
    for k, v in M._B.items():
        assert defs["MOLSDE_B_" + k] == v and v % 4 == 0
    for name in ("GFP_DIST_W", "GFP_COFF_W", "E0_HV", "E0_OB", "E0_BT"):
        assert defs["MOLSDE_P_" + name] == getattr(M, "P_" + name)
    assert defs["MOLSDE_E0_BT_FLOATS"] == M.E0_BT_FLOATS == 2 * 32 * 32 * 2 // 4
    # sections are whole numbers of 128-byte lines (TMA bulk copies, 128 B aligned operand tiles); blocks do not overlap
    for v in (M.P_E0_BT, M.P_E0_END, M.P_GAT_SZ, M.P_BASIS_STRIDE, M._G["F0C"], M._G["F3C"], M._G["WEC"], M._G["WQKVS"]):
        assert v % 32 == 0
    assert M.P_E0_BT + 11 * M.E0_BT_FLOATS == M.P_E0_END
    assert M._G["F0C"] + 1024 == M._G["F3C"] and M._G["F3C"] + 1024 == M._G["WEC"] and M._G["WEC"] + 1024 == M._G["BQKVS"]
    assert M._G["BQKVS"] + 128 == M._G["LN1_W"] and M._G["LN2_B"] + 32 == M._G["WP_SZ"] == M._G["WQKVS"]
    assert M._G["WQKVS"] + 2 * 128 * 32 * 2 // 4 == M.P_GAT_SZ
    assert M._B["W1_LO"] == 128 * 64 * 2 // 4 and M._B["W1_LO"] + 4096 == M._B["EPI"] and M._B["EPI"] + 512 == M._B["B2"]
    api = open(os.path.join(REPO, "include", "molsde_b200.h")).read()
    assert int(re.search(r"#define MOLSDE_TILE_LD (\d+)", api).group(1)) == _abi.TILE_LD


def test_state_dict_layout_matches_reference(golden):
    from moleculesde_b200.sde_2d_to_3d import SDEModel2Dto3D_02
    m = SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0,
                          num_diffusion_timesteps=1000, SDE_type="VE", use_extend_graph=True)
    mine = {k: (tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()}
    assert mine == golden["manifest"]["sde2d3d"]


def test_all_module_state_dicts_match_reference_manifest(golden):
    """GIN, SchNet and the dense 3D->2D model expose exactly the reference's state_dict keys, shapes and dtypes (the manifest
    was recorded from the unmodified reference classes), so `model_complete.pth` checkpoints move in both directions; a
    ParamStore re-homing the parameters into its flat buffer does not change that."""
    import torch
    from moleculesde_b200.gnn import GNN
    from moleculesde_b200.pretrain import ParamStore
    from moleculesde_b200.schnet import SchNet
    from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense
    gnn = GNN(5, 300, JK="last", drop_ratio=0.0, gnn_type="GIN")
    sch = SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean", node_class=119)
    m32 = SDEModel3Dto2D_node_adj_dense(dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4,
                                        emb_dim=300, num_linears=3, beta_min=0.1, beta_max=1.0, num_diffusion_timesteps=1000,
                                        SDE_type="VE", num_class_X=119, noise_on_one_hot=True)
    mods = {"gnn": gnn, "schnet": sch, "sde3d2d": m32}

    def manifest(m):
        return {k: (tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()}

    for name, m in mods.items():
        assert manifest(m) == golden["manifest"][name], name
    before = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in mods.items()}
    store = ParamStore(mods, torch.device("cpu"))
    for name, m in mods.items():
        assert manifest(m) == golden["manifest"][name], name + " (after ParamStore)"
        for k, v in m.state_dict().items():
            assert torch.equal(v, before[name][k]), (name, k)
    # the channel-stacked layout: per-channel q/k/v parameters of a dense edge layer are adjacent in the flat buffer
    idx = store.index["sde3d2d"]
    a = idx["edge_score_network.layers.1.attn.0.func_q.layers.1.weight"]
    b = idx["edge_score_network.layers.1.attn.1.func_q.layers.1.weight"]
    assert b[0] == a[0] + a[1]


def test_flat_layout_stacks_of_the_fused_tape_paths():
    """`pretrain._layout_key` places (a) q | k | v | skip of every GATLayer and (b) the filter-network layers of all SchNet
    interactions side by side in the flat buffer, so that `_stacked` returns zero-copy [128, 32] / [G*128, 51] / [G, 128, 128]
    views (data AND gradient) -- the precondition of the one-GEMM q|k|v|skip projection and of the hoisted SchNet filter stack
    (`pretrain._gat_layer`, `tape_schnet`); state_dict values are untouched by the re-ordering."""
    from moleculesde_b200 import sde_2d_to_3d as M
    from moleculesde_b200.pretrain import ParamStore, _QKVS, _stacked
    from moleculesde_b200.schnet import SchNet
    torch.manual_seed(3)
    m23 = M.SDEModel2Dto3D_02(300, 32, None, 0.2, 1.0, 1000, "VE", use_extend_graph=True)
    sch = SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean", node_class=119)
    before = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("sde2d3d", m23), ("schnet", sch))}
    store = ParamStore({"schnet": sch, "sde2d3d": m23}, torch.device("cpu"))
    for n, m in (("sde2d3d", m23), ("schnet", sch)):
        for k, v in m.state_dict().items():
            assert torch.equal(v, before[n][k]), (n, k)
    P = store.vars("sde2d3d")
    for m in range(2):
        for c in range(2):
            pf = f"score_network.gnn_layers.{m}.{c}."
            W = _stacked(P, [pf + f"MHA.{n}.weight" for n in _QKVS], (128, 32))
            b = _stacked(P, [pf + f"MHA.{n}.bias" for n in _QKVS], (128,))
            assert W is not None and b is not None, pf
            for j, n in enumerate(_QKVS):
                lin = getattr(m23.score_network.gnn_layers[m][c].MHA, n)
                assert torch.equal(W.data[32 * j:32 * j + 32], lin.weight.data) and torch.equal(b.data[32 * j:32 * j + 32], lin.bias.data)
                assert W.data[32 * j:].data_ptr() == lin.weight.data_ptr() and W.grad[32 * j:].data_ptr() == P[pf + f"MHA.{n}.weight"].grad.data_ptr()
    assert _stacked(P, [f"score_network.gnn_layers.0.0.MHA.{n}.weight" for n in _QKVS], (128, 32)) is \
        _stacked(P, [f"score_network.gnn_layers.0.0.MHA.{n}.weight" for n in _QKVS], (128, 32)), "cached per parameter dict"
    S = store.vars("schnet")
    G = 6
    W0 = _stacked(S, [f"interactions.{i}.mlp.0.weight" for i in range(G)], (G * 128, 51))
    b0 = _stacked(S, [f"interactions.{i}.mlp.0.bias" for i in range(G)], (G * 128,))
    W2 = _stacked(S, [f"interactions.{i}.mlp.2.weight" for i in range(G)], (G, 128, 128))
    b2 = _stacked(S, [f"interactions.{i}.mlp.2.bias" for i in range(G)], (G * 128,))
    assert None not in (W0, b0, W2, b2)
    for i in range(G):
        mlp = sch.interactions[i].mlp
        assert torch.equal(W0.data[128 * i:128 * i + 128], mlp[0].weight.data) and torch.equal(W2.data[i], mlp[2].weight.data)
        assert torch.equal(b0.data[128 * i:128 * i + 128], mlp[0].bias.data) and torch.equal(b2.data[128 * i:128 * i + 128], mlp[2].bias.data)
    # parameters that are NOT adjacent give None (the tapes then keep the per-layer path)
    assert _stacked(S, ["interactions.0.mlp.0.weight", "interactions.0.mlp.2.weight"], (256, 51)) is None


def test_packed_blob_roundtrip(golden):
    from conftest import sd_from_manifest
    from moleculesde_b200 import sde_2d_to_3d as M
    m = M.SDEModel2Dto3D_02(300, 32, None, 0.2, 1.0, 1000, "VE", use_extend_graph=True)
    sd = sd_from_manifest(golden["manifest"]["sde2d3d"], 1)
    m.load_state_dict(sd)
    pk = m.packed_params()
    blob = pk["blob"]
    assert blob.numel() == M.P_TOTAL

    def tile(off, R, K):   # hi + lo of a tcgen05 B tile pair stored at float offset `off`
        n = R * K // 2
        return M.unpack_umma_tile_f16(blob[off:off + n], R, K) + M.unpack_umma_tile_f16(blob[off + n:off + 2 * n], R, K)

    # Fourier sub-block tiles: the fused project.0 o coff_mlp hidden layer reproduces the two-layer reference expression when the
    # features are fed in the kernel's K order [sin f(16h..) | cos f(16h..)] per sub-block
    g_i, g_j, ang = torch.randn(5, 128), torch.randn(5, 128), torch.randn(5, 2)   # gfp features [sin 32 | cos 32] x 2 per node
    emb_i = g_i @ sd["coff_mlp.weight"].t() + sd["coff_mlp.bias"]
    emb_j = g_j @ sd["coff_mlp.weight"].t() + sd["coff_mlp.bias"]
    ref = torch.cat([ang, emb_i, emb_j], -1) @ sd["project.layers.0.weight"].t() + sd["project.layers.0.bias"]
    feats = torch.cat([g_i, g_j], -1)                     # [5, 256] = blocks 1..4, each [sin 32 | cos 32]
    got = torch.zeros(5, 32)
    for blk in range(1, 5):
        for half in range(2):
            k0 = (blk - 1) * 64 + half * 16
            a = torch.cat([feats[:, k0:k0 + 16], feats[:, k0 + 32:k0 + 48]], dim=1)
            got += a @ tile(M.P_E0_BT + (2 * blk + half) * M.E0_BT_FLOATS, 32, 32).t()
    hv = blob[M.P_E0_HV:M.P_E0_HV + 128].view(32, 4)
    got = got + hv[:, 0] + ang[:, :1] * hv[:, 1] + ang[:, 1:] * hv[:, 2]
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
    assert torch.all(hv[:, 3] == 0)
    # input_mlp over gfp(d) (sub-blocks 0, 1) and project.1 (tile 10)
    w_in = sd["input_mlp.layers.0.weight"]
    t0, t1 = tile(M.P_E0_BT, 32, 32), tile(M.P_E0_BT + M.E0_BT_FLOATS, 32, 32)
    assert float((torch.cat([t0[:, :16], t1[:, :16], t0[:, 16:], t1[:, 16:]], 1) - w_in).abs().max()) <= 2.0 ** -21 * float(w_in.abs().max())
    w_p1 = sd["project.layers.1.weight"]
    assert float((tile(M.P_E0_BT + 10 * M.E0_BT_FLOATS, 32, 32) - w_p1).abs().max()) <= 2.0 ** -21 * float(w_p1.abs().max())
    ob = blob[M.P_E0_OB:M.P_E0_OB + 64].view(32, 2)
    assert torch.equal(ob[:, 0], sd["input_mlp.layers.0.bias"]) and torch.equal(ob[:, 1], sd["project.layers.1.bias"])
    base = M.P_GAT0 + 3 * M.P_GAT_SZ
    we = tile(base + M._G["WEC"], 32, 32)
    ref_we = sd["score_network.gnn_layers.1.1.MHA.lin_edge.weight"]
    assert float((we - ref_we).abs().max()) <= 2.0 ** -21 * float(ref_we.abs().max())
    # node GEMMs: q | k | v | skip as ONE B tile [128 n][32 k]; FFN.0 / FFN.3 tiles; bias vector in the same order
    wq = tile(base + M._G["WQKVS"], 128, 32)
    for part, key in enumerate(("lin_query", "lin_key", "lin_value", "lin_skip")):
        ref_w = sd[f"score_network.gnn_layers.1.1.MHA.{key}.weight"]
        assert float((wq[32 * part:32 * part + 32] - ref_w).abs().max()) <= 2.0 ** -21 * float(ref_w.abs().max()), key
        assert torch.equal(blob[base + M._G["BQKVS"] + 32 * part:base + M._G["BQKVS"] + 32 * part + 32],
                           sd[f"score_network.gnn_layers.1.1.MHA.{key}.bias"])
    for off, key in ((M._G["F0C"], "FFN.0"), (M._G["F3C"], "FFN.3")):
        ref_w = sd[f"score_network.gnn_layers.1.1.{key}.weight"]
        assert float((tile(base + off, 32, 32) - ref_w).abs().max()) <= 2.0 ** -21 * float(ref_w.abs().max())
    assert torch.equal(blob[base + M._G["LN2_B"]:base + M._G["LN2_B"] + 32], sd["score_network.gnn_layers.1.1.norm2.bias"])
    # basis MLP: tcgen05 B tile [128 n][64 k] (hi + lo == weight to 2^-22) and the epilogue table {b1, w2[0..2]} per hidden unit
    base = M.P_BASIS0 + M.P_BASIS_STRIDE
    w1 = sd["score_network.basis_mlp_modules.1.0.weight"]  # [128 (n), 64 (k)]
    assert float((tile(base + M._B["W1_HI"], 128, 64) - w1).abs().max()) <= 2.0 ** -21 * float(w1.abs().max())
    hi_words = blob[base + M._B["W1_HI"]:base + M._B["W1_HI"] + 4096].view(torch.float16)
    for n, k in ((0, 0), (5, 3), (8, 4), (77, 41), (127, 63)):   # canonical K-major core-matrix addressing (csrc/sde2d3d_params.h)
        idx = (k // 8) * (128 * 8) + (n // 8) * 64 + (n % 8) * 8 + (k % 8)
        assert hi_words[idx] == w1[n, k].half()
    epi = blob[base + M._B["EPI"]:base + M._B["EPI"] + 512].view(128, 4)
    assert torch.equal(epi[:, 0], sd["score_network.basis_mlp_modules.1.0.bias"])
    assert torch.equal(epi[:, 1:].t().contiguous(), sd["score_network.basis_mlp_modules.1.2.weight"])
    assert torch.equal(blob[base + M._B["B2"]:base + M._B["B2"] + 3], sd["score_network.basis_mlp_modules.1.2.bias"])


def test_sde_schedules_match_oracle():
    from moleculesde_b200 import sde as S
    from oracle import model as O
    t = torch.linspace(1, 1e-4, 1000)
    for mine, ref in ((S.VESDE(0.2, 1.0, 1000), O.VESDE(0.2, 1.0, 1000)), (S.VPSDE(0.2, 1.0, 1000), O.VPSDE(0.2, 1.0, 1000))):
        x = torch.randn(1000, 3)
        m1, s1 = mine.marGINal_prob(x, t)
        m2, s2 = ref.marginal_prob(x, t)
        assert torch.equal(m1, m2) and torch.equal(s1, s2)
        f1, g1 = mine.discretize(x, t)
        f2, g2 = ref.discretize(x, t)
        assert torch.equal(f1, f2) and torch.equal(g1, g2)
        tab = mine.step_table(t)
        assert torch.equal(tab[:, 0], s2) and torch.equal(tab[:, 1], g2)
        assert torch.equal(tab[:, 3], ref.corrector_alpha(t))
        torch.testing.assert_close(tab[:, 2:3] * x - x, f2, rtol=0, atol=0)


def test_tile_planner(golden_batch):
    from moleculesde_b200 import _abi
    from moleculesde_b200.graph import CSR
    from moleculesde_b200.plan import build_plan
    _, batch = golden_batch
    ei = batch.extended_edge_index
    N = batch.positions.size(0)
    deg = torch.bincount(ei[0], minlength=N)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]).int()
    csr = CSR(rowptr, ei[1].int())
    plan = build_plan(csr, batch.ptr)
    tt = plan.tile_tgt_ptr.numpy()
    ct = plan.chunk_tile_ptr.numpy()
    assert tt[0] == 0 and tt[-1] == N and np.all(np.diff(tt) > 0)
    rp = rowptr.numpy()
    assert np.all(rp[tt[1:]] - rp[tt[:-1]] <= _abi.TILE_EDGES)
    # chunks are unions of whole molecules within the per-CTA limits
    mol_bounds = set(batch.ptr.tolist())
    for c in range(plan.num_chunks):
        a, b = tt[ct[c]], tt[ct[c + 1]]
        assert a in mol_bounds and b in mol_bounds and b - a <= _abi.CHUNK_MAX_NODES
        assert ct[c + 1] - ct[c] <= _abi.MAX_CHUNK_TILES
    # fixed groups: one chunk per group
    plan2 = build_plan(csr, batch.ptr, [0, 3, 8])
    assert plan2.num_chunks == 2
    with pytest.raises(_abi.MolsdeError):
        big = torch.tensor([0, 300])
        build_plan(CSR(torch.zeros(301, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)), big, [0, 1])


def test_no_cpu_fallback():
    from moleculesde_b200 import _abi
    from moleculesde_b200.graph import radius_graph
    with pytest.raises(_abi.MolsdeError):
        radius_graph(torch.zeros(4, 3), 10.0, torch.zeros(4, dtype=torch.long), 1)


def test_edge_segments_without_batch_vector():
    """`GNN.forward(x, edge_index, edge_attr)` carries no `batch`: the sub-graph boundaries recovered from the collated edge list
    (graph.edge_segments) must partition nodes and edges consistently (every edge inside its segment's node range)."""
    import torch
    from moleculesde_b200.data import Batch, synth_molecules
    from moleculesde_b200.graph import edge_segments
    mols = synth_molecules(40, 5, "pcqm")
    mols[7].edge_index = mols[7].edge_index[:, :0]          # a molecule without bonds: its atoms join the preceding segment
    mols[7].edge_attr = mols[7].edge_attr[:0]
    b = Batch.from_data_list(mols)
    N = b.x.size(0)
    node_ptr, edge_ptr, G = edge_segments(b.edge_index, N)
    assert node_ptr[0] == 0 and node_ptr[-1] == N and edge_ptr[0] == 0 and edge_ptr[-1] == b.edge_index.size(1)
    assert torch.all(node_ptr[1:] > node_ptr[:-1]) and torch.all(edge_ptr[1:] > edge_ptr[:-1])
    assert 39 <= G <= 2 * 40, "about one segment per molecule (more only for disconnected molecules)"
    for g in range(G):
        e = b.edge_index[:, int(edge_ptr[g]):int(edge_ptr[g + 1])]
        assert int(e.min()) >= int(node_ptr[g]) and int(e.max()) < int(node_ptr[g + 1])
    # flipped / by-target-sorted edge lists (what the by-source CSR is built from) segment the same way
    order = torch.argsort(b.edge_index[1], stable=True)
    flipped = torch.stack([b.edge_index[1][order], b.edge_index[0][order]])
    n2, e2, G2 = edge_segments(flipped, N)
    for g in range(G2):
        e = flipped[:, int(e2[g]):int(e2[g + 1])]
        assert int(e.min()) >= int(n2[g]) and int(e.max()) < int(n2[g + 1])
    # no edges at all
    n3, e3, G3 = edge_segments(b.edge_index[:, :0], N)
    assert G3 == 1 and n3.tolist() == [0, N] and e3.tolist() == [0, 0]


def test_edge_segments_randomised_edge_orders():
    """Segments stay valid when a molecule's edges are stored in any order and when a molecule has several components: a cut may
    only appear where the node ranges before / after it are disjoint (coarser segments are fine, wrong ones are not)."""
    import torch
    from moleculesde_b200.data import Batch, synth_molecules
    from moleculesde_b200.graph import edge_segments
    g = torch.Generator().manual_seed(11)
    for seed in range(6):
        mols = synth_molecules(12, 100 + seed, "pcqm")
        for m in mols:
            E = m.edge_index.size(1)
            keep = torch.rand(E // 2, generator=g) > 0.25            # drop ~25 % of the bonds (both directions): fragments appear
            keep = keep.repeat_interleave(2)
            perm = torch.randperm(int(keep.sum()), generator=g)      # and shuffle the edge order inside the molecule
            m.edge_index = m.edge_index[:, keep][:, perm]
            m.edge_attr = m.edge_attr[keep][perm]
        b = Batch.from_data_list(mols)
        N, E = b.x.size(0), b.edge_index.size(1)
        node_ptr, edge_ptr, G = edge_segments(b.edge_index, N)
        assert int(node_ptr[0]) == 0 and int(node_ptr[-1]) == N and int(edge_ptr[-1]) == E
        seg_of_edge = torch.bucketize(torch.arange(E), edge_ptr[1:].long(), right=True)
        lo, hi = node_ptr[seg_of_edge].long(), node_ptr[seg_of_edge + 1].long()
        assert torch.all((b.edge_index >= lo) & (b.edge_index < hi)), "every edge lies inside the node range of its segment"
        assert G <= E and G >= 1


def test_umma_tile_split_randomised():
    """fp16 hi/lo operand tiles of the tcgen05 GEMMs: hi + lo reproduces the fp32 weight to 2^-22 relative (fp16's narrow exponent:
    absolute 2^-25 below 2^-14), the canonical core-matrix addressing round-trips, zeros stay zeros."""
    from moleculesde_b200.sde_2d_to_3d import umma_tile_split_words, unpack_umma_tile_f16
    g = torch.Generator().manual_seed(0)
    for R, K, scale in ((32, 32, 1.0), (128, 64, 0.05), (128, 32, 30.0), (8, 8, 1e-3)):
        w = torch.randn(R, K, generator=g) * scale
        words = umma_tile_split_words(w)
        n = R * K // 2
        assert words.numel() == 2 * n and words.dtype == torch.float32
        back = unpack_umma_tile_f16(words[:n], R, K) + unpack_umma_tile_f16(words[n:], R, K)
        assert float((back - w).abs().max()) <= max(2.0 ** -21 * float(w.abs().max()), 2.0 ** -24)
        hi = unpack_umma_tile_f16(words[:n], R, K)
        assert torch.equal(hi, w.half().float())
    assert torch.equal(umma_tile_split_words(torch.zeros(8, 8)), torch.zeros(64))



def test_build_plan_host_matches_python_specification():
    """`molsde_build_plan_host` (C, sequential) == the pure-Python plan on random batches: free packing and fixed sampling groups,
    including the error cases (a node with > 128 incoming edges, a group beyond 224 atoms)."""
    import numpy as np
    from moleculesde_b200 import _abi
    from moleculesde_b200.graph import CSR
    from moleculesde_b200.plan import build_plan, build_plan_py
    rng = np.random.default_rng(0)
    for trial in range(30):
        B = int(rng.integers(1, 40))
        sizes = rng.integers(1, 60 if trial % 3 else 120, size=B)
        nptr = np.concatenate([[0], np.cumsum(sizes)])
        N = int(nptr[-1])
        deg = rng.integers(0, 40 if trial % 4 else 100, size=N)
        rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        csr = CSR(torch.from_numpy(rowptr), torch.zeros(int(rowptr[-1]), dtype=torch.int32))
        node_ptr = torch.from_numpy(nptr.astype(np.int32))
        groups = None
        if trial % 2:
            cuts = sorted(set([0, B] + rng.integers(0, B + 1, size=3).tolist()))
            groups = cuts
        res = []
        for fn in (build_plan, build_plan_py):
            try:
                p = fn(csr, node_ptr, groups)
                res.append((p.num_chunks, p.num_tiles, p.max_chunk_tiles, p.chunk_tile_ptr.tolist(), p.tile_tgt_ptr.tolist(),
                            p.chunk_order.tolist()))
            except _abi.MolsdeError as e:
                res.append(("error", str(e).split(" ")[0]))
        assert res[0] == res[1], (trial, res[0][:3], res[1][:3])


def test_fastcall_matches_ctypes_calling_convention(built_lib):
    """`_molsde_fastcall` (csrc/fastcall.c) passes interleaved int32 / int64 / uint64 / pointer / float arguments -- more integer-class
    arguments than registers, so some travel on the stack -- exactly like ctypes does, None becomes NULL, ints are accepted for float
    parameters, wrong arity raises; and `_abi.lib()` rebinds the status-returning entry points through it."""
    import ctypes
    import numpy as np
    from moleculesde_b200 import _abi
    L = _abi.lib()
    assert type(L).__name__ == "_FastLib" and type(L.molsde_debug_echo).__name__ == "Bound"
    assert type(L.molsde_version).__name__ != "Bound"                    # returns a string: stays on ctypes
    assert type(L.molsde_sde2d3d_pc_sample).__name__ != "Bound"          # takes structures by reference: stays on ctypes
    raw = ctypes.CDLL(built_lib)
    raw.molsde_debug_echo.argtypes = L._cdll.molsde_debug_echo.argtypes
    raw.molsde_debug_echo.restype = ctypes.c_int32
    args = (-(2 ** 40) - 3, 1.25, -7, 0x7F00DEADBEE0, -3.5, 2 ** 41 + 1, 123456, 2 ** 63 + 5, -9, 0.1, 77, -2 ** 31)
    a, b = np.zeros(16), np.zeros(16)
    assert L.molsde_debug_echo(*args, a.ctypes.data) == 7
    assert raw.molsde_debug_echo(*args, b.ctypes.data) == 7
    assert np.array_equal(a, b) and a[0] == args[0] and a[9] == np.float32(0.1) and a[11] == -2 ** 31
    c = np.zeros(16)
    assert L.molsde_debug_echo(1, 2, 3, None, 4, 5, 6, 7, 8, 9, 10, 11, c.ctypes.data) == 7   # ints for floats, None for the pointer
    assert c[1] == 2.0 and c[3] == 0.0 and c[4] == 4.0 and c[9] == 9.0
    assert L.molsde_debug_echo(*args, None) == -1                                               # MOLSDE_ERR_INVALID
    with pytest.raises(TypeError):
        L.molsde_debug_echo(1, 2, 3)
