"""CPU oracle for the MoleculeSDE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product path
(``moleculesde_b200``) never imports it and fails loudly when its CUDA library is
missing.

Contents
--------
``ref_ops``      restatement of the third-party ops the reference calls but does not
                 vendor (torch_scatter / torch_sparse / torch_cluster / torch_geometric,
                 pinned only by the reference README: pyg 2.0.2, torch 1.9.1).
``model``        restatement of the reference's own hot-path modules
                 (SchNet, SDEModel2Dto3D_02, EquivariantScoreNetwork, VE/VP SDEs,
                 predictor-corrector sampler, EBM_node_dot_prod, dense 3D->2D nets),
                 each function citing the reference file:line it follows.
``shims/``       tiny stand-in packages (torch_geometric, torch_scatter, ...) built on
                 ``ref_ops`` so that the UNMODIFIED reference sources under
                 ``/root/reference`` can be imported in the build container to generate
                 the golden fixtures in ``tests/golden`` (see ``tests/golden/make_golden.py``).

Parity status
-------------
The reference ships no tests or golden vectors.  ``model`` is pinned
against outputs of the reference's own source files executed here over ``shims``
(fixtures in ``tests/golden``).  The third-party op semantics inside ``ref_ops`` are
restated from the published behaviour of those libraries and are covered by
brute-force conformance tests only: **third-party boundary parity is unpinned by the
reference**.
"""
