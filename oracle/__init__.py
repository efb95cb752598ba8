"""CPU oracle for the MuCon hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``mucon_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / CPU baseline.

Parity pin: the reference (yassersouri/MuCon) has no tests, golden vectors or
fixtures for this path.  The oracle is therefore pinned
against the reference *implementation itself*, imported unmodified from
``/root/reference/src`` in the build container:
  * ``tests/golden/make_golden.py`` runs the reference and freezes its outputs
    into ``tests/golden/*.npz`` (committed);
  * ``tests/test_oracle_vs_reference.py`` re-runs the comparison live whenever
    ``/root/reference`` exists (it does not exist on the GPU box).

Modules
  hyp_viterbi    hypothesis-table decoder, same algorithm and data flow as the
                 reference decoder (core/viterbi/viterbi.py) -- this is the
                 "reference CPU path" that bench.py times.
  dense_viterbi  dense (K x N x J) restatement with explicit dtypes; yields the
                 back-pointer table the CUDA kernel is compared against.
  poisson        Poisson length-model table / parameters (core/viterbi/length_model.py).
  masks          create_masks restatements (mucon/masks.py).
  backbone       torch fp32 functional restatement of the temporal backbone.
  coracle        ctypes binding of oracle/oracle.c (C dense restatement).
"""
