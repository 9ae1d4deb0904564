"""CPU oracle for the SubGNN subgraph message-passing hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / the timed CPU baseline.  Nothing under ``subgnn_b200/`` imports it;
the product path fails loudly when ``libsubgnn_b200.so`` is missing.

It is a restatement (numpy / plain torch-CPU / pure Python) of the reference
algorithms, each function citing the ``/root/reference`` file:line it follows.

Pinning status (see DESIGN.md "Oracle"):
  * walks, border nodes, k-hop border sets, degree sequences, N/P sampling,
    SP-min similarity, SG_MPN, LSTM walk encoder, SubGNN.forward / training
    step: PINNED against the unmodified reference modules imported under
    ``sys.modules`` stubs in the build container (``oracle/ref_loader.py``),
    same seeds => identical outputs; vectors committed under ``tests/golden``
    by ``tests/golden/make_golden.py``.
  * fastdtw (third-party ``fastdtw==0.3.4``, pinned in SubGNN.yml:109, source
    NOT under /root/reference): restated from the published pure-Python
    algorithm; hand-derived known-answer vectors only => "parity unpinned"
    for that one function.
"""
