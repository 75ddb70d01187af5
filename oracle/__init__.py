"""CPU oracle for the StableKeypoints hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker / the CPU baseline.  The product (``stablekeypoints_b200``) never
imports from here and fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference's own Python code, imported read-only from
``/root/reference`` in the build container (``oracle/ref_shim.py``) and committed as small
fixtures under ``tests/golden/`` by ``tests/golden/make_golden.py``.  The third-party model
underneath (``diffusers==0.8.0``, requirements.yaml:174) is absent from the container and is
restated from its published architecture in ``oracle/sd15.py``: that part is "parity unpinned"
against real diffusers.
"""
