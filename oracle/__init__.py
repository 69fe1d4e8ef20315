"""CPU oracle for the ModelCompose composition hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Plain torch-CPU / numpy restatements of the reference's algorithms (each function cites
the reference file:line it follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product package ``modelcompose_b200`` never does (it fails loudly without its CUDA library).

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the authoring
container by ``tests/golden/make_golden.py`` (which executes the unmodified reference
sources from /root/reference through the shims in ``tests/golden/_reference_loader.py``)
and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks every
oracle function against those fixtures on every run.  The third-party arithmetic the
reference delegates to (peft 0.4.0 LoRA layer construction, transformers 4.31 rotary
helpers) is restated in the shims and is itself unpinned by any reference test — see
DESIGN.md "parity pinning".
"""
