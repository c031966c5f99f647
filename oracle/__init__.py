"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the B200 hot path.

This package is a plain-PyTorch fp32 *restatement* of the reference's
(mrluin/TextualDegRemoval) restoration networks, written functionally over a
``state_dict`` so that it can be checked tensor-for-tensor against the reference
modules (``oracle/ref_loader.py`` + ``oracle/make_golden.py``; fixtures under
``tests/golden/``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``textualdegremoval_b200``) never imports anything from here and has no CPU
fallback: it raises if the CUDA extension is missing.

Parity status: PINNED for Restormer / RestormerRefFusion / NAFNet (restatement
checked against the imported reference modules, fixtures committed).
"""
