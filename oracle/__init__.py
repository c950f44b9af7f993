"""CPU oracle for the TRIS Stage-1 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tris_b200/`` imports this package.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs (as the checker / CPU arm, never as
the product path).

Parity status: the reference repository (fawnliu/TRIS @ c6666b3) ships no tests,
golden vectors or fixtures for this path, so the restatement is pinned against
outputs of the *unmodified reference modules imported in the build container*
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).  Pretrained CLIP
weights are not available offline; all comparisons use the deterministic
random weights of ``oracle.weights``.
"""
