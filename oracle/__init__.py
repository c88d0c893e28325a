"""CPU oracle for the LICOS learned-codec hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic of this path lives in the third-party package
``compressai`` (un-pinned in /root/reference/environment.yml:28-29, effective
version >= 1.2.0), which is neither vendored under /root/reference nor
installable in this image, and the reference's own tests hold no golden
vectors for it.  This package restates the published
CompressAI algorithms; its self-checks are in tests/test_oracle_*.py.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from here.  The product package
``licos_b200`` never does.
"""
