"""
CPU oracle for the threshold-Paillier modexp hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with CPython big integers (and optionally GMP through
``ctypes``), the arithmetic of the reference's hot path.  Nothing in the product package
``protocols.distributed_keygen_b200`` may import it: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker / the
timed CPU baseline.

Parity status: PINNED for the in-tree reference functions.  ``tests/golden/make_golden.py``
imports the reference's own ``PaillierSharedKey`` / ``DistributedPaillier`` classes from
``/root/reference`` (with shims for the un-vendored third-party packages, see that script) and
records their outputs on the reference's 24 golden key fixtures; ``tests/test_oracle.py`` checks
this oracle against those vectors.  The third-party ``pow_mod`` / ``mod_inv``
(``tno.mpc.encryption_schemes.utils ~=0.10``: ``gmpy2.powmod`` / ``gmpy2.invert`` when gmpy2 is
installed, CPython ``pow`` otherwise) are un-vendored; they return canonical residues, so CPython
``pow`` and GMP ``mpz_powm`` (both checked against each other here) define their value exactly.
The third-party ``Paillier.encrypt`` (``r^N mod N^2`` then ``(1+mN)·r^N``) is restated from its
published algorithm; it is pinned only through the reference tests' own criterion
``decrypt(encrypt(m)) == m`` on the golden keys.
"""
