"""TEST INFRASTRUCTURE ONLY.

CPU checkers for the DynaAlign all-pairs similarity hot path:

* ``oracle.ref``    -- ctypes binding of ``oracle/_ref/libdynaref.so`` (the reference's own C++,
                       compiled unmodified from /root/reference by ``oracle/Makefile``).
* ``oracle.port``   -- ctypes binding of ``oracle/libdynaoracle.so`` (our plain-C restatement).
* ``oracle.minhash_r`` -- numpy restatement of the pure-R pipeline in R/minHash.R.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product (``dynaalign_b200``) never does.
"""
