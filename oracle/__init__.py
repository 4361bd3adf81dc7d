"""CPU oracle for CerberusNet's cost-volume hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The product
package ``cerberusnet_b200`` never imports it (tests/test_boundary.py enforces that).

Two restatements of the same reference semantics:

* ``oracle.c_oracle``      -- plain C (``costvolume_oracle.c``), ctypes-bound, numpy in/out.
* ``oracle.torch_oracle``  -- pure PyTorch, runs on CPU (and on CUDA for A/B on the GPU box);
                              autograd supplies the backward oracle.

Parity pin: the reference has no tests or golden vectors for this path; both restatements are
pinned against fixtures generated from the reference's own Python (tests/golden/make_golden.py)
and, on the GPU box, against the reference CUDA op compiled unmodified into ``oracle/_ref/``.
"""
