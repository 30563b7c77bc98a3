"""CPU oracle for the SAC update hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker (or as the timed
CPU baseline) -- never as the thing shipped.  The product (``sac`` package +
``libsacx.so``) must not import this package and has no CPU fallback.

Contents
--------
``mt_sample``   restatement of CPython's ``random.sample`` index stream
                (reference: sac/replay_buffer.py:32-39 -> CPython 3.12
                ``Random.sample`` / ``_randbelow_with_getrandbits`` / MT19937).
``sac_numpy``   NumPy restatement of the update arithmetic with hand-derived
                gradients (reference: sac/models.py:30-33,73-87,115-149;
                sac/agent.py:195-300; torch ``Normal.log_prob``, ``softplus``,
                ``_single_tensor_adam``).
``torch_port``  torch-eager (autograd) restatement used as the CPU baseline
                and pinned bit-for-bit against the real reference.

Parity pin: the reference ships no tests or golden vectors ("parity unpinned" by
its own tests).  The pin used instead is the reference itself,
executed in the build container from /root/reference by
``tests/golden/make_golden.py``; its recorded inputs/outputs are committed under
``tests/golden/*.npz`` and every oracle here is checked against them.
"""
