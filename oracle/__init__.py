"""CPU oracle for the RRNCO construction-rollout hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain-torch (CPU, fp32, optional
fp64 twin) restatement of the reference's eager rollout: the three envs, the
RRNet decoder, the decoding strategies, the policy loop and the instance
sub-sampler.  Every function cites the reference file:line it follows
(paths relative to the upstream repo root).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it -- and there only as
the checker / the CPU arm.  Nothing under ``rrnco_b200/`` imports it; the
product path fails loudly when the CUDA library is missing.

Parity pinning: the reference is pure Python but needs rl4co / tensordict /
torchrl, none of which are installed.  ``oracle/shims`` holds minimal stand-ins
for exactly the rl4co/tensordict/torchrl symbols the reference imports (their
semantics recalled from rl4co 0.6.0); with those on
``sys.path`` the reference's OWN env / decoder / decoding / policy / sampler
files are executed unmodified from ``/root/reference`` by
``tests/golden/make_golden.py`` and their outputs committed as fixtures under
``tests/golden/``.  The oracle is checked against those fixtures in the
``-m "not gpu"`` suite.  So: pinned to the reference's own code for everything
under ``/root/reference``; the rl4co helpers underneath are recalled, not
vendored ("parity pinned modulo recalled rl4co helpers").
"""
