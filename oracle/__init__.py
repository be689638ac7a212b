"""CPU oracle for the HECTOR-style force-and-moment MPC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``biped_mpc_py_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU baseline - never as the product path.

What it restates (all citations are into /root/reference/bipedalLocomotionMPC.py):

* ``reference_mpc.py``  - the QP assembly (MPC.py:50-286), the torque map
  (MPC.py:306-365, 426-470) and the forward kinematics (MPC.py:367-424), float64,
  bug-for-bug (see the quirk list in SURVEY.md section 8a).
* ``qp_exact.py``       - the solve.  The reference calls ``cvxopt.solvers.qp``
  (MPC.py:289-297), a third-party dense interior-point code that is *unpinned*
  (no requirements file) and *not installable here* (no wheel, no network).  The
  QP is strictly convex (H is diagonal with minimum entry 2e-4, MPC.py:278-281),
  so its optimum is unique and solver independent: the oracle returns that
  optimum (interior point -> active-set polish -> KKT certificate).

* ``highs_check.py``    - an independent second opinion on the solve: scipy's bundled HiGHS QP solver on the same
  dense QP (objective agreement to 1e-8, never better than the certified optimum).
* ``rollout.py``        - the closed-loop rules R1-R7 on the CPU (checker of ``bmpc_rollout``).

Parity pinning: the reference has no tests or golden vectors.  The oracle is
pinned two ways: (1) ``oracle/gen_golden.py`` imports the *real* reference module
in the build container (with a stub ``cvxopt`` that captures the matrices the
reference builds) and commits the captured QP data and torque outputs under
``tests/golden/``; (2) the KKT-certified known answers G1-G5 of SURVEY.md 8c.
The cvxopt arithmetic itself cannot be executed, so the *solver* half of parity is
anchored on the certificate (any certified point is the reference QP's unique
optimum), not on cvxopt output.
"""
