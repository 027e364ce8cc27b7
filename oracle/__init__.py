"""CPU oracle for the batched trajectory cost-and-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``motion_planning_baselines_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or as the timed CPU baseline -- never as the product path.

What is pinned and what is not
------------------------------
* PINNED (by golden fixtures generated from the UNMODIFIED reference planners
  imported from /root/reference, see ``oracle/make_golden.py`` and
  ``tests/golden/``): GP-prior precision and scale_tril construction, prior
  sampling, CostGP / CostGoalPrior / CostCollision glue / CostComposite, the
  importance-sampling term, the Stoch-GPMP / STOMP / MPPI softmax updates, the
  CHOMP gradient step and the GPMP2 linear system + solve.
* PARITY UNPINNED: forward kinematics (``fk_map_collision``), the sphere/box
  signed-distance primitives and the hinge (``field.compute_cost``).  In the
  reference they live in the third-party ``torch_robotics`` package
  (jacarvalho/torch_robotics, version unpinned, absent from /root/reference
  and from this image).  ``oracle/robots.py`` and ``oracle/fields.py`` ARE the
  specification at that boundary; they are duck-typed to the contract the
  reference calls (mp_baselines/planners/costs/cost_functions.py:50-52,
  mp_baselines/planners/costs/factors/field_factor.py:39,56).
"""
