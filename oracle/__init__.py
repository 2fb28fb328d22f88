"""CPU oracle for the generate -> solve -> label path and the classifier forward of rodrgo/deep_dantzig.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deep_dantzig_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the CPU arm that is timed beside the GPU path -- never as the thing shipped.

What is restated (all citations relative to the reference tree):
  * seed schedule + instance generator  -- ``src/data/randomlp_dataset.py:31-43, 76-86`` (numpy legacy RandomState,
    bit-frozen stream), see :mod:`oracle.randomlp`.
  * solve                               -- the reference delegates to Gurobi (``from gurobipy import *``,
    ``src/data/gurobi_lp.py:1``; unpinned, the only version hint is a Gurobi 7.5 documentation URL at
    ``gurobi_lp.py:446``).  Gurobi is proprietary and absent here, so the oracle uses the solver the reference itself
    imports and documents as the equivalent form (``randomlp_dataset.py:4, 71-75``):
    ``scipy.optimize.linprog(c, A_ub=A, b_ub=b, bounds=(None, None), method='highs-ds')``.
  * status codes / active-set labelling -- ``src/data/gurobi_lp.py:435-465`` and ``randomlp_dataset.py:91-128``.
  * classifier forward (both graphs)    -- ``src/ml/models/s2v.py:91-187, 218-323``, see :mod:`oracle.classifier`.

Pinning status:
  * generator: pinned -- it *is* numpy's frozen legacy stream; known answers in ``tests/golden/randomlp_kat.json``.
  * solver:    **parity unpinned** -- the reference holds no golden vectors, tests or fixtures for this path
     and its own solver cannot run here.  The stand-in is cross-checked three ways instead
    (HiGHS dual simplex, HiGHS interior point + crossover, and an independent dense tableau simplex); because every
    instance has a unique optimal vertex any exact solver must produce the same active set.
  * classifier: pinned -- ``tests/golden/s2v_*.npz`` were produced by importing the unmodified reference
    ``ml.models.s2v.Model`` in the build container (``tests/golden/make_s2v_golden.py``).
"""
