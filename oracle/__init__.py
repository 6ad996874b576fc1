"""CPU oracle for the AudioPure purification hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in plain fp32 torch/numpy CPU arithmetic, the algorithm of the
reference (cychomatica/AudioPure) along the path that ``audiopure_b200``
accelerates.  Every function cites the reference ``file:line`` it follows
(paths relative to the reference checkout).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; ``audiopure_b200`` never does, and the
product path raises if its CUDA library is missing instead of falling back
to anything in here.

Parity pinning
--------------
The reference ships no tests, golden vectors or known-answer files
(SURVEY.md section 4), so the oracle is pinned against outputs of the
*unmodified reference code itself*, executed in the build container by
``oracle/make_golden.py`` (which imports ``/root/reference`` with the stub
recipe of SURVEY.md section 8c) and committed as small fixtures under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures.  Two pieces of third-party arithmetic are
absent from the container and are restated from their published algorithm
instead, which leaves them "parity unpinned against the original package":

* ``torchsde==0.2.5`` fixed-step Euler-Maruyama (``diffwave_sde.py:201-204``)
  -> ``oracle.purify.sde_purify``.  The drift/diffusion terms ARE pinned: the
  golden generator evaluates the reference's own ``RevVPSDE.f`` / ``.g``.
* ``statsmodels==0.13.2`` ``proportion_confint(method='beta')``
  (``certified_robust.py:113-117``) -> ``oracle.certify.lower_conf_bound``
  (Clopper-Pearson through ``scipy.stats.beta.ppf``).
"""

from . import schedule, weights, wavenet, purify, mel, certify  # noqa: F401
