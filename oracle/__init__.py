"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement of the reference hot path (sheoyon-jhin/ANCDE: natural cubic
spline coefficients + the dual-NCDE integration loop) used as the parity
checker for the CUDA kernels in ``ancde_b200/csrc``.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``ancde_b200/`` or ``controldiffeq/`` imports it; those fail loudly when the
CUDA library is missing.

Contents
--------
``port.py``             torch-CPU restatement (fp32/fp64) of the algorithm, each
                        function citing the reference file:line it follows.
                        Travels to the GPU box (``/root/reference`` does not).
``torchdiffeq_shim.py`` restatement of the un-vendored third-party dependency
                        ``torchdiffeq==0.0.1`` (ancde.yml:245): fixed-grid
                        solvers (euler / midpoint / rk4 = 3/8 rule), dopri5 and
                        the adjoint.  Written from the published algorithm; the
                        library itself is absent here and the reference holds no
                        test that pins results at this boundary, so this part is
                        "parity unpinned" (see header of that file).
``reference_loader.py`` imports the UNMODIFIED reference python packages from
                        ``/root/reference`` on top of the shim (only where that
                        directory exists: the build container).  Used to pin
                        ``port.py`` and to generate ``tests/golden/*.npz``.
``gen_golden.py``       the script that made the committed golden vectors.

Pinning status
--------------
* spline coefficients: PINNED against the reference's own
  cached fixtures (``processed_data/Stock1/Stock{30,50,70}``) and against the
  imported reference code.
* dual-NCDE loop (a6-a13): PINNED against the imported reference
  ``controldiffeq`` + ``experiments/models`` code run in the build container.
* RK4 / adjoint arithmetic itself (a14, torchdiffeq): PARITY UNPINNED -- the
  shim is a restatement from the published source, validated only by
  self-checks (order-4 convergence, scipy cross-check).
"""
