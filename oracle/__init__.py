"""
oracle/ -- CPU restatement of the MDSuite (SamTov/LAMMPS-Analysis) hot paths.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import, call or execute anything in this package,
and there only as the *checker* (or as the timed CPU baseline), never as part of
the shipped GPU path.  ``lammps_analysis_b200`` never imports ``oracle``.

Every function cites the reference file:line (relative to the upstream repo
root) whose arithmetic it restates.  The reference is 100 % Python on
TensorFlow / TensorFlow-Probability, neither of which is installable here, so
the oracle restates the TF/tfp op semantics in NumPy:

* ``tf.histogram_fixed_width`` (tensorflow/core/kernels/histogram_op.cc, CPU
  functor; tensorflow is an unpinned ``requirements.txt`` dependency
  ``tensorflow>=2.5``): ``int32(min(double(max(v,lo)-lo)/step, nbins-1))`` with
  ``step = double(hi-lo)/double(nbins)``.
* ``tfp.stats.auto_correlation`` (tensorflow_probability/python/stats/
  sample_stats.py, unpinned): zero-pad to ``2**ceil(log2(2N))``, complex128 FFT,
  ``|X|^2``, inverse FFT, keep lags ``0..N-1``, divide lag m by ``N-m``.
* ``tf.math.rint`` / ``tf.math.round``: round half to even.

Parity pinning status:
  - unwrap-with-carry, unwrap-via-indices, ionic current, memory-manager planner
    arithmetic and ``fit_einstein_curve`` are pinned against the reference's own
    known-answer unit tests (``tests/test_oracle_golden.py``).
  - RDF index / mask / minibatch logic and bin counts, the transformations, the
    whole MemoryManager, the DataManager window and batch generators, the Einstein
    ensemble operation, ``fit_einstein_curve`` and ``golden_section_search`` are
    pinned against vectors produced by EXECUTING the reference's own Python source
    under a NumPy TensorFlow shim (``tests/golden/make_reference_goldens.py`` ->
    ``tests/golden/reference_run.json``).
  - TensorFlow's / tfp's own numerics (histogram_fixed_width truncation rule,
    norm summation order, FFT autocorrelation) cannot be executed here (no
    tensorflow, no network for the reference's downloaded golden JSONs): for those
    the oracle is a restatement of the published algorithms => "parity unpinned"
    at that level, beyond the analytic-model tests (random walk D, Langevin VACF)
    the reference itself uses.
"""
