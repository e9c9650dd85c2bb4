"""CPU oracle for the MultiNeRF per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch-CPU / numpy) restatement of the reference algorithm
(`/root/reference/internal/{stepfun,render,coord,math,ref_utils,geopoly,image,
models,train_utils}.py`).  It exists to CHECK the CUDA path; nothing in
`multinerf_b200/` may import it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs use it.

Pinning status (see DESIGN.md "Oracle"):
  * L2 functions (stepfun / render / coord / math / ref_utils / geopoly / image)
    are pinned against (a) the reference's own known-answer tests, re-stated in
    `tests/test_oracle_*.py`, and (b) golden vectors produced by executing the
    REAL reference source files in this container under a numpy stand-in for
    `jax.numpy` (`tests/golden/make_golden.py`, fixtures in `tests/golden/*.npz`).
  * `Model.__call__` / `MLP.__call__` / the train step have no test or golden
    vector in the reference and JAX/Flax cannot be installed here: those are
    pinned through the same jax->numpy stand-in run of the real `models.py`
    forward where the stand-in reaches (see make_golden.py), and are otherwise
    "parity unpinned" (optax.adam / flax initialisers are restated from their
    published definitions).
"""
