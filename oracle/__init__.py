"""CPU oracle for the MultiNeRF per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch-CPU / numpy) restatement of the reference algorithm
(`/root/reference/internal/{stepfun,render,coord,math,ref_utils,geopoly,image,
models,train_utils}.py`).  It exists to CHECK the CUDA path; nothing in
`multinerf_b200/` may import it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs use it.

Pinning status (see DESIGN.md "Oracle"):
  * L2 functions (stepfun / render / coord / math / ref_utils / geopoly / image / camera_utils)
    are pinned against (a) the reference's own known-answer tests, re-stated in
    `tests/test_oracle_*.py`, and (b) golden vectors produced by executing the
    REAL reference source files in this container under a numpy stand-in for
    `jax.numpy` (`tests/golden/make_golden.py`, fixtures in `tests/golden/*.npz`).
  * `Model.__call__` / `MLP.__call__` (all four BASELINE model families: mip-NeRF 360, blender,
    RawNeRF, Ref-NeRF; plus GLO vectors, random backgrounds, bottleneck noise and near-plane annealing) and the loss functions / clip_gradients of the train-step closure are
    pinned the same way: `tests/golden/make_golden_model.py` runs the reference's REAL
    `internal/models.py` and `internal/train_utils.py` under jax/flax/gin stand-ins and
    `tests/test_oracle_model_golden.py` compares `oracle.o_models.model_apply` and
    `oracle.o_train.*` against those outputs (fp32, 2e-5 / 2e-4).
  * Still "parity unpinned": XLA's own float rounding, threefry random streams (randomness is an
    explicit input here), flax initialisers and optax.adam (restated from their published
    definitions).
"""
