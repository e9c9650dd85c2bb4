from . import train_state, checkpoints  # noqa: F401
