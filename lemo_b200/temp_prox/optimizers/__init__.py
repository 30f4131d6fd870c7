from .optim_factory import create_optimizer  # noqa: F401
