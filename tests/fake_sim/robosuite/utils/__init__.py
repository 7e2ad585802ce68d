from . import transform_utils  # noqa: F401
