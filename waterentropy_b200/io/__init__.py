from .universe import Universe  # noqa: F401
