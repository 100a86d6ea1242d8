from .fabric import Fabric  # noqa: F401
