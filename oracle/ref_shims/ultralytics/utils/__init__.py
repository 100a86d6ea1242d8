import logging

LOGGER = logging.getLogger("ultralytics-shim")
from . import checks  # noqa: E402,F401
