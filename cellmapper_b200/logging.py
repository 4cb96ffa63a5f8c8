"""Package logger (same breadcrumbs as the reference's ``src/cellmapper/logging.py``; level via LOGLEVEL)."""

import logging
import os

logger = logging.getLogger("cellmapper_b200")
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(levelname)s %(name)s: %(message)s"))
    logger.addHandler(_h)
logger.setLevel(os.environ.get("LOGLEVEL", "WARNING").upper())
logger.propagate = False
