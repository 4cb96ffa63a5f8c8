"""Development tools only: build the -DCM_DEV_PROBES variant of the library (it exports cm_debug_probe_flags /
cm_debug_probe_prof, which the shipping library does not) and make cellmapper_b200._lib load it.  Import this module
BEFORE cellmapper_b200._lib."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellmapper_b200 import build as _build

os.environ["CM_LIBPATH"] = _build.build(dev_probes=True)
