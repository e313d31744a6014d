"""astropy.cosmology stand-in for make_golden.py: the repo's own dependency-free background
(pysco_b200/cosmology.py), so that the reference and the build share the same E(a)."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(__file__), "..", "..", "..", "..", "pysco_b200", "cosmology.py")
_spec = importlib.util.spec_from_file_location("_psc_cosmology", os.path.abspath(_p))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
Flatw0waCDM = _m.Flatw0waCDM
