"""astropy.constants values read by the reference (utils.py:11, solver.py:14, cosmotable.py:9):
CODATA-2018 / IAU-2015 defaults of astropy >= 4."""


class _Const:
    def __init__(self, value):
        self.value = value


c = _Const(299792458.0)
pc = _Const(3.0856775814913673e16)
G = _Const(6.6743e-11)
