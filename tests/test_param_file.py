"""iostream.read_param_file against the reference's reader (tests/golden/param_test.json = the unmodified
reference's output for tests/golden/param_test.ini, produced in the build container).

The reference's reader leaves trailing blanks on values that are followed by a comment with the pandas version of
this image ('newton ', 'false '), which also keeps such a 'false ' from becoming a bool; those are artefacts, not
the file format, so the comparison is made after stripping (and case-folding booleans), which is what
pysco_b200.iostream.read_param_file does."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _normalise(tname, value):
    if tname == "str":
        v = value.strip()
        if v.casefold() in ("true", "false"):
            return v.casefold() == "true"
        return v
    return value


def test_read_param_file_matches_reference_reader():
    from pysco_b200 import iostream
    ref = json.load(open(os.path.join(GOLD, "param_test.json")))
    mine = iostream.read_param_file(os.path.join(GOLD, "param_test.ini"))
    assert list(mine.index) == list(ref.keys())
    for key, (tname, value) in ref.items():
        want = _normalise(tname, value)
        got = mine[key]
        if isinstance(want, bool):
            assert isinstance(got, (bool, np.bool_)) and bool(got) == want, key
        elif isinstance(want, (int, float)):
            assert type(got).__name__.startswith(type(want).__name__) and got == want, (key, got, want)
        else:
            assert got == want, (key, got, want)
    assert mine["npart"] == 32 ** 3 and isinstance(mine["z_out"], str)
    assert iostream.parse_z_out(mine) == [10, 5, 2, 1, 0.5, 0]


def test_run_rejects_bad_arguments_like_the_reference():
    """main.py:45-46, 67-68: ValueError for a bad `verbose` and for a param that is neither dict nor Series"""
    import pytest
    import pysco_b200
    with pytest.raises(ValueError):
        pysco_b200.run({"verbose": 7})

    class NotADict:                      # indexable like the reference expects, but neither dict nor Series
        def __getitem__(self, key):
            return 0

    with pytest.raises(ValueError):
        pysco_b200.run(NotADict())
