"""Bulk-water recipe on the B200 path.

Drop-in for `waterEntropy/recipes/bulk_water.py:16-90`: every water whose RAD
shell contains only waters contributes its labelled shell, HB labels and
force/torque covariances under the key `("WAT", "WAT")`.
"""
from __future__ import annotations

from waterentropy_b200.entropy.orientations import Orientations
from waterentropy_b200.entropy.vibrations import Vibrations
from waterentropy_b200.recipes.interfacial_solvent import _release, _run


def get_bulk_water_orient_entropy(system, start: int, end: int, step: int, temperature=298):
    """Returns `(Sorient_dict, covariances, vibrations)`."""
    eng = _run(system, list(range(start, end, step)), "bulk", keep_records=False)
    covariances = eng.covariances()
    Sorients = Orientations()
    Sorients.add_device(eng, bulk=True)  # reduced on the device (we_orient_entropy)
    _release(eng)
    vibrations = Vibrations(temperature)
    vibrations.add_data(covariances, diagonalise=True)
    return Sorients.resid_labelled_Sorient, covariances, vibrations
