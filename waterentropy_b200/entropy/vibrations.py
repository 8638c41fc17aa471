"""Vibrational (translational + rotational) entropy from covariance matrices.

Consumer of the hot path's covariance bins; API-compatible with the
reference's `entropy/vibrations.py` (`Vibrations`, `print_Svib_data`).  As
upstream, `add_data` converts the stored covariance matrices to SI units IN
PLACE (`entropy/vibrations.py:84`), so the `CovarianceCollection` a recipe
returns is already unit-scaled, and with `diagonalise=True` only the matrix
diagonals enter the entropy (`:179-180`): S = R * sum_i [a_i/(e^a_i - 1) -
ln(1 - e^-a_i)], a_i = hbar * sqrt(lambda_i / kT) / kT.
"""
import textwrap

import numpy as np
from numpy import linalg as LA

from waterentropy_b200.utils.helpers import nested_dict


class Vibrations:
    BOLTZMANN = 1.38064852e-23  # J / K
    HBAR = 1.0545718e-34  # J s
    AVOGADRO = 6.022e23
    GAS_CONSTANT = 8.314  # J / (K mol)
    CALORIE_TO_JOULE = 4184
    KG_TO_AMU = 6.02214086e26
    PER_MOLE_TO_PER_MOLECULE = 6.02214086e23
    ANGSTROM_TO_METRE = 1e-10
    KJ_TO_J = 1000
    NANOMETRE_TO_METRE = 1e-9

    def __init__(self, temperature):
        self.temperature = temperature
        self.translational_freq = nested_dict()
        self.rotational_freq = nested_dict()
        self.translational_S = nested_dict()
        self.rotational_S = nested_dict()

    # unit factors for squared forces: (energy/length)^2 per amu -> SI
    def _conversion(self, energy, length):
        return ((energy**2) * self.KG_TO_AMU) / ((self.AVOGADRO**2) * (length**2))

    def kJ_conversion(self):
        return self._conversion(self.KJ_TO_J, self.NANOMETRE_TO_METRE)

    def mda_conversion(self):
        return self._conversion(self.KJ_TO_J, self.ANGSTROM_TO_METRE)

    def kcal_conversion(self):
        return self._conversion(self.CALORIE_TO_JOULE, self.ANGSTROM_TO_METRE)

    def add_data(self, covariances, diagonalise=True):
        if diagonalise and self._add_stacked(covariances):
            return
        for key, force_cov in covariances.forces.items():
            torque_cov = covariances.torques[key]
            s_f, lam_f = self.get_data(force_cov, diagonalise)
            s_t, lam_t = self.get_data(torque_cov, diagonalise)
            for store, value in ((self.translational_S, s_f), (self.rotational_S, s_t),
                                 (self.translational_freq, lam_f), (self.rotational_freq, lam_t)):
                if key not in store:
                    store[key] = value

    def _add_stacked(self, covariances):
        """All bins at once when the collection's matrices are views of one array (collections decoded from the
        device tables, `engine.decode_cov_tables`): the same element-wise arithmetic as the per-key path below -
        in-place SI scaling of every matrix included - without ~20 us of Python per matrix."""
        stacked = getattr(covariances, "_stacked", None)
        if stacked is None:
            return False
        keys, M = stacked
        if len(keys) != len(covariances.forces) or len(keys) != len(covariances.torques):
            return False
        for k, key in enumerate(keys):  # still the views we made? (a caller may have replaced or merged entries)
            f, t = covariances.forces.get(key), covariances.torques.get(key)
            if f is None or t is None or f.base is not M or t.base is not M or f.shape != (3, 3) or t.shape != (3, 3):
                return False
        if len(keys) == 0:
            return True
        M *= self.mda_conversion()  # in place, like upstream: every view is scaled
        eig = M[:, :, (0, 1, 2), (0, 1, 2)]  # [bin, force / torque, 3] diagonals (a copy)
        lam = np.where(eig >= 0, eig, 0)
        kT = self.BOLTZMANN * self.temperature
        a = (self.HBAR * (lam / kT) ** 0.5) / (self.BOLTZMANN * self.temperature)
        with np.errstate(divide="ignore", invalid="ignore"):
            b = np.where(a != 0, np.exp(a) - 1, 0.0)
            c = np.where(a != 0, np.log(1 - np.exp(-a)), 0.0)
            s = np.where((b != 0) & (c != 0), a / b - c, 0) * self.GAS_CONSTANT
        for k, key in enumerate(keys):
            if key not in self.translational_S:
                self.translational_S[key] = s[k, 0]
            if key not in self.rotational_S:
                self.rotational_S[key] = s[k, 1]
            if key not in self.translational_freq:
                self.translational_freq[key] = eig[k, 0]
            if key not in self.rotational_freq:
                self.rotational_freq[key] = eig[k, 1]
        return True

    def get_data(self, covariance, diagonalise):
        covariance *= self.mda_conversion()  # in place, like upstream
        return self.calculate_entropies(covariance, diagonalise)

    def calculate_frequencies(self, covariance, diagonalise):
        if diagonalise:
            eigenvalues = covariance.diagonal()
        else:
            eigenvalues, _ = LA.eig(covariance)
        lam = np.where(np.isreal(eigenvalues), eigenvalues.real, 0)
        lam = np.where(lam >= 0, lam, 0)
        kT = self.BOLTZMANN * self.temperature
        return (lam / kT) ** 0.5, eigenvalues

    def calculate_entropies(self, covariance, diagonalise):
        frequency, eigenvalues = self.calculate_frequencies(covariance, diagonalise)
        a = (self.HBAR * frequency) / (self.BOLTZMANN * self.temperature)
        with np.errstate(divide="ignore", invalid="ignore"):
            b = np.where(a != 0, np.exp(a) - 1, 0.0)
            c = np.where(a != 0, np.log(1 - np.exp(-a)), 0.0)
            s = np.where((b != 0) & (c != 0), a / b - c, 0)
        return s * self.GAS_CONSTANT, eigenvalues


def print_Svib_data(vibrations: Vibrations, covariances):
    """Print the table in the reference's layout (`entropy/vibrations.py:195-235`)."""
    print(textwrap.dedent("""
    Vibrations
    ==========
    resid: residue ID of the closest solute to the waters
    resname: name of the residue associated with the residue ID
    solvent_name: name of the solvent near the solute
    Strans: sum of 3 translational terms to give total translational entropy
    Srot: sum of 3 rotational terms to give total rotational entropy
    Svib: vibrational entropy (Strans+Srot) of water molecules
    counts: total number of waters around the nearest solute across frames analysed
    """))
    print("resid resname solvent_name Strans Srot Svib counts")
    for key in vibrations.translational_S:
        nearest, solvent = key
        resname, resid = nearest.split("_") if "_" in nearest else (nearest, "N/A")
        s_t = sum(vibrations.translational_S[key])
        s_r = sum(vibrations.rotational_S[key])
        print(f"{resid} {resname} {solvent} {s_t:.4f} {s_r:.4f} {s_t + s_r:.4f} {covariances.counts[key]}")
    print()
