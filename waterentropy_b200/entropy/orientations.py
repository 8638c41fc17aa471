"""Orientational entropy of water from labelled-shell HB statistics.

Consumer of the hot path's count tables; API-compatible with the reference's
`entropy/orientations.py` (`Orientations`, `WaterOrientCalculator`,
`print_Sorient_dicts`) so the recipes can return the same `Sorient_dict`:

    Sorient_dict[resid][resname] = [Sor, count, N_c, N_w, Nc_eff, pbias, Sor_Nc, Sor_Nw]

Physics (reference `entropy/orientations.py:101-217`): for a shell type with
neighbour-label multiplicities N_i, the probabilities of donating to / accepting
from label i among all donations / acceptances are pD_i, pA_i; labels that are
both donated to and accepted from contribute ppD_i*ppA_i (ppD = pD/(pD+pA)) to
the bias, and  S = R ln( Nc_eff^(3/2) * sqrt(pi) * pbias_ave / 2 ),  clipped at 0.
The weighted averages over shell types reproduce the reference's update order
exactly (including its use of the already-updated total count for every column
after the first, `entropy/orientations.py:337-360`), so the table is a pure,
bit-reproducible function of the integer counts.
"""
import textwrap
from collections import Counter

import numpy as np

from waterentropy_b200.utils.helpers import nested_dict

GAS_CONSTANT = 8.314  # J / (K mol)


def _entropy_from_neighbours(n_eff, p_corr):
    """R * max(0, ln(n^(3/2) * sqrt(pi) * p / sigma)) with sigma = 2."""
    s = 0
    if n_eff != 0:
        s = np.log(n_eff ** (3 / 2) * np.pi**0.5 * p_corr / 2)
    return max(s, 0) * GAS_CONSTANT


class WaterOrientCalculator:
    """Entropy of one labelled shell type (reference `:14-238`)."""

    GAS_CONSTANT = GAS_CONSTANT

    def __init__(self):
        self.Sorient = 0
        self.Nc_eff = 0
        self.pbias_ave = 0

    @staticmethod
    def _probabilities(values, field):
        """{label: probability} over the counts stored under `field`."""
        if field not in values:
            return {}
        counts = values[field]
        total = 0
        for c in counts.values():
            total += c
        return {lab: c / total for lab, c in sorted(counts.items())}

    def add_data(self, shell_label, shell_values):
        multiplicity = Counter(shell_label)
        p_don = self._probabilities(shell_values, "donates_to")
        p_acc = self._probabilities(shell_values, "accepts_from")
        n_eff = 0
        biases = []
        for label, n_i in multiplicity.items():
            pd, pa = p_don.get(label, 0), p_acc.get(label, 0)
            if pd != 0 and pa != 0:
                both = pd + pa
                ppd, ppa = pd / both, pa / both
                n_eff += ppd * ppa / 0.25 * n_i
                biases.extend([ppd * ppa] * n_i)
            else:
                biases.extend([0] * n_i)
        bias = sum(biases) / len(biases) if biases else 0
        self.Nc_eff = n_eff
        self.pbias_ave = bias
        self.Sorient = _entropy_from_neighbours(n_eff, bias)

    def get_non_HB_data(self, Nc):
        """Entropy without HB biasing: p_corr = 0.25."""
        return _entropy_from_neighbours(Nc, 0.25)


def _fold(value, weight, mean, weight_seen):
    return (value * weight + mean * weight_seen) / (weight_seen + weight)


class Orientations:
    """`resid_labelled_Sorient[resid][resname]` and `resname_labelled_Sorient[resname]`."""

    def __init__(self):
        self.resid_labelled_Sorient = nested_dict()
        self.resname_labelled_Sorient = nested_dict()

    def add_data(self, hb_labels):
        self.get_orientational_entropy_from_dict(
            hb_labels.labelled_shell_counts, self.resname_labelled_Sorient)
        for resid, by_name in sorted(list(hb_labels.resid_labelled_shell_counts.items())):
            self.get_orientational_entropy_from_dict(by_name, self.resid_labelled_Sorient[resid])

    def get_orientational_entropy_from_dict(self, labelled_dict, Sorient_dict):
        for resname, shells in sorted(list(labelled_dict.items())):
            s_hb = n_c = n_w = n_eff = bias = s_nc = s_nw = 0
            total = 0
            for labels, values in sorted(list(shells.items())):
                w = values["shell_count"]
                calc = WaterOrientCalculator()
                calc.add_data(labels, values)
                s_hb = _fold(calc.Sorient, w, s_hb, total)
                total = total + w
                # upstream folds the remaining columns against the updated total
                s_nc = _fold(calc.get_non_HB_data(len(labels)), w, s_nc, total)
                s_nw = _fold(calc.get_non_HB_data(values["N_w"]), w, s_nw, total)
                n_eff = _fold(calc.Nc_eff, w, n_eff, total)
                bias = _fold(calc.pbias_ave, w, bias, total)
                n_c = _fold(len(labels), w, n_c, total)
                n_w = _fold(values["N_w"], w, n_w, total)
            Sorient_dict[resname] = [s_hb, total, n_c, n_w, n_eff, bias, s_nc, s_nw]


def print_Sorient_dicts(Sorient_dict: dict):
    """Print the table in the reference's layout (`entropy/orientations.py:373-407`)."""
    print(textwrap.dedent("""
    Orientations
    ============
    resid: residue ID of the closest solute to the waters
    resname: name of the residue associated with the residue ID
    Sor: Orientational entropy of water around residue using HB biasing
    count: total number of waters around the residue across frames analysed
    N_c: average number of UAs around waters analysed
    N_w: average number of water UAs around waters analysed
    Nc_eff: average number of effective neighbouts that waters can HB with
    pbias: average of the probability of forming HBs with neighbouring UAs
    Sor_Nc: Orientational entropy of water around residue using Nc (no HB bias)
    Sor_Nw: Orientational entropy of water around residue using Nw (no HB bias)
    """))
    print("resid resname Sor count N_c N_w Nc_eff pbias Sor_Nc Sor_Nw")
    for resid, by_name in sorted(list(Sorient_dict.items())):
        for resname, row in sorted(list(by_name.items())):
            sor, count, n_c, n_w, n_eff, bias, s_nc, s_nw = row
            print(f"{resid} {resname} {sor:.4f} {count} {n_c:.4f} {n_w:.4f} "
                  f"{n_eff:.4f} {bias:.4f} {s_nc:.4f} {s_nw:.4f} ")
    print()
