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
# does this interpreter's sum() compensate (CPython >= 3.12)?  [1.0, 1e100, 1.0, -1e100] sums to 2.0 if so
_SUM_IS_COMPENSATED = sum([1.0, 1e100, 1.0, -1e100]) == 2.0


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

    def add_entries(self, top, bulk, entries):
        """Fill both tables straight from the device label entries (no HBLabelCollection in between)."""
        self.resid_labelled_Sorient, self.resname_labelled_Sorient = sorient_from_entries(top, bulk, entries)

    def add_device(self, engine, bulk=None):
        """Fill both tables on the device (`we_orient_entropy`, csrc/we_orient.cuh): the label table is sorted,
        merged and folded there in the reference's order and only one row per residue comes back.  Agrees with
        `add_entries` / `add_data` to rounding (device `log` / `pow`; tested at rtol 1e-9)."""
        self.resid_labelled_Sorient, self.resname_labelled_Sorient = engine.orient_entropy(bulk)

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


# ---------------------------------------------------------------------------------------------
# The same tables straight from the device label entries, with array operations.
#
# `Orientations().add_data(decode_label_entries(...))` rebuilds the reference's nested dictionaries entry
# by entry (a Python loop over tens of thousands of label keys) only to reduce them to one row per
# residue.  The functions below compute those rows from the structured entry array itself.  Every
# floating-point operation, its order and the fold order (Python's ordering of the label tuples) are
# those of the classes above, so the rows are bit-identical (tests/test_host_logic.py).
def _label_ranks(top, codes):
    """rank of every label code in Python's string order of the label strings"""
    ucodes = np.unique(codes)
    strings = [top.label_string(int(c)) for c in ucodes]
    order = sorted(range(len(strings)), key=strings.__getitem__)
    lut = np.full(65536, -1, dtype=np.int64)
    lut[ucodes[order]] = np.arange(len(order))
    return lut


def _entry_rows(top, e):
    """Per label entry (already merged): the WaterOrientCalculator numbers + the sortable label key."""
    E = len(e)
    n = e["n_labels"].astype(np.int64)
    codes = e["labels"][:, :25].astype(np.int64)
    pos = np.arange(25)[None, :]
    valid = pos < n[:, None]
    lut = _label_ranks(top, codes[valid]) if valid.any() else np.full(65536, -1, dtype=np.int64)
    rank = np.where(valid, lut[codes], 1 << 40)
    first = valid & np.concatenate([np.ones((E, 1), dtype=bool), codes[:, 1:] != codes[:, :-1]], axis=1)
    # multiplicity of every member's label: the codes of an entry are sorted, equal labels are one run
    run = np.cumsum(first, axis=1) - 1 + np.arange(E)[:, None] * 25
    mult = np.bincount(run[valid], minlength=E * 25)[run]
    don = np.where(first, e["donates"].astype(np.int64), 0)
    acc = np.where(first, e["accepts"].astype(np.int64), 0)
    perm = np.argsort(rank, axis=1, kind="stable")  # members in Python's label order; equal labels stay adjacent
    take = lambda a: np.take_along_axis(a, perm, axis=1)
    rank_s, first_s, mult_s, don_s, acc_s, valid_s = take(rank), take(first), take(mult), take(don), take(acc), take(valid)
    for j in range(1, 25):  # every member sees its label's counts
        don_s[:, j] = np.where(first_s[:, j], don_s[:, j], don_s[:, j - 1])
        acc_s[:, j] = np.where(first_s[:, j], acc_s[:, j], acc_s[:, j - 1])
    tot_don = don.sum(axis=1)
    tot_acc = acc.sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        pd = np.where(tot_don[:, None] > 0, don_s / tot_don[:, None], 0.0)
        pa = np.where(tot_acc[:, None] > 0, acc_s / tot_acc[:, None], 0.0)
        live = valid_s & (pd != 0) & (pa != 0)
        both = pd + pa
        prod = np.where(live, (pd / both) * (pa / both), 0.0)
    n_eff = np.cumsum(np.where(first_s & live, prod / 0.25 * mult_s, 0.0), axis=1)[:, -1]
    # sum() over the per-member list: CPython (>= 3.12) adds floats with Neumaier's compensated summation
    f_sum, comp = np.zeros(E), np.zeros(E)
    for j in range(25):
        x = prod[:, j]
        t = f_sum + x
        comp += np.where(np.abs(f_sum) >= np.abs(x), (f_sum - t) + x, (x - t) + f_sum)
        f_sum = t
    bias_sum = np.where((comp != 0) & np.isfinite(comp), f_sum + comp, f_sum) if _SUM_IS_COMPENSATED else \
        np.cumsum(prod, axis=1)[:, -1]
    bias = np.where(n > 0, bias_sum / np.maximum(n, 1), 0.0)
    s_hb = np.array([_entropy_from_neighbours(a, b) for a, b in zip(n_eff.tolist(), bias.tolist())], dtype=np.float64)
    plain = np.array([_entropy_from_neighbours(k, 0.25) for k in range(0, 64)], dtype=np.float64)
    n_w = e["n_w"].astype(np.int64)
    key = np.where(valid_s, rank_s, -1)  # Python compares the tuples element-wise; a prefix is smaller
    return dict(s_hb=s_hb, n_eff=n_eff, bias=bias, s_nc=plain[n], s_nw=plain[n_w], n=n, n_w=n_w,
                w=e["shell_count"].astype(np.int64), key=key)


def _merge_entries(e, group_cols, key):
    """Entries with the same (group, label key) become one: counts add, N_w of the first insert stays."""
    E = len(e)
    seen = np.argsort(np.argsort(~e["first_seen_inv"], kind="stable"), kind="stable")
    # label tuples compare like Python tuples when the ranks (+1, 0 = past the end) are read as one
    # big-endian byte string: a single sort key instead of 25
    packed = np.ascontiguousarray((key + 1).astype(">u2")).view("S50").reshape(E)
    order = np.lexsort([seen, packed] + list(reversed(group_cols)))
    cols = [c[order] for c in group_cols]
    k = key[order]
    new = np.ones(E, dtype=bool)
    if E > 1:
        same = np.all(k[1:] == k[:-1], axis=1)
        for c in cols:
            same &= c[1:] == c[:-1]
        new[1:] = ~same
    starts = np.nonzero(new)[0]
    out = e[order][starts].copy()
    if len(starts) != E:
        src = e[order]
        out["shell_count"] = np.add.reduceat(src["shell_count"], starts)
        out["donates"] = np.add.reduceat(src["donates"], starts, axis=0)
        out["accepts"] = np.add.reduceat(src["accepts"], starts, axis=0)
    return out, [c[starts] for c in cols]


def _fold_groups(rows, group_cols):
    """The running averages of `get_orientational_entropy_from_dict`, all groups at once: groups are
    contiguous and ordered by label key; step k folds the k-th entry of every group that has one."""
    E = len(rows["w"])
    new = np.zeros(E, dtype=bool)
    new[:1] = True
    for c in group_cols:
        new[1:] |= c[1:] != c[:-1]
    starts = np.nonzero(new)[0]
    lens = np.diff(np.append(starts, E))
    G = len(starts)
    total = np.zeros(G, dtype=np.int64)
    st = {k: np.zeros(G, dtype=np.float64) for k in ("s_hb", "s_nc", "s_nw", "n_eff", "bias", "n_c", "n_w")}
    for k in range(int(lens.max()) if G else 0):
        g = np.nonzero(lens > k)[0]
        i = starts[g] + k
        w = rows["w"][i]
        t0 = total[g]
        st["s_hb"][g] = (rows["s_hb"][i] * w + st["s_hb"][g] * t0) / (t0 + w)
        t1 = t0 + w
        total[g] = t1
        for name, val in (("s_nc", rows["s_nc"][i]), ("s_nw", rows["s_nw"][i]), ("n_eff", rows["n_eff"][i]),
                          ("bias", rows["bias"][i]), ("n_c", rows["n"][i]), ("n_w", rows["n_w"][i])):
            st[name][g] = (val * w + st[name][g] * t1) / (t1 + w)  # upstream folds these against the updated total
    table = np.stack([st["s_hb"], total.astype(np.float64), st["n_c"], st["n_w"], st["n_eff"], st["bias"],
                      st["s_nc"], st["s_nw"]], axis=1)
    return starts, total, table


def sorient_from_entries(top, bulk, entries):
    """(resid_labelled_Sorient, resname_labelled_Sorient) from device label entries - what
    `Orientations().add_data(decode_label_entries(top, recipe, entries))` leaves in the two attributes."""
    by_resid, by_resname = nested_dict(), nested_dict()
    if len(entries) == 0:
        return by_resid, by_resname
    e = entries[np.argsort(~entries["first_seen_inv"], kind="stable")]
    n = e["n_labels"].astype(np.int64)
    codes = e["labels"][:, :25].astype(np.int64)
    valid = np.arange(25)[None, :] < n[:, None]
    lut = _label_ranks(top, codes[valid]) if valid.any() else np.full(65536, -1, dtype=np.int64)
    key = np.sort(np.where(valid, lut[codes], 1 << 40), axis=1)
    key = np.where(key == (1 << 40), -1, key)
    names = np.array(top.resname_table, dtype=object)
    name_rank = np.argsort(np.argsort(np.array([str(x) for x in names]), kind="stable"), kind="stable")
    rname = name_rank[e["nearest_resname_id"].astype(np.int64)]
    resid = e["nearest_resid"].astype(np.int64)
    for table, cols in ((by_resname, [rname]), (by_resid, [rname] if bulk else [resid, rname])):
        merged, gcols = _merge_entries(e, cols, key)
        rows = _entry_rows(top, merged)
        starts, total, tab = _fold_groups(rows, gcols)
        for g, s0 in enumerate(starts):
            name = top.resname_table[int(merged["nearest_resname_id"][s0])]
            row = tab[g].tolist()
            row[1] = int(total[g])
            if table is by_resname:
                table[name] = row
            else:
                table[name if bulk else int(merged["nearest_resid"][s0])][name] = row
    return by_resid, by_resname


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
