"""Labelled-shell count tables that feed the orientational entropy.

API-compatible with the reference's `HBLabelCollection`
(`analysis/HB_labels.py:15-296`): two nested dictionaries

    labelled_shell_counts[resname][labels] = entry
    resid_labelled_shell_counts[resid][resname][labels] = entry
    entry = {"shell_count": n, "N_w": N_w,
             "donates_to": {label: n}, "accepts_from": {label: n}}

with `labels` the alphabetically sorted tuple of shell labels.  On the B200
path the entries come from the device hash table (`add_counts`); `add_data`
and `merge` keep the reference semantics (N_w is frozen by the first insert).
"""
from waterentropy_b200.utils.helpers import nested_dict


class HBLabelCollection:
    def __init__(self):
        self.labelled_shell_counts = nested_dict()
        self.resid_labelled_shell_counts = nested_dict()

    def _entries(self, resid, resname, key):
        return (self.labelled_shell_counts[resname][key],
                self.resid_labelled_shell_counts[resid][resname][key])

    def add_counts(self, resid, resname, labels, N_w, shell_count, donates_to, accepts_from):
        """Add `shell_count` shells with aggregated HB label counts (dicts)."""
        key = tuple(sorted(labels))
        for entry in self._entries(resid, resname, key):
            if "shell_count" in entry:
                entry["shell_count"] += shell_count
            else:
                entry["shell_count"] = shell_count
                entry["N_w"] = N_w
            for field, src in (("donates_to", donates_to), ("accepts_from", accepts_from)):
                for lab, c in src.items():
                    if c:
                        entry[field][lab] = entry[field].get(lab, 0) + c

    def add_data(self, resid, resname, labelled_shell, donates_to, accepts_from, N_w):
        """One shell with its lists of donated-to / accepted-from labels."""
        don, acc = {}, {}
        for lab in donates_to:
            don[lab] = don.get(lab, 0) + 1
        for lab in accepts_from:
            acc[lab] = acc.get(lab, 0) + 1
        self.add_counts(resid, resname, labelled_shell, N_w, 1, don, acc)

    @staticmethod
    def _fold(dst, src):
        for field, val in src.items():
            if field in ("donates_to", "accepts_from"):
                for lab, c in val.items():
                    dst[field][lab] = dst[field].get(lab, 0) + c
            elif field in dst and field != "N_w":
                dst[field] += val
            else:
                dst[field] = val

    def merge(self, other):
        for resid, by_name in other.resid_labelled_shell_counts.items():
            for resname, by_key in by_name.items():
                for key, src in by_key.items():
                    self._fold(self.resid_labelled_shell_counts[resid][resname][key], src)
        for resname, by_key in other.labelled_shell_counts.items():
            for key, src in by_key.items():
                self._fold(self.labelled_shell_counts[resname][key], src)
