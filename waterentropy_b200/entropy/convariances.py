"""Force / torque covariance container handed to the vibrational-entropy code.

API-compatible with the reference's `CovarianceCollection`
(`entropy/convariances.py:10-108`, module name spelled as upstream): three
dictionaries keyed `(nearest, molecule_name)` holding the running-mean 3x3
force matrix, torque matrix and the sample count.  On the B200 path the means
are produced from per-bin sums accumulated on the device
(`WaterEntropyEngine.covariances()`); `add_data` / `merge` keep the reference
semantics for host-side use.
"""
import numpy as np

from waterentropy_b200.utils.helpers import nested_dict


class CovarianceCollection:
    def __init__(self):
        self.forces = nested_dict()
        self.torques = nested_dict()
        self.counts = nested_dict()

    def add_data(self, nearest, molecule_name: str, force: np.ndarray, torque: np.ndarray):
        """Fold one molecule's matrices into the running means (count last)."""
        key = (nearest, molecule_name)
        if key in self.forces:
            n = self.counts[key]
            self.forces[key] = (self.forces[key] * n + force) / (n + 1)
            self.torques[key] = (self.torques[key] * n + torque) / (n + 1)
            self.counts[key] = n + 1
        else:
            self.forces[key] = force
            self.torques[key] = torque
            self.counts[key] = 1

    def set_sums(self, key, force_sum, torque_sum, count: int):
        """Install a bin from device-side sums (mean = sum / count)."""
        self.forces[key] = np.asarray(force_sum, dtype=np.float64) / count
        self.torques[key] = np.asarray(torque_sum, dtype=np.float64) / count
        self.counts[key] = int(count)

    def merge(self, other):
        """Count-weighted merge of another collection into this one."""
        for key in other.forces:
            if key not in self.forces:
                self.forces[key] = other.forces[key]
                self.torques[key] = other.torques[key]
                self.counts[key] = other.counts[key]
                continue
            a, b = self.counts[key], other.counts[key]
            self.forces[key] = (self.forces[key] * a + other.forces[key] * b) / (a + b)
            self.torques[key] = (self.torques[key] * a + other.torques[key] * b) / (a + b)
            self.counts[key] = a + b
