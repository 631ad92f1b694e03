"""TEST INFRASTRUCTURE — CPU restatement of the reference's data path (mlx_data/dataloader.py), NumPy only.
Only tests/ and oracle/ref_runner.py may import this.  Pinned by tests/test_ref_pin.py: the unmodified
mlx_data/dataloader.py runs under oracle/mlx_stub and its normalised properties, padded items and shuffled ragged batches
must equal this restatement's (tests/golden/ref_*.npz, keys ds/*)."""
import numpy as np


class MoleculeDatasetOracle:
    def __init__(self, tokenized_molecules, properties, max_length=120, pad_token=0, properties_mean=None,
                 properties_std=None):
        self.molecules = tokenized_molecules                                   # dataloader.py:31
        self.max_length = max_length
        self.pad_token = pad_token
        self.properties = np.array(properties, dtype=np.float32)               # :36
        if properties_mean is not None and properties_std is not None:         # :39-47
            self.properties_mean = np.array(properties_mean, dtype=np.float32)
            self.properties_std = np.array(properties_std, dtype=np.float32)
        else:
            self.properties_mean = self.properties.mean(axis=0, keepdims=True)
            self.properties_std = self.properties.std(axis=0, keepdims=True)
        if self.properties_mean.ndim == 1:                                     # :50-53
            self.properties_mean = self.properties_mean[np.newaxis, :]
        if self.properties_std.ndim == 1:
            self.properties_std = self.properties_std[np.newaxis, :]
        self.properties_std = np.where(self.properties_std < 1e-8, 1.0, self.properties_std)   # :56-60
        self.properties_normalized = (self.properties - self.properties_mean) / self.properties_std   # :63-65

    def __len__(self):
        return len(self.molecules)

    def __getitem__(self, idx):
        mol = list(self.molecules[idx])                                        # :72
        props = self.properties_normalized[idx]
        if len(mol) < self.max_length:                                         # :76-79
            mol = mol + [self.pad_token] * (self.max_length - len(mol))
        else:
            mol = mol[:self.max_length]
        return {"molecule": np.array(mol, dtype=np.uint32), "properties": np.array(props, dtype=np.float32)}   # :81-84

    def to_batches(self, batch_size, shuffle=True):
        indices = np.arange(len(self))                                         # :91
        if shuffle:
            np.random.shuffle(indices)                                         # :93-94
        for i in range(0, len(self), batch_size):                              # :96
            batch_indices = indices[i:i + batch_size]
            mols = [self[int(j)]["molecule"] for j in batch_indices]           # :102-105
            props = [self[int(j)]["properties"] for j in batch_indices]
            yield np.stack(mols), np.stack(props)                              # :108-111
