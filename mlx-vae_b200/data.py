"""Synthetic SELFIES-shaped batches for benchmarks and smoke runs.

Shape/dtype contract of the reference's data path (mlx_data/dataloader.py:76-82): right-padded with pad_token=0 to a
fixed max_length, end token 2 at the last real position (models/decoder.py:25-26), tokens unsigned 32-bit in the
reference / int32 here, one z-scored property column (TPSA, dataloader.py:63-65)."""
import numpy as np


def synthetic_batch(B: int, T: int, vocab_size: int = 80, num_conditions: int = 1, latent_dim: int = 128,
                    seed: int = 67, tf_ratio: float = 0.9, end_token: int = 2):
    """Lengths U{min(16,T)..T}; body tokens U{3..V-1}; returns (x int32 [B,T], cond f32 [B,C], eps f32 [B,L],
    tf_mask bool [T]) — tf_mask[t] plays the host coin of decoder.py:180."""
    rng = np.random.default_rng(seed)
    lo = min(16, T)
    lens = rng.integers(lo, T + 1, size=B)
    body = rng.integers(3, vocab_size, size=(B, T)).astype(np.int32)
    pos = np.arange(T)[None, :]
    x = np.where(pos < (lens[:, None] - 1), body, 0).astype(np.int32)
    x[np.arange(B), lens - 1] = end_token
    cond = rng.standard_normal((B, num_conditions)).astype(np.float32)
    eps = rng.standard_normal((B, latent_dim)).astype(np.float32)
    tf_mask = rng.random(T) < tf_ratio
    return x, cond, eps, tf_mask


class MoleculeDataset:
    """Drop-in for ``mlx_data/dataloader.py:4-111`` with the batching moved to the device.

    Same constructor, normalisation and batch contract as the reference:
      * properties are z-scored with the statistics passed in (the training set's, ``train.py:113-114``) or this
        dataset's own, ``std < 1e-8 -> 1`` (dataloader.py:39-65);
      * every sequence is right-padded with ``pad_token`` / truncated to ``max_length`` (:76-79);
      * ``to_batches(batch_size, shuffle)`` draws ONE ``np.random.shuffle`` of ``arange(N)`` from the global NumPy RNG
        (:90-94: same permutation as the reference for the same seed), walks it in ``batch_size`` strides and keeps
        the ragged last batch (:96-97).
    The reference builds each batch with a Python loop over samples (one ``mx.array`` per molecule); here the padded
    token matrix and the normalised properties are uploaded ONCE and a batch is a single device gather, so the loader
    keeps up with a step that consumes ~400 k molecules/s.  Tokens are int32 (reference: uint32)."""

    def __init__(self, tokenized_molecules, properties, max_length: int = 120, pad_token: int = 0,
                 properties_mean=None, properties_std=None, *, device=None):
        import torch
        self.molecules = tokenized_molecules
        self.max_length = int(max_length)
        self.pad_token = int(pad_token)
        self.properties = np.array(properties, dtype=np.float32)
        if self.properties.ndim == 1:
            self.properties = self.properties[:, None]
        if properties_mean is not None and properties_std is not None:
            self.properties_mean = np.array(properties_mean, dtype=np.float32)
            self.properties_std = np.array(properties_std, dtype=np.float32)
        else:
            self.properties_mean = self.properties.mean(axis=0, keepdims=True)
            self.properties_std = self.properties.std(axis=0, keepdims=True)
        if self.properties_mean.ndim == 1:
            self.properties_mean = self.properties_mean[np.newaxis, :]
        if self.properties_std.ndim == 1:
            self.properties_std = self.properties_std[np.newaxis, :]
        self.properties_std = np.where(self.properties_std < 1e-8, 1.0, self.properties_std).astype(np.float32)
        self.properties_normalized = ((self.properties - self.properties_mean) / self.properties_std).astype(np.float32)
        n = len(tokenized_molecules)
        tok = np.full((n, self.max_length), self.pad_token, dtype=np.int32)
        for i, mol in enumerate(tokenized_molecules):
            m = min(len(mol), self.max_length)
            tok[i, :m] = np.asarray(mol[:m], dtype=np.int64)
        self.tokens = tok
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)
        self._dev = None

    def __len__(self) -> int:
        return len(self.molecules)

    def __getitem__(self, idx: int) -> dict:
        import torch
        return {"molecule": torch.as_tensor(self.tokens[idx]), "properties": torch.as_tensor(self.properties_normalized[idx])}

    def _resident(self):
        import torch
        if self._dev is None:
            self._dev = (torch.as_tensor(self.tokens).to(self.device), torch.as_tensor(self.properties_normalized).to(self.device))
        return self._dev

    def to_batches(self, batch_size: int, shuffle: bool = True):
        import torch
        indices = np.arange(len(self))
        if shuffle:
            np.random.shuffle(indices)                                   # dataloader.py:93-94, global NumPy RNG
        tok, props = self._resident()
        idx_dev = torch.as_tensor(indices, dtype=torch.long).to(self.device)
        for i in range(0, len(self), batch_size):
            sel = idx_dev[i:i + batch_size]                              # ragged last batch kept (:96-97)
            yield tok.index_select(0, sel), props.index_select(0, sel)


# ---- the dataset file of train.py (``mlx_data/chembl_cns_selfies.json``) ------------------------------------------------
def make_chembl_standin(n_molecules: int = 4096, max_length: int = 128, vocab_size: int = 80, seed: int = 67,
                        end_token: int = 2) -> dict:
    """A SCHEMA-COMPATIBLE synthetic stand-in for the reference's bundled ``chembl_cns_selfies.json``, which is absent
    from the reference mount (``.MISSING_LARGE_BLOBS:1``).  Schema as read by ``train.py:79-83, :102``:
    ``{"molecules": [{"tpsa": float}, ...], "tokenized_sequences": [[int, ...], ...], "max_length": int}``.
    Content is synthetic (SELFIES-shaped: body tokens 3..V-1, end token 2; CNS-drug-like lengths ~ 20..70 tokens,
    TPSA ~ 20..120 A^2), NOT chemistry: it exercises the data path and the throughput, not model quality."""
    rng = np.random.default_rng(seed)
    lens = np.clip(rng.normal(42.0, 12.0, size=n_molecules).round().astype(np.int64), 8, max_length)
    seqs = [[int(t) for t in rng.integers(3, vocab_size, size=int(n) - 1)] + [end_token] for n in lens]
    tpsa = np.clip(rng.gamma(6.0, 10.0, size=n_molecules), 3.0, 200.0)
    return {"molecules": [{"tpsa": float(v)} for v in tpsa], "tokenized_sequences": seqs, "max_length": int(max_length),
            "note": "synthetic stand-in, schema of mlx_data/chembl_cns_selfies.json (train.py:79-124)"}


def load_splits(data, *, device=None, train_split: float = 0.8, val_split: float = 0.1):
    """``train.py:79-124``: read the dataset dict (or a path to the JSON), shuffle the indices with the GLOBAL NumPy RNG
    (the caller seeds it: ``np.random.seed(67)``, train.py:75), split 80/10/10 and build three ``MoleculeDataset`` objects,
    the validation and test sets normalised with the TRAINING set's statistics (:113-114, :122-123)."""
    if isinstance(data, (str, bytes)) or hasattr(data, "__fspath__"):
        import json
        with open(data, "r") as f:
            data = json.load(f)
    properties = np.array([[mol["tpsa"]] for mol in data["molecules"]], dtype=np.float32)
    sequences = data["tokenized_sequences"]
    indices = np.arange(len(sequences))
    np.random.shuffle(indices)
    n_total = len(sequences)
    n_train, n_val = int(train_split * n_total), int(val_split * n_total)
    parts = (indices[:n_train], indices[n_train:n_train + n_val], indices[n_train + n_val:])
    train = MoleculeDataset([sequences[i] for i in parts[0]], properties[parts[0]], max_length=data["max_length"],
                            pad_token=0, device=device)
    rest = [MoleculeDataset([sequences[i] for i in idx], properties[idx], max_length=data["max_length"], pad_token=0,
                            properties_mean=train.properties_mean, properties_std=train.properties_std, device=device)
            for idx in parts[1:]]
    return train, rest[0], rest[1]
