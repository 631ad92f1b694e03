"""CPU tests of the device-side data pipeline (mlx_vae_b200.data.MoleculeDataset) against the NumPy
restatement of mlx_data/dataloader.py: padding / truncation, z-scoring with own and with injected statistics, zero-variance
columns, shuffle order from the global NumPy RNG, ragged last batch, empty dataset."""
import numpy as np
import pytest

import dataset_oracle as DO


def make(n=37, seed=0, maxlen=12):
    rng = np.random.default_rng(seed)
    mols = [list(rng.integers(3, 80, size=int(rng.integers(1, 2 * maxlen))).tolist()) for _ in range(n)]
    props = np.stack([rng.normal(70, 25, n), np.full(n, 3.0)], axis=1)       # second column has zero variance
    return mols, props


def test_dataset_matches_reference_restatement():
    from mlx_vae_b200.data import MoleculeDataset
    mols, props = make()
    ref = DO.MoleculeDatasetOracle(mols, props, max_length=12, pad_token=0)
    ds = MoleculeDataset(mols, props, max_length=12, pad_token=0, device="cpu")
    assert len(ds) == len(ref) == 37
    np.testing.assert_allclose(ds.properties_mean, ref.properties_mean)
    np.testing.assert_allclose(ds.properties_std, ref.properties_std)
    for i in (0, 5, 36):
        a, b = ds[i], ref[i]
        assert np.array_equal(a["molecule"].numpy().astype(np.uint32), b["molecule"])
        np.testing.assert_allclose(a["properties"].numpy(), b["properties"], rtol=1e-6)
    # validation set reuses the training statistics (train.py:113-114)
    m2, p2 = make(n=9, seed=1)
    ref2 = DO.MoleculeDatasetOracle(m2, p2, 12, 0, ref.properties_mean, ref.properties_std)
    ds2 = MoleculeDataset(m2, p2, 12, 0, ds.properties_mean, ds.properties_std, device="cpu")
    np.testing.assert_allclose(ds2.properties_normalized, ref2.properties_normalized, rtol=1e-6)
    for shuffle in (False, True):
        np.random.seed(67)
        rb = list(ref.to_batches(8, shuffle=shuffle))
        np.random.seed(67)
        gb = list(ds.to_batches(8, shuffle=shuffle))
        assert len(rb) == len(gb) == 5 and gb[-1][0].shape[0] == 5          # ragged last batch kept
        for (rm, rp), (gm, gp) in zip(rb, gb):
            assert gm.dtype.is_floating_point is False and tuple(gm.shape) == rm.shape
            assert np.array_equal(gm.numpy().astype(np.uint32), rm)
            np.testing.assert_allclose(gp.numpy(), rp, rtol=1e-6)


def test_empty_and_one_dimensional_properties():
    from mlx_vae_b200.data import MoleculeDataset
    ds = MoleculeDataset([], np.zeros((0, 1)), max_length=5, device="cpu", properties_mean=[0.0], properties_std=[1.0])
    assert len(ds) == 0 and list(ds.to_batches(4)) == []
    ds = MoleculeDataset([[3, 4, 2]], np.array([[50.0]]), max_length=5, device="cpu")
    m, p = next(ds.to_batches(4, shuffle=False))
    assert m.tolist() == [[3, 4, 2, 0, 0]] and p.shape == (1, 1) and float(p[0, 0]) == 0.0   # std 0 -> 1


def test_load_splits_follows_train_py():
    """train.py:75-124: seed 67, one shuffle of arange(N) from the global NumPy RNG, 80/10/10, train statistics reused."""
    from mlx_vae_b200 import data
    js = data.make_chembl_standin(333, max_length=40, seed=5)
    assert set(js) >= {"molecules", "tokenized_sequences", "max_length"} and len(js["molecules"]) == 333
    assert all(s[-1] == 2 and 0 not in s and len(s) <= 40 for s in js["tokenized_sequences"])
    np.random.seed(67)
    train, val, test = data.load_splits(js, device="cpu")
    np.random.seed(67)
    idx = np.arange(333); np.random.shuffle(idx)
    n_train, n_val = int(0.8 * 333), int(0.1 * 333)
    assert (len(train), len(val), len(test)) == (n_train, n_val, 333 - n_train - n_val)
    assert train.molecules[0] == js["tokenized_sequences"][idx[0]] and test.molecules[-1] == js["tokenized_sequences"][idx[-1]]
    props = np.array([[m["tpsa"]] for m in js["molecules"]], dtype=np.float32)
    sel = idx[n_train:n_train + n_val]
    ref = DO.MoleculeDatasetOracle([js["tokenized_sequences"][i] for i in sel], props[sel], max_length=40,
                                   properties_mean=train.properties_mean, properties_std=train.properties_std)
    assert np.array_equal(val.properties_normalized, ref.properties_normalized)
    assert np.array_equal(val.properties_mean, train.properties_mean)
