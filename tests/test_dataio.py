"""Data formats either side of the path (host side only): city .npz and test-set .npz, as the reference writes them
(create_dataset.py:169-174, scripts/generate_data.py:201-224,457) and reads them (generator_lazy.py:134-140, test.py:152-177)."""
import numpy as np
import pytest
import torch

from oracle import synth


def test_city_npz_round_trip(tmp_path):
    from rrnco_b200.dataio import load_city_npz
    city = synth.make_city(5, 60)
    p = tmp_path / "X_data.npz"
    np.savez_compressed(p, distance=city["distance"], duration=city["duration"], points=city["points"])
    got = load_city_npz(str(p))
    for k in ("distance", "duration", "points"):
        assert got[k].dtype == np.float64 and np.array_equal(got[k], city[k])
    np.savez_compressed(tmp_path / "bad.npz", distance=city["distance"][:, :10], points=city["points"])
    with pytest.raises(ValueError):
        load_city_npz(str(tmp_path / "bad.npz"))
    np.savez_compressed(tmp_path / "bad2.npz", duration=city["duration"])
    with pytest.raises(KeyError):
        load_city_npz(str(tmp_path / "bad2.npz"))


def test_rcvrp_test_set_npz(tmp_path):
    from rrnco_b200.dataio import iter_batches, load_npz_to_tensordict, prepare_test_td
    B, n = 10, 7
    rng = np.random.RandomState(0)
    data = {"depot": rng.rand(B, 2).astype(np.float32), "locs": rng.rand(B, n, 2).astype(np.float32),
            "demand": rng.randint(1, 10, (B, n)).astype(np.float32), "capacity": np.full(B, 30, np.float32),
            "distance_matrix": rng.rand(B, n + 1, n + 1).astype(np.float32)}
    p = tmp_path / "rcvrp.npz"
    np.savez(p, **data)
    td = prepare_test_td(load_npz_to_tensordict(str(p)), "rcvrp")
    assert td.batch_size[0] == B
    assert torch.equal(td["demand"], torch.from_numpy(data["demand"]) / 30) and (td["capacity"] == 1).all()
    assert torch.equal(td["distance_matrix"], torch.from_numpy(data["distance_matrix"]))
    sizes = [b.batch_size[0] for b in iter_batches(td, 4)]
    assert sizes == [4, 4, 2]
    last = list(iter_batches(td, 4))[-1]
    assert torch.equal(last["locs"], td["locs"][8:])
    with pytest.raises(KeyError):
        prepare_test_td(load_npz_to_tensordict(str(p)), "rcvrptw")
    with pytest.raises(ValueError):
        prepare_test_td(td, "cvrp")
    np.savez(tmp_path / "ragged.npz", a=np.zeros((3, 2)), b=np.zeros((4, 2)))
    with pytest.raises(ValueError):
        load_npz_to_tensordict(str(tmp_path / "ragged.npz"))
