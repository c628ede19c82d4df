"""Host-side (CPU) checks of the instance-generation laws, the x8 augmentation and the sampler's outlier cleaning against
fixtures recorded from the UNMODIFIED reference (tests/golden/make_golden.py: generator_*.npz, augment.npz,
sampler_outliers.npz).  The laws are plain element-wise torch ops, so they run on any device; the gather itself is a
CUDA kernel and is covered by the `gpu` tests."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].shape != () else z[k].item() for k in z.files}


def part(z, prefix):
    return {k[len(prefix):]: v for k, v in z.items() if k.startswith(prefix)}


def assert_td_equal(got, want, skip=()):
    assert set(want) - set(skip) <= set(got.keys()), set(want) - set(got.keys())
    for k, w in want.items():
        if k in skip:
            continue
        g = got[k]
        assert g.shape == w.shape and g.dtype == w.dtype, (k, g.shape, w.shape, g.dtype, w.dtype)
        assert torch.equal(g, w), (k, (g.float() - w.float()).abs().max())


def test_rcvrp_generator_law_bit_exact():
    """LazyRCVRPGenerator._process_real_world_data rrnco/envs/rcvrp/generator_lazy.py:275-304."""
    import rrnco_b200 as rb
    z = load("generator_rcvrp.npz")
    n = z["out.locs"].shape[1]
    gen = rb.LazyRCVRPGenerator(num_loc=n, device="cpu")
    assert gen.capacity == z["capacity_value"]
    td = gen._process_real_world_data(part(z, "in."), [6], draws=part(z, "draw."))
    assert_td_equal(td, part(z, "out."))


def test_atsp_generator_law_bit_exact():
    """LazyATSPGenerator._process_real_world_data rrnco/envs/atsp/generator_lazy.py:239-260 (locs eps 1e-6)."""
    import rrnco_b200 as rb
    z = load("generator_atsp.npz")
    td = rb.LazyATSPGenerator(num_loc=z["out.locs"].shape[1], device="cpu")._process_real_world_data(part(z, "in."), [6])
    assert_td_equal(td, part(z, "out."))


@pytest.mark.parametrize("preset", ["vrptw", "ovrpbltw"])
def test_rcvrptw_generator_law_and_variant_subsampling(preset):
    """LazyRMTVRPGenerator._process_real_world_data rrnco/envs/rmtvrp/generator_lazy.py:350-419 (locs eps 1e-8, duration
    min-max with the zero-range guard -- instance 2 has a constant duration matrix), demand / time-window laws
    rmtvrp/generator.py:445-469,515-562, subsample_problems :352-432.  Bit-exact except `distance_limit`, where upstream
    calls torch.cdist (<= 1e-6 relative)."""
    import rrnco_b200 as rb
    z = load(f"generator_rcvrptw_{preset}.npz")
    n = z["out.locs"].shape[1] - 1
    gen = rb.LazyRMTVRPGenerator(num_loc=n, variant_preset=preset, device="cpu")
    assert gen.capacity == z["capacity_value"]
    td = gen._process_real_world_data(part(z, "in."), [6], draws=part(z, "draw."))
    want = part(z, "out.")
    assert_td_equal(td, want, skip=("distance_limit",))
    assert torch.allclose(td["distance_limit"], want["distance_limit"], rtol=1e-6, atol=0)
    assert (td["duration_matrix"][2] == 0).all()  # zero range: (d - min) / 1
    sub = gen.subsample_problems(td)
    wsub = part(z, "sub.")
    assert_td_equal(sub, wsub, skip=("distance_limit",))
    assert torch.allclose(sub["distance_limit"], wsub["distance_limit"], rtol=1e-6, atol=0)
    if preset == "vrptw":  # configs/env/rcvrptw.yaml: closed routes, no limit, no backhauls, time windows kept
        assert not sub["open_route"].any() and torch.isinf(sub["distance_limit"]).all() and (sub["demand_backhaul"] == 0).all()
        assert torch.isfinite(sub["time_windows"][:, 1:, 1]).all()


def test_state_augmentation_dihedral8_matches_reference():
    """StateAugmentation(dihedral8, no_aug_coords=False) rrnco/models/utils/transforms.py:15-37,142-154 (test.py:28,188);
    the un-materialised form transforms the coordinates identically and leaves the matrices at B rows."""
    import rrnco_b200 as rb
    z = load("augment.npz")
    td = rb.TensorDictLite({"locs": z["locs_in"], "distance_matrix": z["dm_in"]}, batch_size=[3])
    out = rb.StateAugmentation(num_augment=8, augment_fn="dihedral8", first_aug_identity=True, no_aug_coords=False)(td)
    assert out.batch_size == torch.Size([24])
    assert torch.equal(out["locs"], z["locs_out"]) and torch.equal(out["distance_matrix"], z["dm_out"])
    shared = rb.StateAugmentation(num_augment=8, augment_fn="dihedral8", no_aug_coords=False, share_instance_data=True)(td)
    assert shared.batch_size == torch.Size([24]) and torch.equal(shared["locs"], z["locs_out"])
    assert shared["distance_matrix"].data_ptr() == td["distance_matrix"].data_ptr()  # not copied
    # copy a of instance b reads matrix row (a * B + b) % B = b
    r = torch.arange(24)
    assert torch.equal(shared["distance_matrix"][r % 3], z["dm_out"])
    # RRNet's training default (rl.py:54): no coordinate transform, pure batchify
    plain = rb.StateAugmentation(num_augment=8, augment_fn="dihedral8")(td)
    assert torch.equal(plain["locs"], rb.batchify(z["locs_in"], 8))
    with pytest.raises(AssertionError):
        rb.StateAugmentation(num_augment=4, augment_fn="dihedral8")


def test_sampler_outlier_removal_matches_reference():
    """Real_World_Sampler.sample on a city with unreachable pairs (> 1e5): rrnco/envs/rmtvrp/sampler.py:41-60."""
    import rrnco_b200 as rb
    z = np.load(os.path.join(GOLDEN, "sampler_outliers.npz"))
    city = {k[5:]: z[k] for k in z.files if k.startswith("city.")}
    clean = rb.remove_outlier_points(city)
    assert clean["distance"].max() <= 1e5 and len(clean["points"]) < len(city["points"])
    np.random.seed(99)
    idx = rb.Real_World_Sampler(with_duration=True, device="cpu").uniform_sample(4, len(clean["points"]), 9)
    assert np.array_equal(clean["points"][idx], z["points"])
    assert np.array_equal(clean["distance"][idx[:, :, None], idx[:, None, :]], z["distance_matrix"])
    assert np.array_equal(clean["duration"][idx[:, :, None], idx[:, None, :]], z["duration_matrix"])
    untouched = {k: v for k, v in city.items()}
    untouched["distance"] = np.minimum(city["distance"], 10.0)
    assert rb.remove_outlier_points(untouched) is untouched  # nothing above 1e5: the data is used as it is


def test_device_index_sampling_law():
    """uniform_sample_device: n distinct indices per instance, every index equally likely, order random (the law of
    np.random.choice(L, n, replace=False), rcvrp/sampler.py:97-104)."""
    import rrnco_b200 as rb
    g = torch.Generator().manual_seed(5)
    L, n, B = 50, 11, 4000
    idx = rb.Real_World_Sampler.uniform_sample_device(B, L, n, "cpu", g).long()
    assert idx.shape == (B, n) and idx.min() >= 0 and idx.max() < L
    assert (idx.sort(1)[0][:, 1:] != idx.sort(1)[0][:, :-1]).all()  # distinct
    counts = torch.bincount(idx.reshape(-1), minlength=L).double()
    exp = B * n / L
    chi2 = ((counts - exp) ** 2 / exp).sum().item()
    assert chi2 < (L - 1) + 5 * (2 * (L - 1)) ** 0.5, chi2
    first = torch.bincount(idx[:, 0], minlength=L).double()  # first position (the depot) uniform as well
    chi2 = ((first - B / L) ** 2 / (B / L)).sum().item()
    assert chi2 < (L - 1) + 5 * (2 * (L - 1)) ** 0.5, chi2
