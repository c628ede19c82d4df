"""CPU-side checks of the boundary: the C-ABI library builds/loads and exports every symbol the header
declares, the Python mirror keeps the reference's names, and the product path never routes through the
oracle or a CPU fallback."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from rrnco_b200.build import build_library
    return build_library()


def test_header_symbols_exported(lib_path):
    header = open(os.path.join(ROOT, "include", "rrnco_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rrnco_[a-z_0-9]+)\s*\(", header))
    assert {"rrnco_rollout", "rrnco_rcvrp_step", "rrnco_gather_submatrix", "rrnco_decoder_logits"} <= declared
    handle = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/rrnco_b200.h but not exported"
    from rrnco_b200 import _lib
    assert declared == set(_lib.exported_symbols()), "ctypes signatures out of sync with the header"


def test_abi_version_and_strerror(lib_path):
    from rrnco_b200 import _lib
    h = _lib.lib()
    assert h.rrnco_abi_version() == 1
    assert h.rrnco_strerror(0) == b"ok"
    assert b"unsupported" in h.rrnco_strerror(-2)
    assert h.rrnco_set_precision(2) == -1 and h.rrnco_set_precision(3) == 0
    assert h.rrnco_rollout_workspace_bytes(1, 101, 4, 101) >= 2 * 4 * 101 * 8


def test_gumbel_uniform_is_in_the_open_interval(lib_path):
    """rrnco_u01 (host twin of the device mapping): never 0 or 1, so -log(-log(u)) is finite for every Philox word."""
    from rrnco_b200 import _lib
    h = _lib.lib()
    one = np.float32(1.0)
    for x in (0, 1, 0x1FF, 0x200, 0x7FFFFFFF, 0xFFFFFE00, 0xFFFFFFFF):
        u = np.float32(h.rrnco_u01(x))
        assert np.float32(0.0) < u < one, (hex(x), u)
        assert np.isfinite(-np.log(-np.log(u)))
    assert np.float32(h.rrnco_u01(0xFFFFFFFF)) == np.float32(1.0 - 2.0 ** -24)
    assert np.float32(h.rrnco_u01(0)) == np.float32(2.0 ** -24)


def test_torch_library_ops_registered(lib_path):
    """csrc/torch_ops.cpp: the C-ABI entry points as dispatcher ops `torch.ops.rrnco_b200.*` (TORCH_LIBRARY over the same
    shared library); loading and schema lookup need no GPU."""
    from rrnco_b200.build import build_torch_ops
    from rrnco_b200 import torch_ops
    assert os.path.exists(build_torch_ops())
    assert torch_ops.load()
    ops = torch_ops.ops()
    for name in ("minmax_normalize", "gather_submatrix", "atsp_step", "rcvrp_step", "tour_reward", "select_action", "rollout"):
        schema = str(getattr(ops, name).default._schema)
        assert schema.startswith(f"rrnco_b200::{name}("), schema
    from rrnco_b200 import _lib
    assert ops.rollout_workspace_bytes(1, 101, 4, 101) == _lib.lib().rrnco_rollout_workspace_bytes(1, 101, 4, 101)
    if not torch.cuda.is_available():  # CUDA-only kernels: the dispatcher refuses host tensors, no silent fallback
        with pytest.raises((NotImplementedError, RuntimeError)):
            ops.minmax_normalize(torch.rand(2, 4, 4))


def test_bad_arguments_rejected_without_gpu(lib_path):
    from rrnco_b200 import _lib
    h = _lib.lib()
    # null pointers / sizes are validated before any CUDA call
    assert h.rrnco_minmax_normalize(4, 10, None, None, None, None, None) == -1
    assert h.rrnco_rcvrp_step(8, 1, 8, None, None, None, 1, None, None, None, None, None, None, None, None, None) == -1
    assert h.rrnco_minmax_normalize(0, 10, None, None, None, None, None) == 0  # empty batch is a no-op


def test_no_cpu_fallback():
    import rrnco_b200 as rb
    td = rb.TensorDictLite({"distance_matrix": torch.rand(2, 5, 5)}, batch_size=[2])
    env = rb.ATSPEnv(generator_params={"num_loc": 5})
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            env.reset(td)  # cannot move to CUDA, and there is no CPU path
    from rrnco_b200._lib import RRNCOError, ptr
    with pytest.raises(RRNCOError):
        ptr(torch.zeros(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rrnco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    code = "import sys; import rrnco_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_tdlite_batchify_matches_reference_semantics():
    import rrnco_b200 as rb
    from oracle import td as otd
    x = torch.arange(24).view(6, 4)
    assert torch.equal(rb.batchify(x, 3), otd.batchify(x, 3))
    y = rb.batchify(x, (2, 3))
    assert torch.equal(y, otd.batchify(x, (2, 3)))
    assert torch.equal(rb.unbatchify(y, (2, 3)), otd.unbatchify(y, (2, 3)))
    assert rb.unbatchify(y, (2, 3)).shape == (6, 2, 3, 4)
    td = rb.TensorDictLite({"a": x}, batch_size=[6])
    assert rb.batchify(td, 2)["a"].shape == (12, 4) and rb.batchify(td, 2).batch_size == torch.Size([12])
    assert torch.equal(rb.batchify(x, 0), x)  # n_aug = 0 in training is skipped (rl.py:104-106)


@pytest.mark.parametrize("name", ["atsp", "rcvrp", "rcvrptw"])
def test_decoder_state_dict_matches_reference_names(name):
    import rrnco_b200 as rb
    z = np.load(os.path.join(ROOT, "tests", "golden", f"policy_{name}.npz"))
    ref = {k[6:]: z[k].shape for k in z.files if k.startswith("param.")}
    dec = rb.RRNetDecoder(env_name=name)
    mine = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    assert mine == ref
    dec.load_state_dict({k: torch.from_numpy(z["param." + k]) for k in ref}, strict=True)


def test_unsupported_options_raise():
    import rrnco_b200 as rb
    with pytest.raises(NotImplementedError):
        rb.RRNetDecoder(embed_dim=256)
    with pytest.raises(ValueError):
        rb.RRNetDecoder(env_name="tsp")
    with pytest.raises(ValueError):
        rb.RRNetPolicy(env_name="rcvrp")  # the encoder must be supplied
    with pytest.raises(NotImplementedError):
        rb.RMTVRPEnv(select_start_nodes_fn="random")
