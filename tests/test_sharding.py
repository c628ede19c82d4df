"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard a batch and all-gather the costs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rrnco_b200.sharding import gather_costs, shard_bounds, shard_range, shard_td


def test_shard_bounds_cover_batch():
    for n in (0, 1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(10, 1, 4) == (3, 6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_items, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    batch = {"demand": torch.rand(n_items, 5, generator=g), "distance_matrix": torch.rand(n_items, 6, 6, generator=g)}
    mine = shard_td(batch, rank, world)
    # stand-in for the rollout: a per-instance "best cost" that depends only on the instance
    local = mine["distance_matrix"].sum((1, 2)) + mine["demand"].sum(1)
    full = gather_costs(local, n_items)
    want = batch["distance_matrix"].sum((1, 2)) + batch["demand"].sum(1)
    ok = torch.equal(full, want)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the timing reduction bench.py uses (max over ranks)
    ok = ok and t.item() == world
    torch.save(ok, os.path.join(out_dir, f"ok{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 9])
def test_gloo_two_ranks_shard_and_gather(tmp_path, n_items):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    assert all(torch.load(os.path.join(tmp_path, f"ok{r}.pt")) for r in range(world))


def _grad_worker(rank, world, port, n_items, out_dir):
    """Two ranks, shards of different sizes: REINFORCE gradients of the shards, all-reduced, equal the single-process gradient."""
    from rrnco_b200.sharding import allreduce_gradients, shard_range
    from rrnco_b200.training import pomo_shared_baseline_loss
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = 5
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(S, n_items, 7, generator=g)          # [starts, instances, features] (r = s * n_inst + b order)
    reward = torch.randn(S, n_items, generator=g)

    def model():
        torch.manual_seed(1)
        return torch.nn.Sequential(torch.nn.Linear(7, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))

    def backward(net, lo, hi):
        ll = net(feats[:, lo:hi]).squeeze(-1)                  # stand-in log-likelihoods with a graph to the parameters
        pomo_shared_baseline_loss(reward[:, lo:hi].reshape(-1), ll.reshape(-1), S).backward()

    ref = model()
    backward(ref, 0, n_items)
    net = model()
    lo, hi = shard_range(n_items, rank, world)
    backward(net, lo, hi)
    allreduce_gradients(net.parameters(), hi - lo, n_items)
    ok = all(torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7) for p, q in zip(net.parameters(), ref.parameters()))
    torch.save(ok, os.path.join(out_dir, f"gok{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 9])
def test_gloo_two_ranks_gradient_allreduce(tmp_path, n_items):
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    assert all(torch.load(os.path.join(tmp_path, f"gok{r}.pt")) for r in range(world))
