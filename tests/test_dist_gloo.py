"""CPU, world_size 2 over gloo: the batch-sharding logic around the drop-ins.  The per-pair
loss is injected (the CPU oracle port) because the product Chamfer has no CPU path; what is tested is the
sharding, the loss all-reduce and the gradient all-reduce, against the un-sharded single-process result."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from oracle import oracle as O
    D = importlib.import_module("gan-rl_3d_b200.distributed")
    r, lr, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    B = 5                                              # odd on purpose: shards of 3 and 2
    pc_in = O.make_clouds(B, 24, "sphere", 1)
    target = O.make_clouds(B, 20, "sphere", 2)
    torch.manual_seed(0)
    net = torch.nn.Linear(3, 3)                        # stands in for the autoencoder: same weights on all ranks

    def pair_fn(p, t, bidirectional=True):
        return O.ref_port_chamfer(p, t, bidirectional)

    # single-process truth over the whole batch (utils/losses.py:75 + train_rl_gan_net.py:236-240)
    net.zero_grad()
    full = torch.mean(pair_fn(net(pc_in), target))
    full.backward()
    want_grads = [p.grad.clone() for p in net.parameters()]

    lo, hi = D.shard_bounds(B, rank, world)
    net.zero_grad()
    loss = D.sharded_chamfer_loss(net(pc_in[lo:hi]), target[lo:hi], B, pair_fn=pair_fn)
    loss.backward()
    calls = D.allreduce_gradients(net.parameters())
    gathered = D.gather_rows(torch.full((2, 3), float(rank)))
    ok = (abs(loss.item() - full.item()) < 1e-6
          and all(torch.allclose(p.grad, g, rtol=1e-5, atol=1e-7) for p, g in zip(net.parameters(), want_grads))
          and calls == 1 and gathered.shape == (4, 3) and gathered[2:].eq(1).all().item())
    q.put((rank, bool(ok), loss.item(), full.item(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_and_gradients_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[4] for r in results) == [(0, 3), (3, 5)]
    assert all(r[1] for r in results), results


def test_shard_bounds_cover_batch_without_overlap():
    D = importlib.import_module("gan-rl_3d_b200.distributed")
    for n in (0, 1, 7, 32, 1024):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_single_process_is_identity():
    D = importlib.import_module("gan-rl_3d_b200.distributed")
    v = torch.tensor([1.0, 2.0, 3.0], requires_grad=True)
    out = D.sharded_mean_loss(v, 3)
    out.backward()
    assert out.item() == pytest.approx(2.0) and torch.allclose(v.grad, torch.full((3,), 1 / 3))
    assert D.allreduce_gradients([v]) == 0
