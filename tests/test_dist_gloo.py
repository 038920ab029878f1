"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path (3dinfomax_b200/dist.py).

  * all_gather_rows is differentiable and its backward is a reduce-scatter(sum);
  * local rows x gathered columns with (row_offset, total_rows) compose to the single-process global loss and to the
    same embedding gradients (checked with the oracle's loss on CPU tensors — the CUDA loss kernel is covered by
    tests/gpu_cases.py::case_ntxent with the same row_offset contract).
"""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_loss(z1_local, z2_all, C, row_offset, total_rows, tau):
    """oracle restatement of the sharded loss: rows = local molecules, columns = all molecules' conformers"""
    B = z1_local.shape[0]
    z2v = z2_all.view(-1, C, z2_all.shape[1])
    sim = torch.einsum("ik,juk->iju", z1_local, z2v)
    sim = sim / (z1_local.norm(dim=1)[:, None, None] * z2v.norm(dim=2)[None, :, :])
    sim = torch.exp(sim / tau).sum(dim=2)
    pos = sim[torch.arange(B), row_offset + torch.arange(B)]
    return -torch.log(pos / (sim.sum(dim=1) - pos)).sum() / total_rows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D = importlib.import_module("3dinfomax_b200.dist")
        torch.manual_seed(0)
        B, C, dim, tau = 8, 3, 16, 0.1
        z1 = torch.randn(B, dim)
        z2 = torch.randn(B * C, dim)
        lo, hi = D.shard_bounds(B, rank, world)
        a = z1[lo:hi].clone().requires_grad_(True)
        b = z2[lo * C:hi * C].clone().requires_grad_(True)
        gathered = D.all_gather_rows(b)
        assert gathered.shape == (B * C, dim) and torch.equal(gathered.detach(), z2)
        loss = _local_loss(a, gathered, C, lo, B, tau)
        loss.backward()
        total = loss.detach().clone()
        dist.all_reduce(total)
        # single-process reference on the global batch
        A, Bz = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        ref = O.ntxent_multiple_positives(A, Bz, tau=tau)
        ref.backward()
        ok = (abs(total.item() - ref.item()) < 1e-6 and torch.allclose(a.grad, A.grad[lo:hi], atol=1e-6)
              and torch.allclose(b.grad, Bz.grad[lo * C:hi * C], atol=1e-6))
        q.put((rank, ok, total.item(), ref.item()))
    finally:
        dist.destroy_process_group()


def test_sharded_contrastive_loss_matches_global_loss():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res


def test_shard_bounds():
    D = importlib.import_module("3dinfomax_b200.dist")
    assert [D.shard_bounds(2048, r, 8) for r in (0, 7)] == [(0, 256), (1792, 2048)]
    assert D.shard_bounds(10, 1, 4) == (2, 4)
