"""CPU, world_size 2 over gloo: the replica-sharding plan (sample_ddp.py partitioning), the optional frame
gather and the max-over-ranks timing reduction that bench.py uses under torchrun."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latte_b200 import sharding as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 3
        total, iters = S.plan_iterations(10, n, world)
        assert (total, iters) == (12, 2)
        mine = [i for k in range(iters) for i in S.video_indices(k, n, rank, world)]
        # frames tagged with their global index: the gather must restore global order
        frames = torch.tensor(S.video_indices(0, n, rank, world), dtype=torch.uint8).reshape(n, 1, 1, 1, 1).expand(n, 2, 4, 4, 3).contiguous()
        allf = S.gather_frames(frames)
        assert allf.shape == (n * world, 2, 4, 4, 3)
        assert allf[:, 0, 0, 0, 0].tolist() == list(range(n * world))
        t = S.max_over_ranks(1.0 + rank)
        assert t == float(world)
        assert S.rank_seed(7, rank, world) == 7 * world + rank
        q.put((rank, mine))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_gather_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # disjoint cover of the padded job, interleaved exactly like sample_ddp.py:171-173
    assert sorted(got[0] + got[1]) == list(range(12))
    assert got[0][:3] == [0, 2, 4] and got[1][:3] == [1, 3, 5]


def test_single_process_defaults():
    assert S.world() == (0, 1)
    assert S.max_over_ranks(3.5) == 3.5
    x = torch.zeros(2, 1, 2, 2, 3, dtype=torch.uint8)
    assert S.gather_frames(x) is x


# ------------------------------------------------------------------------------------------------ training: DDP over the autograd boundary
class _TrainWrap(torch.nn.Module):
    """What `Latte.forward` does in training mode, with the test backend (the native backend needs a GPU): the engine's forward with
    the hand-written backward attached as ONE autograd node whose outputs are the parameter gradients."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x, t, y):
        from latte_b200 import training
        from oracle.train_ops_oracle import TorchOps
        return training.train_forward(self.net, TorchOps(torch.float32), torch.float32, x, training.conditioning(self.net, t, y))


def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from latte_b200 import Latte
        torch.manual_seed(0)                                  # same parameters on both ranks
        net = Latte(input_size=16, hidden_size=128, depth=2, num_heads=2, num_frames=8, num_classes=11, extras=2)
        with torch.no_grad():
            for p in net.parameters():
                if p.requires_grad and float(p.abs().max()) == 0.0:
                    p.normal_(0, 0.02)
        g = torch.Generator().manual_seed(100 + rank)         # different data per rank
        x = torch.randn(2, 8, 4, 16, 16, generator=g)
        t = torch.randint(0, 1000, (2,), generator=g)
        y = torch.randint(0, 11, (2,), generator=g)
        dout = torch.randn(2, 8, 8, 16, 16, generator=g)
        local = _TrainWrap(net)
        local(x, t, y).backward(dout)
        mine = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        net.zero_grad(set_to_none=True)
        ddp = torch.nn.parallel.DistributedDataParallel(local)
        ddp(x, t, y).backward(dout)
        synced = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        assert set(mine) == set(synced) and len(mine) > 20
        worst = 0.0
        for k in mine:                                        # DDP's hooks must have averaged exactly these gradients
            both = [torch.zeros_like(mine[k]) for _ in range(world)]
            dist.all_gather(both, mine[k])
            want = sum(both) / world
            worst = max(worst, (synced[k] - want).abs().max().item() / (want.abs().max().item() + 1e-12))
        q.put((rank, worst))
    finally:
        dist.destroy_process_group()


def test_training_autograd_boundary_under_ddp():
    """world_size 2 over gloo: wrapping the module in DistributedDataParallel (train.py:125) averages the gradients that the
    engine's single autograd node hands to the parameters -- the reducer's hooks see ordinary leaf gradients."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert max(got.values()) < 1e-5, got
