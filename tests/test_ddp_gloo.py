"""World-size-2 gloo test (CPU) of the multi-GPU host logic: the flat-bucket gradient all-reduce must equal the
mean of the per-rank gradients, and the .grad views must survive an optimizer step + zero_grad."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker_bucketed(rank, world, port, out):
    """BucketedGradAllReduce: hooks + per-bucket async all-reduce; three steps (bucket order is frozen after the first)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from occuseg_b200.ddp import BucketedGradAllReduce
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(),
                              torch.nn.Linear(64, 3))
    unused = torch.nn.Parameter(torch.zeros(10))             # a parameter that never receives a gradient
    params = list(net.parameters()) + [unused]
    opt = torch.optim.SGD(params, lr=0.1)
    red = BucketedGradAllReduce(params, world, bucket_mb=0.004)
    ok = True
    for it in range(3):
        g = torch.Generator().manual_seed(100 * it + rank)
        x = torch.randn(11, 5, generator=g)
        # the same forward/backward on a detached copy gives this rank's local gradient
        import copy
        twin = copy.deepcopy(net)
        twin.zero_grad(set_to_none=True)
        twin(x).square().mean().backward()
        local = [p.grad.clone() for p in twin.parameters()]
        net(x).square().mean().backward()
        red.finish()
        for p, l in zip(net.parameters(), local):
            bucket = [torch.zeros_like(l) for _ in range(world)]
            dist.all_gather(bucket, l)
            ok &= torch.allclose(p.grad, sum(bucket) / world, atol=1e-6)
        ok &= bool((unused.grad == 0).all())
        opt.step()
        opt.zero_grad(set_to_none=False)
        ok &= red.check_views()
    ok &= len(red.buckets) >= 2 and red.frozen
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    ok &= torch.equal(both[0], both[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_bucketed, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from occuseg_b200.ddp import FlatGradAllReduce
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    red = FlatGradAllReduce(net.parameters(), world)
    ok = True
    for it in range(2):
        g = torch.Generator().manual_seed(100 * it + rank)
        x = torch.randn(11, 5, generator=g)
        net(x).square().mean().backward()
        local = [p.grad.clone() for p in net.parameters()]
        red.all_reduce()
        # reference: gather every rank's local gradient and average
        for p, l in zip(net.parameters(), local):
            bucket = [torch.zeros_like(l) for _ in range(world)]
            dist.all_gather(bucket, l)
            ok &= torch.allclose(p.grad, sum(bucket) / world, atol=1e-7)
        opt.step()
        opt.zero_grad(set_to_none=False)
        ok &= red.check_views()
    # weights stay identical across ranks
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    ok &= torch.equal(both[0], both[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
