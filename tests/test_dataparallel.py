"""world_size-2 gloo (CPU) tests of the data-parallel plumbing: the bucketed gradient exchange
must reproduce the single-process gradient of the concatenated batch, whatever the bucket size."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _net():
    torch.manual_seed(7)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1), nn.ReLU(),
                         nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(8, 5))


def _data():
    g = torch.Generator().manual_seed(11)
    return torch.randn(8, 3, 12, 12, generator=g), (torch.rand(8, 5, generator=g) < 0.3).float()


def _worker(rank, world, port, bucket_mb, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from chexpert_b200.dataparallel import GradientBuckets
    net = _net()
    if rank == 1:                         # replicas must be re-synchronised from rank 0 by the constructor
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    gb = GradientBuckets(net, bucket_mb=bucket_mb)
    x, t = _data()
    xs, ts = x.chunk(world)[rank], t.chunk(world)[rank]
    grads = []
    for _ in range(2):                    # two steps: reset() must really clear the buckets
        gb.reset()
        loss = nn.functional.binary_cross_entropy_with_logits(net(xs), ts, reduction='none').sum(1).mean(0)
        loss.backward()
        gb.finish()
        grads = [p.grad.clone() for p in net.parameters()]
    torch.save({'grads': grads, 'nb': len(gb.buckets), 'params': [p.detach().clone() for p in net.parameters()]},
               os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.parametrize('bucket_mb', [25.0, 0.0005])
def test_bucketed_allreduce_matches_full_batch(tmp_path, bucket_mb):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), bucket_mb, str(tmp_path)), nprocs=world, join=True)
    net = _net()
    x, t = _data()
    nn.functional.binary_cross_entropy_with_logits(net(x), t, reduction='none').sum(1).mean(0).backward()
    want = [p.grad for p in net.parameters()]
    r0, r1 = (torch.load(tmp_path / f'r{r}.pt') for r in range(world))
    if bucket_mb < 1:
        assert r0['nb'] > 1               # the tiny cap really splits the parameters into several buckets
    for a, b in zip(r0['params'], r1['params']):
        assert torch.equal(a, b)
    for g0, g1, w in zip(r0['grads'], r1['grads'], want):
        assert torch.equal(g0, g1)
        assert torch.allclose(g0, w, rtol=1e-5, atol=1e-7)


def test_single_process_is_a_noop():
    from chexpert_b200.dataparallel import GradientBuckets
    net = _net()
    gb = GradientBuckets(net)
    x, t = _data()
    gb.reset()
    nn.functional.binary_cross_entropy_with_logits(net(x), t).backward()
    gb.finish()
    ref = _net()
    nn.functional.binary_cross_entropy_with_logits(ref(x), t).backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad)
