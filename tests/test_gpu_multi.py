"""Two-rank NCCL checks of the data-parallel path on real GPUs (needs >= 2 devices; skipped otherwise):
SynchronizedBatchNorm2d over 2 ranks x batch b == single-device BN over batch 2b, through both the
NVLink peer-memory exchange (csrc/p2p.cu) and the NCCL fallback, and the flat gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, use_nccl_stats, ret):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if use_nccl_stats:
        os.environ["SSG_SYNCBN_NCCL"] = "1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ssunet_gan_b200 as ssg
        from ssunet_gan_b200 import batchnorm, nn_layers, ops, replicate
        ssg.set_compute_dtype(torch.float32)
        g = torch.Generator().manual_seed(3)
        C = 24
        xb = torch.randn(6, C, 9, 7, generator=g) * 1.5 + 0.3
        gy = torch.randn(6, C, 9, 7, generator=g)
        gamma = 1 + 0.1 * torch.randn(C, generator=g)
        beta = 0.1 * torch.randn(C, generator=g)

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.bn = batchnorm.SynchronizedBatchNorm2d(C)

            def forward(self, x):
                return ops.to_nchw_f32(self.bn(x))

        net = Net()
        with torch.no_grad():
            net.bn.weight.copy_(gamma); net.bn.bias.copy_(beta)
        net.cuda().train()
        dp = replicate.DataParallelWithCallback(net)
        xs = xb.chunk(world)[rank].cuda().requires_grad_(True)
        outs = []
        for it in range(3):                       # several exchanges: both receive slots and the epoch counter are exercised
            y = dp(xs)
            (y * gy.chunk(world)[rank].cuda()).sum().backward()
            outs.append((y.detach().cpu(), xs.grad.detach().cpu().clone(), net.bn.weight.grad.detach().cpu().clone()))
            xs.grad = None
            net.bn.weight.grad.zero_(); net.bn.bias.grad.zero_()
        # single-device reference on the whole batch (the reference's clamp(eps) == +eps here: var >> eps)
        xr = xb.clone().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        yr = torch.nn.functional.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5)
        (yr * gy).sum().backward()
        ok = True
        for y, dx, dgam in outs:
            ok &= torch.allclose(y, yr.detach().chunk(world)[rank], atol=2e-5, rtol=1e-4)
            ok &= torch.allclose(dx, xr.grad.chunk(world)[rank], atol=2e-5, rtol=1e-3)
            ok &= torch.allclose(dgam, gr.grad / world, atol=1e-4, rtol=1e-3)     # wrapper averages parameter gradients
        used_p2p = ops.PeerStatReducer.for_group(dist.group.WORLD) is not None
        rm = net.bn.running_mean.detach().cpu()
        gathered = [torch.zeros_like(rm) for _ in range(world)]
        dist.all_gather_object(gathered, rm)
        same_stats = all(torch.equal(gathered[0], t) for t in gathered)          # rank-ordered reduction: bit-identical
        ret[rank] = (bool(ok), used_p2p, same_stats)
    finally:
        torch.cuda.synchronize()
        dist.destroy_process_group()


@pytest.mark.parametrize("use_nccl_stats", [False, True])
def test_syncbn_two_ranks_matches_full_batch(use_nccl_stats):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), use_nccl_stats, ret), nprocs=world, join=True)
    for rank in range(world):
        ok, used_p2p, same_stats = ret[rank]
        assert ok, "SyncBN over 2 ranks does not reproduce full-batch BN (rank %d)" % rank
        assert same_stats, "running statistics differ between ranks"
        assert used_p2p == (not use_nccl_stats), "expected the %s statistics exchange" % ("NCCL" if use_nccl_stats else "peer-memory")
