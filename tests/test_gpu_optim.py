"""Host-side contracts of the fused optimiser and the captured step (round-2 advisor findings): device / aliasing checks, the
packed-operand registry, torch.optim.Adam-compatible checkpoints, hyper-parameter changes after CUDA-graph capture."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset():
    import ssunet_gan_b200 as ssg
    yield
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")


def _tiny_net():
    from ssunet_gan_b200 import nn_layers
    torch.manual_seed(3)
    return torch.nn.Sequential(nn_layers.Conv2d(64, 64, 3, padding=1), nn_layers.Conv2d(64, 128, 3, padding=1))


def test_fused_adam_requires_cuda_parameters():
    """The reference builds Adam before .cuda() (train_seg_gan.py:452,474); here that order must fail loudly, not hand host
    pointers to device kernels."""
    from ssunet_gan_b200 import _lib, optim
    net = _tiny_net()
    with pytest.raises(_lib.SsgError, match="cuda"):
        optim.FusedClampAdam(net.parameters(), lr=1e-3)


def test_fused_adam_readopts_drifted_parameters_and_matches_torch_adam():
    from ssunet_gan_b200 import optim
    net = _tiny_net().cuda()
    ref = _tiny_net().cuda()
    opt = optim.FusedClampAdam(net.parameters(), lr=1e-3)
    topt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(3):
        grads = [torch.randn(p.shape, device="cuda", generator=g) for p in net.parameters()]
        if it == 1:
            # what Module._apply does (.cuda() / .to() / .float()): parameters get NEW storage, the Python tags survive
            for p in net.parameters():
                p.data = p.data.clone()
                p.grad = p.grad.clone()
        opt.zero_grad()
        for p, q, gr in zip(net.parameters(), ref.parameters(), grads):
            p.grad.copy_(gr)
            q.grad = gr.clone()
        opt.step()
        topt.step()
        for p, off in zip(opt._params, opt._offsets):
            assert p.data_ptr() == opt.flat_p.data_ptr() + 4 * off and p.grad.data_ptr() == opt.flat_g.data_ptr() + 4 * off
    for p, q in zip(net.parameters(), ref.parameters()):
        assert float((p - q).abs().max()) < 2e-6
    # checkpoints interchange with torch.optim.Adam's layout
    sd, tsd = opt.state_dict(), topt.state_dict()
    assert sorted(sd["state"].keys()) == sorted(tsd["state"].keys())
    for i in sd["state"]:
        assert float(sd["state"][i]["step"]) == float(tsd["state"][i]["step"]) == 3.0
        assert float((sd["state"][i]["exp_avg"] - tsd["state"][i]["exp_avg"]).abs().max()) < 1e-6
        assert float((sd["state"][i]["exp_avg_sq"] - tsd["state"][i]["exp_avg_sq"]).abs().max()) < 1e-6
    net2 = _tiny_net().cuda()
    net2.load_state_dict(net.state_dict())
    opt2 = optim.FusedClampAdam(net2.parameters(), lr=1e-3)
    opt2.load_state_dict(tsd)                      # resume FROM a torch.optim.Adam checkpoint
    assert opt2._step == 3
    grads = [torch.randn(p.shape, device="cuda", generator=g) for p in net.parameters()]
    for p, q, r, gr in zip(net.parameters(), ref.parameters(), net2.parameters(), grads):
        q.grad = gr.clone()
        r.grad.copy_(gr)
    topt.step()
    opt2.step()
    for q, r in zip(ref.parameters(), net2.parameters()):
        assert float((q - r).abs().max()) < 2e-6


def test_pack_registry_tracks_the_optimiser_and_external_changes():
    """Packed conv operands live in persistent buffers refreshed by ONE launch after the step; a change made by somebody else
    (load_state_dict, an in-place edit) is picked up on the next use, in place (captured graphs keep the address)."""
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import ops, optim
    from ssunet_gan_b200._lib import W_RSKC
    ssg.set_compute_dtype(torch.bfloat16)
    net = _tiny_net().cuda()
    opt = optim.FusedClampAdam(net.parameters(), lr=1e-2)
    w = net[0].weight

    def fresh():
        out = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
        ops.call("ssg_pack_conv_weight_pad", w.detach(), out, ops.dtype_code(torch.bfloat16), W_RSKC, 64, 64, 3, 3, 64, 64, None)
        return out

    a = ops.packed_weight(w, W_RSKC, torch.bfloat16)
    assert torch.equal(a, fresh()) and len(opt.packs.entries) == 1
    ptr = a.data_ptr()
    opt.zero_grad()
    for p in net.parameters():
        p.grad.fill_(1.0)
    launches = ops._lib.launch_count
    opt.step()                                   # Adam + ONE multi-pack launch
    assert ops._lib.launch_count - launches == 2
    b = ops.packed_weight(w, W_RSKC, torch.bfloat16)
    assert b.data_ptr() == ptr and torch.equal(b, fresh())
    assert ops._lib.launch_count - launches == 3          # (only the fresh() reference pack above was launched since)
    with torch.no_grad():
        w.mul_(0.5)                              # external change: version counter moves
    c = ops.packed_weight(w, W_RSKC, torch.bfloat16)
    assert c.data_ptr() == ptr and torch.equal(c, fresh())
    # another optimiser's step must not invalidate this registry (round 2 bug: a global epoch made every step re-pack everything)
    other = optim.FusedClampAdam(_tiny_net().cuda().parameters(), lr=1e-2)      # (construction re-homes parameters: global invalidation)
    ops.packed_weight(w, W_RSKC, torch.bfloat16)
    other.zero_grad()
    other.step()
    launches = ops._lib.launch_count
    ops.packed_weight(w, W_RSKC, torch.bfloat16)
    assert ops._lib.launch_count == launches


def test_graphed_step_follows_param_group_changes():
    """lr / clip live in device memory and are refreshed before every replay: an LR scheduler keeps working after capture,
    and the host-side step count advances."""
    import ssunet_oracle as O
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import models_seg_gan, optim, train_step
    ssg.set_compute_dtype(torch.bfloat16)
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
    g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
    d = models_seg_gan.Discriminator(3)
    d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3)))
    g.cuda().train(); d.cuda().train()
    og = optim.FusedClampAdam(g.parameters(), lr=2e-5)
    od = optim.FusedClampAdam(d.parameters(), lr=2e-5)
    x, t = O.synthetic_batch(2, 3, 64, 64, seed=1234, blobby=True)
    step = train_step.GraphedGanStep(g, d, og, od, (2, 3, 64, 64))
    w = g.net.final.weight
    w0 = w.detach().clone()
    r = step(x.cuda(), t.cuda())
    iou, dice = step.metrics()
    assert 0.0 < iou < 1.0 and 0.0 < float(dice) < 1.0 and np.isfinite(float(r["loss"]))
    d1 = float((w.detach() - w0).abs().max())
    assert 0 < d1 <= 2.01e-5 and og._step == 1 and od._step == 1
    for grp in og.param_groups:
        grp["lr"] = 2e-3                          # what a scheduler does
    w1 = w.detach().clone()
    step(x.cuda(), t.cuda())
    d2 = float((w.detach() - w1).abs().max())
    assert 1e-3 < d2 <= 2.1e-3, d2                # the replayed Adam kernel used the new learning rate
    assert og._step == 2
    sd = og.state_dict()
    assert float(sd["state"][0]["step"]) == 2.0


def test_multi_pack_equals_per_tensor_pack():
    """`ssg_pack_conv_weights_multi` (one launch, 32 x 32 channel tiles through shared memory) against the per-tensor kernel: all
    three layouts, 1 x 1 and 3 x 3, channel counts off the tile size, zero-padded channel extents."""
    from ssunet_gan_b200 import ops
    from ssunet_gan_b200._lib import W_RSKC, W_RSCK, W_RSCK_FLIP
    g = torch.Generator(device="cuda").manual_seed(5)
    reg = ops.PackRegistry()
    cases = []
    for (cout, cin, k, cout_p, cin_p) in [(64, 64, 3, 64, 64), (70, 40, 3, 72, 40), (3, 64, 3, 8, 64), (128, 3, 3, 128, 8),
                                          (200, 96, 1, 200, 96), (33, 31, 1, 40, 32), (8, 8, 3, 8, 8)]:
        w = torch.randn(cout, cin, k, k, device="cuda", generator=g)
        for layout in (W_RSKC, W_RSCK, W_RSCK_FLIP):
            for dt in (torch.bfloat16, torch.float32):
                cases.append((w, layout, dt, cout_p, cin_p, reg.get(w, layout, dt, cout_p, cin_p)))
    for c in cases:
        c[-1].fill_(7.0)               # `get` packed each operand with the per-tensor kernel: wipe it, the multi kernel must rewrite all
    reg.refresh()
    for w, layout, dt, cout_p, cin_p, got in cases:
        cout, cin, k, _ = w.shape
        ref = torch.empty(cout_p * cin_p * k * k, dtype=dt, device="cuda")
        ops.call("ssg_pack_conv_weight_pad", w, ref, ops.dtype_code(dt), layout, cout, cin, k, k, cout_p, cin_p, None)
        assert torch.equal(got, ref), (tuple(w.shape), layout, dt, cout_p, cin_p)
