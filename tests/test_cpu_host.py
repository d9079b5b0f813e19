"""CPU-side checks: the C-ABI library loads and exports everything the header declares, the product
refuses to run without CUDA, host-side logic (pairwise-sum tree, SyncMaster rendezvous, data-parallel
gradient averaging and SyncBN statistics exchange over gloo with world_size 2)."""
import ctypes
import os
import re
import socket
import threading

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ssunet_gan_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "ssunet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ssg_\w+)\s*\(", hdr)))
    assert len(declared) >= 50
    for name in declared:
        assert hasattr(L, name), "missing export %s" % name
    assert sorted(_lib.exported_symbols()) == declared
    assert L.ssg_version() >= 100


def test_product_has_no_cpu_path():
    from ssunet_gan_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.SsgError):
        _lib.call("ssg_add", None, None, None, 0, 0)
    with pytest.raises(_lib.SsgError):
        ops.conv2d(torch.zeros(1, 3, 8, 8), torch.zeros(4, 3, 3, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ssunet-gan_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "ssunet_oracle" not in src and "oracle/" not in src and "/root/reference" not in src, f


def test_pairwise_tree_host_helpers_match_numpy():
    from ssunet_gan_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(0)
    for n in (1, 5, 8, 127, 128, 129, 255, 256, 257, 1000, 4099, 65536, 100003, 786432):
        a = (rs.rand(n).astype(np.float32) * 3).astype(np.float32)
        cnt = L.ssg_pairwise_leaves_host(n, None, 0)
        offs = np.empty(cnt + 1, dtype=np.int64)
        assert L.ssg_pairwise_leaves_host(n, offs.ctypes.data_as(ctypes.c_void_p), cnt + 1) == cnt
        assert offs[0] == 0 and offs[-1] == n and np.all(np.diff(offs) > 0) and np.all(np.diff(offs) <= 128)
        leaf = np.array([a[offs[i]:offs[i + 1]].sum() for i in range(cnt)], dtype=np.float32)
        got = L.ssg_pairwise_combine_host(leaf.ctypes.data_as(ctypes.c_void_p), n)
        assert np.float32(got) == a.sum(), n


def test_sync_master_rendezvous():
    """comm.py semantics: one message per slave, callback on the master, reply to each, ACKs drained."""
    from ssunet_gan_b200.comm import SyncMaster

    def callback(msgs):
        total = sum(m for _, m in msgs)
        return [(ident, total + ident) for ident, _ in msgs]

    master = SyncMaster(callback)
    pipes = [master.register_slave(i) for i in (1, 2, 3)]
    out = {}

    def slave(p):
        out[p.identifier] = p.run_slave(10 * p.identifier)

    ts = [threading.Thread(target=slave, args=(p,)) for p in pipes]
    [t.start() for t in ts]
    out[0] = master.run_master(5)
    [t.join() for t in ts]
    assert out == {0: 65, 1: 66, 2: 67, 3: 68}
    assert master.nr_slaves == 3
    master.register_slave(1)      # new round re-initialises the registry
    assert master.nr_slaves == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ssunet_oracle as O
        from ssunet_gan_b200 import batchnorm, replicate
        torch.manual_seed(100 + rank)          # ranks start from DIFFERENT weights: the wrapper must broadcast rank 0's
        net = torch.nn.Sequential(torch.nn.Linear(6, 4), batchnorm.SynchronizedBatchNorm1d(4), torch.nn.Linear(4, 2))
        dp = replicate.DataParallelWithCallback(net)
        w0 = net[0].weight.detach().clone()
        gathered = [torch.zeros_like(w0) for _ in range(world)]
        dist.all_gather(gathered, w0)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        sbn = net[1]
        par = (sbn._is_parallel, sbn._parallel_id, sbn._world)
        # gradient averaging == gradient of the global-batch mean loss
        g = torch.Generator().manual_seed(5)
        xs = torch.randn(8, 6, generator=g)
        lin = torch.nn.Linear(6, 3)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(3, 6, generator=g)); lin.bias.zero_()
        dpl = replicate.DataParallelWithCallback(lin)
        shard = xs.chunk(world)[rank]
        dpl(shard).pow(2).mean().backward()
        full = torch.nn.Linear(6, 3)
        with torch.no_grad():
            full.weight.copy_(lin.weight); full.bias.zero_()
        full(xs).pow(2).mean().backward()
        grad_ok = torch.allclose(lin.weight.grad, full.weight.grad, atol=1e-6)
        # SyncBN statistics exchange: all-reduced (sum, ssum) over ranks == single-device BN on the whole batch
        xb = torch.randn(6, 8, 5, 7, generator=torch.Generator().manual_seed(9)) * 2 + 0.5
        sd = {"bn.weight": torch.ones(8), "bn.bias": torch.zeros(8), "bn.running_mean": torch.zeros(8),
              "bn.running_var": torch.ones(8), "bn.num_batches_tracked": torch.zeros((), dtype=torch.int64)}

        def sync(s, ss, n):
            t = torch.cat([s, ss, torch.tensor([float(n)])]).double()
            dist.all_reduce(t)
            return t[:8].float(), t[8:16].float(), int(t[16])

        y = O.batch_norm(sd, "bn", xb.chunk(world)[rank], True, sync_stats=sync)
        ref = torch.nn.functional.batch_norm(xb, None, None, None, None, True, 0.1, 1e-5).chunk(world)[rank]
        bn_ok = torch.allclose(y, ref, atol=2e-5)
        ret[rank] = (same, par, grad_ok, bn_ok)
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        same, par, grad_ok, bn_ok = ret[rank]
        assert same, "parameters were not broadcast from rank 0"
        assert par == (True, rank, world)
        assert grad_ok, "gradient averaging does not reproduce the global-batch gradient"
        assert bn_ok, "SyncBN statistics exchange does not reproduce full-batch BN"


def test_dataset_decodes_reference_layout(tmp_path):
    """dataset.Dataset keeps the reference's file layout and decode order (dataset.py:95-144): images/<id><ext> via cv2.imread
    (BGR uint8 HWC), masks/<class>/<id><ext> stacked along the last axis; conversion to float / CHW is the device feed's job."""
    cv2 = pytest.importorskip("cv2")
    import numpy as np
    from ssunet_gan_b200 import dataset
    rng = np.random.RandomState(0)
    img_dir, mask_dir = tmp_path / "images", tmp_path / "masks"
    img_dir.mkdir()
    ids = ["a01", "b02"]
    imgs, masks = {}, {}
    for i in ids:
        imgs[i] = rng.randint(0, 256, size=(12, 10, 3)).astype(np.uint8)
        cv2.imwrite(str(img_dir / (i + ".png")), imgs[i])
        masks[i] = []
        for c in range(3):
            (mask_dir / str(c)).mkdir(parents=True, exist_ok=True)
            m = rng.choice(np.array([0, 255], dtype=np.uint8), size=(12, 10))
            cv2.imwrite(str(mask_dir / str(c) / (i + ".png")), m)
            masks[i].append(m)
    ds = dataset.Dataset(ids, str(img_dir), str(mask_dir), ".png", ".png", num_classes=3)
    assert len(ds) == 2
    ori, img, mask, extra, meta = ds[1]
    assert meta == {"img_id": "b02"} and extra == []
    assert img.dtype == np.uint8 and np.array_equal(img, imgs["b02"]) and np.array_equal(ori, imgs["b02"])
    assert mask.shape == (12, 10, 3) and all(np.array_equal(mask[..., c], masks["b02"][c]) for c in range(3))
    # a host transform (albumentations-style callable) is still honoured before the feed
    ds2 = dataset.Dataset(ids, str(img_dir), str(mask_dir), ".png", ".png", 3, transform=lambda image, mask: {"image": image[::-1], "mask": mask[::-1]})
    _, img2, mask2, _, _ = ds2[0]
    assert np.array_equal(img2, imgs["a01"][::-1]) and np.array_equal(mask2[..., 2], masks["a01"][2][::-1])
    # single-class / single-band layout (dataset.py:103-113)
    cv2.imwrite(str(mask_dir / "a01.png"), masks["a01"][0])
    ds1 = dataset.Dataset(["a01"], str(img_dir), str(mask_dir), ".png", ".png", num_classes=1, input_channels=1)
    _, g1, m1, _, _ = ds1[0]
    assert g1.shape == (12, 10, 1) and m1.shape == (12, 10, 1) and np.array_equal(m1[..., 0], masks["a01"][0])


def test_device_feed_refuses_cpu():
    import torch
    from ssunet_gan_b200 import dataset
    from ssunet_gan_b200._lib import SsgError
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(SsgError):
        dataset.DeviceFeed()


def test_cvresize_header_and_product_tables_match_cv2(tmp_path, golden_dir):
    """The per-pixel fixed-point function the CUDA resize / vote kernels call (csrc/cvresize.h), compiled for the host and fed
    with the PRODUCT's coefficient tables, reproduces cv2.resize on the fixtures written by cv2 itself."""
    import subprocess
    import numpy as np
    from ssunet_gan_b200 import aerial_image_segmentation_api as api
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "cvresize_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(root, "tests", "cvresize_host.cpp")], check=True)
    L = ctypes.CDLL(so)
    L.host_resize_u8.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
    z = np.load(os.path.join(golden_dir, "tiles_resize_bridge.npz"))
    for tag in ("r_half", "r_up", "r_down", "r_quarter"):
        src, want = np.ascontiguousarray(z[tag + "_src"]), z[tag + "_dst"]
        h, w, c = src.shape
        oh, ow = want.shape[:2]
        xt = api._linear_table(w, ow, "x", "cpu").numpy()
        yt = api._linear_table(h, oh, "y", "cpu").numpy()
        assert xt.dtype == np.int32 and xt.shape == (ow, 4) and bool((xt[:, 2] + xt[:, 3] == 2048).all())
        dst = np.empty((oh, ow, c), dtype=np.uint8)
        L.host_resize_u8(src.ctypes.data, dst.ctypes.data, 1, h, w, c, oh, ow, xt.ctypes.data, yt.ctypes.data)
        assert np.array_equal(dst, want), tag


def test_grad_sink_protocol():
    """ops.GradSink: the first member writes, later members accumulate, only the LAST one reports a gradient; the sink resets
    itself afterwards (a second backward through the same graph starts clean)."""
    from ssunet_gan_b200 import ops
    log = []
    sink = ops.GradSink(expected=3)

    def member(tag, val):
        return sink.contribute(lambda: log.append(("write", tag)) or [val], lambda buf: (log.append(("acc", tag)), buf.__setitem__(0, buf[0] + val)))

    for rnd in range(2):
        assert member("a", 1) is None
        assert member("b", 10) is None
        out = member("c", 100)
        assert out == [111] and sink.buf is None and sink.arrived == 0
    assert log == [("write", "a"), ("acc", "b"), ("acc", "c")] * 2
    # no sink without the tensor-core path / without gradients
    import torch
    assert ops.grad_sink_for(torch.zeros(2, requires_grad=False)) is None
    ops.set_grad_sink(False)
    try:
        assert ops.grad_sink_for(torch.zeros(2, requires_grad=True)) is None
    finally:
        ops.set_grad_sink(True)


def test_cat_pair_is_a_tuple_of_two_sources():
    import torch
    from ssunet_gan_b200 import ops
    a, b = torch.zeros(1, 64, 2, 2, requires_grad=True), torch.zeros(1, 8, 2, 2)
    pair = ops.CatPair((a, b))
    assert pair[0] is a and pair[1] is b and len(pair) == 2 and pair.requires_grad
    assert not ops.CatPair((b, b)).requires_grad
    # on the CPU (no tensor-core path) concat_channels refuses rather than silently materialising with torch
    with pytest.raises(ops._lib.SsgError):
        ops.concat_channels(a, b, virtual=True)


def test_reference_arm_reproduces_the_survey_scalars():
    """bench.py's CPU arm drives the reference's OWN modules (baseline/_ref, staged by baseline/make_ref.py) through the literal
    loop body; on 2 x 3 x 128 x 128 / seed 1234 it must give the six scalars SURVEY.md §8c recorded from the reference."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ssg_ref_step", os.path.join(root, "baseline", "ref_step.py"))
    ref_step = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_step)
    if not ref_step.available():
        pytest.skip("baseline/_ref is not staged (run python baseline/make_ref.py where /root/reference is visible)")
    import torch

    def make_batch(b, c, h, w, seed=0):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(b, c, h, w, generator=g)
        t = (torch.rand(b, 3, h, w, generator=g) > 0.5).float()
        return x, t

    times, first, init = ref_step.timed_steps(2, 128, steps=1, make_batch=make_batch)
    want = {"loss": 0.9099170, "content": 1.5091802, "adv_g": 0.6497294, "adv_d": 1.3694998, "iou": 0.32644978, "dice": 0.4970036}
    for k, v in want.items():
        assert abs(first[k] - v) < 2e-6 * max(1.0, abs(v)), (k, first[k], v)
    assert len(init[0]) == 269 and len(times) == 1


def test_eval_bn_fold_math_on_cpu():
    """archs._fold_eval_bn (inference: bn1 folded into conv1, archs.py:229-231 with running statistics) is plain tensor algebra:
    conv(x, W * s) + (beta - mean * s) == bn_eval(conv(x, W)); and the cache follows in-place weight changes."""
    import torch.nn.functional as F
    from ssunet_gan_b200 import archs
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(5, 7, 3, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(7)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    bn.eval()
    x = torch.randn(2, 5, 9, 11)
    w, b = archs._fold_eval_bn(conv, bn)
    with torch.no_grad():
        assert torch.allclose(F.conv2d(x, w, b, 1, 1), bn(conv(x)), atol=1e-5)
        assert archs._fold_eval_bn(conv, bn)[0] is w                      # cached
        conv.weight.mul_(2.0)                                             # version counter moves: recomputed
        w2, b2 = archs._fold_eval_bn(conv, bn)
        assert w2 is not w and torch.allclose(F.conv2d(x, w2, b2, 1, 1), bn(conv(x)), atol=1e-5)


def test_stride2_dgrad_parity_class_table_on_cpu():
    """The tap -> (output-parity class, dy offset) table of csrc/conv_tc_halo.cu's stride-2 data-gradient kernel
    (`s2d_class`, `s2d_aoff`, `s2d_first`), restated in numpy and checked against autograd of F.conv2d(stride 2, pad 1):
    dx[2i + py, 2j + px] = sum over the taps of class (py, px) of dy[i + (r == 0), j + (s == 0)] . W[r][s]  (models_seg_gan.py:38-39)."""
    import torch.nn.functional as F
    torch.manual_seed(7)
    n, cin, cout, h, w = 2, 5, 4, 12, 8
    x = torch.randn(n, cin, h, w, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(cout, cin, 3, 3, dtype=torch.float64)
    y = F.conv2d(x, wt, None, 2, 1)
    dy = torch.randn_like(y)
    y.backward(dy)
    oh, ow = h // 2, w // 2
    dyp = np.zeros((n, cout, oh + 1, ow + 1))                       # the halo box: one extra row / column, zero outside the image
    dyp[:, :, :oh, :ow] = dy.numpy()
    wn = wt.numpy()
    dx = np.zeros((n, cin, h, w))
    seen = {}
    for t in range(9):
        r, s = divmod(t, 3)
        cls = (2 if r != 1 else 0) + (1 if s != 1 else 0)           # s2d_class
        di, dj = int(r == 0), int(s == 0)                           # s2d_aoff = (di * 9 + dj) pixels of the 17 x 9 halo box
        first = t in (0, 1, 3, 4)                                   # s2d_first
        assert first == (cls not in seen)
        seen[cls] = True
        py, px = cls >> 1, cls & 1
        contrib = np.einsum("nkij,kc->ncij", dyp[:, :, di:di + oh, dj:dj + ow], wn[:, :, r, s])
        dx[:, :, py::2, px::2] += contrib
    assert sorted(seen) == [0, 1, 2, 3]
    assert np.allclose(dx, x.grad.numpy(), atol=1e-10)
