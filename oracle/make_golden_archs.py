"""Generate tests/golden/archs_*.{npz,json} by running the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_archs.py

SURVEY.md §8f rows 2 and 4: the other seven networks of `archs.__all__` (+ ProgUNet, NestedUNet under deep
supervision), the supervised trainer's loop body (train.py:81-120) and the validation loop body (train.py:152-176).
The reference modules are imported from /root/reference/scripts, loaded with the portable weights of
oracle/ssunet_oracle.py and driven on the portable synthetic batches; only outputs go into the fixtures.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/scripts")
warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True

import ssunet_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(8)

CASES = [("UNet", False, 32), ("NestedUNet", False, 32), ("NestedUNet", True, 32), ("SSUNet", False, 32), ("UNet_ori", False, 32),
         ("UNet_B_SS", False, 32), ("AttUNet", False, 32), ("UNet_R_SS", False, 64), ("ProgUNet", False, 32)]


def csum(t):
    t = t.detach().double()
    return np.array([float(t.sum()), float(t.abs().sum()), float((t * t).sum())])


def case_tag(name, ds):
    return name + ("_ds" if ds else "")


def main():
    import archs, losses, metrics  # noqa
    crit = losses.BCEDiceLoss()
    layout = {}
    for idx, (name, ds, hw) in enumerate(CASES):
        tag = case_tag(name, ds)
        torch.manual_seed(41)
        net = archs.__dict__[name](3, 3, ds)
        sd0 = net.state_dict()
        layout[tag] = [[k, list(v.shape)] for k, v in sd0.items()]
        layout[tag + ":init_seed41"] = {k: csum(v).tolist() for k, v in sd0.items() if v.is_floating_point()}
        spec = [(k, tuple(v.shape)) for k, v in sd0.items()]
        rec = {}
        # ---- train-mode forward + BCEDice (averaged under deep supervision, train.py:87-92) + backward
        net.load_state_dict(O.portable_state_dict(spec, salt=idx + 1))
        net.train()
        x, t = O.synthetic_batch(2, 3, hw, hw, seed=1234)
        out = net(x)
        outs = out if isinstance(out, list) else [out]
        if name == "ProgUNet":      # heads at four resolutions (not a train.py configuration): any scalar that reaches every head
            loss = crit(outs[0], t) + sum(o.square().mean() for o in outs[1:])
        else:
            loss = sum(crit(o, t) for o in outs) / len(outs)
        loss.backward()
        for i, o in enumerate(outs):
            rec["logits%d" % i] = o.detach().numpy()
        rec["loss"] = np.float64(loss.item())
        gr = {k: csum(p.grad) for k, p in net.named_parameters() if p.grad is not None}
        rec["grad_keys"] = np.array(list(gr.keys()))
        rec["grad_csum"] = np.stack(list(gr.values()))
        rec["nograd_keys"] = np.array([k for k, p in net.named_parameters() if p.grad is None])
        sd1 = net.state_dict()
        bn_keys = [k for k in sd1 if k.endswith("running_var")]
        rec["bn_keys"] = np.array(bn_keys)
        rec["bn_running_var_csum"] = np.stack([csum(sd1[k]) for k in bn_keys])
        # ---- eval-mode forward on fresh running statistics, other input
        net.load_state_dict(O.portable_state_dict(spec, salt=idx + 1))
        net.eval()
        xe, te = O.synthetic_batch(1, 3, hw, hw, seed=77, blobby=True)
        with torch.no_grad():
            oe = net(xe)
        oes = oe if isinstance(oe, list) else [oe]
        for i, o in enumerate(oes):
            rec["eval_logits%d" % i] = o.numpy()
        # validation loop body (train.py:155-175) on the eval output
        if name == "ProgUNet":
            vloss, viou, vdice = crit(oes[0], te), metrics.iou_score(oes[0], te), metrics.dice_coef(oes[0], te)
        elif ds or isinstance(oe, list):
            vloss = sum(crit(o, te) for o in oes) / len(oes)
            viou = metrics.iou_score(oes[-1], te)
            vdice = metrics.dice_coef(oes[-1], te)
        else:
            o = oe.clone()
            o[torch.isnan(o)] = 0
            vloss = crit(o, te)
            viou = metrics.iou_score(o[:, 1:3].clone(), te[:, 1:3].clone())
            vdice = metrics.dice_coef(o[:, 1:3].clone(), te[:, 1:3].clone())
        rec["val_scalars"] = np.array([float(vloss), float(viou), float(vdice)], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, "archs_%s.npz" % tag), **rec)
        print("%-14s loss %.6f  params %d  no-grad params %d" % (tag, loss.item(), sum(p.numel() for p in net.parameters()),
                                                                   len(rec["nograd_keys"])))
    with open(os.path.join(OUT, "archs_layout.json"), "w") as f:
        json.dump(layout, f)

    # ---------------- supervised trainer loop body, 2 iterations (train.py:81-120) ----------------
    # config_v1.json: Adam lr 1e-4, weight_decay 1e-7, clip 0.7; portable BN gammas (~1 +- 0.1) exceed the clip, so the
    # clamp between forward and backward is exercised.
    for name, ds, hw in (("UNet_R_SS_v2", False, 64), ("NestedUNet", True, 32)):
        tag = case_tag(name, ds)
        net = archs.__dict__[name](3, 3, ds)
        spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
        if name != "UNet_R_SS_v2":
            assert spec == [(k, tuple(s)) for k, s in layout[tag]]
        net.load_state_dict(O.portable_state_dict(spec, salt=31))
        net.train()
        params = filter(lambda p: p.requires_grad, net.parameters())
        opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-7)
        clip = 0.7
        rec = {}
        for it in range(2):
            inp, tar = O.synthetic_batch(2, 3, hw, hw, seed=4321 + it, blobby=(it == 1))
            # ---- literal restatement of train.py:85-116 driving the reference modules ----
            if ds:
                outputs = net(inp)
                loss = 0
                for output in outputs:
                    loss += crit(output, tar)
                loss /= len(outputs)
                iou = metrics.iou_score(outputs[-1], tar)
                dice = metrics.dice_coef(outputs[-1], tar)
                last = outputs[-1]
            else:
                output = net(inp)
                output[torch.isnan(output)] = 0
                out_m = output[:, 1:3, :, :].clone()
                tar_m = tar[:, 1:3, :, :].clone()
                loss = crit(output, tar)
                iou = metrics.iou_score(out_m, tar_m)
                dice = metrics.dice_coef(out_m, tar_m)
                last = output
            for p in net.parameters():
                p.data.clamp_(-clip, clip)
            opt.zero_grad()
            loss.backward()
            opt.step()
            rec["it%d_scalars" % it] = np.array([loss.item(), float(iou), float(dice)], dtype=np.float64)
            rec["it%d_logits" % it] = last.detach().numpy()
        sdn = net.state_dict()
        rec["keys"] = np.array(list(sdn.keys()))
        rec["csum"] = np.stack([csum(v) for v in sdn.values()])
        probe = "final.weight" if "final.weight" in sdn else "final4.weight"
        rec["probe_key"] = np.array(probe)
        rec["probe"] = sdn[probe].numpy()
        np.savez_compressed(os.path.join(OUT, "supervised_step_%s.npz" % tag), **rec)
        print("supervised %-14s it0 %s it1 %s" % (tag, rec["it0_scalars"], rec["it1_scalars"]))
    for fn in sorted(os.listdir(OUT)):
        if fn.startswith(("archs_", "supervised_")):
            print("  %-40s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
