"""The reference's scripts, replayed call by call against the mirror on a GPU (the GPU box has no reference checkout, so
the unchanged files cannot be exec'd there; these tests execute their call sequences with the reference's own
argument values): test.py:21-49, evaluation.py:24-64 (incl. a torchsummary-style hook pass) and the train.py:95-131 loop,
plus the multi-GPU launcher on one rank."""
import math
import os

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import synth

pytestmark = pytest.mark.gpu

DATA = ("[name]\nmodel_name=coco\n\n[train-configure]\nepochs=1\nsteps=150,250\nbatch_size=8\nsubdivisions=2\nlearning_rate=0.001\n\n"
        "[model-configure]\npre_weights=None\nclasses=80\nwidth=352\nheight=352\nanchor_num=3\n"
        "anchors=12.64,19.39, 37.88,51.48, 55.71,138.31, 126.91,78.23, 131.57,214.55, 279.92,258.87\n\n"
        "[data-configure]\ntrain=/nonexistent/train.txt\nval=/nonexistent/val.txt\nnames=%s\n")


@pytest.fixture()
def workdir(tmp_path, golden_dir):
    names = tmp_path / "coco.names"
    names.write_text("\n".join(["person", "bicycle"] + ["c%d" % i for i in range(2, 80)]) + "\n")
    data = tmp_path / "coco.data"
    data.write_text(DATA % names)
    w = dict(np.load(os.path.join(golden_dir, "modelzoo_weights.npz")))
    weights = tmp_path / "model.pth"
    torch.save({k: torch.from_numpy(v) for k, v in w.items()}, weights)
    return str(data), str(weights)


def test_test_py_sequence(workdir, golden_dir):
    """test.py:21-49 with the modelzoo weights on img/000139.jpg (stored pre-resized): person .87, bicycle .46, person .32."""
    import model.detector
    import utils.utils
    data, weights = workdir
    cfg = utils.utils.load_datafile(data)
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model_ = model.detector.Detector(cfg["classes"], cfg["anchor_num"], True).to(device)
    model_.load_state_dict(torch.load(weights, map_location=device))
    model_.eval()
    g = np.load(os.path.join(golden_dir, "images_modelzoo.npz"))
    img = torch.from_numpy(g["000139_u8"])                       # what cv2.imread + resize + transpose produce (test.py:33-37)
    img = img.to(device).float() / 255.0
    preds = model_(img)
    output = utils.utils.handel_preds(preds, cfg, device)
    output_boxes = utils.utils.non_max_suppression(output, conf_thres=0.3, iou_thres=0.4)
    LABEL_NAMES = [line.strip() for line in open(cfg["names"])]
    drawn = [(LABEL_NAMES[int(b[5])], "%.2f" % b[4]) for b in (box.tolist() for box in output_boxes[0])]
    assert drawn == [("person", "0.87"), ("bicycle", "0.46"), ("person", "0.32")]


def test_evaluation_py_sequence(workdir):
    """evaluation.py:24-64: loader with the reference's collate semantics, summary-style hooks, two evaluation passes."""
    import model.detector
    import utils.utils
    import train_dist
    data, weights = workdir
    cfg = utils.utils.load_datafile(data)
    val_dataset = train_dist.SyntheticDetection(6, cfg["width"], cfg["height"], cfg["classes"], seed=5)
    val_dataloader = torch.utils.data.DataLoader(val_dataset, batch_size=4, shuffle=False, collate_fn=train_dist.collate_fn,
                                                 num_workers=0, pin_memory=True, drop_last=False)
    device = torch.device("cuda")
    model_ = model.detector.Detector(cfg["classes"], cfg["anchor_num"], True).to(device)
    model_.load_state_dict(torch.load(weights, map_location=device))
    model_.eval()
    # torchsummary.summary(model, input_size=(3, H, W)): forward hooks on every module, one forward of a batch of 2, hooks removed
    seen, hooks = [], []
    for mod in model_.modules():
        hooks.append(mod.register_forward_hook(lambda mod_, i, o: seen.append(type(mod_).__name__)))
    model_(torch.rand(2, 3, cfg["height"], cfg["width"]).to(device))
    for h in hooks:
        h.remove()
    assert "Detector" in seen
    r1 = utils.utils.evaluation(val_dataloader, cfg, model_, device)
    r2 = utils.utils.evaluation(val_dataloader, cfg, model_, device, 0.3)
    for r in (r1, r2):
        assert r is None or (len(r) == 4 and all(0.0 <= float(v) <= 1.0 for v in r))


def test_train_py_loop_three_iterations(workdir):
    """train.py:70-131 literally (plain torch SGD, zero_grad with set_to_none, warm-up, subdivisions) for three iterations."""
    import model.detector
    import utils.loss
    import utils.utils
    import train_dist
    from torch import optim
    data, weights = workdir
    cfg = utils.utils.load_datafile(data)
    batch_size = int(cfg["batch_size"] / cfg["subdivisions"])
    ds = train_dist.SyntheticDetection(3 * batch_size, cfg["width"], cfg["height"], cfg["classes"])
    train_dataloader = torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=True, collate_fn=train_dist.collate_fn,
                                                   num_workers=0, pin_memory=True, drop_last=True)
    device = torch.device("cuda")
    model_ = model.detector.Detector(cfg["classes"], cfg["anchor_num"], True).to(device)
    model_.load_state_dict(torch.load(weights, map_location=device), strict=False)
    before = model_.output_cls_layers.weight.detach().clone()
    optimizer = optim.SGD(params=model_.parameters(), lr=cfg["learning_rate"], momentum=0.949, weight_decay=0.0005)
    scheduler = optim.lr_scheduler.MultiStepLR(optimizer, milestones=cfg["steps"], gamma=0.1)
    batch_num, totals = 0, []
    for epoch in range(cfg["epochs"]):
        model_.train()
        for imgs, targets in train_dataloader:
            imgs = imgs.to(device).float() / 255.0
            targets = targets.to(device)
            preds = model_(imgs)
            iou_loss, obj_loss, cls_loss, total_loss = utils.loss.compute_loss(preds, targets, cfg, device)
            total_loss.backward()
            for g in optimizer.param_groups:
                warmup_num = 5 * len(train_dataloader)
                if batch_num <= warmup_num:
                    scale = math.pow(batch_num / warmup_num, 4)
                    g["lr"] = cfg["learning_rate"] * scale
                lr = g["lr"]
            if batch_num % cfg["subdivisions"] == 0:
                optimizer.step()
                optimizer.zero_grad()
            info = "Epoch:%d LR:%f CIou:%f Obj:%f Cls:%f Total:%f" % (epoch, lr, float(iou_loss.detach()), float(obj_loss.detach()), float(cls_loss.detach()), float(total_loss.detach()))
            assert "nan" not in info
            totals.append(float(total_loss))
            batch_num += 1
        scheduler.step()
    assert batch_num == 3 and all(np.isfinite(totals))
    assert not torch.equal(before, model_.output_cls_layers.weight.detach())
    model_.eval()                                                # the epoch-end evaluation forward sees the updated statistics
    assert all(torch.isfinite(p).all() for p in model_(torch.rand(1, 3, 352, 352).to(device)))


def test_launcher_single_rank(workdir, monkeypatch):
    import train_dist
    data, _ = workdir
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    net = train_dist.main(["--data", data, "--synthetic", "16", "--max-iters", "4"])
    assert all(torch.isfinite(p).all() for p in net.parameters())
