"""GPU half of SURVEY 8f.2. The reference tree does not travel to the GPU box, so main.py itself cannot run there
(tests/test_harness.py runs it, unchanged, in the build container). This test walks the same path with the same
harness pieces on the fused models: torch.hub.load through the hub shim (models/vision_all.py:156), reset_parameters
walk (:157-158,322-329), head attach (:168-174) / separate lineareval head (main.py:184-201), the synthetic STL10
behind the reference's transform chain (utils_datasets.py:554-583: Resize bicubic, RandomCrop(pad, fill=128), flip,
ToTensor, Normalize), torch SGD momentum 0.9 + LambdaLR step schedule (utils_network.py:119-126,218-225,531-545) and
the loop body of Network.run_one_epoch (utils_network.py:406-452) incl. classification_count_correct (:85-95)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _loaders(image_size, bs, n_train, n_test):
    import torchvision.transforms as T
    from torchvision.transforms import InterpolationMode
    from vit_torch_b200.compat import harness
    norm = dict(mean=[0.44671062065972217, 0.43980983983523964, 0.40664644709967324],
                std=[0.2603409782662331, 0.25657727311344447, 0.27126738145225493])     # utils_datasets.py:586-589
    tr = [T.Resize(image_size, InterpolationMode.BICUBIC), T.RandomCrop(image_size, padding=max(2, image_size // 12), fill=128),
          T.RandomHorizontalFlip(), T.ToTensor(), T.Normalize(**norm)]
    te = [T.Resize(image_size, InterpolationMode.BICUBIC), T.ToTensor(), T.Normalize(**norm)]
    train = torch.utils.data.Subset(harness.SyntheticSTL10(split="train", transform=T.Compose(tr)), torch.arange(n_train))
    test = torch.utils.data.Subset(harness.SyntheticSTL10(split="test", transform=T.Compose(te)), torch.arange(n_test))
    return (torch.utils.data.DataLoader(train, batch_size=bs, shuffle=True, num_workers=2),
            torch.utils.data.DataLoader(test, batch_size=bs, shuffle=False, num_workers=2))


def _run_one_epoch(model, frozen, loader, opt, loss_fn, training, device="cuda"):
    losses, corrects = [], []
    for inputs, labels in loader:
        inputs, labels = inputs.to(device), labels.to(device)
        x = inputs
        with torch.no_grad():
            for m in frozen:
                x = m(x)
        if training:
            x = model(x)
        else:
            with torch.no_grad():
                x = model(x)
        loss = loss_fn(x, labels)
        if training:
            opt.zero_grad()
            loss.backward()
            opt.step()
        with torch.no_grad():
            corrects.extend(np.array((torch.argmax(x, dim=-1) == labels).cpu().numpy()).reshape(-1).tolist())
        losses.append(float(loss.item()))
    return sum(losses) / len(losses), float(np.mean(corrects))


@pytest.mark.parametrize("mode", ["finetune", "lineareval"])
def test_main_flow_on_fused_models(tmp_path, mode):
    from vit_torch_b200 import compat, models, train, zoo
    home = str(tmp_path)
    compat.install_hub_shim(home)
    torch.hub.set_dir(home + "/hub")
    try:
        torch.manual_seed(0)
        backbone = torch.hub.load("facebookresearch/dino:main", "dino_vits16", pretrained=False)
        assert isinstance(backbone, models.DinoVisionTransformer)
        train.reset_parameters_like_zoo(backbone)
        fc = [256, 128, 32, 10]
        if mode == "finetune":          # DINO forward never applies .head: opt in, as CaiT / DeiT do
            backbone.head = zoo.get_classifier_head(384, fc)
            backbone.apply_head = True
            model, frozen = backbone, []
        else:
            out_dim = backbone.cuda()(torch.rand(1, 3, 96, 96).cuda()).shape[-1]       # get_output_shape, main.py:193-194
            model, frozen = zoo.get_classifier_head(out_dim, fc), [backbone]
        model.cuda()
        for m in frozen:
            m.cuda()
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda e: 0.5 ** np.floor(e / 10))
        tr_loader, te_loader = _loaders(96, 16, 96, 32)
        hist = []
        for _ in range(3):
            tl, ta = _run_one_epoch(model, frozen, tr_loader, opt, nn.CrossEntropyLoss(), True)
            sched.step()
            vl, va = _run_one_epoch(model, frozen, te_loader, opt, nn.CrossEntropyLoss(), False)
            hist.append((tl, ta, vl, va))
        print(mode, hist)
        assert all(np.isfinite(h[0]) and np.isfinite(h[2]) for h in hist)
        assert hist[-1][0] < hist[0][0]                       # the training loss falls on the class-conditional data
        if mode == "lineareval":
            assert all(p.grad is None for p in backbone.parameters())
    finally:
        torch.hub._hub_dir = None
