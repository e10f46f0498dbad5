"""Input pipeline (reference vtab.py:79-82): Resize((224,224), interpolation=3) -> ToTensor ->
Normalize.  The oracle for this row is Pillow + torchvision themselves (both in the image): the host-built tap
tables and the numpy restatement are pinned against Pillow on the CPU; the CUDA kernels are compared bit-for-bit
with the reference's transform pipeline on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image_classification"))

SIZES = [(32, 32), (375, 500), (224, 300), (224, 224), (600, 400), (97, 1031), (500, 224), (8, 8), (225, 223)]


def _image(h, w, seed):
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = (127.5 + 127.5 * np.sin(xx / 7.0 + seed)[..., None] * np.cos(yy / 11.0)[..., None] * np.ones(3)).astype(np.uint8)
    return np.where(rng.random((h, w, 1)) < 0.5, noise, smooth).astype(np.uint8)   # edges, saturation, smooth ramps


def _reference_transform():
    from torchvision import transforms
    return transforms.Compose([
        transforms.Resize((224, 224), interpolation=transforms.InterpolationMode.BICUBIC),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])


@pytest.mark.parametrize("h,w", SIZES)
def test_tap_tables_and_restatement_match_pillow(h, w):
    from PIL import Image
    from cara_b200.preprocess import resample_tables, resize_reference
    img = _image(h, w, h * 1000 + w)
    ref = np.asarray(Image.fromarray(img).resize((224, 224), Image.BICUBIC))
    assert np.array_equal(resize_reference(img), ref)
    b, k, ks = resample_tables(w, 224)
    assert b.shape == (224, 2) and k.shape == (224, ks) and int(b[:, 0].min()) >= 0 and int((b[:, 0] + b[:, 1]).max()) <= w
    assert np.all(np.abs(k.sum(1) - (1 << 22)) <= ks)          # every output pixel's weights sum to one (fixed point)


def test_decode_only_loader_and_cli_flag(tmp_path, monkeypatch):
    """get_data(gpu_preprocess=True) on a real file list yields decoded uint8 arrays of their own sizes."""
    from PIL import Image
    import vtab
    root = tmp_path / "data" / "vtab-1k" / "cifar"
    os.makedirs(root / "images")
    lines = []
    for i, (h, w) in enumerate([(32, 32), (40, 24), (32, 32), (17, 50)]):
        Image.fromarray(_image(h, w, i)).save(root / "images" / ("%d.png" % i))
        lines.append("images/%d.png %d" % (i, i % 3))
    for name in ("train800val200.txt", "test.txt"):
        (root / name).write_text("\n".join(lines) + "\n")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(torch.utils.data, "DataLoader",
                        lambda ds, **kw: torch.utils.data.dataloader.DataLoader(ds, **{**kw, "num_workers": 0, "pin_memory": False}))
    _, val = vtab.get_data("cifar", evaluate=True, batch_size=2, gpu_preprocess=True)
    x, y = next(iter(val))
    assert isinstance(x, list) and [a.shape for a in x] == [(32, 32, 3), (40, 24, 3), (32, 32, 3), (17, 50, 3)]
    assert x[0].dtype == np.uint8 and y.tolist() == [0, 1, 2, 0]
    _, val = vtab.get_data("cifar", evaluate=True, batch_size=2, gpu_preprocess=False)
    x2, _ = next(iter(val))
    assert x2.shape == (4, 3, 224, 224) and x2.dtype == torch.float32
    tf = _reference_transform()
    assert torch.equal(x2[1], tf(Image.fromarray(x[1])))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", SIZES)
def test_gpu_resize_normalize_is_bit_identical(h, w):
    from PIL import Image
    from cara_b200.preprocess import GpuPreprocessor
    tf = _reference_transform()
    imgs = np.stack([_image(h, w, 7 * i + h + w) for i in range(3)])
    ref = torch.stack([tf(Image.fromarray(im)) for im in imgs])
    ref_u8 = np.stack([np.asarray(Image.fromarray(im).resize((224, 224), Image.BICUBIC)) for im in imgs])
    out, u8 = GpuPreprocessor("cuda:0")(imgs, return_uint8=True)
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.gpu
def test_gpu_preprocess_mixed_sizes_and_model_input():
    """A decoded batch of mixed sizes through vtab.to_device equals the reference's CPU batch."""
    from PIL import Image
    import vtab
    tf = _reference_transform()
    imgs = [_image(h, w, i) for i, (h, w) in enumerate([(32, 32), (375, 500), (32, 32), (300, 224), (224, 224)])]
    x, y = vtab.to_device(imgs, torch.arange(5))
    assert x.shape == (5, 3, 224, 224) and x.is_cuda and y.is_cuda
    assert torch.equal(x.cpu(), torch.stack([tf(Image.fromarray(im)) for im in imgs]))
