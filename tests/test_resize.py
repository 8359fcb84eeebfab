"""Antialiased bicubic resize: the coefficient tables and the two fixed-point passes restated in
bridgelang_b200/resize.py must reproduce PIL.Image.resize — the routine torchvision's Resize reaches inside the
reference's image transform (dinosiglip_vit.py:91-111, processing_prismatic.py:128-145) — BIT FOR BIT.  The GPU kernel
is then compared with the same NumPy statement and with PIL directly (-m gpu)."""

import numpy as np
import pytest
import torch
from PIL import Image

from bridgelang_b200.resize import resample_coeffs, resize_u8_reference

SIZES = [(256, 256), (480, 640), (224, 224), (100, 300), (512, 384), (720, 1280), (50, 40), (225, 223)]


@pytest.mark.parametrize("hw", SIZES)
def test_tables_and_passes_match_pil_bit_for_bit(hw):
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    img = rng.integers(0, 256, (*hw, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((224, 224), Image.BICUBIC))
    got = resize_u8_reference(img[None], (224, 224))[0]
    assert np.array_equal(got, want)


def test_non_square_targets_and_bilinear():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (300, 500, 3), dtype=np.uint8)
    for (hd, wd) in ((224, 373), (384, 384), (336, 336), (150, 250)):
        want = np.asarray(Image.fromarray(img).resize((wd, hd), Image.BICUBIC))
        assert np.array_equal(resize_u8_reference(img[None], (hd, wd))[0], want)
    want = np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
    assert np.array_equal(resize_u8_reference(img[None], (224, 224), "bilinear")[0], want)


def test_coefficient_rows_are_normalised():
    kk, bounds, ksize = resample_coeffs(640, 224)
    assert kk.shape == (224, ksize) and bounds.shape == (224, 2)
    assert np.all(np.abs(kk.sum(axis=1) - (1 << 22)) <= ksize)        # rounding of each tap only
    assert np.all(bounds[:, 0] >= 0) and np.all(bounds[:, 0] + bounds[:, 1] <= 640)


@pytest.mark.gpu
@pytest.mark.parametrize("hw", SIZES)
def test_device_resize_matches_pil_bit_for_bit(hw):
    from bridgelang_b200 import ops
    rng = np.random.default_rng(hw[0] + 13 * hw[1])
    imgs = rng.integers(0, 256, (3, *hw, 3), dtype=np.uint8)
    got = ops.resize_u8(torch.from_numpy(imgs).cuda(), (224, 224)).cpu().numpy()
    for i in range(3):
        want = np.asarray(Image.fromarray(imgs[i]).resize((224, 224), Image.BICUBIC))
        assert np.array_equal(got[i], want), (hw, i)


@pytest.mark.gpu
def test_processor_device_paths_are_bit_identical_to_the_host_transform():
    """PrismaticImageProcessor: raw frame → device resize → LUT normalise == the host transform's pixel_values in bf16,
    for all three resize strategies (letterbox pads on the host; resize-crop crops on the device)."""
    import bridgelang_b200 as blb
    from bridgelang_b200.weights import DINO_MEAN, DINO_STD, SIGLIP_MEAN, SIGLIP_STD
    rng = np.random.default_rng(5)
    images = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for (h, w) in ((256, 256), (480, 640), (300, 200))]
    for strategy in ("resize-naive", "letterbox", "resize-crop"):
        proc = blb.PrismaticImageProcessor(True, strategy, [(3, 224, 224)] * 2, ["bicubic"] * 2,
                                           [DINO_MEAN, SIGLIP_MEAN], [DINO_STD, SIGLIP_STD])
        want = proc.preprocess(images, return_tensors="pt")["pixel_values"].to(torch.bfloat16)
        for device_resize in (False, True):
            got = proc.preprocess_to_device(images, device_resize=device_resize)
            assert got.shape == (3, 6, 224, 224) and torch.equal(got.cpu(), want), (strategy, device_resize)
        fr = proc.frames_to_device(images)
        assert fr.dtype == torch.uint8 and fr.shape == (3, 224, 224, 3)
