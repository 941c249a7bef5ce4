"""RGB <-> LAB pinned by the reference's own pixels.

``example/Sat2Aerx1G2RGB/*.png`` and ``example/Sat2Aerx1G2LAB/*.png`` show the same target tiles, the second after
``Basic._arr2lab`` -> float32 tensor -> ``Basic._lab2img`` (src/dataset.py:94-104,148-159).  tests/golden/lab_pin.npz holds the
centre 160x160 crops (oracle/make_lab_pin.py).  The truncating uint8 conversion makes the pair sensitive to the last bits of
the round trip: half of the values come back as k-1."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lab_pin.npz")
NAMES = ("test_0", "train_0", "val_0")


@pytest.fixture(scope="module")
def pin():
    return np.load(GOLDEN)


def test_fixture_is_the_known_answer_the_survey_describes(pin):
    for n in NAMES:
        d = pin[n + "_tar_rgb"].astype(int) - pin[n + "_tar_lab"].astype(int)
        assert set(np.unique(d)) == {0, 1}                       # tar_LAB = tar_RGB - {0, 1}
        assert 0.3 < (d == 1).mean() < 0.7


def test_oracle_reproduces_the_reference_tiles_bit_for_bit(pin):
    from oracle import srcgan_oracle as O
    for n in NAMES:
        rgb, lab = pin[n + "_tar_rgb"], pin[n + "_tar_lab"]
        t = O.arr2lab(rgb)
        assert t.dtype == np.float32 and t.shape == (3,) + rgb.shape[:2]
        assert 0.0 <= t[0].min() and t[0].max() <= 1.0 and 0.0 < t[1:].min() and t[1:].max() < 1.0
        assert np.array_equal(O.lab2img(t), lab), n


@pytest.mark.gpu
def test_device_lab_glue_reproduces_the_reference_tiles_bit_for_bit(pin):
    """srcgan_rgb2lab_u8 / srcgan_lab2rgb_u8 (float64 arithmetic on the device) against the reference's pixels and the
    oracle's tensors; the fp32 tensor<->tensor kernels against the same tiles within the truncation step."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import color
    imgs = np.stack([pin[n + "_tar_rgb"] for n in NAMES])
    want = np.stack([pin[n + "_tar_lab"] for n in NAMES])
    dev = torch.from_numpy(imgs).cuda()
    lab = color.image_to_lab(dev)
    ref_lab = np.stack([O.arr2lab(x) for x in imgs])
    assert lab.dtype == torch.float32 and tuple(lab.shape) == (3, 3, 160, 160)
    # float64 pow / cbrt on the device vs libm: the float32-rounded tensors may differ in the last place only
    got_lab = lab.cpu().numpy()
    assert np.abs(got_lab - ref_lab).max() <= 1.2e-7, np.abs(got_lab - ref_lab).max()
    assert (got_lab == ref_lab).mean() > 0.999, (got_lab == ref_lab).mean()
    back = color.lab_to_image(torch.from_numpy(ref_lab).cuda()).cpu().numpy()
    assert back.shape == want.shape and back.dtype == np.uint8
    # the reference's G2LAB tiles, bit for bit (float64 pow on the device vs libm can differ in the last place, which the
    # truncation exposes only when v*255 is within ~1e-13 of an integer)
    assert (back == want).mean() >= 0.99999, ((back == want).mean(), np.abs(back.astype(int) - want.astype(int)).max())
    full = color.lab_to_image(lab).cpu().numpy()                           # whole round trip on the device
    assert (full == want).mean() > 0.999 and np.abs(full.astype(int) - want.astype(int)).max() <= 1, (full == want).mean()
    assert np.array_equal(color.tensor2img(torch.from_numpy(ref_lab[:1]).cuda(), mode="LAB").cpu().numpy(),
                          want[0].transpose(2, 0, 1))
    # the fp32 tensor kernels (training-side feed): same transform to float32 accuracy; after the truncation k or k-1
    x = torch.from_numpy(imgs).cuda().permute(0, 3, 1, 2).float() / 255.0
    lab32 = color.rgb2lab(x, True)
    assert np.abs(lab32.cpu().numpy() - ref_lab).max() < 2e-5
    rgb32 = (color.lab2rgb(lab32, True) * 255.0)
    assert np.abs(rgb32.cpu().numpy().transpose(0, 2, 3, 1) - imgs.astype(np.float32)).max() < 2e-2
