"""The plain-C restatement (oracle/scalar.c) equals torch grid_sample / cv2.moments."""
import cv2
import numpy as np
import torch

from oracle import d2, scalar_c


def test_c_paste_equals_grid_sample_and_moments():
    rng = np.random.default_rng(2)
    torch.manual_seed(2)
    H, W = 120, 150
    checked = 0
    for trial in range(30):
        m = torch.rand(28, 28)
        if trial % 3 == 1:
            m = torch.full((28, 28), 0.5)
        if trial % 5 == 2:
            m = (m > 0.4).float()
        x0, y0 = rng.random(2) * 120 - 10
        w, h = np.exp(rng.random(2) * 6 - 2)
        b = d2.Boxes(torch.tensor([[x0, y0, x0 + w, y0 + h]], dtype=torch.float32))
        b.clip((H, W))
        if not bool(b.nonempty()[0]):
            continue
        ref = d2.paste_masks_in_image(m[None], b.tensor, (H, W))[0].numpy()
        out = scalar_c.paste_window(m.numpy(), b.tensor[0].numpy(), 0, W, 0, H)
        assert np.array_equal(out.astype(bool), ref)
        mm = cv2.moments(ref.astype(np.uint8), binaryImage=True)
        raw = scalar_c.raw_moments(out)
        for k, name in enumerate(("m00", "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03")):
            assert raw[k] == int(round(mm[name]))
        checked += 1
    assert checked > 20
