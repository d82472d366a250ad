"""Oracle restatement of the reference's mask clean-up + RLE export (TEST INFRASTRUCTURE).

Follows /root/reference/nn_inference.py:
  * ``rle_decode``        :237-251
  * ``rle_encoding``      :253-263  (column-major, 1-based (start, length) pairs)
  * ``postprocess_masks`` :265-306  (fill holes, dilate + erode, remove overlaps in list order,
                                     drop masks that fall into more than one piece)
  * the export loop       :319-336  (one ``ImageId, EncodedPixels`` row per returned mask)

Third-party pieces the reference calls and this image does not have (scikit-image, unpinned in
the reference) are restated through the SciPy routines scikit-image itself wraps:
  * ``skimage.morphology.dilation / erosion`` with the default footprint
    (``ndi.generate_binary_structure(2, 1)``, the 3 x 3 cross) are
    ``ndi.grey_dilation / grey_erosion(image, footprint=...)`` with the default ``mode='reflect'``
    (a reflected border repeats the edge pixel, so neighbours outside the image never change
    a max / min that already contains the pixel itself);
  * ``skimage.measure.label`` (default connectivity = ndim, i.e. 8-connected in 2-D) has as many
    labels as ``ndi.label`` with the full 3 x 3 structure.
``scipy.ndimage.binary_fill_holes`` is called directly, as the reference does.
Parity unpinned by the reference's own tests (it has none); KATs in tests/test_oracle_cleanup.py.

The quirks of the reference are kept (they are what its output is):
  * ``ori_score.all() < score_threshold`` compares a bool with 0.5: the image is skipped iff
    some score is exactly zero (:274);
  * ``np.sum(ori_mask, axis=(0, 1))`` on the N x H x W array sums over instances and rows and
    leaves a W-vector; ``keep_ind`` therefore counts image COLUMNS holding more than
    ``min_crys_size`` mask pixels, and when there are fewer such columns than instances the
    instance list is truncated to that many entries (:277-284);
  * ``overlap`` accumulates the cleaned masks before their own overlaps are cut (:297-298).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.ndimage as ndi

_CROSS = ndi.generate_binary_structure(2, 1)
_FULL = np.ones((3, 3), dtype=bool)


def rle_decode(mask_rle: str, shape: Tuple[int, int]) -> np.ndarray:
    """nn_inference.py:237-251."""
    s = mask_rle.split()
    starts, lengths = [np.asarray(x, dtype=int) for x in (s[0:][::2], s[1:][::2])]
    starts -= 1
    ends = starts + lengths
    img = np.zeros(shape[0] * shape[1], dtype=np.uint8)
    for lo, hi in zip(starts, ends):
        img[lo:hi] = 1
    return img.reshape(shape)


def rle_encoding(x: np.ndarray) -> List[int]:
    """nn_inference.py:253-263 (vectorised: same list)."""
    dots = np.where(x.T.flatten() == 1)[0]
    if dots.size == 0:
        return []
    brk = np.flatnonzero(np.diff(dots) > 1)
    starts = np.concatenate(([dots[0]], dots[brk + 1]))
    ends = np.concatenate((dots[brk], [dots[-1]]))
    out = np.empty(2 * starts.size, dtype=np.int64)
    out[0::2] = starts + 1
    out[1::2] = ends - starts + 1
    return out.tolist()


def rle_encoding_literal(x: np.ndarray) -> List[int]:
    """The loop exactly as written (:253-263); used to pin the vectorised form."""
    dots = np.where(x.T.flatten() == 1)[0]
    run_lengths: List[int] = []
    prev = -2
    for b in dots:
        if b > prev + 1:
            run_lengths.extend((int(b) + 1, 0))
        run_lengths[-1] += 1
        prev = b
    return run_lengths


def dilation(mask: np.ndarray) -> np.ndarray:
    return ndi.grey_dilation(mask, footprint=_CROSS)


def erosion(mask: np.ndarray) -> np.ndarray:
    return ndi.grey_erosion(mask, footprint=_CROSS)


def label_count(mask: np.ndarray) -> int:
    return int(ndi.label(mask, structure=_FULL)[1])


def postprocess_masks(ori_mask: np.ndarray, ori_score: np.ndarray, image_hw: Tuple[int, int],
                      min_crys_size: int = 2) -> Optional[List[np.ndarray]]:
    """nn_inference.py:265-306.  ``ori_mask`` N x H x W bool, ``ori_score`` N float."""
    height, width = image_hw
    score_threshold = 0.5
    if len(ori_mask) == 0 or ori_score.all() < score_threshold:
        return None
    keep_ind = np.where(np.sum(ori_mask, axis=(0, 1)) > min_crys_size)[0]
    if len(keep_ind) < len(ori_mask):
        if keep_ind.shape[0] != 0:
            ori_mask = ori_mask[:keep_ind.shape[0]]
            ori_score = ori_score[:keep_ind.shape[0]]
        else:
            ori_mask = []
            ori_score = []
    overlap = np.zeros([height, width])
    masks = []
    for i in range(len(ori_mask)):
        mask = ndi.binary_fill_holes(ori_mask[i]).astype(np.uint8)
        mask = erosion(dilation(mask))
        overlap += mask
        mask[overlap > 1] = 0
        if label_count(mask) > 1:
            mask[()] = 0
        masks.append(mask)
    return masks


def export_rows(names: Sequence[str], masks_per_image: Sequence[np.ndarray],
                scores_per_image: Sequence[np.ndarray], image_hw: Tuple[int, int]):
    """nn_inference.py:319-332: (ImageId, EncodedPixels) rows of a folder of images."""
    conv = lambda l: ' '.join(map(str, l))          # noqa: E731  (:317)
    img_id, encoded = [], []
    for name, m, s in zip(names, masks_per_image, scores_per_image):
        masks = postprocess_masks(np.asarray(m), np.asarray(s), image_hw)
        if masks:
            for i in range(len(masks)):
                img_id.append(name.replace('.tif', ''))
                encoded.append(conv(rle_encoding(masks[i])))
    return img_id, encoded
