"""CPU tests of the clean-up / RLE oracle (oracle/cleanup.py) against hand-checkable cases and
the loops exactly as the reference writes them (nn_inference.py:237-306)."""
import numpy as np
import pytest

from oracle import cleanup as C


def test_rle_encoding_matches_the_literal_loop_and_decodes():
    rng = np.random.default_rng(0)
    for shape in ((7, 5), (16, 16), (3, 40)):
        for p in (0.1, 0.5, 0.9):
            x = (rng.random(shape) < p).astype(np.uint8)
            enc = C.rle_encoding(x)
            assert enc == C.rle_encoding_literal(x)
            # rle_decode fills row-major: decoding into the transposed shape undoes the .T of the encoder
            dec = C.rle_decode(' '.join(map(str, enc)), (shape[1], shape[0])).T
            assert np.array_equal(dec, x)
    assert C.rle_encoding(np.zeros((4, 4), np.uint8)) == []
    # KAT: 4 x 3 image, column 0 rows 1..3 and column 1 rows 0..1 are ONE run (consecutive flat indices)
    x = np.zeros((4, 3), np.uint8)
    x[1:, 0] = 1
    x[:2, 1] = 1
    x[3, 2] = 1
    assert C.rle_encoding(x) == [2, 5, 12, 1]


def test_postprocess_masks_kats():
    H, W = 12, 14
    ring = np.zeros((H, W), bool)
    ring[2:8, 2:8] = True
    ring[4:6, 4:6] = False                       # a hole: filled
    sq = np.zeros((H, W), bool)
    sq[5:10, 6:12] = True                        # overlaps the ring's square: loses the overlap
    scores = np.array([0.9, 0.8])
    out = C.postprocess_masks(np.stack([ring, sq]), scores, (H, W))
    assert out[0][2:8, 2:8].all() and out[0].sum() == 36
    want = sq.copy()
    want[2:8, 2:8] = False
    assert np.array_equal(out[1].astype(bool), want)
    # closing by the cross: a one-pixel-wide slit cut three pixels deep is closed except for its
    # mouth (whose upper neighbour stays empty); a one-pixel notch on an edge is not closed
    slit = np.zeros((H, W), bool)
    slit[3:9, 3:9] = True
    slit[3:6, 5] = False
    slit[6, 3] = False
    o = C.postprocess_masks(slit[None], np.array([0.9]), (H, W))
    want = np.zeros((H, W), bool)
    want[3:9, 3:9] = True
    want[3, 5] = False
    want[6, 3] = False
    assert np.array_equal(o[0].astype(bool), want)
    # an overlap that splits the later mask into two pieces empties it
    bar = np.zeros((H, W), bool)
    bar[5:7, 1:13] = True
    blocker = np.zeros((H, W), bool)
    blocker[3:9, 5:9] = True
    o = C.postprocess_masks(np.stack([blocker, bar]), scores, (H, W))
    assert o[0].sum() == 24 and o[1].sum() == 0
    # quirks: a zero score skips the image; fewer occupied columns than instances truncates
    assert C.postprocess_masks(np.stack([ring, sq]), np.array([0.9, 0.0]), (H, W)) is None
    thin = np.zeros((3, H, W), bool)
    thin[:, 2:10, 4] = True                      # all three in ONE column -> keep_ind has 1 entry
    o = C.postprocess_masks(thin, np.array([0.9, 0.8, 0.7]), (H, W))
    assert len(o) == 1
    o = C.postprocess_masks(np.zeros((2, H, W), bool), scores, (H, W))
    assert o == []
    # masks touching the image border: the reflected border never erodes them
    edge = np.zeros((H, W), bool)
    edge[0:4, 0:5] = True
    o = C.postprocess_masks(edge[None], np.array([0.9]), (H, W))
    assert np.array_equal(o[0].astype(bool), edge)
    # ... and a mask one pixel away from the border GROWS onto it: the dilation reaches the
    # border row and the erosion there has no outside neighbour to take it away again
    near = np.zeros((H, W), bool)
    near[5:11, 6:12] = True
    o = C.postprocess_masks(near[None], np.array([0.9]), (H, W))
    grown = near.copy()
    grown[11, 7:11] = True
    assert np.array_equal(o[0].astype(bool), grown)


def test_export_rows_shape():
    H, W = 8, 9
    m = np.zeros((2, H, W), bool)
    m[0, 2:5, 2:5] = True
    m[1, 5:7, 5:8] = True
    ids, enc = C.export_rows(["a.tif", "b.tif"], [m, m[:0]], [np.array([0.9, 0.8]), np.zeros(0)], (H, W))
    assert ids == ["a", "a"]
    assert enc[0] == "19 3 27 3 35 3"
