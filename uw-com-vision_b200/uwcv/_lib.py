"""ctypes binding of libuwcv.so (include/uwcv.h).  There is no fallback: if the
shared object is missing the import of any product entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build
from .build import LIB_PATH

_lib = None
_path = LIB_PATH


def use_library_variant(variant: str) -> str:
    """Development only (tools/): load lib/libuwcv_<variant>.so ("tuning": environment sweep
    knobs, "check": device-side bounds traps) instead of the release library.  Must be called
    before the first entry point is used; the product path never calls it."""
    global _path
    if _lib is not None:
        raise RuntimeError("the library is already loaded")
    _path = _build.build(variant=variant)
    return _path

MAX_PEERS = 16


class Gather(C.Structure):
    """uwcv_gather of include/uwcv.h."""
    _fields_ = [("world", C.c_int32), ("dst_plus_1", C.c_int32), ("row_base", C.c_int64),
                ("rows_i", C.c_void_p * MAX_PEERS), ("rows_f", C.c_void_p * MAX_PEERS)]

E_CAPACITY = -7


class UwcvError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        super().__init__(f"{where}: uwcv error {code}: {strerror(code)}")


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_path):
        raise RuntimeError(
            f"{_path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the uwcv hot path has no CPU or PyTorch fallback)")
    L = C.CDLL(_path)
    vp, i64, i32, f32, f64, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_size_t
    L.uwcv_version.restype = C.c_int
    L.uwcv_strerror.restype = C.c_char_p
    L.uwcv_strerror.argtypes = [C.c_int]
    L.uwcv_plane_row_words.restype = C.c_int
    L.uwcv_plane_row_words.argtypes = [C.c_int]
    L.uwcv_workspace_bytes.restype = sz
    L.uwcv_workspace_bytes.argtypes = [i64, i64]
    L.uwcv_planes_alloc.restype = C.c_int
    L.uwcv_planes_alloc.argtypes = [sz, C.POINTER(vp), C.POINTER(C.c_int)]
    L.uwcv_planes_free.restype = C.c_int
    L.uwcv_planes_free.argtypes = [vp]
    L.uwcv_paste_measure.restype = C.c_int
    L.uwcv_paste_measure.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, f32, f64,
                                     vp, vp, vp, vp, sz, vp, vp]
    L.uwcv_paste_measure_stages.restype = C.c_int
    L.uwcv_paste_measure_stages.argtypes = L.uwcv_paste_measure.argtypes + [i32]
    L.uwcv_paste_measure_range.restype = C.c_int
    L.uwcv_paste_measure_range.argtypes = L.uwcv_paste_measure_stages.argtypes + [i64, i64]
    L.uwcv_paste_measure_heads.restype = C.c_int
    L.uwcv_paste_measure_heads.argtypes = [vp, i32, i32, i32] + L.uwcv_paste_measure_range.argtypes[1:]
    L.uwcv_paste_measure_gather.restype = C.c_int
    L.uwcv_paste_measure_gather.argtypes = L.uwcv_paste_measure_heads.argtypes + [C.POINTER(Gather)]
    L.uwcv_mask_column_totals.restype = C.c_int
    L.uwcv_mask_column_totals.argtypes = [vp, sz, i64, i32, vp, vp, vp]
    L.uwcv_clean_masks.restype = C.c_int
    L.uwcv_clean_masks.argtypes = [vp, sz, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.uwcv_rle_write.restype = C.c_int
    L.uwcv_rle_write.argtypes = [vp, sz, i64, i32, i32, vp, vp, vp]
    L.uwcv_rle_text_prep.restype = C.c_int
    L.uwcv_rle_text_prep.argtypes = [i64, vp, vp, vp, vp]
    L.uwcv_rle_text_write.restype = C.c_int
    L.uwcv_rle_text_write.argtypes = [i64, vp, vp, vp, vp, vp]
    L.uwcv_ingest.restype = C.c_int
    L.uwcv_ingest.argtypes = [vp, vp, sz, vp]
    L.uwcv_mask_pixel_boxes.restype = C.c_int
    L.uwcv_mask_pixel_boxes.argtypes = [vp, i64, i32, i32, vp, vp]
    L.uwcv_pack_mask_tiles.restype = C.c_int
    L.uwcv_pack_mask_tiles.argtypes = [vp, i64, i32, i32, vp, sz, vp]
    L.uwcv_tiles_to_masks.restype = C.c_int
    L.uwcv_tiles_to_masks.argtypes = [vp, sz, i64, i32, i32, vp, vp]
    L.uwcv_unpack_planes.restype = C.c_int
    L.uwcv_unpack_planes.argtypes = [vp, i64, i32, i32, vp, vp]
    L.uwcv_union_workspace_bytes.restype = sz
    L.uwcv_union_workspace_bytes.argtypes = [i64, i64]
    L.uwcv_union_measure.restype = C.c_int
    L.uwcv_union_measure.argtypes = [vp, sz, i64, vp, vp, vp, i64, vp, i64, vp, sz, i64, i64, f64,
                                     vp, vp, vp, vp]
    L.uwcv_union_group_workspace_bytes.restype = sz
    L.uwcv_union_group_workspace_bytes.argtypes = [i64, i32]
    L.uwcv_union_group.restype = C.c_int
    L.uwcv_union_group.argtypes = [vp, i64, i32, i64, vp, sz, vp, vp, vp, i64, vp, vp]
    L.uwcv_union_measure_grouped.restype = C.c_int
    L.uwcv_union_measure_grouped.argtypes = [vp, sz, i64, vp, vp, vp, vp, vp, i64, vp, sz, i64, i64, f64,
                                             vp, vp, vp, vp]
    L.uwcv_nms_workspace_bytes.restype = sz
    L.uwcv_nms_workspace_bytes.argtypes = [C.POINTER(C.c_int64), i32, i32]
    L.uwcv_nms_filter.restype = C.c_int
    L.uwcv_nms_filter.argtypes = [vp, vp, vp, C.POINTER(C.c_int64), i32, i32, f32, f64, i32,
                                  vp, vp, vp, sz, vp]
    _lib = L
    return L


EXPORTS = ("uwcv_version", "uwcv_strerror", "uwcv_plane_row_words", "uwcv_workspace_bytes",
           "uwcv_planes_alloc", "uwcv_planes_free",
           "uwcv_paste_measure", "uwcv_paste_measure_stages", "uwcv_paste_measure_range", "uwcv_paste_measure_heads", "uwcv_paste_measure_gather",
           "uwcv_unpack_planes", "uwcv_ingest", "uwcv_mask_column_totals", "uwcv_clean_masks", "uwcv_rle_write",
           "uwcv_rle_text_prep", "uwcv_rle_text_write",
           "uwcv_mask_pixel_boxes", "uwcv_pack_mask_tiles", "uwcv_tiles_to_masks",
           "uwcv_union_workspace_bytes", "uwcv_union_measure", "uwcv_union_group_workspace_bytes",
           "uwcv_union_group", "uwcv_union_measure_grouped", "uwcv_nms_workspace_bytes",
           "uwcv_nms_filter")


def strerror(code: int) -> str:
    return lib().uwcv_strerror(int(code)).decode()


def check(code: int, where: str) -> None:
    if code != 0:
        raise UwcvError(code, where)
