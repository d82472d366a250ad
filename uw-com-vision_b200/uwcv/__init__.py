"""uwcv -- B200-native post-inference measurement path of uw-com-vision.

Public surface (mirrors the reference-side names, see api.py):
    measure_instances, MeasurementStream, paste_masks_in_image, detector_postprocess,
    fast_rcnn_inference_single_image, get_counts, MeasurementTable, Instances, Boxes
"""
from .structures import Boxes, Instances                      # noqa: F401
from .schema import (INT_COLUMNS, FLOAT_COLUMNS, CSV_COLUMNS, CLASS_NAMES,   # noqa: F401
                     CLASS_KEYWORDS)
from .api import (Engine, MeasurementTable, MeasurementStream, PendingTable,
                  submit_measure_instances, measure_instances, paste_masks_in_image,  # noqa: F401
                  detector_postprocess, fast_rcnn_inference_single_image, get_counts,
                  scale_clip_boxes, tile_words)
from .grouping import (group_by_class, write_classes_csv, moving_average, report_class,  # noqa: F401
                       write_results_csv, write_shape_descriptor_csv, smoothed_columns)
from .dist import (shard_indices, shard_instances_by_tile, take_instances,   # noqa: F401
                   all_gather_table, sort_rows)
from .union import UnionTable, measure_union                  # noqa: F401
from .heads import SingleForward, fast_rcnn_inference         # noqa: F401
from .cleanup import RleExport, export_rle, postprocess_masks, write_rle_csv     # noqa: F401

__all__ = [
    "Boxes", "Instances", "Engine", "MeasurementTable", "MeasurementStream", "PendingTable",
    "submit_measure_instances", "measure_instances",
    "paste_masks_in_image", "detector_postprocess", "fast_rcnn_inference_single_image",
    "get_counts", "group_by_class", "write_classes_csv", "moving_average", "report_class",
    "write_results_csv", "write_shape_descriptor_csv", "smoothed_columns", "shard_indices", "shard_instances_by_tile", "take_instances", "all_gather_table", "sort_rows", "UnionTable", "measure_union", "SingleForward", "fast_rcnn_inference", "RleExport", "export_rle", "postprocess_masks", "write_rle_csv",
]
