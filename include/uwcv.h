/*
 * uwcv.h -- C ABI of libuwcv.so, the B200 (sm_100a) implementation of the
 * post-inference hot path of Deam0on/uw-com-vision.
 *
 * The reference has no plugin / FFI API of its own (it is four Python scripts); the
 * boundary this library replaces is the Detectron2 data contract its scripts consume
 * (SURVEY.md section 8(b)).  Each entry point names the reference interface it stands
 * in for; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.
 *
 * Conventions
 *   - every function returns 0 on success or a negative uwcv error code
 *     (uwcv_strerror); no exception crosses the ABI;
 *   - all array arguments are DEVICE pointers on the current CUDA device, 16-byte
 *     aligned, owned by the caller (workspace included); unless stated otherwise the
 *     library never allocates, frees or synchronises: it only enqueues work on
 *     `stream` (a cudaStream_t passed as void*);
 *   - re-entrant, no global state: safe from several host threads on different
 *     streams / devices.
 */
#ifndef UWCV_H_
#define UWCV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UWCV_MASK_SIDE 28   /* mask head output side (Detectron2 ROI_MASK_HEAD.POOLER_RESOLUTION*2) */
#define UWCV_NUM_INT   20   /* int64 columns of a measurement row   */
#define UWCV_NUM_FLOAT 30   /* float64 columns of a measurement row */

/* stage bits of uwcv_paste_measure_stages / _range / _heads / _gather */
#define UWCV_STAGE_LAYOUT  1   /* tile geometry + offsets, marks cleared                              */
#define UWCV_STAGE_PASTE   2   /* paste + threshold + bit-pack + raw moments (+ planes, fused)        */
#define UWCV_STAGE_TRACE   4   /* border trace + descriptors + float columns (+ fused gather)         */
#define UWCV_STAGE_PLANES  8   /* full-frame planes written from the tiles (split pipeline)           */
#define UWCV_STAGE_SPLIT  16   /* modifier: stage 2 leaves the planes to stage 8                      */

#define UWCV_OK            0
#define UWCV_E_NULL       -1  /* required pointer is NULL                         */
#define UWCV_E_SHAPE      -2  /* negative / zero / inconsistent sizes             */
#define UWCV_E_ALIGN      -3  /* pointer not 16-byte aligned                      */
#define UWCV_E_THRESH     -4  /* mask threshold must be > 0       */
#define UWCV_E_WORKSPACE  -5  /* workspace smaller than uwcv_workspace_bytes(N,0) */
#define UWCV_E_LAUNCH     -6  /* CUDA reported a launch error                     */
#define UWCV_E_CAPACITY   -7  /* (device status) tile words exceed the workspace  */
#define UWCV_E_TOO_LARGE  -8  /* image side > 32768 or candidate count too large; as a device
                                 status: a tile whose raw moments could exceed int64          */

/* Library version, major * 10000 + minor * 100 + patch. */
int uwcv_version(void);

/* Static string for an error code. */
const char* uwcv_strerror(int code);

/* Words (uint32) per row of a full-frame bit-plane for image width W:
 * ceil(W / 32) rounded up to a multiple of 4 (16-byte rows). */
int uwcv_plane_row_words(int W);

/*
 * uwcv_planes_alloc / uwcv_planes_free -- device memory for the full-frame bit-planes as a
 * COMPRESSIBLE allocation on the current device (CUDA virtual memory management with generic
 * compression).  The planes are zeros almost everywhere; in compressible pages B200 compresses
 * them between L2 and HBM, so the same kernels write them 14 % and read them 40 % faster
 * (profiles/r02_fill_compress_microbench.txt).  Purely optional: every entry point works on any
 * device pointer, and these two are the only calls of the library that allocate.  `bytes` is
 * rounded up to the allocation granularity (2 MiB); *compressed (may be NULL) reports whether the
 * driver granted compression (0: ordinary pages, still valid).  The memory is NOT initialised.
 * uwcv_planes_free is the caller's cudaFree for such a pointer: no work may still use it.
 * Stands in for: the `torch.zeros(N, img_h, img_w)` behind Detectron2's paste_masks_in_image
 * (layers/mask_ops.py), reached from nn_inference.py:372.
 * Errors: UWCV_E_NULL, UWCV_E_SHAPE (bytes == 0 / foreign pointer), UWCV_E_WORKSPACE (out of
 * memory), UWCV_E_LAUNCH (no driver / VMM call failed).
 */
int uwcv_planes_alloc(size_t bytes, void** ptr, int* compressed);
int uwcv_planes_free(void* ptr);

/* Bytes of workspace uwcv_paste_measure needs for N instances whose tiles hold
 * `tile_words` 32-pixel words in total (28 bytes per word + descriptors).  The exact
 * word count of a call is reported back in status[1]; a caller that does not know it
 * passes a generous value and checks status[0]. */
size_t uwcv_workspace_bytes(int64_t N, int64_t tile_words);

/*
 * uwcv_paste_measure -- fused paste + threshold + bit-pack + per-instance measurement.
 *
 * Stands in for (reference file:line):
 *   detectron2 ROIMasks.to_bitmasks / layers.mask_ops.paste_masks_in_image reached from
 *     predictor(im) at nn_inference.py:372 (via detector_postprocess),
 *   pred_masks.to("cpu").numpy() at nn_inference.py:376,
 *   and the measurement block nn_inference.py:405-459 applied per instance.
 *
 *   masks      [N, 28, 28] float32 mask probabilities (after sigmoid, predicted class)
 *   boxes      [N, 4] float32 XYXY in OUTPUT-image pixels, already scaled / clipped /
 *              non-empty-filtered exactly as detector_postprocess does (keep those
 *              three torch ops on the caller side so the floats are bit-identical)
 *   image_idx  [N] int32 or NULL (0)      -> row column image_idx
 *   inst_idx   [N] int32 or NULL (0..N-1) -> row column inst_idx
 *   classes    [N] int64 or NULL (0)      -> row column class_id
 *   scores     [N] float32 or NULL (0)    -> row column score
 *   H, W       output image size (all instances of a call share it); <= 32768
 *   thr        mask threshold, > 0 (Detectron2 default 0.5; compared as out >= thr)
 *   pixels_per_metric   nn_inference.py:409 (0.85)
 *   bitplanes  NULL, or [N, H, uwcv_plane_row_words(W)] uint32: Detectron2-literal
 *              full-frame masks, 1 bit per pixel (bit b of word w = pixel x = 32 w + b)
 *   rows_i     [N, UWCV_NUM_INT] int64 out     (column order: SURVEY.md 8(b))
 *   rows_f     [N, UWCV_NUM_FLOAT] float64 out
 *   workspace / ws_bytes   >= uwcv_workspace_bytes(N, tile_words)
 *   status     [4] int64 out (device): [0] 0, UWCV_E_CAPACITY or UWCV_E_TOO_LARGE (a tile so large
 *              that m30 / m03 of an all-set mask would pass 2^63: raw moments are exact int64
 *              sums, so full-frame masks are limited to about 6 000 pixels a side), [1] tile words needed,
 *              [2] tile rows needed, [3] number of tiles above the one-instance-per-warp limit.  On a non-zero status no row is written.
 */
int uwcv_paste_measure(const float* masks, const float* boxes, const int32_t* image_idx,
                       const int32_t* inst_idx, const int64_t* classes, const float* scores,
                       int64_t N, int H, int W, float thr, double pixels_per_metric,
                       uint32_t* bitplanes, int64_t* rows_i, double* rows_f,
                       void* workspace, size_t ws_bytes, int64_t* status, void* stream);

/*
 * uwcv_paste_measure_stages -- the same call split by kernel, for profiling and for
 * callers that interleave their own work: `stages` is a bit mask, 1 = tile layout,
 * 2 = paste + threshold + bit-pack + raw moments, 4 = border trace + descriptors.
 * Stages must be issued in that order on one stream with identical arguments;
 * uwcv_paste_measure is stages = 7.
 *
 * Split pipeline: with bit 16 set stage 2 writes the tiles and the integer rows only (as with
 * bitplanes == NULL) and stage 8 writes the full-frame bit-planes FROM THE TILES with a kernel
 * that only moves data (96 threads, ~30 registers per CTA; csrc/plane_fill.cu).  Stages 8 and 4
 * both depend on stage 2 only, so a caller may issue 1 | 2 | 16 on one stream and, behind an
 * event, 8 and 4 on two others: the HBM-bound plane fill then shares the SMs with the border
 * trace of the same call and the tile arithmetic of the next one.  Same planes, same rows
 * (Detectron2 paste_masks_in_image, reached from nn_inference.py:372).
 */
int uwcv_paste_measure_stages(const float* masks, const float* boxes, const int32_t* image_idx,
                              const int32_t* inst_idx, const int64_t* classes,
                              const float* scores, int64_t N, int H, int W, float thr,
                              double pixels_per_metric, uint32_t* bitplanes, int64_t* rows_i,
                              double* rows_f, void* workspace, size_t ws_bytes, int64_t* status,
                              void* stream, int stages);

/*
 * uwcv_paste_measure_range -- as uwcv_paste_measure_stages, restricted to the instances
 * [first, first + count) of the N-instance call (stage 1, the layout, always covers all N).
 * Lets a caller start pasting the first images of a batch while the mask probabilities of
 * the later ones are still in flight from the host, and trace all of them in one launch.
 */
int uwcv_paste_measure_range(const float* masks, const float* boxes, const int32_t* image_idx,
                             const int32_t* inst_idx, const int64_t* classes, const float* scores,
                             int64_t N, int H, int W, float thr, double pixels_per_metric,
                             uint32_t* bitplanes, int64_t* rows_i, double* rows_f, void* workspace,
                             size_t ws_bytes, int64_t* status, void* stream, int stages,
                             int64_t first, int64_t count);

/*
 * uwcv_paste_measure_heads -- as uwcv_paste_measure_range, fed straight from the mask head
 * (the single-forward path): `masks` is [N, mask_channels, 28, 28] float32
 * and instance i uses channel classes[i] + channel_offset (mask_channels == 1: channel 0, the
 * class-agnostic head).  With is_logits != 0 the values are the head's raw logits and the
 * sigmoid is applied while the tile is staged, as torch's CUDA sigmoid computes it
 * (1 / (1 + expf(-x)), IEEE add and divide): no N x 1 x 28 x 28 probability tensor is
 * materialised.  A channel outside [0, mask_channels) yields an empty mask.
 *
 * Stands in for detectron2 modeling/roi_heads/mask_head.py::mask_rcnn_inference
 * (pred_mask_logits[arange(N), pred_classes][:, None].sigmoid()) followed by
 * detector_postprocess, i.e. what predictor(im) does after the mask head at
 * nn_inference.py:372.  classes must be non-NULL when mask_channels > 1.
 */
int uwcv_paste_measure_heads(const float* masks, int mask_channels, int channel_offset,
                             int is_logits, const float* boxes, const int32_t* image_idx,
                             const int32_t* inst_idx, const int64_t* classes, const float* scores,
                             int64_t N, int H, int W, float thr, double pixels_per_metric,
                             uint32_t* bitplanes, int64_t* rows_i, double* rows_f, void* workspace,
                             size_t ws_bytes, int64_t* status, void* stream, int stages,
                             int64_t first, int64_t count);

/*
 * uwcv_paste_measure_gather -- as uwcv_paste_measure_heads, with the all-gather of the
 * measurement table fused into the border-trace kernel (the one collective of
 * the path).  Every rank passes the tables of ALL ranks -- device pointers that are valid on
 * this GPU: peer-mapped / symmetric memory over NVLink, this rank's own table included -- and
 * its row offset; the trace kernel stores its finished rows (coalesced 16-byte stores, one
 * contiguous block per warp) into rows [row_base, row_base + N) of every table.  The tables are
 * complete on every rank once all ranks have passed a barrier issued on `stream` after this
 * call (the caller's: e.g. a symmetric-memory signal barrier); no NCCL kernel is involved.
 * gather == NULL or gather->world == 0: no gather.  Stage 4 must cover the whole call.
 * gather->dst_plus_1 = p + 1 restricts the stores to rank p's table (the other pointers may be
 * NULL): a gather to one destination costs one copy of the rows over NVLink instead of `world`.
 *
 * Stands in for the dist all-gather the image-sharded job ends with (there is none in the
 * single-process reference; nn_inference.py:485-498 loops over all images in one process).
 */
#define UWCV_MAX_PEERS 16
typedef struct uwcv_gather {
  int32_t world;                       /* number of ranks, <= UWCV_MAX_PEERS */
  int32_t dst_plus_1;                  /* 0: store into every rank's table (all-gather);
                                          p + 1: only into the table of rank p (gather to p) */
  int64_t row_base;                    /* first row of this rank in the gathered tables */
  int64_t* rows_i[UWCV_MAX_PEERS];     /* [total_rows, UWCV_NUM_INT] of rank p   */
  double* rows_f[UWCV_MAX_PEERS];      /* [total_rows, UWCV_NUM_FLOAT] of rank p */
} uwcv_gather;

int uwcv_paste_measure_gather(const float* masks, int mask_channels, int channel_offset,
                              int is_logits, const float* boxes, const int32_t* image_idx,
                              const int32_t* inst_idx, const int64_t* classes, const float* scores,
                              int64_t N, int H, int W, float thr, double pixels_per_metric,
                              uint32_t* bitplanes, int64_t* rows_i, double* rows_f, void* workspace,
                              size_t ws_bytes, int64_t* status, void* stream, int stages,
                              int64_t first, int64_t count, const uwcv_gather* gather);

/*
 * uwcv_unpack_planes -- expand bit-planes into the Detectron2-literal N x H x W bool
 * tensor (one byte per pixel), for callers that read pred_masks as such
 * (nn_inference.py:326, :376).
 */
int uwcv_unpack_planes(const uint32_t* bitplanes, int64_t N, int H, int W, uint8_t* out,
                       void* stream);

/*
 * uwcv_ingest -- copy `bytes` (a multiple of 16, else UWCV_E_SHAPE) from
 * pinned, device-mapped HOST memory (cudaHostAlloc / torch pin_memory under unified addressing)
 * to device memory with a kernel instead of the copy engine.  For the small per-call arrays
 * (boxes, scores, classes, indices): a cudaMemcpyAsync of 1 MB issued behind the 200 MB of mask
 * probabilities of a batch is served after them and delays the first kernel by milliseconds.
 */
int uwcv_ingest(const void* src_host_mapped, void* dst, size_t bytes, void* stream);

/*
 * Union / connected-component mode -- the reference-literal GetMask_Contours
 * (nn_inference.py:394-459): the masks of the selected class are OR-ed into one image and
 * EVERY external contour of that union is measured (touching instances merge).
 *
 * Call order: uwcv_paste_measure_stages(stages = 3, bitplanes = NULL) on the selected
 * instances, then group them (instances of one image whose 1-pixel-dilated pixel boxes, rows_i
 * columns bbox_*, overlap belong to one group: uwcv_union_group does it on the device, or the
 * caller on the host), then uwcv_union_measure[_grouped].
 *
 *   paste_workspace / paste_ws_bytes / N   exactly as passed to uwcv_paste_measure_stages
 *   member_group   [N] int32: group of every instance, -1 = not part of any group (empty mask)
 *   group_desc     [G] uwcv_tile: word-aligned window of the frame covering the group's pixel
 *                  boxes; word_off = offset of the group's tile in group_planes (tw * th words)
 *   group_image    [G] int32: image index reported in the rows
 *   group_planes   [3 * group_words] uint32 scratch (mask / visited / sign planes)
 *   rec_workspace  >= uwcv_union_workspace_bytes(rec_cap, ext_rows_cap)
 *   rec_cap        capacity in contours; ext_rows_cap capacity in (contour, row) pairs
 *   rows_i  [rec_cap, 10] int64 out: image_idx, group, start_x, start_y (raster-first pixel of
 *           the component), brect_x, brect_y, brect_w, brect_h (cv2.boundingRect), n_points, valid
 *   rows_f  [rec_cap, 16] float64 out: contour_area, perimeter, rect cx, cy, w, h, angle, Feret,
 *           Aspect_Ratio, Roundness, Circularity, Sphericity, Length, Width, CircularED, Chords
 *   counters [4] int64 out (device): [0] contours found, [1] 0 or UWCV_E_CAPACITY (no row is
 *           valid then; retry with counters[0] / counters[2] as capacities), [2] extreme rows needed
 * Rows are in discovery order; the reference's order is a stable sort by brect_x of the rows
 * taken in decreasing (start_y, start_x) order, and it drops contour_area < 100.
 */
typedef struct uwcv_tile {
  int32_t wx0;        /* first 32-pixel word column */
  int32_t y0;         /* first pixel row            */
  int32_t tw;         /* width in words             */
  int32_t th;         /* height in rows             */
  int64_t word_off;   /* offset of the tile's first word */
  int64_t reserved;
} uwcv_tile;

size_t uwcv_union_workspace_bytes(int64_t rec_cap, int64_t ext_rows_cap);

int uwcv_union_measure(const void* paste_workspace, size_t paste_ws_bytes, int64_t N,
                       const int32_t* member_group, const uwcv_tile* group_desc,
                       const int32_t* group_image, int64_t G, uint32_t* group_planes,
                       int64_t group_words, void* rec_workspace, size_t rec_ws_bytes,
                       int64_t rec_cap, int64_t ext_rows_cap, double pixels_per_metric,
                       int64_t* rows_i, double* rows_f, int64_t* counters, void* stream);

/*
 * Grouping on the device, so that a union-mode call needs no host round trip between the paste
 * and the contour kernels:
 *   uwcv_union_group: paste_rows_i = the int64 rows [N, UWCV_NUM_INT] written by
 *     uwcv_paste_measure_stages(stages = 3); the instances of an image must be consecutive and the
 *     image indices (column 0) must be image_base, image_base + 1, ... (B images).  One CTA per
 *     image forms the groups (minimum-label propagation over the box-overlap graph in shared
 *     memory) and their tiles; outputs member_group [N], group_desc [N] (the first G are used),
 *     group_image [N] and group_counters [4] int64 (device): [0] G, [1] words of all group tiles
 *     (a multiple of 4), [3] 0 or UWCV_E_CAPACITY when [1] > group_words_cap.
 *   uwcv_union_measure_grouped: uwcv_union_measure with G and the plane size read from
 *     group_counters on the device; group_planes holds 3 * group_words_cap words.  On
 *     group_counters[3] != 0 nothing is measured and counters[1] = UWCV_E_CAPACITY: repeat both
 *     calls with group_words_cap >= group_counters[1].
 */
size_t uwcv_union_group_workspace_bytes(int64_t N, int B);
int uwcv_union_group(const int64_t* paste_rows_i, int64_t N, int B, int64_t image_base,
                     void* group_workspace, size_t group_ws_bytes, int32_t* member_group,
                     uwcv_tile* group_desc, int32_t* group_image, int64_t group_words_cap,
                     int64_t* group_counters, void* stream);
int uwcv_union_measure_grouped(const void* paste_workspace, size_t paste_ws_bytes, int64_t N,
                               const int32_t* member_group, const uwcv_tile* group_desc,
                               const int32_t* group_image, const int64_t* group_counters,
                               uint32_t* group_planes, int64_t group_words_cap, void* rec_workspace,
                               size_t rec_ws_bytes, int64_t rec_cap, int64_t ext_rows_cap,
                               double pixels_per_metric, int64_t* rows_i, double* rows_f,
                               int64_t* counters, void* stream);

/*
 * Mask clean-up + RLE export -- stands in for postprocess_masks
 * (nn_inference.py:265-306) and rle_encoding (:253-263) of the reference's export loop (:319-336).
 *
 * Call order on one stream: uwcv_paste_measure_stages(stages = 3) -> [uwcv_mask_column_totals,
 * host decides `limit`] -> uwcv_clean_masks -> exclusive scan of run_counts (caller) ->
 * uwcv_rle_write.  The instances of one image must be consecutive and carry inst_idx = 0, 1, ...
 * in list (score) order; image_slot[i] in [0, B) is the image's position in this call.  The
 * workspace's mask plane is REPLACED by the cleaned masks; its other planes are scratch (do not
 * run stage 4 on it afterwards).
 *
 * uwcv_mask_column_totals: column_totals [B, W] int32 (zeroed by the caller) += number of mask
 *   pixels of every instance in each image column -- what the reference's
 *   np.sum(ori_mask, axis=(0, 1)) yields (:277); the caller derives limit[b] from it.
 * uwcv_clean_masks: per instance with inst_idx < limit[image_slot] (limit == NULL: all):
 *   fill holes (scipy binary_fill_holes), dilate + erode with the 3 x 3 cross (skimage defaults,
 *   reflected border), cut the pixels covered by an earlier cleaned mask of the image, empty the
 *   mask if it has more than one 8-connected piece.  Outputs, [N] each: flags (bit 0 = emptied as
 *   multi-piece, bit 1 = beyond limit: not part of the output), area (pixels of the result),
 *   run_counts (column-major runs of the result, before merging across columns).
 * uwcv_rle_write: runs [2 * run_offsets[N]] int64 = (start, length) pairs, start 1-based in
 *   column-major order (x * H + y + 1), instance i at run_offsets[i] .. run_offsets[i + 1].  A run
 *   that ends on the last image row and the next that starts on row 0 of the following column are
 *   one run for the reference (consecutive flat indices): the caller merges pairs with
 *   start[k + 1] == start[k] + length[k].
 */
int uwcv_mask_column_totals(const void* paste_workspace, size_t ws_bytes, int64_t N, int W,
                            const int32_t* image_slot, int32_t* column_totals, void* stream);
int uwcv_clean_masks(void* paste_workspace, size_t ws_bytes, int64_t N, int H, int W,
                     const int32_t* image_slot, const int32_t* inst_idx, const int32_t* limit,
                     int32_t* flags, int64_t* area, int64_t* run_counts, void* stream);
int uwcv_rle_write(const void* paste_workspace, size_t ws_bytes, int64_t N, int H, int W,
                   const int64_t* run_offsets, int64_t* runs, void* stream);
/*
 * The EncodedPixels text -- conv = ' '.join(map(str, rle_encoding(mask))) (nn_inference.py:317) --
 * written on the device.  runs [total_runs, 2] as uwcv_rle_write left them, run_instance
 * [total_runs] int32 = the instance of every run.  uwcv_rle_text_prep: chars [total_runs] int64 =
 * characters run k contributes ("start length " with the runs that continue it across a column
 * boundary merged into its length; 0 for such a continuation).  The caller takes the exclusive
 * prefix sum (text_offsets) and calls uwcv_rle_text_write, which prints the decimal digits; the
 * text of instance i is bytes [text_offsets[run_offsets[i]], text_offsets[run_offsets[i + 1]] - 1).
 */
int uwcv_rle_text_prep(int64_t total_runs, const int64_t* runs, const int32_t* run_instance,
                       int64_t* chars, void* stream);
int uwcv_rle_text_write(int64_t total_runs, const int64_t* runs, const int32_t* run_instance,
                        const int64_t* text_offsets, uint8_t* text, void* stream);

/*
 * Literal postprocess_masks(ori_mask, ...) entry (nn_inference.py:265, called at :325-327 with
 * Detectron2's pasted pred_masks): N x H x W bool / uint8 masks in, cleaned masks out.
 *   uwcv_mask_pixel_boxes: boxes [N, 4] float32 = [xmin, ymin, xmax + 1, ymax + 1] of the set
 *     pixels of every mask (zeros for an empty one): feed them to uwcv_paste_measure_stages
 *     (stages = 1, the layout) so that every mask gets a tile with a margin;
 *   uwcv_pack_mask_tiles: fills the tile mask plane of that workspace from the bytes (instead of
 *     stage 2); then uwcv_mask_column_totals / uwcv_clean_masks / uwcv_rle_write as above;
 *   uwcv_tiles_to_masks: out [N, H, W] uint8, zeroed by the caller, receives the cleaned masks.
 */
int uwcv_mask_pixel_boxes(const uint8_t* masks, int64_t N, int H, int W, float* boxes, void* stream);
int uwcv_pack_mask_tiles(const uint8_t* masks, int64_t N, int H, int W, void* paste_workspace,
                         size_t ws_bytes, void* stream);
int uwcv_tiles_to_masks(const void* paste_workspace, size_t ws_bytes, int64_t N, int H, int W,
                        uint8_t* out, void* stream);

/* Bytes of workspace for uwcv_nms_filter; image_off is a HOST array [B + 1];
 * num_classes <= 0 means 128.  Linear in the candidate count (about 40 bytes per candidate + 12 per
 * image and class): no suppression matrix is stored, the IoUs are evaluated on the fly. */
size_t uwcv_nms_workspace_bytes(const int64_t* image_off, int B, int num_classes);

/*
 * uwcv_nms_filter -- batched score filter + per-class NMS + top-k.
 *
 * Stands in for detectron2 fast_rcnn_inference_single_image (score > thresh,
 * batched_nms, keep[:topk]) whose thresholds the reference sets at
 * nn_inference.py:226, with torchvision _batched_nms_vanilla semantics.
 *
 *   boxes      [R, 4] float32 XYXY (already clipped), candidates of image b are rows
 *              image_off[b] .. image_off[b+1]-1
 *   scores     [R] float32,  classes [R] int64 in [0, num_classes) (others are dropped)
 *   image_off  HOST pointer, [B + 1] int64, image_off[0] = 0, non-decreasing; at most
 *              262144 candidates per image.  Only read DURING the call (the offsets travel as
 *              kernel arguments): it may be pageable and may be freed on return; the call never
 *              synchronises and can be captured in a CUDA graph
 *   num_classes  K of the box head (<= 8192; <= 0 means 128): the serial part of NMS runs
 *              once per (image, class) in parallel
 *   score_thr  keep candidates with score > score_thr (float32 compare)
 *   iou_thr    suppress when (double)iou > iou_thr
 *   topk       per image, < 0 = unlimited
 *   keep       [R] int64 out: for image b, keep[image_off[b] + r] for r < keep_count[b]
 *              are candidate row indices in score-descending order
 *   keep_count [B] int32 out
 */
int uwcv_nms_filter(const float* boxes, const float* scores, const int64_t* classes,
                    const int64_t* image_off, int B, int num_classes, float score_thr,
                    double iou_thr, int topk, int64_t* keep, int32_t* keep_count, void* workspace,
                    size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* UWCV_H_ */
