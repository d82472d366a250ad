// extern "C" entry points of libuwcv.so (declared in include/uwcv.h).
#include "../../include/uwcv.h"
#include "uwcv_common.cuh"

namespace uwcv {
cudaError_t launch_layout(const float*, int64_t, int, int, const Workspace&, int64_t*, int,
                          cudaStream_t);
cudaError_t launch_paste_measure(const float*, const float*, const int32_t*, const int32_t*,
                                 const int64_t*, int64_t, int64_t, int, int, float, uint32_t*,
                                 int64_t*, const Workspace&, const int64_t*, int, cudaStream_t,
                                 const MaskSource&);
cudaError_t launch_contour_measure(int64_t, int64_t, const float*, double, int64_t*, double*,
                                   const Workspace&, const int64_t*, int, cudaStream_t,
                                   const GatherDst&, int);
cudaError_t launch_plane_fill(uint32_t*, int64_t, int64_t, int, int, const Workspace&, const int64_t*, int,
                              cudaStream_t);
cudaError_t launch_unpack(const uint32_t*, int64_t, int, int, uint8_t*, int, cudaStream_t);
cudaError_t launch_ingest(const void*, void*, size_t, cudaStream_t);
size_t nms_workspace_bytes_host(const int64_t*, int, int);
size_t union_workspace_bytes_host(int64_t, int64_t);
cudaError_t launch_union(int64_t, const Workspace&, const int32_t*, const TileDesc*, const int32_t*,
                         int64_t, uint32_t*, int64_t, void*, int64_t, int64_t, double, int64_t*,
                         double*, int64_t*, int, cudaStream_t, const int64_t*);
size_t union_group_workspace_bytes_host(int64_t, int);
cudaError_t launch_union_group(const int64_t*, int64_t, int, int64_t, void*, int32_t*, TileDesc*, int32_t*,
                               int64_t, int64_t*, cudaStream_t);
cudaError_t launch_nms(const float*, const float*, const int64_t*, const int64_t*, int, int, float,
                       double, int, int64_t*, int32_t*, void*, cudaStream_t);
cudaError_t launch_column_totals(int64_t, int, const Workspace&, const int32_t*, int32_t*,
                                 cudaStream_t);
cudaError_t launch_clean(int64_t, int, int, const Workspace&, const int32_t*, const int32_t*,
                         const int32_t*, int32_t*, int64_t*, int64_t*, cudaStream_t);
cudaError_t launch_rle_write(int64_t, int, int, const Workspace&, const int64_t*, int64_t*,
                             cudaStream_t);
cudaError_t launch_rle_text_prep(int64_t, const int64_t*, const int32_t*, int64_t*, cudaStream_t);
cudaError_t launch_rle_text_write(int64_t, const int64_t*, const int32_t*, const int64_t*, uint8_t*, cudaStream_t);
cudaError_t launch_pixel_boxes(const uint8_t*, int64_t, int, int, float*, cudaStream_t);
cudaError_t launch_pack_tiles(const uint8_t*, int64_t, int, int, const Workspace&, cudaStream_t);
cudaError_t launch_tiles_to_masks(int64_t, int, int, const Workspace&, uint8_t*, cudaStream_t);
}  // namespace uwcv

namespace {
inline bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }
int num_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}
}  // namespace

extern "C" {

int uwcv_version(void) { return 100; }

const char* uwcv_strerror(int code) {
  switch (code) {
    case UWCV_OK: return "ok";
    case UWCV_E_NULL: return "required pointer is NULL";
    case UWCV_E_SHAPE: return "invalid shape or size argument";
    case UWCV_E_ALIGN: return "pointer is not 16-byte aligned";
    case UWCV_E_THRESH: return "mask threshold must be > 0";
    case UWCV_E_WORKSPACE: return "workspace too small";
    case UWCV_E_LAUNCH: return "CUDA launch error";
    case UWCV_E_CAPACITY: return "tile words exceed workspace capacity (see status[1])";
    case UWCV_E_TOO_LARGE: return "image side or candidate count too large";
    default: return "unknown uwcv error";
  }
}

int uwcv_plane_row_words(int W) { return uwcv::plane_row_words(W); }

size_t uwcv_workspace_bytes(int64_t N, int64_t tile_words) {
  if (N < 0) N = 0;
  if (tile_words < 0) tile_words = 0;
  return uwcv::workspace_bytes(N, tile_words + 4);
}

int uwcv_paste_measure(const float* masks, const float* boxes, const int32_t* image_idx,
                       const int32_t* inst_idx, const int64_t* classes, const float* scores,
                       int64_t N, int H, int W, float thr, double pixels_per_metric,
                       uint32_t* bitplanes, int64_t* rows_i, double* rows_f, void* workspace,
                       size_t ws_bytes, int64_t* status, void* stream) {
  return uwcv_paste_measure_range(masks, boxes, image_idx, inst_idx, classes, scores, N, H, W,
                                  thr, pixels_per_metric, bitplanes, rows_i, rows_f, workspace,
                                  ws_bytes, status, stream, 7, 0, N);
}

int uwcv_paste_measure_stages(const float* masks, const float* boxes, const int32_t* image_idx,
                              const int32_t* inst_idx, const int64_t* classes,
                              const float* scores, int64_t N, int H, int W, float thr,
                              double pixels_per_metric, uint32_t* bitplanes, int64_t* rows_i,
                              double* rows_f, void* workspace, size_t ws_bytes, int64_t* status,
                              void* stream, int stages) {
  return uwcv_paste_measure_range(masks, boxes, image_idx, inst_idx, classes, scores, N, H, W,
                                  thr, pixels_per_metric, bitplanes, rows_i, rows_f, workspace,
                                  ws_bytes, status, stream, stages, 0, N);
}

int uwcv_paste_measure_gather(const float* masks, int mask_channels, int channel_offset,
                             int is_logits, const float* boxes, const int32_t* image_idx,
                             const int32_t* inst_idx, const int64_t* classes, const float* scores,
                             int64_t N, int H, int W, float thr, double pixels_per_metric,
                             uint32_t* bitplanes, int64_t* rows_i, double* rows_f, void* workspace,
                             size_t ws_bytes, int64_t* status, void* stream, int stages,
                             int64_t first, int64_t count, const uwcv_gather* gather) {
  if (N < 0 || H <= 0 || W <= 0) return UWCV_E_SHAPE;
  uwcv::GatherDst gd;
  gd.world = 0;
  gd.only = -1;
  gd.row_base = 0;
  if (gather && gather->world > 0) {
    if (gather->world > UWCV_MAX_PEERS || gather->row_base < 0) return UWCV_E_SHAPE;
    if ((stages & 4) && (first != 0 || count != N)) return UWCV_E_SHAPE;   // whole-call traces only
    if (gather->dst_plus_1 < 0 || gather->dst_plus_1 > gather->world) return UWCV_E_SHAPE;
    gd.world = gather->world;
    gd.only = gather->dst_plus_1 - 1;
    gd.row_base = gather->row_base;
    for (int p = 0; p < gather->world; ++p) {
      gd.rows_i[p] = nullptr;
      gd.rows_f[p] = nullptr;
      if (gd.only >= 0 && p != gd.only) continue;     // only the destination's table is touched
      if (!gather->rows_i[p] || !gather->rows_f[p]) return UWCV_E_NULL;
      if (misaligned(gather->rows_i[p]) || misaligned(gather->rows_f[p])) return UWCV_E_ALIGN;
      gd.rows_i[p] = gather->rows_i[p];
      gd.rows_f[p] = gather->rows_f[p];
    }
  }
  if (mask_channels < 1 || mask_channels > 65536) return UWCV_E_SHAPE;
  if (mask_channels > 1 && N > 0 && !classes) return UWCV_E_NULL;
  if (H > 32768 || W > 32768) return UWCV_E_TOO_LARGE;
  if (!(thr > 0.f)) return UWCV_E_THRESH;
  if (!(pixels_per_metric > 0.0)) return UWCV_E_SHAPE;
  if (first < 0 || count < 0 || first + count > N) return UWCV_E_SHAPE;
  if (!status) return UWCV_E_NULL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 0) {
    return cudaMemsetAsync(status, 0, 4 * sizeof(int64_t), st) == cudaSuccess ? UWCV_OK
                                                                              : UWCV_E_LAUNCH;
  }
  if (!masks || !boxes || !rows_i || !rows_f || !workspace) return UWCV_E_NULL;
  if ((stages & UWCV_STAGE_PLANES) && !bitplanes) return UWCV_E_NULL;     // nothing to write the planes into
  if (misaligned(masks) || misaligned(boxes) || misaligned(rows_i) || misaligned(rows_f) ||
      misaligned(workspace) || misaligned(status) || (bitplanes && misaligned(bitplanes)))
    return UWCV_E_ALIGN;
  if (ws_bytes < uwcv::workspace_bytes(N, 4)) return UWCV_E_WORKSPACE;
  const uwcv::Workspace ws = uwcv::carve(workspace, ws_bytes, N);
  uwcv::MaskSource src;
  src.stride = (int64_t)mask_channels * UWCV_MASK_SIDE * UWCV_MASK_SIDE;
  src.channels = mask_channels;
  src.channel_offset = channel_offset;
  src.logits = is_logits ? 1 : 0;
  if ((stages & 1) && uwcv::launch_layout(boxes, N, H, W, ws, status, num_sms(), st) != cudaSuccess)
    return UWCV_E_LAUNCH;
  // split pipeline (stage bit 16): the paste writes tiles and integer rows only, the planes are
  // written from the tiles by stage 8 (plane_fill.cu), which a caller may issue on another stream
  const bool split = (stages & 16) != 0;
  if ((stages & 2) &&
      uwcv::launch_paste_measure(masks, boxes, image_idx, inst_idx, classes, first, count, H, W,
                                 thr, split ? nullptr : bitplanes, rows_i, ws, status, num_sms(), st,
                                 src) != cudaSuccess)
    return UWCV_E_LAUNCH;
  if ((stages & 8) &&
      uwcv::launch_plane_fill(bitplanes, first, count, H, W, ws, status, num_sms(), st) != cudaSuccess)
    return UWCV_E_LAUNCH;
  if ((stages & 4) && uwcv::launch_contour_measure(first, count, scores, pixels_per_metric, rows_i,
                                                   rows_f, ws, status, num_sms(), st, gd,
                                                   split ? 1 : 0) != cudaSuccess)
    return UWCV_E_LAUNCH;
  return UWCV_OK;
}

int uwcv_paste_measure_heads(const float* masks, int mask_channels, int channel_offset,
                             int is_logits, const float* boxes, const int32_t* image_idx,
                             const int32_t* inst_idx, const int64_t* classes, const float* scores,
                             int64_t N, int H, int W, float thr, double pixels_per_metric,
                             uint32_t* bitplanes, int64_t* rows_i, double* rows_f, void* workspace,
                             size_t ws_bytes, int64_t* status, void* stream, int stages,
                             int64_t first, int64_t count) {
  return uwcv_paste_measure_gather(masks, mask_channels, channel_offset, is_logits, boxes, image_idx,
                                   inst_idx, classes, scores, N, H, W, thr, pixels_per_metric,
                                   bitplanes, rows_i, rows_f, workspace, ws_bytes, status, stream,
                                   stages, first, count, nullptr);
}

int uwcv_paste_measure_range(const float* masks, const float* boxes, const int32_t* image_idx,
                             const int32_t* inst_idx, const int64_t* classes, const float* scores,
                             int64_t N, int H, int W, float thr, double pixels_per_metric,
                             uint32_t* bitplanes, int64_t* rows_i, double* rows_f, void* workspace,
                             size_t ws_bytes, int64_t* status, void* stream, int stages,
                             int64_t first, int64_t count) {
  return uwcv_paste_measure_heads(masks, 1, 0, 0, boxes, image_idx, inst_idx, classes, scores, N, H,
                                  W, thr, pixels_per_metric, bitplanes, rows_i, rows_f, workspace,
                                  ws_bytes, status, stream, stages, first, count);
}

int uwcv_unpack_planes(const uint32_t* bitplanes, int64_t N, int H, int W, uint8_t* out,
                       void* stream) {
  if (N < 0 || H <= 0 || W <= 0) return UWCV_E_SHAPE;
  if (N == 0) return UWCV_OK;
  if (!bitplanes || !out) return UWCV_E_NULL;
  if (misaligned(bitplanes)) return UWCV_E_ALIGN;
  return uwcv::launch_unpack(bitplanes, N, H, W, out, num_sms(),
                             reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_mask_column_totals(const void* paste_workspace, size_t ws_bytes, int64_t N, int W,
                            const int32_t* image_slot, int32_t* column_totals, void* stream) {
  if (N < 0 || W <= 0) return UWCV_E_SHAPE;
  if (N == 0) return UWCV_OK;
  if (!paste_workspace || !image_slot || !column_totals) return UWCV_E_NULL;
  if (misaligned(paste_workspace)) return UWCV_E_ALIGN;
  if (ws_bytes < uwcv::workspace_bytes(N, 4)) return UWCV_E_WORKSPACE;
  const uwcv::Workspace ws = uwcv::carve(const_cast<void*>(paste_workspace), ws_bytes, N);
  return uwcv::launch_column_totals(N, W, ws, image_slot, column_totals,
                                    reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_clean_masks(void* paste_workspace, size_t ws_bytes, int64_t N, int H, int W,
                     const int32_t* image_slot, const int32_t* inst_idx, const int32_t* limit,
                     int32_t* flags, int64_t* area, int64_t* run_counts, void* stream) {
  if (N < 0 || H <= 0 || W <= 0) return UWCV_E_SHAPE;
  if (H > 32768 || W > 32768) return UWCV_E_TOO_LARGE;
  if (N == 0) return UWCV_OK;
  if (!paste_workspace || !image_slot || !inst_idx || !flags || !area || !run_counts)
    return UWCV_E_NULL;
  if (misaligned(paste_workspace)) return UWCV_E_ALIGN;
  if (ws_bytes < uwcv::workspace_bytes(N, 4)) return UWCV_E_WORKSPACE;
  const uwcv::Workspace ws = uwcv::carve(paste_workspace, ws_bytes, N);
  return uwcv::launch_clean(N, H, W, ws, image_slot, inst_idx, limit, flags, area, run_counts,
                            reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_rle_write(const void* paste_workspace, size_t ws_bytes, int64_t N, int H, int W,
                   const int64_t* run_offsets, int64_t* runs, void* stream) {
  if (N < 0 || H <= 0 || W <= 0) return UWCV_E_SHAPE;
  if (N == 0) return UWCV_OK;
  if (!paste_workspace || !run_offsets || !runs) return UWCV_E_NULL;
  if (misaligned(paste_workspace)) return UWCV_E_ALIGN;
  if (ws_bytes < uwcv::workspace_bytes(N, 4)) return UWCV_E_WORKSPACE;
  const uwcv::Workspace ws = uwcv::carve(const_cast<void*>(paste_workspace), ws_bytes, N);
  return uwcv::launch_rle_write(N, H, W, ws, run_offsets, runs,
                                reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_rle_text_prep(int64_t total_runs, const int64_t* runs, const int32_t* run_instance,
                       int64_t* chars, void* stream) {
  if (total_runs < 0) return UWCV_E_SHAPE;
  if (total_runs == 0) return UWCV_OK;
  if (!runs || !run_instance || !chars) return UWCV_E_NULL;
  return uwcv::launch_rle_text_prep(total_runs, runs, run_instance, chars,
                                    reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_rle_text_write(int64_t total_runs, const int64_t* runs, const int32_t* run_instance,
                        const int64_t* text_offsets, uint8_t* text, void* stream) {
  if (total_runs < 0) return UWCV_E_SHAPE;
  if (total_runs == 0) return UWCV_OK;
  if (!runs || !run_instance || !text_offsets || !text) return UWCV_E_NULL;
  return uwcv::launch_rle_text_write(total_runs, runs, run_instance, text_offsets, text,
                                     reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_ingest(const void* src_host_mapped, void* dst, size_t bytes, void* stream) {
  if (bytes == 0) return UWCV_OK;
  if (!src_host_mapped || !dst) return UWCV_E_NULL;
  if (misaligned(src_host_mapped) || misaligned(dst)) return UWCV_E_ALIGN;
  if (bytes % 16) return UWCV_E_SHAPE;                 // whole 16-byte words: nothing is read or written past the buffers
  return uwcv::launch_ingest(src_host_mapped, dst, bytes, reinterpret_cast<cudaStream_t>(stream)) ==
                 cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_mask_pixel_boxes(const uint8_t* masks, int64_t N, int H, int W, float* boxes, void* stream) {
  if (N < 0 || H <= 0 || W <= 0) return UWCV_E_SHAPE;
  if (H > 32768 || W > 32768) return UWCV_E_TOO_LARGE;
  if (N == 0) return UWCV_OK;
  if (!masks || !boxes) return UWCV_E_NULL;
  return uwcv::launch_pixel_boxes(masks, N, H, W, boxes, reinterpret_cast<cudaStream_t>(stream)) ==
                 cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_pack_mask_tiles(const uint8_t* masks, int64_t N, int H, int W, void* paste_workspace,
                         size_t ws_bytes, void* stream) {
  if (N < 0 || H <= 0 || W <= 0) return UWCV_E_SHAPE;
  if (N == 0) return UWCV_OK;
  if (!masks || !paste_workspace) return UWCV_E_NULL;
  if (misaligned(paste_workspace)) return UWCV_E_ALIGN;
  if (ws_bytes < uwcv::workspace_bytes(N, 4)) return UWCV_E_WORKSPACE;
  const uwcv::Workspace ws = uwcv::carve(paste_workspace, ws_bytes, N);
  return uwcv::launch_pack_tiles(masks, N, H, W, ws, reinterpret_cast<cudaStream_t>(stream)) ==
                 cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_tiles_to_masks(const void* paste_workspace, size_t ws_bytes, int64_t N, int H, int W,
                        uint8_t* out, void* stream) {
  if (N < 0 || H <= 0 || W <= 0) return UWCV_E_SHAPE;
  if (N == 0) return UWCV_OK;
  if (!paste_workspace || !out) return UWCV_E_NULL;
  if (misaligned(paste_workspace)) return UWCV_E_ALIGN;
  if (ws_bytes < uwcv::workspace_bytes(N, 4)) return UWCV_E_WORKSPACE;
  const uwcv::Workspace ws = uwcv::carve(const_cast<void*>(paste_workspace), ws_bytes, N);
  return uwcv::launch_tiles_to_masks(N, H, W, ws, out, reinterpret_cast<cudaStream_t>(stream)) ==
                 cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

size_t uwcv_union_workspace_bytes(int64_t rec_cap, int64_t ext_rows_cap) {
  if (rec_cap < 0) rec_cap = 0;
  if (ext_rows_cap < 0) ext_rows_cap = 0;
  return uwcv::union_workspace_bytes_host(rec_cap, ext_rows_cap);
}

static int union_measure_impl(const void* paste_workspace, size_t paste_ws_bytes, int64_t N,
                              const int32_t* member_group, const uwcv_tile* group_desc,
                              const int32_t* group_image, int64_t G, uint32_t* group_planes,
                              int64_t group_words, void* rec_workspace, size_t rec_ws_bytes,
                              int64_t rec_cap, int64_t ext_rows_cap, double pixels_per_metric,
                              int64_t* rows_i, double* rows_f, int64_t* counters, void* stream,
                              const int64_t* group_counters) {
  static_assert(sizeof(uwcv_tile) == sizeof(uwcv::TileDesc), "uwcv_tile mirrors TileDesc");
  if (N < 0 || G < 0 || group_words < 0 || rec_cap < 0 || ext_rows_cap < 0) return UWCV_E_SHAPE;
  if (!(pixels_per_metric > 0.0)) return UWCV_E_SHAPE;
  if (!counters) return UWCV_E_NULL;
  if (N > 0 && G > 0) {
    if (!paste_workspace || !member_group || !group_desc || !group_image || !group_planes ||
        !rec_workspace || !rows_i || !rows_f)
      return UWCV_E_NULL;
    if (misaligned(paste_workspace) || misaligned(group_desc) || misaligned(group_planes) ||
        misaligned(rec_workspace) || misaligned(rows_i) || misaligned(rows_f))
      return UWCV_E_ALIGN;
    if (rec_ws_bytes < uwcv::union_workspace_bytes_host(rec_cap, ext_rows_cap))
      return UWCV_E_WORKSPACE;
    if (paste_ws_bytes < uwcv::workspace_bytes(N, 4)) return UWCV_E_WORKSPACE;
  }
  const uwcv::Workspace ws = uwcv::carve(const_cast<void*>(paste_workspace), paste_ws_bytes, N);
  return uwcv::launch_union(N, ws, member_group,
                            reinterpret_cast<const uwcv::TileDesc*>(group_desc), group_image, G,
                            group_planes, group_words, rec_workspace, rec_cap, ext_rows_cap,
                            pixels_per_metric, rows_i, rows_f, counters, num_sms(),
                            reinterpret_cast<cudaStream_t>(stream), group_counters) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_union_measure(const void* paste_workspace, size_t paste_ws_bytes, int64_t N,
                       const int32_t* member_group, const uwcv_tile* group_desc,
                       const int32_t* group_image, int64_t G, uint32_t* group_planes,
                       int64_t group_words, void* rec_workspace, size_t rec_ws_bytes,
                       int64_t rec_cap, int64_t ext_rows_cap, double pixels_per_metric,
                       int64_t* rows_i, double* rows_f, int64_t* counters, void* stream) {
  return union_measure_impl(paste_workspace, paste_ws_bytes, N, member_group, group_desc, group_image, G,
                            group_planes, group_words, rec_workspace, rec_ws_bytes, rec_cap, ext_rows_cap,
                            pixels_per_metric, rows_i, rows_f, counters, stream, nullptr);
}

size_t uwcv_union_group_workspace_bytes(int64_t N, int B) {
  return uwcv::union_group_workspace_bytes_host(N < 0 ? 0 : N, B < 0 ? 0 : B);
}

int uwcv_union_group(const int64_t* paste_rows_i, int64_t N, int B, int64_t image_base, void* group_workspace,
                     size_t group_ws_bytes, int32_t* member_group, uwcv_tile* group_desc,
                     int32_t* group_image, int64_t group_words_cap, int64_t* group_counters, void* stream) {
  if (N < 0 || B < 0 || group_words_cap < 0) return UWCV_E_SHAPE;
  if (!group_counters) return UWCV_E_NULL;
  if (N > 0 && B > 0) {
    if (!paste_rows_i || !group_workspace || !member_group || !group_desc || !group_image) return UWCV_E_NULL;
    if (misaligned(paste_rows_i) || misaligned(group_workspace) || misaligned(group_desc)) return UWCV_E_ALIGN;
    if (B > 65535) return UWCV_E_TOO_LARGE;
    if (group_ws_bytes < uwcv::union_group_workspace_bytes_host(N, B)) return UWCV_E_WORKSPACE;
  }
  return uwcv::launch_union_group(paste_rows_i, N, B, image_base, group_workspace, member_group,
                                  reinterpret_cast<uwcv::TileDesc*>(group_desc), group_image, group_words_cap,
                                  group_counters, reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

int uwcv_union_measure_grouped(const void* paste_workspace, size_t paste_ws_bytes, int64_t N,
                               const int32_t* member_group, const uwcv_tile* group_desc,
                               const int32_t* group_image, const int64_t* group_counters,
                               uint32_t* group_planes, int64_t group_words_cap, void* rec_workspace,
                               size_t rec_ws_bytes, int64_t rec_cap, int64_t ext_rows_cap,
                               double pixels_per_metric, int64_t* rows_i, double* rows_f,
                               int64_t* counters, void* stream) {
  if (!group_counters) return UWCV_E_NULL;
  return union_measure_impl(paste_workspace, paste_ws_bytes, N, member_group, group_desc, group_image,
                            /*G capacity*/ N, group_planes, group_words_cap, rec_workspace, rec_ws_bytes,
                            rec_cap, ext_rows_cap, pixels_per_metric, rows_i, rows_f, counters, stream,
                            group_counters);
}

size_t uwcv_nms_workspace_bytes(const int64_t* image_off, int B, int num_classes) {
  if (!image_off || B <= 0) return 256;
  if (num_classes <= 0) num_classes = 128;
  return uwcv::nms_workspace_bytes_host(image_off, B, num_classes);
}

int uwcv_nms_filter(const float* boxes, const float* scores, const int64_t* classes,
                    const int64_t* image_off, int B, int num_classes, float score_thr,
                    double iou_thr, int topk, int64_t* keep, int32_t* keep_count, void* workspace,
                    size_t ws_bytes, void* stream) {
  if (B < 0) return UWCV_E_SHAPE;
  if (num_classes <= 0) num_classes = 128;
  if (num_classes > 8192) return UWCV_E_TOO_LARGE;
  if (B == 0) return UWCV_OK;
  if (!image_off || !keep_count) return UWCV_E_NULL;
  if (image_off[0] != 0) return UWCV_E_SHAPE;
  for (int b = 0; b < B; ++b) {
    const int64_t n = image_off[b + 1] - image_off[b];
    if (n < 0) return UWCV_E_SHAPE;
    if (n > 262144) return UWCV_E_TOO_LARGE;
  }
  if (B > 65535) return UWCV_E_TOO_LARGE;
  const int64_t R = image_off[B];
  if (R > 0 && (!boxes || !scores || !classes || !keep)) return UWCV_E_NULL;
  if (!workspace) return UWCV_E_NULL;
  if (misaligned(boxes) || misaligned(workspace)) return UWCV_E_ALIGN;
  if (ws_bytes < uwcv::nms_workspace_bytes_host(image_off, B, num_classes)) return UWCV_E_WORKSPACE;
  if (topk < 0) topk = 0x7fffffff;
  return uwcv::launch_nms(boxes, scores, classes, image_off, B, num_classes, score_thr, iou_thr, topk, keep,
                          keep_count, workspace, reinterpret_cast<cudaStream_t>(stream)) ==
                 cudaSuccess
             ? UWCV_OK : UWCV_E_LAUNCH;
}

}  // extern "C"
