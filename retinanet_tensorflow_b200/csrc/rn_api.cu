// Library plumbing, pyramid description (host) and anchor generation.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>

#include "rn_common.cuh"
#include "rn_tiles.cuh"

namespace rn {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// One level's constants, following dataset.position_grid (dataset.py:16-20) literally:
// cell = to_float(1 / size); linspace(cell/2, 1 - cell/2, size).
static void fill_axis(int n, float* start, float* step) {
  const float cell = (float)(1.0 / (double)n);
  const float s = cell / 2.0f;
  const float e = 1.0f - cell / 2.0f;
  *start = s;
  *step = n > 1 ? (e - s) / (float)(n - 1) : 0.0f;
}

static int finish_pyramid(rn_pyramid_t* p) {
  long long off = 0;
  int tiles = 0;
  for (int l = 0; l < p->num_levels; ++l) {
    rn_level_t& L = p->level[l];
    RN_REQUIRE(L.grid_h > 0 && L.grid_w > 0, RN_ERR_BAD_SHAPE, "level %d has an empty grid", l);
    fill_axis(L.grid_h, &L.y0, &L.y_step);
    fill_axis(L.grid_w, &L.x0, &L.x_step);
    L.max_half_h = 0.f;
    L.max_half_w = 0.f;
    for (int a = 0; a < L.num_anchors; ++a) {
      L.max_half_h = L.anchor_h[a] * 0.5f > L.max_half_h ? L.anchor_h[a] * 0.5f : L.max_half_h;
      L.max_half_w = L.anchor_w[a] * 0.5f > L.max_half_w ? L.anchor_w[a] * 0.5f : L.max_half_w;
    }
    L.tiles_x = (L.grid_w + kTileW - 1) / kTileW;
    L.tiles_y = (L.grid_h + kTileH - 1) / kTileH;
    L.tile_start = tiles;
    tiles += L.tiles_x * L.tiles_y;
    L.anchor_offset = off;
    off += (long long)L.grid_h * L.grid_w * L.num_anchors;
  }
  RN_REQUIRE(off < (1ll << 31), RN_ERR_BAD_SHAPE, "more than 2^31 anchors per image");
  p->anchors_per_image = off;
  p->tiles_per_image = tiles;
  return RN_OK;
}

}  // namespace rn

using namespace rn;

extern "C" {

int rn_version(void) { return RN_VERSION; }
const char* rn_last_error(void) { return g_error; }
int64_t rn_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

const char* rn_status_string(int status) {
  switch (status) {
    case RN_OK: return "ok";
    case RN_ERR_BAD_ARG: return "bad argument";
    case RN_ERR_BAD_SHAPE: return "bad shape";
    case RN_ERR_CAPACITY: return "capacity exceeded";
    case RN_ERR_CUDA: return "CUDA error";
    case RN_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

int rn_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  RN_REQUIRE(e == cudaSuccess, RN_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  return RN_OK;
}

int rn_pyramid_build(int image_h, int image_w, int num_levels, const int32_t* strides,
                     const double* anchor_sizes_px, int num_anchors, int size_mode, rn_pyramid_t* out) {
  RN_REQUIRE(out && strides && anchor_sizes_px, RN_ERR_BAD_ARG, "rn_pyramid_build: NULL argument");
  RN_REQUIRE(image_h > 0 && image_w > 0, RN_ERR_BAD_SHAPE, "image size must be positive");
  RN_REQUIRE(num_levels >= 1 && num_levels <= RN_MAX_LEVELS, RN_ERR_BAD_SHAPE, "num_levels must be in 1..%d",
             RN_MAX_LEVELS);
  RN_REQUIRE(num_anchors >= 1 && num_anchors <= RN_MAX_ANCHORS, RN_ERR_BAD_SHAPE, "num_anchors must be in 1..%d",
             RN_MAX_ANCHORS);
  RN_REQUIRE(size_mode == RN_SIZE_TF1_INT_TRUNC || size_mode == RN_SIZE_FLOAT, RN_ERR_BAD_ARG, "unknown size_mode");
  memset(out, 0, sizeof(*out));
  out->num_levels = num_levels;
  out->image_h = image_h;
  out->image_w = image_w;
  for (int l = 0; l < num_levels; ++l) {
    RN_REQUIRE(strides[l] > 0, RN_ERR_BAD_ARG, "stride must be positive");
    rn_level_t& L = out->level[l];
    // tf.to_int32(tf.ceil(image_size / factor))  (dataset.py:55): float64 true division
    L.grid_h = (int32_t)std::ceil((double)image_h / (double)strides[l]);
    L.grid_w = (int32_t)std::ceil((double)image_w / (double)strides[l]);
    L.num_anchors = num_anchors;
    for (int a = 0; a < num_anchors; ++a) {
      double h = anchor_sizes_px[((size_t)l * num_anchors + a) * 2 + 0];
      double w = anchor_sizes_px[((size_t)l * num_anchors + a) * 2 + 1];
      if (size_mode == RN_SIZE_TF1_INT_TRUNC) {  // ndarray.astype(int32) then f64 divide (dataset.py:53)
        h = std::trunc(h);
        w = std::trunc(w);
      }
      L.anchor_h[a] = (float)(h / (double)image_h);
      L.anchor_w[a] = (float)(w / (double)image_w);
    }
  }
  return finish_pyramid(out);
}

int rn_pyramid_from_grids(int num_levels, const int32_t* grid_hw, const float* anchor_sizes_norm, int num_anchors,
                          rn_pyramid_t* out) {
  RN_REQUIRE(out && grid_hw && anchor_sizes_norm, RN_ERR_BAD_ARG, "rn_pyramid_from_grids: NULL argument");
  RN_REQUIRE(num_levels >= 1 && num_levels <= RN_MAX_LEVELS, RN_ERR_BAD_SHAPE, "num_levels must be in 1..%d",
             RN_MAX_LEVELS);
  RN_REQUIRE(num_anchors >= 1 && num_anchors <= RN_MAX_ANCHORS, RN_ERR_BAD_SHAPE, "num_anchors must be in 1..%d",
             RN_MAX_ANCHORS);
  memset(out, 0, sizeof(*out));
  out->num_levels = num_levels;
  for (int l = 0; l < num_levels; ++l) {
    rn_level_t& L = out->level[l];
    L.grid_h = grid_hw[2 * l];
    L.grid_w = grid_hw[2 * l + 1];
    L.num_anchors = num_anchors;
    for (int a = 0; a < num_anchors; ++a) {
      L.anchor_h[a] = anchor_sizes_norm[((size_t)l * num_anchors + a) * 2 + 0];
      L.anchor_w[a] = anchor_sizes_norm[((size_t)l * num_anchors + a) * 2 + 1];
    }
  }
  return finish_pyramid(out);
}

int rn_rows_from_pyramid(const rn_pyramid_t* pyr, int batch, rn_rows_t* out) {
  RN_REQUIRE(pyr && out, RN_ERR_BAD_ARG, "rn_rows_from_pyramid: NULL argument");
  RN_REQUIRE(batch >= 0, RN_ERR_BAD_ARG, "negative batch");
  memset(out, 0, sizeof(*out));
  out->num_segments = pyr->num_levels;
  out->images = batch;
  for (int l = 0; l < pyr->num_levels; ++l) {
    const rn_level_t& L = pyr->level[l];
    out->rows[l] = (int64_t)batch * L.grid_h * L.grid_w * L.num_anchors;
  }
  return RN_OK;
}

int rn_maps_from_merged(const rn_pyramid_t* pyr, int batch, void* d_base, size_t bytes_per_anchor, rn_maps_t* out) {
  RN_REQUIRE(pyr && out, RN_ERR_BAD_ARG, "rn_maps_from_merged: NULL argument");
  memset(out, 0, sizeof(*out));
  if (d_base == nullptr) return RN_OK;
  char* base = (char*)d_base;
  for (int l = 0; l < pyr->num_levels; ++l) {
    out->p[l] = base + (size_t)batch * (size_t)pyr->level[l].anchor_offset * bytes_per_anchor;
  }
  return RN_OK;
}

}  // extern "C"

// ---- anchor generation ------------------------------------------------------------------------
namespace rn {

__global__ void anchor_generate_kernel(const __grid_constant__ rn_pyramid_t pyr, float4* __restrict__ boxes,
                                       float4* __restrict__ centres) {
  const long long total = pyr.anchors_per_image;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int l = 0;
    for (int k = 1; k < pyr.num_levels; ++k) l += (i >= pyr.level[k].anchor_offset) ? 1 : 0;
    const rn_level_t& L = pyr.level[l];
    const int local = (int)(i - L.anchor_offset);
    const int a = local % L.num_anchors;
    const int cell = local / L.num_anchors;
    const int y = cell / L.grid_w, x = cell % L.grid_w;
    const float cy = cell_centre(L.y0, L.y_step, y);
    const float cx = cell_centre(L.x0, L.x_step, x);
    const float h = L.anchor_h[a], w = L.anchor_w[a];
    if (centres) centres[i] = make_float4(cy, cx, h, w);
    if (boxes) {
      const float hh = h * 0.5f, hw = w * 0.5f;  // size / 2 is exact
      boxes[i] = make_float4(fsub(cy, hh), fsub(cx, hw), fadd(cy, hh), fadd(cx, hw));
    }
  }
}

}  // namespace rn

extern "C" int rn_anchor_generate(const rn_pyramid_t* pyr, float* d_boxes, float* d_centres, void* stream) {
  RN_REQUIRE(pyr, RN_ERR_BAD_ARG, "rn_anchor_generate: NULL pyramid");
  RN_REQUIRE(d_boxes || d_centres, RN_ERR_BAD_ARG, "rn_anchor_generate: no output");
  const long long total = pyr->anchors_per_image;
  if (total == 0) return RN_OK;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)rn::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  rn::anchor_generate_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*pyr, (float4*)d_boxes,
                                                                                     (float4*)d_centres);
  RN_CUDA_CHECK_LAUNCH("rn_anchor_generate");
  return RN_OK;
}
