// The two ends of the path on the device:
//   rn_shapes_render   data_loaders/shapes.py:26-53,113-131 -- the synthetic Shapes images (background colour +
//                      up to a few filled squares / circles / triangles, drawn in list order) rasterised straight
//                      into an HBM image batch [B, H, W, 3] u8: no PNG on disk, no decode, no host->device copy
//                      of pixels.  The shape SPECS stay host-generated (Python's Mersenne Twister stream, a handful
//                      of integers per image: retinanet_tensorflow_b200/shapes.py mirrors random_image exactly).
//   rn_draw_boxes      utils.draw_bounding_boxes (utils.py:120-148) / train.py:68-85 -- 1-pixel box outlines in the
//                      class colour drawn over an image batch from the DEVICE detection lists (det_boxes /
//                      det_classes / det_count), for summaries, without bringing detections to the host.  The text
//                      label of the reference (cv2.putText) is not drawn.
//
// Rasterisation rules follow OpenCV's filled primitives as closely as closed forms allow: rectangle = both corners
// inclusive (exact); circle = OpenCV's midpoint-circle spans, reproduced row by row (exact); triangle = scanline
// fill of the int32-truncated vertices with rounded edge crossings, both ends inclusive (agrees with cv2.fillPoly
// except for isolated boundary pixels; tests/test_gpu_shapes.py bounds the difference).
#include "rn_common.cuh"

namespace rn {

constexpr int kMaxShapes = 8;

struct ShapeSpec {       // one shape: kind 0 square, 1 triangle, 2 circle (the reference's class order)
  int kind, x, y, s;
  int r, g, b, pad;
};

// half-width of OpenCV's filled circle on the row at distance |dy| from the centre (cv::circle, thickness < 0):
// the midpoint algorithm walks dx = 0..; the widest span drawn for a row is what the fill leaves.
__device__ __forceinline__ int circle_half_width(int radius, int ady) {
  if (ady > radius) return -1;
  // cv::Circle: err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1; rows y0 +- dy get [x0 - dx, x0 + dx],
  // rows y0 +- dx get [x0 - dy, x0 + dy]; iterate until dx < dy.  Evaluate it for this row.
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  int best = -1;
  while (dx >= dy) {
    if (dy == ady) best = max(best, dx);
    if (dx == ady) best = max(best, dy);
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
    ++dy;
  }
  return best;
}

__global__ void __launch_bounds__(256)
shapes_render_kernel(int batch, int height, int width, const ShapeSpec* __restrict__ specs, const int* __restrict__ counts,
                     const unsigned char* __restrict__ bg, int max_shapes, unsigned char* __restrict__ image) {
  __shared__ ShapeSpec s_spec[kMaxShapes];
  __shared__ int s_tri[kMaxShapes][6];   // truncated triangle vertices (x0,y0,x1,y1,x2,y2)
  const int b = blockIdx.y;
  const int n = min(counts[b], max_shapes);
  if (threadIdx.x < n) {
    const ShapeSpec sp = specs[(size_t)b * max_shapes + threadIdx.x];
    s_spec[threadIdx.x] = sp;
    if (sp.kind == 1) {
      // points = [(x, y - s), (x - s / sin(60 deg), y + s), (x + s / sin(60 deg), y + s)] as int32 (truncation)
      const double k = (double)sp.s / 0.8660254037844386;
      s_tri[threadIdx.x][0] = sp.x;
      s_tri[threadIdx.x][1] = sp.y - sp.s;
      s_tri[threadIdx.x][2] = (int)((double)sp.x - k);
      s_tri[threadIdx.x][3] = sp.y + sp.s;
      s_tri[threadIdx.x][4] = (int)((double)sp.x + k);
      s_tri[threadIdx.x][5] = sp.y + sp.s;
    }
  }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= height * width) return;
  const int py = pix / width, px = pix - py * width;
  int r = bg[3 * b], g = bg[3 * b + 1], bl = bg[3 * b + 2];
  for (int i = 0; i < n; ++i) {   // later shapes paint over earlier ones
    const ShapeSpec sp = s_spec[i];
    bool in = false;
    if (sp.kind == 0) {
      in = px >= sp.x - sp.s && px <= sp.x + sp.s && py >= sp.y - sp.s && py <= sp.y + sp.s;
    } else if (sp.kind == 2) {
      const int hw = circle_half_width(sp.s, abs(py - sp.y));
      in = hw >= 0 && abs(px - sp.x) <= hw;
    } else {
      const int x0 = s_tri[i][0], y0 = s_tri[i][1], x1 = s_tri[i][2], y1 = s_tri[i][3], x2 = s_tri[i][4];
      if (py >= y0 && py <= y1 && y1 > y0) {
        // edges apex->left and apex->right, crossing at this scanline (rounded), base row included
        const double t = (double)(py - y0) / (double)(y1 - y0);
        const int xl = (int)floor((double)x0 + t * (double)(x1 - x0) + 0.5);
        const int xr = (int)floor((double)x0 + t * (double)(x2 - x0) + 0.5);
        in = px >= xl && px <= xr;
      } else if (y1 == y0) {
        in = py == y0 && px >= x1 && px <= x2;
      }
    }
    if (in) {
      r = sp.r;
      g = sp.g;
      bl = sp.b;
    }
  }
  unsigned char* o = image + ((size_t)b * height * width + pix) * 3;
  o[0] = (unsigned char)r;
  o[1] = (unsigned char)g;
  o[2] = (unsigned char)bl;
}

// one warp per (image, detection): walks the rectangle's perimeter
__global__ void __launch_bounds__(256)
draw_boxes_kernel(int batch, int height, int width, int capacity, const float4* __restrict__ boxes,
                  const int32_t* __restrict__ classes, const int32_t* __restrict__ counts,
                  const unsigned char* __restrict__ palette, int num_classes, unsigned char* __restrict__ image) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = warp / capacity, k = warp - b * capacity;
  if (b >= batch || k >= min(counts[b], capacity)) return;
  const float4 bx = boxes[(size_t)b * capacity + k];
  const int c = classes[(size_t)b * capacity + k];
  if (c < 0 || c >= num_classes) return;
  // boxes = (boxes * [H, W, H, W]).round().astype(int32)   (utils.py:125-127; np.round = half to even)
  const int y0 = __float2int_rn(bx.x * (float)height), x0 = __float2int_rn(bx.y * (float)width);
  const int y1 = __float2int_rn(bx.z * (float)height), x1 = __float2int_rn(bx.w * (float)width);
  const unsigned char cr = palette[3 * c], cg = palette[3 * c + 1], cb = palette[3 * c + 2];
  auto put = [&](int y, int x) {
    if (y < 0 || y >= height || x < 0 || x >= width) return;
    unsigned char* o = image + (((size_t)b * height + y) * width + x) * 3;
    o[0] = cr;
    o[1] = cg;
    o[2] = cb;
  };
  const int ya = min(y0, y1), yb = max(y0, y1), xa = min(x0, x1), xb = max(x0, x1);
  for (int x = xa + lane; x <= xb; x += 32) {
    put(ya, x);
    put(yb, x);
  }
  for (int y = ya + lane; y <= yb; y += 32) {
    put(y, xa);
    put(y, xb);
  }
}

}  // namespace rn

using namespace rn;

extern "C" int rn_shapes_render(int batch, int height, int width, int max_shapes, const int32_t* d_specs,
                                const int32_t* d_counts, const uint8_t* d_background, uint8_t* d_image, void* stream) {
  RN_REQUIRE(batch >= 1 && height >= 1 && width >= 1, RN_ERR_BAD_ARG, "rn_shapes_render: bad shape");
  RN_REQUIRE(max_shapes >= 1 && max_shapes <= kMaxShapes, RN_ERR_BAD_ARG, "rn_shapes_render: max_shapes must be in 1..%d",
             kMaxShapes);
  RN_REQUIRE(d_specs && d_counts && d_background && d_image, RN_ERR_BAD_ARG, "rn_shapes_render: NULL pointer");
  RN_REQUIRE((long long)height * width < (1ll << 31) && batch <= 65535, RN_ERR_BAD_SHAPE, "rn_shapes_render: image too large");
  const dim3 grid((unsigned)(((long long)height * width + 255) / 256), (unsigned)batch);
  shapes_render_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(batch, height, width,
                                                              reinterpret_cast<const ShapeSpec*>(d_specs), d_counts,
                                                              d_background, max_shapes, d_image);
  RN_CUDA_CHECK_LAUNCH("rn_shapes_render");
  return RN_OK;
}

extern "C" int rn_draw_boxes(int batch, int height, int width, int capacity, const float* d_boxes,
                             const int32_t* d_classes, const int32_t* d_counts, const uint8_t* d_palette,
                             int num_classes, uint8_t* d_image, void* stream) {
  RN_REQUIRE(batch >= 1 && height >= 1 && width >= 1 && capacity >= 1 && num_classes >= 1, RN_ERR_BAD_ARG,
             "rn_draw_boxes: bad shape");
  RN_REQUIRE(d_boxes && d_classes && d_counts && d_palette && d_image, RN_ERR_BAD_ARG, "rn_draw_boxes: NULL pointer");
  const long long warps = (long long)batch * capacity;
  RN_REQUIRE(warps * 32 < (1ll << 31) * 256, RN_ERR_BAD_SHAPE, "rn_draw_boxes: too many boxes");
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  draw_boxes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(batch, height, width, capacity,
                                                           reinterpret_cast<const float4*>(d_boxes), d_classes, d_counts,
                                                           d_palette, num_classes, d_image);
  RN_CUDA_CHECK_LAUNCH("rn_draw_boxes");
  return RN_OK;
}
