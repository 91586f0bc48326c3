// Order-preserving row compaction: tf.boolean_mask(x, trainable_mask) per level + concat over levels
// (utils.postprocess_and_mask, utils.py:258-284; utils.merge_outputs, utils.py:151-153).
//
// The reference materialises the masked `[M, K]` tensors of every head output; the fused path
// (rn_head_loss) never does, but the drop-in `utils.process_labels_and_logits` has to return them.
// Rows are kept in (level, b, y, x, a) order.  Three small kernels:
//   mask_count_kernel   kept rows per chunk of kChunk rows (reads only the 1-byte masks)
//   mask_scan_kernel    exclusive prefix over the chunks (one CTA) + total M
//   mask_gather_kernel  one CTA per chunk: ranks of its kept rows (block scan), then a coalesced copy of those rows
//                       in 16-byte (or 4-byte) units straight to their final position
// The scan result is reusable: every tensor that shares the mask (labels, logits, probabilities, regressions)
// is one more gather launch.  HBM traffic of a gather = kept bytes read + written (+ the masks).
#include "rn_common.cuh"

namespace rn {

constexpr int kChunk = 1024;      // rows per chunk
constexpr int kCompactThreads = 256;

struct ChunkMap {
  unsigned start[RN_MAX_LEVELS + 1];   // first chunk of each segment
};

__device__ __forceinline__ int chunk_segment(const Segs& segs, const ChunkMap& cm, unsigned chunk) {
  int s = 0;
#pragma unroll
  for (int i = 1; i < RN_MAX_LEVELS; ++i) s += (i < segs.num && chunk >= cm.start[i]) ? 1 : 0;
  return s;
}

__global__ void __launch_bounds__(kCompactThreads)
mask_count_kernel(const __grid_constant__ Segs segs, const __grid_constant__ ChunkMap cm, const __grid_constant__ Ptrs mask,
                  unsigned* __restrict__ chunk_count) {
  __shared__ unsigned s_cnt[kCompactThreads / 32];
  const unsigned chunk = blockIdx.x;
  const int s = chunk_segment(segs, cm, chunk);
  const long long row0 = (long long)(chunk - cm.start[s]) * kChunk;
  const long long left = segs.rows[s] - row0;
  const int n = left < kChunk ? (int)left : kChunk;
  const unsigned char* m = reinterpret_cast<const unsigned char*>(mask.p[s]) + row0;
  unsigned c = 0;
  for (int r = threadIdx.x; r < n; r += kCompactThreads) c += m[r] != 0;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < kCompactThreads / 32; ++w) t += s_cnt[w];
    chunk_count[chunk] = t;
  }
}

// exclusive scan of chunk_count in place (-> chunk offsets), total to *d_total
__global__ void __launch_bounds__(1024) mask_scan_kernel(unsigned n_chunks, unsigned* __restrict__ chunk, long long* d_total) {
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned base = 0; base < n_chunks; base += 1024) {
    const unsigned i = base + threadIdx.x;
    const unsigned v = i < n_chunks ? chunk[i] : 0u;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const unsigned before = s_carry + (warp ? s_warp[warp - 1] : 0u) + (x - v);
    if (i < n_chunks) chunk[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *d_total = (long long)s_carry;
}

template <typename Unit>
__global__ void __launch_bounds__(kCompactThreads)
mask_gather_kernel(const __grid_constant__ Segs segs, const __grid_constant__ ChunkMap cm, const __grid_constant__ Ptrs mask,
                   const __grid_constant__ Ptrs src, const unsigned* __restrict__ chunk_offset, Unit* __restrict__ out,
                   unsigned units_per_row) {
  __shared__ unsigned short s_list[kChunk];   // kept rows of this chunk (row-in-chunk), in order
  __shared__ unsigned s_warp[kCompactThreads / 32];
  const unsigned chunk = blockIdx.x;
  const int s = chunk_segment(segs, cm, chunk);
  const long long row0 = (long long)(chunk - cm.start[s]) * kChunk;
  const long long left = segs.rows[s] - row0;
  const int n = left < kChunk ? (int)left : kChunk;
  const unsigned char* m = reinterpret_cast<const unsigned char*>(mask.p[s]) + row0;
  // each thread owns kChunk / threads consecutive rows, so ranks follow row order
  constexpr int kPer = kChunk / kCompactThreads;
  const int r0 = threadIdx.x * kPer;
  unsigned keep = 0;
#pragma unroll
  for (int k = 0; k < kPer; ++k)
    if (r0 + k < n && m[r0 + k] != 0) keep |= 1u << k;
  const unsigned c = __popc(keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  unsigned before = x - c, total = 0;
  for (int w = 0; w < kCompactThreads / 32; ++w) {
    const unsigned v = s_warp[w];
    before += w < warp ? v : 0u;
    total += v;
  }
#pragma unroll
  for (int k = 0; k < kPer; ++k)
    if (keep & (1u << k)) s_list[before++] = (unsigned short)(r0 + k);
  __syncthreads();
  const Unit* in = reinterpret_cast<const Unit*>(src.p[s]) + (size_t)row0 * units_per_row;
  Unit* o = out + (size_t)chunk_offset[chunk] * units_per_row;
  const unsigned n_units = total * units_per_row;
  if (total == (unsigned)n) {   // nothing dropped in this chunk (the common case): a straight copy
    for (unsigned u = threadIdx.x; u < n_units; u += kCompactThreads) __stcs(o + u, __ldcs(in + u));
  } else {
    for (unsigned u = threadIdx.x; u < n_units; u += kCompactThreads) {
      const unsigned r = u / units_per_row, col = u - r * units_per_row;
      __stcs(o + u, __ldcs(in + (size_t)s_list[r] * units_per_row + col));
    }
  }
}

static ChunkMap make_chunks(const Segs& segs, unsigned long long* total) {
  ChunkMap cm;
  unsigned long long t = 0;
  for (int i = 0; i < RN_MAX_LEVELS; ++i) {
    cm.start[i] = (unsigned)t;
    if (i < segs.num) t += (unsigned long long)((segs.rows[i] + kChunk - 1) / kChunk);
  }
  cm.start[RN_MAX_LEVELS] = (unsigned)t;
  *total = t;
  return cm;
}

}  // namespace rn

using namespace rn;

extern "C" size_t rn_mask_workspace_bytes(const rn_rows_t* rows) {
  if (!rows) return 0;
  unsigned long long chunks = 0;
  for (int i = 0; i < rows->num_segments && i < RN_MAX_LEVELS; ++i) chunks += (unsigned long long)((rows->rows[i] + kChunk - 1) / kChunk);
  return (size_t)(chunks + 1) * sizeof(unsigned);
}

extern "C" int rn_mask_scan(const rn_rows_t* rows, const rn_maps_t* mask, void* d_workspace, size_t workspace_bytes,
                            int64_t* d_count, void* stream) {
  RN_REQUIRE(rows && maps_present(mask) && d_workspace && d_count, RN_ERR_BAD_ARG, "rn_mask_scan: NULL input");
  RN_REQUIRE(rows->num_segments >= 1 && rows->num_segments <= RN_MAX_LEVELS, RN_ERR_BAD_SHAPE, "rn_mask_scan: bad segments");
  RN_REQUIRE(workspace_bytes >= rn_mask_workspace_bytes(rows), RN_ERR_CAPACITY, "rn_mask_scan: workspace too small");
  const Segs segs = make_segs(rows);
  unsigned long long chunks;
  const ChunkMap cm = make_chunks(segs, &chunks);
  RN_REQUIRE(chunks < (1ull << 31) && segs.start[RN_MAX_LEVELS] < (1ll << 32), RN_ERR_UNSUPPORTED, "rn_mask_scan: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* counts = reinterpret_cast<unsigned*>(d_workspace);
  if (chunks > 0) {
    mask_count_kernel<<<(unsigned)chunks, kCompactThreads, 0, st>>>(segs, cm, make_ptrs(mask), counts);
    RN_CUDA_CHECK_LAUNCH("rn_mask_scan(count)");
  }
  mask_scan_kernel<<<1, 1024, 0, st>>>((unsigned)chunks, counts, reinterpret_cast<long long*>(d_count));
  RN_CUDA_CHECK_LAUNCH("rn_mask_scan(scan)");
  return RN_OK;
}

extern "C" int rn_mask_gather(const rn_rows_t* rows, int row_bytes, const rn_maps_t* src, const rn_maps_t* mask,
                              const void* d_workspace, void* d_out, void* stream) {
  RN_REQUIRE(rows && maps_present(src) && maps_present(mask) && d_workspace && d_out, RN_ERR_BAD_ARG,
             "rn_mask_gather: NULL input");
  RN_REQUIRE(row_bytes >= 1, RN_ERR_BAD_ARG, "rn_mask_gather: row_bytes must be >= 1");
  const Segs segs = make_segs(rows);
  unsigned long long chunks;
  const ChunkMap cm = make_chunks(segs, &chunks);
  if (chunks == 0) return RN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned* offs = reinterpret_cast<const unsigned*>(d_workspace);
  bool a16 = row_bytes % 16 == 0 && (((uintptr_t)d_out) & 15) == 0;
  bool a4 = row_bytes % 4 == 0 && (((uintptr_t)d_out) & 3) == 0;
  for (int i = 0; i < rows->num_segments; ++i) {
    if (((uintptr_t)src->p[i]) & 15) a16 = false;
    if (((uintptr_t)src->p[i]) & 3) a4 = false;
  }
  if (a16)
    mask_gather_kernel<uint4><<<(unsigned)chunks, kCompactThreads, 0, st>>>(segs, cm, make_ptrs(mask), make_ptrs(src), offs,
                                                                          reinterpret_cast<uint4*>(d_out), row_bytes / 16);
  else if (a4)
    mask_gather_kernel<uint32_t><<<(unsigned)chunks, kCompactThreads, 0, st>>>(
        segs, cm, make_ptrs(mask), make_ptrs(src), offs, reinterpret_cast<uint32_t*>(d_out), row_bytes / 4);
  else
    mask_gather_kernel<unsigned char><<<(unsigned)chunks, kCompactThreads, 0, st>>>(
        segs, cm, make_ptrs(mask), make_ptrs(src), offs, reinterpret_cast<unsigned char*>(d_out), row_bytes);
  RN_CUDA_CHECK_LAUNCH("rn_mask_gather");
  return RN_OK;
}
