// Micro-benchmark: what does a zero-compute kernel with the fast loss kernel's access pattern reach?
// (4.11 GB read + 4.11 GB written, the COCO batch-64 logits / gradient tensors.)  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu && ./stream_pattern
#include <cstdio>
#include <cuda_runtime.h>

template <int U>   // grid-stride: U 16-byte loads first, then U stores (the loss kernel's main loop, U = 4)
__global__ void __launch_bounds__(256, 4) k_gridstride(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned gs = gridDim.x * blockDim.x;
  unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  for (; u + (U - 1) * gs < n; u += U * gs) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * gs);
#pragma unroll
    for (int k = 0; k < U; ++k) __stcs(out + u + k * gs, v[k]);
  }
  for (; u < n; u += gs) __stcs(out + u, __ldcs(in + u));
}

template <int U>   // each CTA owns a contiguous chunk; per iteration it covers U * 4 KB contiguous
__global__ void __launch_bounds__(256, 4) k_chunked(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned per = (n + gridDim.x - 1) / gridDim.x;
  const unsigned lo = blockIdx.x * per, hi = min(n, lo + per);
  unsigned u = lo + threadIdx.x;
  for (; u + (U - 1) * 256 < hi; u += U * 256) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * 256);
#pragma unroll
    for (int k = 0; k < U; ++k) __stcs(out + u + k * 256, v[k]);
  }
  for (; u < hi; u += 256) __stcs(out + u, __ldcs(in + u));
}

template <int U>   // persistent grid over CTA tiles: a CTA covers U * 4 KB contiguous per iteration, tiles are
                    // handed out round-robin, so the wavefront of all CTAs is one compact region
__global__ void __launch_bounds__(256, 4) k_tiles(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned tile = U * 256;
  const unsigned ntiles = n / tile;
  for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const unsigned u = t * tile + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * 256);
#pragma unroll
    for (int k = 0; k < U; ++k) __stcs(out + u + k * 256, v[k]);
  }
  for (unsigned u = ntiles * tile + blockIdx.x * 256 + threadIdx.x; u < n; u += gridDim.x * 256) __stcs(out + u, __ldcs(in + u));
}

template <int U>   // same tiles, one CTA per tile (non-persistent)
__global__ void __launch_bounds__(256, 4) k_tiles_flat(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned u = blockIdx.x * (U * 256) + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; ++k) if (u + k * 256 < n) v[k] = __ldcs(in + u + k * 256);
#pragma unroll
  for (int k = 0; k < U; ++k) if (u + k * 256 < n) __stcs(out + u + k * 256, v[k]);
}

template <int U>   // persistent, warp-contiguous: a warp covers U * 512 B contiguous per iteration
__global__ void __launch_bounds__(256, 4) k_warptiles(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const unsigned tile = U * 32, ntiles = n / tile;
  for (unsigned t = gw; t < ntiles; t += nw) {
    const unsigned u = t * tile + lane;
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * 32);
#pragma unroll
    for (int k = 0; k < U; ++k) __stcs(out + u + k * 32, v[k]);
  }
}

template <int U>   // T round-robin tiles per CTA, grid = ntiles / T (between persistent and one-tile-per-CTA)
__global__ void __launch_bounds__(256, 4) k_tiles_multi(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n, unsigned T) {
  const unsigned tile = U * 256, ntiles = n / tile;
  for (unsigned j = 0, t = blockIdx.x; j < T && t < ntiles; ++j, t += gridDim.x) {
    const unsigned u = t * tile + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * 256);
#pragma unroll
    for (int k = 0; k < U; ++k) __stcs(out + u + k * 256, v[k]);
  }
}

template <int U>   // persistent, but CTA i first does (i % 4) single units so that the CTAs' load/store phases differ
__global__ void __launch_bounds__(256, 4) k_staggered(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned gs = gridDim.x * blockDim.x;
  unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  for (unsigned j = 0; j < (blockIdx.x & 3u) && u < n; ++j, u += gs) __stcs(out + u, __ldcs(in + u));
  for (; u + (U - 1) * gs < n; u += U * gs) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * gs);
#pragma unroll
    for (int k = 0; k < U; ++k) __stcs(out + u + k * gs, v[k]);
  }
  for (; u < n; u += gs) __stcs(out + u, __ldcs(in + u));
}

template <int U>   // persistent, software pipelined: the next batch's loads are issued before this batch's stores
__global__ void __launch_bounds__(256, 4) k_pipelined(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned gs = gridDim.x * blockDim.x;
  unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u + (U - 1) * gs >= n) {
    for (; u < n; u += gs) __stcs(out + u, __ldcs(in + u));
    return;
  }
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * gs);
  for (;;) {
    const unsigned un = u + U * gs;
    const bool more = un + (U - 1) * gs < n;
    uint4 w[U];
    if (more) {
#pragma unroll
      for (int k = 0; k < U; ++k) w[k] = __ldcs(in + un + k * gs);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) __stcs(out + u + k * gs, v[k]);
    u = un;
    if (!more) break;
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = w[k];
  }
  for (; u < n; u += gs) __stcs(out + u, __ldcs(in + u));
}

template <int U>   // persistent, loads and stores interleaved one by one (store k as soon as load k is back)
__global__ void __launch_bounds__(256, 4) k_interleaved(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned gs = gridDim.x * blockDim.x;
  unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  for (; u + (U - 1) * gs < n; u += U * gs) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * gs);
#pragma unroll
    for (int k = 0; k < U; ++k) {
      __stcs(out + u + k * gs, v[k]);
      asm volatile("" ::: "memory");
    }
  }
  for (; u < n; u += gs) __stcs(out + u, __ldcs(in + u));
}

// persistent warps, dynamic warp tiles: a warp takes G consecutive tiles of U * 512 B per atomic (the next ticket is
// requested before the current one is worked on), so fast SMs simply take more tiles
template <int U, int G>
__global__ void __launch_bounds__(256, 4) k_dynamic(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n,
                                                    unsigned* counter) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned tile = U * 32, ntiles = n / tile;   // (tail units ignored by the benchmark)
  unsigned next = 0;
  if (lane == 0) next = atomicAdd(counter, G);
  next = __shfl_sync(0xffffffffu, next, 0);
  while (next < ntiles) {
    const unsigned t0 = next;
    if (lane == 0) next = atomicAdd(counter, G);
    next = __shfl_sync(0xffffffffu, next, 0);
    const unsigned t1 = min(t0 + G, ntiles);
    for (unsigned t = t0; t < t1; ++t) {
      const unsigned u = t * tile + lane;
      uint4 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) v[k] = __ldcs(in + u + k * 32);
#pragma unroll
      for (int k = 0; k < U; ++k) __stcs(out + u + k * 32, v[k]);
    }
  }
}

// persistent grid-stride 4+4 that records when every CTA finished (globaltimer ns)
__global__ void __launch_bounds__(256, 4) k_gridstride_timed(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n,
                                                             unsigned long long* t_end, unsigned* smid) {
  const unsigned gs = gridDim.x * blockDim.x;
  unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  for (; u + 3 * gs < n; u += 4 * gs) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldcs(in + u + k * gs);
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(out + u + k * gs, v[k]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t;
    unsigned id;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    t_end[blockIdx.x] = t;
    smid[blockIdx.x] = id;
  }
}

__global__ void k_flat(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {   // one thread per 16 B
  const unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n) out[u] = in[u];
}

template <int U>
__global__ void __launch_bounds__(256, 4) k_readonly(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned n) {
  const unsigned gs = gridDim.x * blockDim.x;
  unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned acc = 0;
  for (; u + (U - 1) * gs < n; u += U * gs) {
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint4 v = __ldcs(in + u + k * gs);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345678u) out[0] = make_uint4(acc, 0, 0, 0);
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const unsigned n = 64u * 200700u * 80u / 4u;   // float4 units of the COCO batch-64 logits tensor
  const size_t bytes = (size_t)n * 16;
  uint4 *in, *out;
  cudaMalloc(&in, bytes);
  cudaMalloc(&out, bytes);
  cudaMemset(in, 1, bytes);
  cudaMemset(out, 0, bytes);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const unsigned g = sms * 4;
  const int reps = 20;
  auto report = [&](const char* name, float ms, double gb) { printf("%-44s %.3f ms  %.0f GB/s\n", name, ms, gb / (ms * 1e-3) / 1e9); };
  const double rw = 2.0 * bytes, ro = 1.0 * bytes;
  report("cudaMemcpyAsync D2D", time_ms([&] { cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice); }, reps), rw);
  report("grid-stride, 4 loads then 4 stores (K2)", time_ms([&] { k_gridstride<4><<<g, 256>>>(in, out, n); }, reps), rw);
  report("grid-stride, 8 loads then 8 stores", time_ms([&] { k_gridstride<8><<<g, 256>>>(in, out, n); }, reps), rw);
  report("grid-stride, 2 loads then 2 stores", time_ms([&] { k_gridstride<2><<<g, 256>>>(in, out, n); }, reps), rw);
  report("grid-stride x4, 8 CTAs/SM grid", time_ms([&] { k_gridstride<4><<<g * 2, 256>>>(in, out, n); }, reps), rw);
  report("chunk per CTA, 4 loads then 4 stores", time_ms([&] { k_chunked<4><<<g, 256>>>(in, out, n); }, reps), rw);
  report("chunk per CTA, 8 loads then 8 stores", time_ms([&] { k_chunked<8><<<g, 256>>>(in, out, n); }, reps), rw);
  report("CTA tiles round-robin, 4+4 (persistent)", time_ms([&] { k_tiles<4><<<g, 256>>>(in, out, n); }, reps), rw);
  report("CTA tiles round-robin, 8+8 (persistent)", time_ms([&] { k_tiles<8><<<g, 256>>>(in, out, n); }, reps), rw);
  report("CTA tiles round-robin, 2+2 (persistent)", time_ms([&] { k_tiles<2><<<g, 256>>>(in, out, n); }, reps), rw);
  report("CTA tiles, one CTA per tile, 4+4", time_ms([&] { k_tiles_flat<4><<<(n + 1023) / 1024, 256>>>(in, out, n); }, reps), rw);
  report("warp tiles round-robin, 4+4 (persistent)", time_ms([&] { k_warptiles<4><<<g, 256>>>(in, out, n); }, reps), rw);
  for (unsigned T : {2u, 4u, 8u, 16u, 32u, 64u}) {
    char name[64];
    snprintf(name, sizeof name, "CTA tiles 4+4, %u tiles per CTA", T);
    const unsigned ntiles = n / 1024;
    report(name, time_ms([&] { k_tiles_multi<4><<<(ntiles + T - 1) / T, 256>>>(in, out, n, T); }, reps), rw);
  }
  for (unsigned V : {2u, 4u, 8u, 16u, 32u, 64u, 128u}) {
    char name[64];
    snprintf(name, sizeof name, "grid-stride 4+4, grid = %u x resident wave", V);
    report(name, time_ms([&] { k_gridstride<4><<<g * V, 256>>>(in, out, n); }, reps), rw);
  }
  unsigned* counter;
  cudaMalloc(&counter, 4);
  report("dynamic warp tiles 5+5, 8 tiles per ticket", time_ms([&] { cudaMemsetAsync(counter, 0, 4); k_dynamic<5, 8><<<g, 256>>>(in, out, n, counter); }, reps), rw);
  report("dynamic warp tiles 5+5, 2 tiles per ticket", time_ms([&] { cudaMemsetAsync(counter, 0, 4); k_dynamic<5, 2><<<g, 256>>>(in, out, n, counter); }, reps), rw);
  report("dynamic warp tiles 5+5, 32 tiles per ticket", time_ms([&] { cudaMemsetAsync(counter, 0, 4); k_dynamic<5, 32><<<g, 256>>>(in, out, n, counter); }, reps), rw);
  report("dynamic warp tiles 4+4, 8 tiles per ticket", time_ms([&] { cudaMemsetAsync(counter, 0, 4); k_dynamic<4, 8><<<g, 256>>>(in, out, n, counter); }, reps), rw);
  report("dynamic warp tiles 2+2, 16 tiles per ticket", time_ms([&] { cudaMemsetAsync(counter, 0, 4); k_dynamic<2, 16><<<g, 256>>>(in, out, n, counter); }, reps), rw);
  {
    unsigned long long* t_end;
    unsigned* smid;
    cudaMalloc(&t_end, g * 8);
    cudaMalloc(&smid, g * 4);
    for (int i = 0; i < 3; ++i) k_gridstride_timed<<<g, 256>>>(in, out, n, t_end, smid);
    cudaDeviceSynchronize();
    unsigned long long* h = new unsigned long long[g];
    unsigned* hs = new unsigned[g];
    cudaMemcpy(h, t_end, g * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, smid, g * 4, cudaMemcpyDeviceToHost);
    unsigned long long lo = ~0ull, hi = 0;
    for (unsigned i = 0; i < g; ++i) { lo = h[i] < lo ? h[i] : lo; hi = h[i] > hi ? h[i] : hi; }
    printf("persistent 4+4: CTA finish times spread %.1f us (first to last); by SM (us after the first):\n", (hi - lo) / 1e3);
    double per_sm[256] = {0};
    for (unsigned i = 0; i < g; ++i) if (hs[i] < 256) per_sm[hs[i]] = per_sm[hs[i]] > (h[i] - lo) / 1e3 ? per_sm[hs[i]] : (h[i] - lo) / 1e3;
    for (int i = 0; i < sms; ++i) printf("%.0f%s", per_sm[i], (i % 37 == 36) ? "\n" : " ");
    printf("\n");
  }
  report("persistent 4+4, staggered start", time_ms([&] { k_staggered<4><<<g, 256>>>(in, out, n); }, reps), rw);
  report("persistent 4+4, software pipelined", time_ms([&] { k_pipelined<4><<<g, 256>>>(in, out, n); }, reps), rw);
  report("persistent 2+2, software pipelined", time_ms([&] { k_pipelined<2><<<g, 256>>>(in, out, n); }, reps), rw);
  report("one thread per 16 B", time_ms([&] { k_flat<<<(n + 255) / 256, 256>>>(in, out, n); }, reps), rw);
  report("read only, grid-stride x4", time_ms([&] { k_readonly<4><<<g, 256>>>(in, out, n); }, reps), ro);
  report("read only, grid-stride x8", time_ms([&] { k_readonly<8><<<g, 256>>>(in, out, n); }, reps), ro);
  printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
