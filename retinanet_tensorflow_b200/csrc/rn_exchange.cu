// Multi-GPU exchange of the statistics vector over NVLink peer memory (the one
// exchange step of the path is an all-reduce(sum) of ~1 KB).  Instead of an NCCL all-reduce per
// phase, every rank owns a "mailbox" in peer-mapped (symmetric) memory:
//
//   flags : [slots][world] uint64   epoch of the last message from rank r in that slot
//   data  : [slots][world][len_max] double
//
// rn_stats_exchange launches ONE small CTA that (1) stores this rank's slice into slot[rank] of
// every peer's mailbox with plain st.global over NVLink, (2) publishes the epoch with a
// system-scope release, (3) spins (bounded) until the flags of all ranks in its OWN mailbox carry
// the epoch, (4) sums the world messages in rank order and writes the result back into the stats
// vector.  The sum order is fixed, so all ranks obtain bit-identical normalisers.  Slots are
// double-buffered by epoch parity: a rank can only be one phase-instance ahead of a peer, so a
// message is never overwritten before it has been read (a rank posts epoch e+1 only after it has
// read every message of epoch e, and nobody can post e+2 before everybody has posted e+1).
//
// The epoch lives in DEVICE memory (one counter per phase, incremented by the kernel itself), so a
// step that contains exchanges can be captured into a CUDA graph and replayed.  A peer that does
// not arrive within the time limit (default 10 minutes; rn_set_option("exchange_timeout_ms"))
// is a FATAL condition: the slice is overwritten with NaN -- every loss and gradient that depends
// on it becomes NaN -- and RN_FLAG_EXCHANGE_TIMEOUT is raised; the exchange never silently falls
// back to local statistics.
#include "rn_common.cuh"

namespace rn {

static long long g_timeout_ms = 600000;   // rn_set_option("exchange_timeout_ms", ...)

struct ExchangeParams {
  double* stats;            // local stats vector (+ offset applied by the host)
  int len, len_max;
  int world, rank;
  int slot, slots;              // slot: base slot of the phase (epoch parity is added by the kernel) when epochs != 0
  unsigned long long epoch;     // host-supplied epoch (epochs == nullptr)
  unsigned long long* epochs;   // device counter of this phase (incremented here), or nullptr
  long long timeout_ns;
  void* const* peers;       // device array [world] of mailbox base pointers (peer-mapped)
  uint32_t* status;
};

__device__ __forceinline__ unsigned long long* mb_flags(void* base, int slots, int world, int slot) {
  return reinterpret_cast<unsigned long long*>(base) + (size_t)slot * world;
}
__device__ __forceinline__ double* mb_data(void* base, int slots, int world, int len_max, int slot, int src) {
  const size_t flag_bytes = (((size_t)slots * world * sizeof(unsigned long long)) + 255) & ~(size_t)255;
  return reinterpret_cast<double*>(reinterpret_cast<char*>(base) + flag_bytes) +
         ((size_t)slot * world + src) * len_max;
}

__device__ __forceinline__ long long now_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256) stats_exchange_kernel(const __grid_constant__ ExchangeParams p) {
  __shared__ int s_timeout;
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) {
    s_timeout = 0;
    s_epoch = p.epochs ? ++(*p.epochs) : p.epoch;
  }
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const int slot = p.epochs ? p.slot + (int)(epoch & 1ull) : p.slot;
  // (1) my message -> every rank's mailbox (including my own)
  for (int dst = 0; dst < p.world; ++dst) {
    double* out = mb_data(p.peers[dst], p.slots, p.world, p.len_max, slot, p.rank);
    for (int i = threadIdx.x; i < p.len; i += blockDim.x) out[i] = p.stats[i];
  }
  __threadfence_system();
  __syncthreads();
  // (2) publish
  if ((int)threadIdx.x < p.world) {
    __threadfence_system();
    unsigned long long* f = mb_flags(p.peers[threadIdx.x], p.slots, p.world, slot) + p.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
  }
  // (3) wait for everybody's message in MY mailbox (>=: a peer that is already one instance ahead in the other
  //     parity slot never rewrites this one before we have read it, see the header comment)
  if ((int)threadIdx.x < p.world) {
    const unsigned long long* f = mb_flags(p.peers[p.rank], p.slots, p.world, slot) + threadIdx.x;
    const long long t0 = now_ns();
    unsigned long long v;
    unsigned spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= epoch) break;
      if ((++spins & 1023u) == 0 && now_ns() - t0 > p.timeout_ns) {
        s_timeout = 1;
        break;
      }
    } while (true);
    __threadfence_system();
  }
  __syncthreads();
  if (s_timeout) {   // fatal: poison the slice so nothing downstream looks valid
    for (int i = threadIdx.x; i < p.len; i += blockDim.x) p.stats[i] = __longlong_as_double(0x7ff8000000000000ll);
    if (threadIdx.x == 0 && p.status) atomicOr(p.status, RN_FLAG_EXCHANGE_TIMEOUT);
    return;
  }
  // (4) reduce in rank order (messages were written by remote GPUs: bypass L1)
  for (int i = threadIdx.x; i < p.len; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < p.world; ++r)
      acc += __ldcg(mb_data(p.peers[p.rank], p.slots, p.world, p.len_max, slot, r) + i);
    p.stats[i] = acc;
  }
}

void set_exchange_timeout_ms(long long ms) { g_timeout_ms = ms > 0 ? ms : 1; }

}  // namespace rn

using namespace rn;

extern "C" size_t rn_stats_mailbox_bytes(int world, int len_max, int slots) {
  if (world < 1 || len_max < 1 || slots < 1) return 0;
  const size_t flag_bytes = (((size_t)slots * world * sizeof(unsigned long long)) + 255) & ~(size_t)255;
  return flag_bytes + (size_t)slots * world * len_max * sizeof(double);
}

extern "C" int rn_stats_exchange(double* d_stats, int offset, int len, int len_max, int world, int rank,
                                 const void* const* d_peer_mailboxes, int slot, int slots, uint64_t epoch,
                                 uint32_t* d_status, void* stream) {
  RN_REQUIRE(d_stats && d_peer_mailboxes, RN_ERR_BAD_ARG, "rn_stats_exchange: NULL pointer");
  RN_REQUIRE(world >= 1 && rank >= 0 && rank < world, RN_ERR_BAD_ARG, "rn_stats_exchange: bad rank/world");
  RN_REQUIRE(world <= 256, RN_ERR_UNSUPPORTED, "rn_stats_exchange: world > 256");
  RN_REQUIRE(len >= 1 && len <= len_max && offset >= 0, RN_ERR_BAD_ARG, "rn_stats_exchange: bad slice");
  RN_REQUIRE(slot >= 0 && slot < slots && epoch > 0, RN_ERR_BAD_ARG, "rn_stats_exchange: bad slot/epoch");
  ExchangeParams p;
  p.epochs = nullptr;
  p.timeout_ns = g_timeout_ms * 1000000ll;
  p.stats = d_stats + offset;
  p.len = len;
  p.len_max = len_max;
  p.world = world;
  p.rank = rank;
  p.slot = slot;
  p.slots = slots;
  p.epoch = epoch;
  p.peers = const_cast<void* const*>(d_peer_mailboxes);
  p.status = d_status;
  stats_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p);
  RN_CUDA_CHECK_LAUNCH("rn_stats_exchange");
  return RN_OK;
}

extern "C" int rn_stats_exchange_auto(double* d_stats, int offset, int len, int len_max, int world, int rank,
                                      const void* const* d_peer_mailboxes, int phase, int phases, uint64_t* d_epochs,
                                      uint32_t* d_status, void* stream) {
  RN_REQUIRE(d_stats && d_peer_mailboxes && d_epochs, RN_ERR_BAD_ARG, "rn_stats_exchange_auto: NULL pointer");
  RN_REQUIRE(world >= 1 && rank >= 0 && rank < world, RN_ERR_BAD_ARG, "rn_stats_exchange_auto: bad rank/world");
  RN_REQUIRE(world <= 256, RN_ERR_UNSUPPORTED, "rn_stats_exchange_auto: world > 256");
  RN_REQUIRE(len >= 1 && len <= len_max && offset >= 0, RN_ERR_BAD_ARG, "rn_stats_exchange_auto: bad slice");
  RN_REQUIRE(phase >= 0 && phase < phases, RN_ERR_BAD_ARG, "rn_stats_exchange_auto: bad phase");
  ExchangeParams p;
  p.stats = d_stats + offset;
  p.len = len;
  p.len_max = len_max;
  p.world = world;
  p.rank = rank;
  p.slot = 2 * phase;
  p.slots = 2 * phases;
  p.epoch = 0;
  p.epochs = reinterpret_cast<unsigned long long*>(d_epochs) + phase;
  p.timeout_ns = g_timeout_ms * 1000000ll;
  p.peers = const_cast<void* const*>(d_peer_mailboxes);
  p.status = d_status;
  stats_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p);
  RN_CUDA_CHECK_LAUNCH("rn_stats_exchange_auto");
  return RN_OK;
}
