// One-shot all-reduce over NVLink peer memory, fused with the squared-norm pass of the gradient clip
// (the data-parallel exchanges of one training iteration are <= 361 KB each --
// three flat gradient buffers, the BatchNorm statistic sums of the two scene CNNs, the 36 x 36 patch statistics and the
// per-generator draw counts -- i.e. latency-bound; the reference has no multi-GPU path at all).
//
// Every rank owns one region of a SYMMETRIC arena (same offset on every rank, mapped into every peer's address space:
// torch.distributed._symmetric_memory, i.e. cuMemCreate + fabric handles over NVSwitch).  One launch per rank does
//   1. copy the local operand into the rank's own region (skipped when a packing kernel already wrote it there),
//   2. a flag barrier between block b of all ranks (release / acquire at system scope through peer atomics),
//   3. out[i] = sum over ranks r = 0 .. W-1 of region_r[i], read straight from the peers' HBM through NVLink in FIXED
//      rank order -- so every rank computes bit-identical sums (replicated weights stay replicated) --
//   4. (fp32) accumulates sum out[i]^2 into a device double: the global gradient norm that mggan_clip_adamw reads.
// Block b writes and reads the same element range on every rank, so the block-level barrier is all the ordering needed.
// There is no exit barrier: the host hands every call of an iteration its own region (bump allocation, reset per
// iteration), so a region is rewritten one iteration later at the earliest, behind the entry barriers of all the calls in
// between.  The flag protocol (0 -> 1 by the peer, 1 -> 0 by the owner, compare-and-swap both ways) carries no epoch, so
// the launch can be captured in a CUDA graph and replayed.  Waits are bounded (trap instead of hanging the device).
#include "common.cuh"
#include <cstdint>
#include <type_traits>

#define MGGAN_PEER_MAX 16
#define MGGAN_PEER_BLOCKS 64          // flag rows per arena: grid <= 64 blocks

struct MgganPeerTable {
    void* region[MGGAN_PEER_MAX];          // this call's region on rank r (peer-mapped device pointers)
    unsigned int* flags[MGGAN_PEER_MAX];   // rank r's flag array [MGGAN_PEER_BLOCKS][MGGAN_PEER_MAX]
    int rank, world;
};

namespace {

__device__ __forceinline__ unsigned int cas_release_sys(unsigned int* p, unsigned int cmp, unsigned int val) {
    unsigned int old;
    asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ unsigned int cas_acquire_sys(unsigned int* p, unsigned int cmp, unsigned int val) {
    unsigned int old;
    asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return old;
}

constexpr unsigned int SPIN_LIMIT = 1u << 22;     // ~ seconds of peer round trips: a missing rank traps, it does not hang

// Block b of every rank meets here.  Thread t < world (t != rank) signals peer t and waits for peer t's signal.
__device__ __forceinline__ void peer_barrier(const MgganPeerTable& tb) {
    __syncthreads();                       // this block's writes to its own region are done (CTA scope) ...
    const int t = threadIdx.x;
    if (t < tb.world && t != tb.rank) {
        unsigned int* theirs = tb.flags[t] + blockIdx.x * MGGAN_PEER_MAX + tb.rank;
        unsigned int spins = 0;
        while (cas_release_sys(theirs, 0u, 1u) != 0u)          // ... and published (release is cumulative over the barrier)
            if (++spins > SPIN_LIMIT) __trap();
        unsigned int* mine = tb.flags[tb.rank] + blockIdx.x * MGGAN_PEER_MAX + t;
        spins = 0;
        while (cas_acquire_sys(mine, 1u, 0u) != 1u)
            if (++spins > SPIN_LIMIT) __trap();
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ T ld_peer(const T* p) { return __ldcv(p); }      // never from a stale L1 line

// One element (or one float4) from every rank, all loads in flight before the first add, summed in rank order.
template <typename T>
__device__ __forceinline__ T sum_ranks(const MgganPeerTable& tb, long long i) {
    T v[MGGAN_PEER_MAX];
#pragma unroll
    for (int r = 0; r < MGGAN_PEER_MAX; ++r)
        if (r < tb.world) v[r] = ld_peer(static_cast<const T*>(tb.region[r]) + i);
    T s = v[0];
#pragma unroll
    for (int r = 1; r < MGGAN_PEER_MAX; ++r)
        if (r < tb.world) s += v[r];
    return s;
}
__device__ __forceinline__ float4 sum_ranks4(const MgganPeerTable& tb, long long i4) {
    float4 v[MGGAN_PEER_MAX];
#pragma unroll
    for (int r = 0; r < MGGAN_PEER_MAX; ++r)
        if (r < tb.world) v[r] = ld_peer(reinterpret_cast<const float4*>(tb.region[r]) + i4);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < MGGAN_PEER_MAX; ++r)
        if (r < tb.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    return s;
}

template <typename T, bool SQNORM>
__global__ void __launch_bounds__(MGGAN_THREADS)
peer_allreduce_kernel(MgganPeerTable tb, const T* __restrict__ in, long long n, T* __restrict__ out,
                      double* __restrict__ sqnorm, int vec4) {
    const long long per = ((n + gridDim.x - 1) / gridDim.x + 3) & ~3LL;      // chunk per block, multiple of 4 elements
    const long long lo = (long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    T* own = static_cast<T*>(tb.region[tb.rank]);
    if (in != nullptr)
        for (long long i = lo + threadIdx.x; i < hi; i += MGGAN_THREADS) own[i] = in[i];
    peer_barrier(tb);
    float acc = 0.f;
    long long i0 = lo;
    if constexpr (std::is_same<T, float>::value) if (vec4) {      // fp32: 16-byte loads (regions 256-byte aligned, lo % 4 == 0, `out` checked by the host)
        const long long n4 = lo < hi ? (hi - lo) / 4 : 0;
        for (long long q = threadIdx.x; q < n4; q += MGGAN_THREADS) {
            const float4 s = sum_ranks4(tb, lo / 4 + q);
            reinterpret_cast<float4*>(out)[lo / 4 + q] = s;
            acc = fmaf(s.x, s.x, fmaf(s.y, s.y, fmaf(s.z, s.z, fmaf(s.w, s.w, acc))));
        }
        i0 = lo + n4 * 4;
    }
    (void)vec4;
    for (long long i = i0 + threadIdx.x; i < hi; i += MGGAN_THREADS) {
        const T s = sum_ranks<T>(tb, i);
        out[i] = s;
        if (SQNORM) acc = fmaf((float)s, (float)s, acc);
    }
    if (SQNORM) {
        __shared__ float sred[MGGAN_THREADS / 32];
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < MGGAN_THREADS / 32; ++w) s += (double)sred[w];
            if (s != 0.0) atomicAdd(sqnorm, s);
        }
    }
}

template <typename T>
int launch(const MgganPeerTable& tb, const void* in, long long n, void* out, double* sqnorm, cudaStream_t stream) {
    int blocks = (int)((n + 2047) / 2048);
    if (blocks < 1) blocks = 1;
    if (blocks > MGGAN_PEER_BLOCKS) blocks = MGGAN_PEER_BLOCKS;
    const int vec4 = (reinterpret_cast<uintptr_t>(out) & 15) == 0 ? 1 : 0;
    if (sqnorm != nullptr)
        peer_allreduce_kernel<T, true><<<blocks, MGGAN_THREADS, 0, stream>>>(tb, static_cast<const T*>(in), n,
                                                                           static_cast<T*>(out), sqnorm, vec4);
    else
        peer_allreduce_kernel<T, false><<<blocks, MGGAN_THREADS, 0, stream>>>(tb, static_cast<const T*>(in), n,
                                                                            static_cast<T*>(out), nullptr, vec4);
    return mggan_check_launch("peer_allreduce");
}

}  // namespace

// dtype: 0 = float32 (sqnorm optional), 1 = float64, 2 = int32.  in == NULL: the operand already sits in this rank's region.
// Every rank of the table must launch the same call (same n, dtype) in the same order on its stream.
extern "C" int mggan_peer_allreduce(const MgganPeerTable* table, int dtype, const void* in, long long n, void* out,
                                    double* sqnorm, cudaStream_t stream) {
    MGGAN_REQUIRE(table != nullptr && table->world >= 1 && table->world <= MGGAN_PEER_MAX && table->rank >= 0 &&
                      table->rank < table->world,
                  "mggan_peer_allreduce: bad peer table");
    MGGAN_REQUIRE(n >= 0 && out != nullptr, "mggan_peer_allreduce: bad arguments");
    MGGAN_REQUIRE(sqnorm == nullptr || dtype == 0, "mggan_peer_allreduce: the squared norm is produced for float32 only");
    if (n == 0) return MGGAN_OK;
    if (dtype == 0) return launch<float>(*table, in, n, out, sqnorm, stream);
    if (dtype == 1) return launch<double>(*table, in, n, out, nullptr, stream);
    if (dtype == 2) return launch<int>(*table, in, n, out, nullptr, stream);
    return mggan_set_error(MGGAN_ERR_INVALID, "mggan_peer_allreduce: dtype %d", dtype);
}
