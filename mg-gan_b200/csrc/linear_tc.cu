// GEMM variant 3 behind mggan_linear_* (opt-in: mggan_set_gemm_variant(3)): the dense layers of the path on the tcgen05
// tensor cores with fp32-level accuracy.  Same contract as gemm_kernel (linear.cu / gemm_args.cuh):
//     C(m, n) (+)= sum_k A(m, k) [* act'(Ay(m, k))] B(n, k)        forward, input gradient and weight gradient of a Linear
// The K range is walked in slabs of 32; a slab of A (128 rows) and of B (N <= 256 rows) is split x = hi + lo (hi = x
// truncated to TF32, lo = x - hi exact in fp32) and staged as four (rows x 32) planes in the UMMA canonical K-major
// no-swizzle layout -- exactly the operand shape, descriptors and 3-product issue sequence of the decoder's gate contraction
// (decoder_tc.cu, measured error < 2e-6 relative) -- and 12 tcgen05.mma kind::tf32 (M = 128, N = padded N, K = 8)
// accumulate the slab into one TMEM tile.  One CTA = one warpgroup = one 128-row tile of C; thread r stages row r of A
// (its own 128-byte line of a K-fast operand, or a coalesced column of an M-fast one), rows r and r + 128 of B, and owns
// TMEM lane r in the epilogue (bias, activation, store or split-K atomicAdd; column sums of A for the bias gradient are
// thread-local).  Written after the round's GPU budget was spent: NOT the default until tests/test_gpu_zf_gemm_variants.py has
// run on a B200.  The PTX wrappers are copies of decoder_tc.cu's on purpose (that file is measured and stays untouched).
#include "gemm_args.cuh"

namespace {

constexpr int TROWS = 128;          // UMMA M
constexpr int KC = 32;              // K slab
constexpr int TC_THREADS = 128;
constexpr int A_PLANE = TROWS * KC; // floats per (128 x 32) plane

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a lost arrival traps
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no swizzle: 8-row x 16-byte core matrices, LBO = 128 B along K, SBO = 1024 B along rows (decoder_tc.cu).
__device__ __forceinline__ int oper_off(int r, int k) { return (r >> 3) * 256 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }
__device__ __forceinline__ uint64_t oper_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(128 >> 4) << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46);
}
constexpr uint64_t DESC_KSTEP = 256 >> 4;
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Stage 4 consecutive k of one operand row as (hi, lo) float4s.
__device__ __forceinline__ void stage4(float* hi_plane, float* lo_plane, int row, int k4, const float (&v)[4]) {
    const float4 h = make_float4(tf32_hi(v[0]), tf32_hi(v[1]), tf32_hi(v[2]), tf32_hi(v[3]));
    st4(hi_plane + oper_off(row, k4), h);
    st4(lo_plane + oper_off(row, k4), make_float4(v[0] - h.x, v[1] - h.y, v[2] - h.z, v[3] - h.w));
}

__global__ void __launch_bounds__(TC_THREADS)
gemm_tc_kernel(GemmArgs g, int npad, uint32_t tmem_cols) {
    extern __shared__ __align__(1024) float smem[];
    float* sAhi = smem;
    float* sAlo = sAhi + A_PLANE;
    float* sBhi = sAlo + A_PLANE;
    float* sBlo = sBhi + npad * KC;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sBlo + npad * KC);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(sBar);

    int k_begin = 0, k_end = g.K;
    if (g.splitk > 1) {
        int per = (g.K + g.splitk - 1) / g.splitk;
        per = (per + KC - 1) / KC * KC;
        k_begin = blockIdx.z * per;
        k_end = min(g.K, k_begin + per);
        if (k_begin >= k_end) return;                 // uniform for the CTA: nothing allocated yet
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(sTmem), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;
    const uint32_t tmem_row = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((static_cast<uint32_t>(npad) >> 3) << 17) | ((128u >> 4) << 24);

    const int m = blockIdx.x * TROWS + tid;
    const bool m_ok = m < g.M;
    float csum = 0.f;
    uint32_t phase = 0;
    bool first = true;
    for (int k0 = k_begin; k0 < k_end; k0 += KC) {
        // ---- A slab: this thread's row
#pragma unroll
        for (int kq = 0; kq < KC / 4; ++kq) {
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = k0 + kq * 4 + c;
                float x = 0.f;
                if (m_ok && k < k_end) {
                    const long long off = m * g.sam + k * g.sak;
                    x = __ldg(g.A + off);
                    if (g.Ay != nullptr) x *= act_bwd(__ldg(g.Ay + off), g.act_in, g.slope);
                }
                v[c] = x;
            }
            csum += (v[0] + v[1]) + (v[2] + v[3]);
            stage4(sAhi, sAlo, tid, kq * 4, v);
        }
        // ---- B slab: rows tid and tid + 128 of the (padded) N
        for (int n = tid; n < npad; n += TC_THREADS) {
            const bool n_ok = n < g.N;
#pragma unroll
            for (int kq = 0; kq < KC / 4; ++kq) {
                float v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = k0 + kq * 4 + c;
                    v[c] = (n_ok && k < k_end) ? __ldg(g.B + n * g.sbn + k * g.sbk) : 0.f;
                }
                stage4(sBhi, sBlo, n, kq * 4, v);
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t ah = oper_desc(smem_u32(sAhi)), al = oper_desc(smem_u32(sAlo));
            const uint64_t bh = oper_desc(smem_u32(sBhi)), bl = oper_desc(smem_u32(sBlo));
#pragma unroll
            for (int kk = 0; kk < KC / 8; ++kk) {
                umma_tf32(tmem_d, ah + kk * DESC_KSTEP, bh + kk * DESC_KSTEP, idesc, (first && kk == 0) ? 0u : 1u);
                umma_tf32(tmem_d, al + kk * DESC_KSTEP, bh + kk * DESC_KSTEP, idesc, 1u);
                umma_tf32(tmem_d, ah + kk * DESC_KSTEP, bl + kk * DESC_KSTEP, idesc, 1u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);          // the slab's MMAs have read shared memory: it can be restaged
        phase ^= 1u;
        tc_fence_after();
        first = false;
    }
    // ---- epilogue: thread = row m = TMEM lane
    for (int c16 = 0; c16 < npad; c16 += 16) {
        float v[16];
        tmem_ld16(tmem_row + c16, v);           // whole warps execute the load; only valid elements are written
        if (!m_ok) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = c16 + j;
            if (n >= g.N) continue;
            float* dst = g.C + m * g.scm + n * g.scn;
            if (g.splitk > 1) {
                atomicAdd(dst, v[j]);
            } else {
                float y = v[j];
                if (g.bias != nullptr) y += __ldg(g.bias + n);
                *dst = act_fwd(y, g.act, g.slope);
            }
        }
    }
    if (g.colsum != nullptr && m_ok) atomicAdd(g.colsum + m, csum);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

}  // namespace

int mggan_gemm_tc_launch(const GemmArgs& g, cudaStream_t stream) {
    const int npad = (g.N + 15) / 16 * 16;
    if (npad > 256) return -1;
    uint32_t cols = 32;
    while ((int)cols < npad) cols <<= 1;
    const size_t smem = sizeof(float) * (2 * A_PLANE + 2 * (size_t)npad * KC) + 16;
    cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((g.M + TROWS - 1) / TROWS, 1, g.splitk > 1 ? g.splitk : 1);
    gemm_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(g, npad, cols);
    return mggan_check_launch("linear (tensor cores)");
}
