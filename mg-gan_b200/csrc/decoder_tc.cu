// Tensor-core form of the multi-generator trajectory decoder forward (same operator, inputs, outputs and saved
// activations as decoder.cu: RelativeDecoder.forward mggan/model/modules/common_modules.py:97-131 driven by
// MultiGenerator.forward_all mggan/model/modules/standard.py:227-265 and the gather of standard.py:190-214).
//
// BASELINE.json north_star: "tensor cores used only for the batched (G.B.k) x 4H x H decoder contraction".  That
// contraction -- gates = W_hh h_{t-1} for a 128-sequence tile, (128 x 32) . (32 x 128) per time step -- runs here
// on the 5th-generation tensor cores (tcgen05.mma, kind::tf32, M = 128, N = 128, K = 8 per instruction) with the
// accumulator in tensor memory.  The parity bar is 1e-3 relative fp32 after a 12-step autoregressive recurrence,
// which plain TF32 (10-bit mantissa) does not meet, so both operands are split x = hi + lo with hi = x truncated to
// TF32 and lo = x - hi (exact in fp32), and three products are accumulated in the same TMEM tile:
//     W_hh h  ~=  Wh.hh + Wh.hl + Wl.hh          (dropped term Wl.hl ~ 2^-22 relative; measured ~1e-6)
//
// One CTA = one warpgroup (128 threads) = one 128-row tile; thread r owns row r = TMEM lane r, so the LSTM cell,
// hidden2pos and the xy accumulation of a sequence are thread-local (no shuffles, no shared-memory round trip for
// the gates).  Per time step:
//     wait(mbarrier)                       <- tcgen05.commit of the step's 12 MMAs
//     tcgen05.ld 16 columns at a time      gate columns are permuted to (unit, gate) order: one load = 4 units
//     gates += Wx dxdy_{t-1} + b ; cell ; h_t -> registers, saved activations -> global
//     h_t -> shared memory as (hi, lo) in the UMMA canonical K-major layout ; fence.proxy.async ; barrier
//     thread 0 issues the MMAs of step t+1 ; the other threads run hidden2pos on h_t while the tensor core works
// Three CTAs are resident per SM (74 KB shared memory, 128 TMEM columns each), so one tile's cell math overlaps the
// others' MMAs and TMEM loads.
#include "common.cuh"

namespace {

constexpr int H = 32;          // decoder hidden size
constexpr int M1 = 16;         // hidden2pos mid width
constexpr int TROWS = 128;     // rows per tile = UMMA M
constexpr int TC_THREADS = 128;
constexpr int ZMAX = 16;
constexpr int OPER_FLOATS = TROWS * H;     // one (128 x 32) operand plane (hi or lo): 16 KB
constexpr uint32_t TMEM_COLS = 128;

// ---- PTX wrappers (sm_100a) ---------------------------------------------------------------------------------
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
// Bounded wait: a lost arrival traps (launch error reported by the C ABI) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
// D[tmem] (+)= A[smem] . B[smem]^T, TF32 inputs, FP32 accumulate; issued by one thread for the CTA.
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
// 16 consecutive fp32 columns of this thread's TMEM lane.
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

// ---- UMMA operand layout ------------------------------------------------------------------------------------
// K-major, no swizzle ("interleaved" canonical layout, cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>):
// 8-row x 16-byte core matrices; core matrices adjacent in K are LBO = 128 B apart, adjacent 8-row groups are
// SBO = 1024 B apart.  Element (row r, k) of a (rows x 32) fp32 operand sits at float offset:
__device__ __forceinline__ int oper_off(int r, int k) { return (r >> 3) * 256 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14), LBO >> 4 in
// [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 (no swizzle) in [61,64).
__device__ __forceinline__ uint64_t oper_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(128 >> 4) << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46);
}
// One K = 8 step covers two core matrices = 256 B of the operand: +16 in the (>> 4) start-address field.
constexpr uint64_t DESC_KSTEP = 256 >> 4;

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both
// K-major, N >> 3 in bits [17,23), M >> 4 in bits [24,29).
constexpr uint32_t IDESC_M128_N128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// Gate non-linearities on the MUFU pipe, 4 / 5 instructions each (the cell math of 128 rows x 32 units x 5 activations
// per step is what bounds this kernel once the contraction is on the tensor core): flush-to-zero ex2 / rcp approximations,
// |error| ~1e-7, exact saturation at +-inf.
__device__ __forceinline__ float ex2_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_tc(float x) { return rcp_(1.f + ex2_(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_tc(float x) { return fmaf(-2.f, rcp_(1.f + ex2_(2.8853900817779268f * x)), 1.f); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// D = Ah.Bh + Al.Bh + Ah.Bl over K = 32 (4 K-steps x 3 products), then commit to the mbarrier.
__device__ __forceinline__ void issue_gate_mma(uint32_t tmem_d, uint32_t sAhi, uint32_t sAlo, uint32_t sBhi, uint32_t sBlo, uint32_t bar) {
    const uint64_t ah = oper_desc(sAhi), al = oper_desc(sAlo), bh = oper_desc(sBhi), bl = oper_desc(sBlo);
#pragma unroll
    for (int kk = 0; kk < H / 8; ++kk) {
        umma_tf32(tmem_d, ah + kk * DESC_KSTEP, bh + kk * DESC_KSTEP, IDESC_M128_N128, kk > 0 ? 1u : 0u);
        umma_tf32(tmem_d, al + kk * DESC_KSTEP, bh + kk * DESC_KSTEP, IDESC_M128_N128, 1u);
        umma_tf32(tmem_d, ah + kk * DESC_KSTEP, bl + kk * DESC_KSTEP, IDESC_M128_N128, 1u);
    }
    umma_commit(bar);
}

struct DecWeights {
    const float* Wz; const float* Wx; const float* b; const float* Whh; const float* W1h; const float* W1s;
    const float* b1; const float* W2; const float* b2;
};
struct DecSeq {
    const int* tile_gen; const int* seq_agent; const int* seq_noise; const int* seq_out;
};

// Shared memory carve-up (floats)
constexpr int S_AHI = 0;
constexpr int S_ALO = S_AHI + OPER_FLOATS;
constexpr int S_BHI = S_ALO + OPER_FLOATS;
constexpr int S_BLO = S_BHI + OPER_FLOATS;
constexpr int S_WXB = S_BLO + OPER_FLOATS;          // [unit][gate] float4 (wx0, wx1, b, 0)
constexpr int S_W1HT = S_WXB + H * 4 * 4;           // [unit][M1]   W1h transposed
constexpr int S_W1S = S_W1HT + H * M1;              // [M1][H]
constexpr int S_WZT = S_W1S + M1 * H;               // [ZMAX][H]    Wz transposed
constexpr int S_W2 = S_WZT + ZMAX * H;              // [2][M1] | b1[M1] | b2[2] | pad
constexpr int S_END = S_W2 + 2 * M1 + M1 + 4;
constexpr size_t TC_SMEM_BYTES = sizeof(float) * S_END + 16;      // + mbarrier (8 B) + TMEM base (4 B)

__global__ void __launch_bounds__(TC_THREADS, 3)
decoder_fwd_tc_kernel(DecSeq sq, int n_super, const float* __restrict__ A, const float* __restrict__ social,
                      const float* __restrict__ last_xy, const float* __restrict__ last_dxdy,
                      const float* __restrict__ noise, int Z, DecWeights w, int T, int n_cols,
                      float* __restrict__ out_abs, float* __restrict__ out_rel, float* __restrict__ acts,
                      float* __restrict__ u1save, float* __restrict__ h0save) {
    extern __shared__ __align__(1024) float smem[];
    float* sAhi = smem + S_AHI;
    float* sAlo = smem + S_ALO;
    float* sBhi = smem + S_BHI;
    float* sBlo = smem + S_BLO;
    float4* sWxb = reinterpret_cast<float4*>(smem + S_WXB);
    float* sW1hT = smem + S_W1HT;
    float* sW1s = smem + S_W1S;
    float* sWzT = smem + S_WZT;
    float* sW2 = smem + S_W2;
    float* sB1 = sW2 + 2 * M1;
    float* sB2 = sB1 + M1;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + S_END);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(sBar);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(sTmem), TMEM_COLS);
        tmem_relinquish();
    }
    for (int i = tid; i < ZMAX * H; i += TC_THREADS) {
        const int z = i / H, u = i - z * H;
        sWzT[i] = z < Z ? __ldg(w.Wz + u * Z + z) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;
    const uint32_t tmem_row = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);     // this warp's 32-lane quarter
    uint32_t phase = 0;

    const int per = (n_super + gridDim.x - 1) / gridDim.x;
    const int s_begin = blockIdx.x * per, s_end = min(n_super, s_begin + per);
    int cur_g = -1;
    for (int st = s_begin; st < s_end; ++st) {
        const int g = sq.tile_gen[2 * st];
        if (g < 0) continue;
        // every thread is past the previous tile's last TMEM load and shared-memory read (loop-end barrier)
        if (g != cur_g) {
            cur_g = g;
            const float* Wg = w.Whh + (size_t)g * 4 * H * H;
            for (int i = tid; i < 4 * H * H; i += TC_THREADS) {
                const int n = i >> 5, k = i & 31;            // n = original gate row q*H + u
                const int np = (n & 31) * 4 + (n >> 5);      // permuted column: (unit, gate)
                const float v = __ldg(Wg + i);
                const float hi = tf32_hi(v);
                sBhi[oper_off(np, k)] = hi;
                sBlo[oper_off(np, k)] = v - hi;
            }
            {
                const int u = tid >> 2, q = tid & 3;         // 128 threads = 32 units x 4 gates
                const size_t o = (size_t)g * 4 * H + q * H + u;
                sWxb[tid] = make_float4(__ldg(w.Wx + o * 2), __ldg(w.Wx + o * 2 + 1), __ldg(w.b + o), 0.f);
            }
            for (int i = tid; i < M1 * H; i += TC_THREADS) {
                const int m = i / H, u = i - m * H;
                sW1hT[u * M1 + m] = __ldg(w.W1h + (size_t)g * M1 * H + i);
                sW1s[i] = __ldg(w.W1s + (size_t)g * M1 * H + i);
            }
            if (tid < 2 * M1) sW2[tid] = __ldg(w.W2 + (size_t)g * 2 * M1 + tid);
            if (tid < M1) sB1[tid] = __ldg(w.b1 + (size_t)g * M1 + tid);
            if (tid < 2) sB2[tid] = __ldg(w.b2 + g * 2 + tid);
            __syncthreads();
        }
        const size_t row = (size_t)st * TROWS + tid;
        const int ag = sq.seq_agent[row];
        const bool valid = ag >= 0;
        const int pcol = valid ? sq.seq_out[row] : -1;

        // ---- per-row constants: bs = b1 + W1s social ; last observation
        float bs[M1];
#pragma unroll
        for (int m = 0; m < M1; ++m) bs[m] = sB1[m];
        float xy0 = 0.f, xy1 = 0.f, d0 = 0.f, d1 = 0.f;
        if (valid) {
            const float4* sp = reinterpret_cast<const float4*>(social + (size_t)ag * H);
#pragma unroll
            for (int k4 = 0; k4 < H / 4; ++k4) {
                const float4 s4 = __ldg(sp + k4);
#pragma unroll
                for (int m = 0; m < M1; ++m) {
                    const float4 w4 = ld4(sW1s + m * H + k4 * 4);
                    bs[m] = fmaf(s4.x, w4.x, fmaf(s4.y, w4.y, fmaf(s4.z, w4.z, fmaf(s4.w, w4.w, bs[m]))));
                }
            }
            const float2 p = __ldg(reinterpret_cast<const float2*>(last_xy + (size_t)ag * 2));
            const float2 v = __ldg(reinterpret_cast<const float2*>(last_dxdy + (size_t)ag * 2));
            xy0 = p.x; xy1 = p.y; d0 = v.x; d1 = v.y;
        }
        // ---- h0 = A[agent] + Wz z ; c0 = 0
        float c[H];
        {
            float zv[ZMAX];
#pragma unroll
            for (int z = 0; z < ZMAX; ++z) zv[z] = 0.f;
            if (valid) {
                const float* zp = noise + (size_t)sq.seq_noise[row] * Z;
#pragma unroll
                for (int z = 0; z < ZMAX; ++z)
                    if (z < Z) zv[z] = __ldg(zp + z);
            }
#pragma unroll
            for (int u4 = 0; u4 < H / 4; ++u4) {
                float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    h4 = __ldg(reinterpret_cast<const float4*>(A + (size_t)ag * H) + u4);
#pragma unroll
                    for (int z = 0; z < ZMAX; ++z) {
                        const float4 wz = ld4(sWzT + z * H + u4 * 4);
                        h4.x = fmaf(wz.x, zv[z], h4.x); h4.y = fmaf(wz.y, zv[z], h4.y);
                        h4.z = fmaf(wz.z, zv[z], h4.z); h4.w = fmaf(wz.w, zv[z], h4.w);
                    }
                    if (h0save != nullptr) st4(h0save + dec_h0_off(row, u4 * 4), h4);
                }
                const float4 hi = make_float4(tf32_hi(h4.x), tf32_hi(h4.y), tf32_hi(h4.z), tf32_hi(h4.w));
                st4(sAhi + oper_off(tid, u4 * 4), hi);
                st4(sAlo + oper_off(tid, u4 * 4), make_float4(h4.x - hi.x, h4.y - hi.y, h4.z - hi.z, h4.w - hi.w));
                c[u4 * 4] = 0.f; c[u4 * 4 + 1] = 0.f; c[u4 * 4 + 2] = 0.f; c[u4 * 4 + 3] = 0.f;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gate_mma(tmem_d, smem_u32(sAhi), smem_u32(sAlo), smem_u32(sBhi), smem_u32(sBlo), bar);
        }

        for (int t = 0; t < T; ++t) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();
            float uacc[M1];
#pragma unroll
            for (int m = 0; m < M1; ++m) uacc[m] = bs[m];
            float* arow = (acts != nullptr && valid) ? acts + dec_acts_off(n_super, t, row, 0) : nullptr;
#pragma unroll
            for (int u4 = 0; u4 < H / 4; ++u4) {
                float v[16];
                tmem_ld16(tmem_row + u4 * 16, v);
                float hq[4];
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int u = u4 * 4 + uu;
                    const float4 wi = sWxb[u * 4 + 0], wf = sWxb[u * 4 + 1], wg = sWxb[u * 4 + 2], wo = sWxb[u * 4 + 3];
                    const float ig = sigmoid_tc(v[uu * 4 + 0] + fmaf(wi.x, d0, fmaf(wi.y, d1, wi.z)));
                    const float fg = sigmoid_tc(v[uu * 4 + 1] + fmaf(wf.x, d0, fmaf(wf.y, d1, wf.z)));
                    const float gg = tanh_tc(v[uu * 4 + 2] + fmaf(wg.x, d0, fmaf(wg.y, d1, wg.z)));
                    const float og = sigmoid_tc(v[uu * 4 + 3] + fmaf(wo.x, d0, fmaf(wo.y, d1, wo.z)));
                    c[u] = fmaf(fg, c[u], ig * gg);
                    const float tc = tanh_tc(c[u]);
                    const float h = og * tc;
                    hq[uu] = h;
#pragma unroll
                    for (int m4 = 0; m4 < M1 / 4; ++m4) {
                        const float4 w1 = ld4(sW1hT + u * M1 + m4 * 4);
                        uacc[m4 * 4] = fmaf(w1.x, h, uacc[m4 * 4]); uacc[m4 * 4 + 1] = fmaf(w1.y, h, uacc[m4 * 4 + 1]);
                        uacc[m4 * 4 + 2] = fmaf(w1.z, h, uacc[m4 * 4 + 2]); uacc[m4 * 4 + 3] = fmaf(w1.w, h, uacc[m4 * 4 + 3]);
                    }
                }
                if (arow != nullptr) {     // dec_acts_off layout: each st4 of a warp is 512 contiguous bytes (unit pair 2 u4 + {0, 1})
                    float* a = arow + (u4 * 2) * 512;
                    st4(a, make_float4(hq[0], c[u4 * 4], hq[1], c[u4 * 4 + 1]));
                    st4(a + 512, make_float4(hq[2], c[u4 * 4 + 2], hq[3], c[u4 * 4 + 3]));
                }
                if (t + 1 < T) {
                    const float4 hi = make_float4(tf32_hi(hq[0]), tf32_hi(hq[1]), tf32_hi(hq[2]), tf32_hi(hq[3]));
                    st4(sAhi + oper_off(tid, u4 * 4), hi);
                    st4(sAlo + oper_off(tid, u4 * 4), make_float4(hq[0] - hi.x, hq[1] - hi.y, hq[2] - hi.z, hq[3] - hi.w));
                }
            }
            // all TMEM loads of this step are complete (tcgen05.wait::ld); publish h_t and hand the tile to the tensor core
            tc_fence_before();
            fence_async_smem();
            __syncthreads();
            if (tid == 0 && t + 1 < T) {
                tc_fence_after();
                issue_gate_mma(tmem_d, smem_u32(sAhi), smem_u32(sAlo), smem_u32(sBhi), smem_u32(sBlo), bar);
            }
            // ---- hidden2pos on h_t (overlaps the MMAs of step t+1)
            float n0 = sB2[0], n1 = sB2[1];
#pragma unroll
            for (int m = 0; m < M1; ++m) {
                const float a = lrelu_(uacc[m], 0.01f);
                n0 = fmaf(sW2[m], a, n0);
                n1 = fmaf(sW2[M1 + m], a, n1);
            }
            d0 = n0; d1 = n1;
            xy0 += d0; xy1 += d1;
            if (pcol >= 0) {
                if (u1save != nullptr) {
                    float* us = u1save + dec_u1_off(n_super, t, row, 0);
#pragma unroll
                    for (int m4 = 0; m4 < M1 / 4; ++m4)
                        st4(us + m4 * 512, make_float4(uacc[m4 * 4], uacc[m4 * 4 + 1], uacc[m4 * 4 + 2], uacc[m4 * 4 + 3]));
                }
                const size_t o = ((size_t)t * n_cols + pcol) * 2;
                *reinterpret_cast<float2*>(out_rel + o) = make_float2(d0, d1);
                *reinterpret_cast<float2*>(out_abs + o) = make_float2(xy0, xy1);
            }
        }
        __syncthreads();     // the next tile restages shared memory
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// Self-test of the UMMA plumbing: D (128 x 128) = A (128 x 32) . B (128 x 32)^T with the 3 x TF32 split.
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) float smem[];
    float* sAhi = smem + S_AHI;
    float* sAlo = smem + S_ALO;
    float* sBhi = smem + S_BHI;
    float* sBlo = smem + S_BLO;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + S_END);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(sTmem), TMEM_COLS);
        tmem_relinquish();
    }
    for (int i = tid; i < TROWS * H; i += TC_THREADS) {
        const int r = i >> 5, k = i & 31;
        const float a = A[i], b = B[i];
        const float ah = tf32_hi(a), bh = tf32_hi(b);
        sAhi[oper_off(r, k)] = ah; sAlo[oper_off(r, k)] = a - ah;
        sBhi[oper_off(r, k)] = bh; sBlo[oper_off(r, k)] = b - bh;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;
    if (tid == 0) issue_gate_mma(tmem_d, smem_u32(sAhi), smem_u32(sAlo), smem_u32(sBhi), smem_u32(sBlo), bar);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t tmem_row = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c16 = 0; c16 < 128 / 16; ++c16) {
        float v[16];
        tmem_ld16(tmem_row + c16 * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * 128 + c16 * 16 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

}  // namespace

// Same contract as mggan_decoder_fwd (decoder.cu); requires the work list to be padded to 128-row groups per
// generator (mggan_selection_build / mggan_selection_all do that), i.e. n_tiles even and tiles 2s, 2s+1 of one generator.
extern "C" int mggan_decoder_fwd_tc(int n_tiles, const int* tile_gen, const int* seq_agent, const int* seq_noise,
                                    const int* seq_out, const float* A, const float* social, const float* last_xy,
                                    const float* last_dxdy, const float* noise, int Z, const float* Wz, const float* Wx,
                                    const float* b, const float* Whh, const float* W1h, const float* W1s, const float* b1,
                                    const float* W2, const float* b2, int pred_len, int n_cols, float* out_abs,
                                    float* out_rel, float* acts, float* u1save, float* h0save, cudaStream_t stream) {
    MGGAN_REQUIRE(Z >= 1 && Z <= ZMAX, "mggan_decoder_fwd_tc: noise_dim %d not in [1, %d]", Z, ZMAX);
    MGGAN_REQUIRE(pred_len >= 1 && n_tiles >= 0 && (n_tiles & 1) == 0, "mggan_decoder_fwd_tc: bad pred_len / odd n_tiles");
    MGGAN_REQUIRE((acts == nullptr) == (u1save == nullptr) && (acts == nullptr) == (h0save == nullptr),
                  "mggan_decoder_fwd_tc: save buffers must be all set or all null");
    if (n_tiles == 0) return MGGAN_OK;
    DecSeq sq{tile_gen, seq_agent, seq_noise, seq_out};
    DecWeights w{Wz, Wx, b, Whh, W1h, W1s, b1, W2, b2};
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int n_super = n_tiles / 2;
    cudaFuncSetAttribute(decoder_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    const int grid = n_super < sms * 3 ? n_super : sms * 3;
    decoder_fwd_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(sq, n_super, A, social, last_xy, last_dxdy, noise, Z,
                                                                      w, pred_len, n_cols, out_abs, out_rel, acts, u1save,
                                                                      h0save);
    return mggan_check_launch("decoder_fwd_tc");
}

// D (128,128) = A (128,32) . B (128,32)^T through tcgen05.mma kind::tf32 with the hi/lo split (test hook).
extern "C" int mggan_tc_selftest(const float* A, const float* B, float* D, cudaStream_t stream) {
    cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    tc_selftest_kernel<<<1, TC_THREADS, TC_SMEM_BYTES, stream>>>(A, B, D);
    return mggan_check_launch("tc_selftest");
}
