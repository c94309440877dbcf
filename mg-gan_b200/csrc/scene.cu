// Physical (scene) attention: 2 x [conv3x3 pad1 -> BatchNorm2d -> ReLU -> MaxPool2x2] on the
// per-agent (4,33,33) crop, then a per-position channel-softmax attention -> (N, 64).
// Reference: mggan/model/modules/cnn.py:101-116 (AttentionGlobal.forward), :119-175
// (Conv_Blocks), :178-282 (CNN); BatchNorm runs in train mode with batch statistics over
// (N, H, W), so every conv layer needs a grid-wide reduction before its output can be used.
//
// conv1 is linear in the crop, so BatchNorm-1's batch statistics (and the dense half of conv1's weight
// gradient) follow from data-only patch statistics R (36x36), P (36) computed ONCE per batch of crops:
//   scene_patch_stats        img -> R, P                               (shared by G and D, all three steps)
//   scene_bn1_from_patches   R, P, W1 -> BN1 as an affine map (a, b), running statistics update
//   scene_fused12_fwd        img -> conv1 -> BN1 -> ReLU -> pool -> conv2 -> x2 (stored) + sums of x2
//   scene_bn_finalize        sums -> BN2 affine map
//   scene_attn_fwd           x2 -> BN2 -> ReLU -> pool -> v (64 x C) -> MLP C->32->C -> softmax_c -> sum_c a v
// Backward: scene_attn_bwd -> scene_bn_bwd_finalize -> scene_fused12_bwd -> scene_bn1_bwd_finalize.  The
// gradient w.r.t. a pooled+ReLU'd BatchNorm output is sparse (one position per 2x2 window): (value, 2-bit arg).
// Partial BatchNorm sums are accumulated in fp32 per thread, reduced per CTA and added to the
// global statistics as doubles (one atomicAdd per channel per CTA).
#include "common.cuh"
#include <cstdint>

namespace {

constexpr int IMG = 33, IMG2 = IMG * IMG;      // 1089
constexpr int P1 = 16, P1SQ = 256;             // after first pool
constexpr int P2 = 8, P2SQ = 64;               // after second pool
constexpr int CIN = 4;
constexpr int AH = 32;                         // attention MLP hidden width (cnn.py mlp_dim)
constexpr int LDI = 36;                        // padded image row (35 used)
constexpr int IMGPAD = 35 * LDI;               // one padded channel
// Channel stride of the crop in the backward kernel: = 8 mod 32.  In the sparse conv1 weight-gradient stage the lanes of a
// warp that differ only in the input channel read the same (data-dependent) pixel of 4 channels; with the dense stride
// (1260 = 12 mod 32) channels 0 and 3 sit 4 banks apart and collide with the neighbouring pooled pixels of the other
// (ncu: 63 M of that stage's 134 M shared wavefronts were conflict replays); 8 banks apart they only meet at the margins.
constexpr int IMGPAD_BWD = 35 * LDI + 28;
constexpr int LDP = 18;                        // padded pooled row

// Forward kernel: strides that make the conv2 tensor-core fragments conflict-free (bank = 8 t + g for lanes (g, t)).
struct FwdPad {
    static constexpr int PPAD = LDP * LDP + 4;                               // 328 = 8 mod 32
};
// Backward kernel: dx2 / p1 channel stride 324 (= 4 mod 32) and input-gradient weight rows of 20 floats.
struct BwdPad {
    static constexpr int PPAD = LDP * LDP;
    static constexpr int LDWD = 20;
};
template <int C>
struct FwdLdw2 { static constexpr int value = C == 16 ? 24 : 8; };            // 24 t + g / 8 t + g: 32 distinct banks

// ---- bulk (TMA-engine) staging of one agent's crop --------------------------------------------------------------
// A crop is 4 x 33 x 33 fp32 = 17,424 contiguous bytes (16-byte multiple, 16-byte aligned for every agent of a 16-byte
// aligned batch), so one `cp.async.bulk` (SASS UBLKCP) moves it global -> shared without touching registers and signals an
// mbarrier when the bytes have landed.  The fused conv kernels double-buffer it: agent n+1's crop streams in while agent n
// is convolved, and the kernels read the UNPADDED [4][33][33] layout directly (only row / column -1 of a 3 x 3 window can
// fall outside; it is predicated to zero).
constexpr int RAW = CIN * IMG2;                // floats per crop
constexpr uint32_t RAW_BYTES = RAW * 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void sbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1u << 24)) __trap();       // a lost copy traps (reported by the C ABI) instead of hanging the device
    }
}
// One thread: arm the barrier and start the copy of agent `src`'s crop into `dst`.
__device__ __forceinline__ void stage_crop(const float* __restrict__ img, int src, float* dst, uint64_t* bar) {
    const uint32_t b = smem_addr(bar);
    sbar_expect_tx(b, RAW_BYTES);
    bulk_g2s(smem_addr(dst), img + (size_t)src * RAW, RAW_BYTES, b);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Reduce NV per-thread partial sums over the CTA and add them (as doubles) to global memory.
template <int NV>
__device__ __forceinline__ void block_reduce_to_global(const float (&v)[NV], double* __restrict__ dst, float* sred) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = warp_sum(v[i]);
        if (lane == 0) sred[warp * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < MGGAN_THREADS / 32; ++w) s += (double)sred[w * NV + threadIdx.x];
        atomicAdd(dst + threadIdx.x, s);
    }
    __syncthreads();
}
template <int NV>
__device__ __forceinline__ void block_reduce_to_global_f(const float (&v)[NV], float* __restrict__ dst, float* sred) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = warp_sum(v[i]);
        if (lane == 0) sred[warp * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        float s = 0.f;
        for (int w = 0; w < MGGAN_THREADS / 32; ++w) s += sred[w * NV + threadIdx.x];
        atomicAdd(dst + threadIdx.x, s);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// BatchNorm finalize: statistics -> affine (a, b), mean / invstd for the backward, running stats.
__global__ void scene_bn_finalize_kernel(const double* __restrict__ stats, double count, int C,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         long long* __restrict__ nbt, float momentum, float eps, int training,
                                         float* __restrict__ ab, float* __restrict__ mean_istd) {
    int c = threadIdx.x;
    if (c < C) {
        float mean, var;
        if (training) {
            double m = stats[c] / count;
            double v = stats[C + c] / count - m * m;
            if (v < 0.0) v = 0.0;
            mean = (float)m; var = (float)v;
            double unb = count > 1.0 ? v * count / (count - 1.0) : v;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
        } else {
            mean = running_mean[c]; var = running_var[c];
        }
        float istd = rsqrtf(var + eps);
        istd = istd * (1.5f - 0.5f * (var + eps) * istd * istd);     // one Newton step on rsqrt.approx
        float a = gamma[c] * istd;
        ab[c] = a;
        ab[C + c] = beta[c] - mean * a;
        mean_istd[c] = mean;
        mean_istd[C + c] = istd;
    }
    if (training && threadIdx.x == 0) nbt[0] += 1;
}

// BatchNorm backward finalize: sums (sum dy, sum dy*xhat) -> means m1, m2 and dgamma / dbeta.
__global__ void scene_bn_bwd_finalize_kernel(const double* __restrict__ sums, double count, int C,
                                             float* __restrict__ m12, float* __restrict__ dgamma,
                                             float* __restrict__ dbeta) {
    int c = threadIdx.x;
    if (c < C) {
        m12[c] = (float)(sums[c] / count);
        m12[C + c] = (float)(sums[C + c] / count);
        atomicAdd(dbeta + c, (float)sums[c]);
        atomicAdd(dgamma + c, (float)sums[C + c]);
    }
}

// ------------------------------------------------------------------------------------------
// Patch statistics.  conv1 is linear in the image, so the train-mode BatchNorm-1 statistics and the
// dense part of the conv1 weight gradient can be written through two data-only quantities:
//   P[a]    = sum over (agent, pixel) of patch_a            (36 taps a = (ci, ky, kx), zero padded)
//   R[a][b] = sum over (agent, pixel) of patch_a * patch_b  (36 x 36, symmetric)
// mean(x1_c) = (W_c . P) / n + b_c,  E[x1_c^2] = (W_c^T R W_c + 2 b_c W_c . P) / n + b_c^2, and
// sum x1_c * patch_a = (R W_c)_a + b_c P_a.  They depend on the crops only, so one launch serves every
// scene-CNN forward / backward of a training iteration (G and D, three optimiser steps).
// A CTA owns ONE pair of input channels (A <= B, 10 pairs) for a strided subset of the agents: it stages the two
// padded channels, threads own pixels, and each thread keeps the full 9 x 9 tap block of R for that channel pair in
// registers across all its agents (18 shared-memory loads feed 81 FMAs per pixel), reduced once at the end.
constexpr int NTAP = 36;
constexpr int NPAIRS_CH = CIN * (CIN + 1) / 2;     // 10

constexpr int PS_THREADS = 128;                    // ~150 registers per thread (81 accumulators): 3 CTAs per SM

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }

__global__ void __launch_bounds__(PS_THREADS, 3)
scene_patch_stats_kernel(const float* __restrict__ img, const int* __restrict__ rows, int N, int slots,
                         double* __restrict__ R, double* __restrict__ P) {
    __shared__ __align__(16) float sImg[2][2 * IMGPAD];       // double buffer x channels A and B, [35][36] each, zero halo
    const int bp = blockIdx.x % NPAIRS_CH, slot = blockIdx.x / NPAIRS_CH;
    int cA = 0, cB = 0;
    {
        int q = bp;
        for (cA = 0; cA < CIN; ++cA) {
            if (q < CIN - cA) { cB = cA + q; break; }
            q -= CIN - cA;
        }
    }
    const bool diag = cA == cB;
    for (int i = threadIdx.x; i < 4 * IMGPAD; i += PS_THREADS) (&sImg[0][0])[i] = 0.f;
    float acc[81], ps[9];
#pragma unroll
    for (int q = 0; q < 81; ++q) acc[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) ps[q] = 0.f;
    __syncthreads();

    // asynchronous staging (cp.async, 4-byte: channel rows are only 4-byte aligned) of the next agent's two channels
    // while the current one is processed
    auto stage = [&](int n, int buf) {
        const int src = rows ? rows[n] : n;
        const float* ipA = img + ((size_t)src * CIN + cA) * IMG2;
        const float* ipB = img + ((size_t)src * CIN + cB) * IMG2;
        for (int i = threadIdx.x; i < IMG2; i += PS_THREADS) {
            int y = i / IMG, x = i - y * IMG;
            cp_async4(&sImg[buf][(y + 1) * LDI + x + 1], ipA + i);
            if (!diag) cp_async4(&sImg[buf][IMGPAD + (y + 1) * LDI + x + 1], ipB + i);
        }
        cp_async_commit();
    };
    int buf = 0;
    if (slot < N) stage(slot, 0);
    for (int n = slot; n < N; n += slots) {
        const bool more = n + slots < N;
        if (more) stage(n + slots, buf ^ 1);
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const float* sA = sImg[buf];
        const float* sB = diag ? sA : sA + IMGPAD;
        // lanes = 32 consecutive pixels of one row (conflict-free shared-memory reads), warps = rows; the 33rd
        // column is swept afterwards by the first 33 threads
        for (int it = threadIdx.x >> 5; it < IMG + 2; it += PS_THREADS / 32) {
            int y, x;
            if (it < IMG) { y = it; x = threadIdx.x & 31; }
            else { y = (it - IMG) * 32 + (threadIdx.x & 31); x = IMG - 1; if (y >= IMG) continue; }
            const int base = y * LDI + x;
            float va[9], vb[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) va[t] = sA[base + (t / 3) * LDI + (t % 3)];
            if (diag) {
#pragma unroll
                for (int t = 0; t < 9; ++t) { vb[t] = va[t]; ps[t] += va[t]; }
            } else {
#pragma unroll
                for (int t = 0; t < 9; ++t) vb[t] = sB[base + (t / 3) * LDI + (t % 3)];
            }
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = 0; j < 9; ++j) acc[i * 9 + j] = fmaf(va[i], vb[j], acc[i * 9 + j]);
        }
        __syncthreads();                    // this buffer is overwritten by the staging of the iteration after next
        buf ^= 1;
    }
    // CTA reduction: warp shuffles in double, then one atomicAdd per entry per warp
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 81; ++q) {
        double v = warp_sum_d((double)acc[q]);
        if (lane == 0) {
            const int a = cA * 9 + q / 9, b = cB * 9 + q % 9;
            atomicAdd(R + a * NTAP + b, v);
            if (!diag) atomicAdd(R + b * NTAP + a, v);
        }
    }
    if (diag) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            double v = warp_sum_d((double)ps[q]);
            if (lane == 0) atomicAdd(P + cA * 9 + q, v);
        }
    }
}

// BatchNorm-1 statistics from the patch statistics (double precision).  One CTA: R and W staged in shared memory,
// thread = (channel, tap) partial of W_c^T R W_c and W_c . P, then one thread per channel finishes.
__global__ void __launch_bounds__(MGGAN_THREADS)
scene_bn1_from_patches_kernel(const double* __restrict__ R, const double* __restrict__ P, double count,
                              int C, const float* __restrict__ W, const float* __restrict__ bias,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ running_mean, float* __restrict__ running_var,
                              long long* __restrict__ nbt, float momentum, float eps, int training,
                              float* __restrict__ ab, float* __restrict__ mean_istd) {
    __shared__ double sR[NTAP * NTAP], sP[NTAP], sQ[32 * NTAP], sWP[32 * NTAP];
    __shared__ float sW[32 * NTAP];
    if (training) {
        for (int i = threadIdx.x; i < NTAP * NTAP; i += blockDim.x) sR[i] = R[i];
        for (int i = threadIdx.x; i < NTAP; i += blockDim.x) sP[i] = P[i];
        for (int i = threadIdx.x; i < C * NTAP; i += blockDim.x) sW[i] = __ldg(W + i);
        __syncthreads();
        for (int i = threadIdx.x; i < C * NTAP; i += blockDim.x) {
            const int c = i / NTAP, a = i - c * NTAP;
            const float* w = sW + c * NTAP;
            double rw = 0.0;
#pragma unroll 4
            for (int b = 0; b < NTAP; ++b) rw += sR[a * NTAP + b] * (double)w[b];
            sQ[i] = (double)w[a] * rw;
            sWP[i] = (double)w[a] * sP[a];
        }
        __syncthreads();
    }
    int c = threadIdx.x;
    if (c < C) {
        float mean, var;
        if (training) {
            double wp = 0.0, q = 0.0;
            for (int a = 0; a < NTAP; ++a) { q += sQ[c * NTAP + a]; wp += sWP[c * NTAP + a]; }
            double bc = (double)bias[c];
            double m = wp / count + bc;
            double ex2 = (q + 2.0 * bc * wp) / count + bc * bc;
            double v = ex2 - m * m;
            if (v < 0.0) v = 0.0;
            mean = (float)m; var = (float)v;
            double unb = count > 1.0 ? v * count / (count - 1.0) : v;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
        } else {
            mean = running_mean[c]; var = running_var[c];
        }
        float istd = rsqrtf(var + eps);
        istd = istd * (1.5f - 0.5f * (var + eps) * istd * istd);
        float a = gamma[c] * istd;
        ab[c] = a;
        ab[C + c] = beta[c] - mean * a;
        mean_istd[c] = mean;
        mean_istd[C + c] = istd;
    }
    if (training && threadIdx.x == 0) nbt[0] += 1;
}

// ------------------------------------------------------------------------------------------
// fused forward of both conv blocks: img -> conv1 -> BN1 -> ReLU -> pool -> conv2 -> x2 (+ BN2 sums).
// thread = pooled pixel: its 2x2 window of conv1 outputs is computed from a 4x4 register patch per input
// channel (16 loads feed 36*C FMAs), so conv1 never touches global memory.  Optionally saves the pre-BN
// value at the pool arg (e1) and the arg index | active bit (idx1) for the backward.
template <int C>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
scene_fused12_fwd_kernel(const float* __restrict__ img, const int* __restrict__ rows, int N,
                         const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ ab1,
                         const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ x2,
                         double* __restrict__ stats2, float* __restrict__ e1, unsigned char* __restrict__ idx1) {
    constexpr int PPAD = FwdPad::PPAD;           // channel stride of the pooled map: 8 mod 32 (conflict-free A fragments)
    constexpr int LDW2 = FwdLdw2<C>::value;        // row stride of the conv2 weights [tap][ci][co]: 8 t + g distinct banks
    extern __shared__ __align__(16) float smem[];
    float* sRaw = smem;                          // [2][4][33][33]  double-buffered crop, filled by cp.async.bulk
    float* sW1 = sRaw + 2 * RAW;                 // [36 taps][C]
    float* sP = sW1 + NTAP * C;                  // [C][PPAD]
    float* sW2 = sP + ((C * PPAD + 3) & ~3);     // [9 taps][C in][LDW2]   (C out used)
    float* sAB = sW2 + 9 * C * LDW2;             // [2C]
    float* sred = sAB + 2 * C;                   // [8][2C]
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sred + 8 * 2 * C);      // [2] one mbarrier per crop buffer
    if (threadIdx.x == 0) {
        sbar_init(smem_addr(sBar), 1);
        sbar_init(smem_addr(sBar + 1), 1);
        sbar_fence_init();
    }
    for (int i = threadIdx.x; i < NTAP * C; i += MGGAN_THREADS) {
        int c = i / NTAP, tap = i - c * NTAP;
        sW1[tap * C + c] = __ldg(W1 + i);
    }
    for (int i = threadIdx.x; i < 9 * C * C; i += MGGAN_THREADS) {
        int co = i / (9 * C), r = i - co * 9 * C, ci = r / 9, tap = r - ci * 9;
        sW2[(tap * C + ci) * LDW2 + co] = __ldg(W2 + i);
    }
    if (threadIdx.x < 2 * C) sAB[threadIdx.x] = __ldg(ab1 + threadIdx.x);
    for (int i = threadIdx.x; i < C * PPAD; i += MGGAN_THREADS) sP[i] = 0.f;
    __syncthreads();                              // barriers initialised before anyone arms or polls them
    if (threadIdx.x == 0 && (int)blockIdx.x < N) stage_crop(img, rows ? rows[blockIdx.x] : blockIdx.x, sRaw, sBar);
    constexpr int NT = C / 8;                     // conv2 n-tiles (8 output channels each) = k-steps per tap (8 input channels)
    float st[4 * NT];                             // BatchNorm-2 partial sums: channels 8 j + 2 t + {0, 1}: sum [2j + e], squares [2NT + 2j + e]
#pragma unroll
    for (int c = 0; c < 4 * NT; ++c) st[c] = 0.f;
    const int py = threadIdx.x >> 4, px = threadIdx.x & 15;
    const int warp = threadIdx.x >> 5, g8 = (threadIdx.x & 31) >> 2, t4 = threadIdx.x & 3;

    int j = 0;                                    // agents this CTA has processed: buffer j & 1, barrier phase (j >> 1) & 1
    for (int n = blockIdx.x; n < N; n += gridDim.x, ++j) {
        __syncthreads();                          // previous agent: conv2 has read sP, conv1 has read the other crop buffer
        const float* sImg = sRaw + (j & 1) * RAW;
        if (threadIdx.x == 0 && n + (int)gridDim.x < N) {
            const int nn = n + gridDim.x;
            stage_crop(img, rows ? rows[nn] : nn, sRaw + ((j + 1) & 1) * RAW, sBar + ((j + 1) & 1));
        }
        sbar_wait(smem_addr(sBar + (j & 1)), (j >> 1) & 1);
        {
            float acc[4][C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float bv = __ldg(b1 + c);
                acc[0][c] = bv; acc[1][c] = bv; acc[2][c] = bv; acc[3][c] = bv;
            }
#pragma unroll 1
            for (int ci = 0; ci < CIN; ++ci) {
                float pt[4][4];                   // rows 2 py - 1 .. 2 py + 2, columns 2 px - 1 .. 2 px + 2 of the crop
                const float* bp = sImg + ci * IMG2 + (2 * py - 1) * IMG + 2 * px - 1;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        pt[r][cc] = ((r > 0 || py > 0) && (cc > 0 || px > 0)) ? bp[r * IMG + cc] : 0.f;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float* wp = sW1 + (ci * 9 + ky * 3 + kx) * C;
#pragma unroll
                        for (int c = 0; c < C; c += 4) {
                            float4 w = ld4(wp + c);
#pragma unroll
                            for (int d = 0; d < 4; ++d) {
                                const float v = pt[(d >> 1) + ky][(d & 1) + kx];
                                acc[d][c] = fmaf(v, w.x, acc[d][c]); acc[d][c + 1] = fmaf(v, w.y, acc[d][c + 1]);
                                acc[d][c + 2] = fmaf(v, w.z, acc[d][c + 2]); acc[d][c + 3] = fmaf(v, w.w, acc[d][c + 3]);
                            }
                        }
                    }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float a = sAB[c], b = sAB[C + c];
                float m = fmaf(a, acc[0][c], b), e = acc[0][c]; int arg = 0;
#pragma unroll
                for (int d = 1; d < 4; ++d) {
                    float v = fmaf(a, acc[d][c], b);
                    if (v > m) { m = v; e = acc[d][c]; arg = d; }
                }
                sP[c * PPAD + (py + 1) * LDP + px + 1] = fmaxf(m, 0.f);
                if (e1 != nullptr) {
                    e1[((size_t)n * C + c) * P1SQ + threadIdx.x] = e;
                    idx1[((size_t)n * C + c) * P1SQ + threadIdx.x] = (unsigned char)(arg | (m > 0.f ? 4 : 0));
                }
            }
        }
        __syncthreads();
        {   // conv2 as warp-level 3 x TF32 tensor-core products (common.cuh): m-tile = one output row (16 pixels), n-tile =
            // 8 output channels, k-step = 8 input channels of one tap; every fragment element is a conflict-free LDS.32.
            // The FP32 register-tile form of this stage was half of the kernel's FMA issue and most of its LDS traffic.
            {
                const int y = warp * 2;                   // the warp's two rows are interleaved: 2 NT independent accumulator chains
                float acc[2][NT][4];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float ba = __ldg(b2 + 8 * j + 2 * t4), bb = __ldg(b2 + 8 * j + 2 * t4 + 1);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) { acc[rr][j][0] = ba; acc[rr][j][1] = bb; acc[rr][j][2] = ba; acc[rr][j][3] = bb; }
                }
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int ks = 0; ks < NT; ++ks) {
                        const float* pb = sW2 + (tap * C + ks * 8 + t4) * LDW2 + g8;
                        float b0[NT], b1[NT];
#pragma unroll
                        for (int j = 0; j < NT; ++j) { b0[j] = pb[8 * j]; b1[j] = pb[4 * LDW2 + 8 * j]; }
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const float* pa = sP + (ks * 8 + t4) * PPAD + (y + rr + tap / 3) * LDP + g8 + tap % 3;
                            uint32_t ah[4], al[4];
                            tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8], ah[1], al[1]);
                            tf32_split(pa[4 * PPAD], ah[2], al[2]); tf32_split(pa[4 * PPAD + 8], ah[3], al[3]);
#pragma unroll
                            for (int j = 0; j < NT; ++j) mma_3xtf32(acc[rr][j], ah, al, b0[j], b1[j]);
                        }
                    }
                }
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        float* o = x2 + ((size_t)n * C + 8 * j + 2 * t4) * P1SQ + (y + rr) * P1 + g8;
                        const float* a = acc[rr][j];
                        o[0] = a[0]; o[8] = a[2]; o[P1SQ] = a[1]; o[P1SQ + 8] = a[3];
                        st[2 * j] += a[0] + a[2];
                        st[2 * j + 1] += a[1] + a[3];
                        st[2 * NT + 2 * j] = fmaf(a[0], a[0], fmaf(a[2], a[2], st[2 * NT + 2 * j]));
                        st[2 * NT + 2 * j + 1] = fmaf(a[1], a[1], fmaf(a[3], a[3], st[2 * NT + 2 * j + 1]));
                    }
            }
        }
    }
    if (stats2 != nullptr) {     // lanes that differ in g hold the same channels: shuffle, then shared / global double atomics
        double* sd = reinterpret_cast<double*>(sred);
        __syncthreads();
        if (threadIdx.x < 2 * C) sd[threadIdx.x] = 0.0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4 * NT; ++i) {
            float v = st[i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g8 == 0) {
                const int sq = i >= 2 * NT, jj = (i - (sq ? 2 * NT : 0));
                atomicAdd(sd + (sq ? C : 0) + 8 * (jj >> 1) + 2 * t4 + (jj & 1), (double)v);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * C) atomicAdd(stats2 + threadIdx.x, sd[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------
// fused backward of both conv blocks.  BN2 backward (dense dx2) -> conv2 weight / input gradients ->
// pool / ReLU / BN1 -> sparse dy1 (kept in shared memory) -> (a) BN1 sums, (b) the sparse half of the
// conv1 weight gradient S1[c][tap] = sum dy1 * patch_tap at the pool-arg positions.  The dense half
// follows from the patch statistics in scene_bn1_bwd_finalize.
template <int C>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
scene_fused12_bwd_kernel(const float* __restrict__ img, const int* __restrict__ rows, int N,
                         const float* __restrict__ x2, const float* __restrict__ e1,
                         const unsigned char* __restrict__ idx1, const float* __restrict__ ab1,
                         const float* __restrict__ mean_istd1, const float* __restrict__ ab2,
                         const float* __restrict__ mean_istd2, const float* __restrict__ m12_2,
                         const float* __restrict__ W2, const float* __restrict__ dy2,
                         const unsigned char* __restrict__ idx2, float* __restrict__ dW2, float* __restrict__ dbias2,
                         float* __restrict__ S1, double* __restrict__ sums1) {
    constexpr int NT = C / 8;                            // 8-channel tiles (n-tiles; also k-steps per tap of the input gradient)
    constexpr int UNITS = 9 * NT;                        // (tap, input-channel tile) units of the conv2 weight gradient
    constexpr int UPW = (UNITS + 7) / 8;                 // units per warp (3 or 2)
    constexpr int QG = MGGAN_THREADS / (C * CIN);        // pooled-pixel groups for the sparse conv1 term (4 or 8)
    constexpr int Q_PER = P1SQ / QG;
    constexpr int LDY = P1SQ + 4;                        // channel stride of sDY: neighbouring channels on different banks
    constexpr int PPAD = BwdPad::PPAD;                   // 324 = 4 mod 32: bank 4 g + t (weight gradient) / 8 t + g (input gradient)
    constexpr int LDWD = BwdPad::LDWD;                   // 20: rows co = c0 + 2t of the input-gradient B fragment 8 banks apart
    extern __shared__ __align__(16) float smem[];
    float* sDX = smem;                                   // [C][PPAD]  dx2 with zero halo
    float* sP = sDX + C * PPAD;                          // [C][PPAD]  p1 with zero halo
    float* sWT = sP + ((C * PPAD + 3) & ~3);             // [tap][co][LDWD]  (ci used)
    float* sDY = sWT + 9 * C * LDWD;                     // [C][LDY] dy1 (sparse values, dense layout)
    float* sImg = sDY + C * LDY;                         // [4][35][36]
    float* sPar = sImg + CIN * IMGPAD_BWD;                   // ab1[2C] mi1[2C] ab2[2C] mi2[2C] m12_2[2C]
    float* sred = sPar + 10 * C;                         // [8][2C]
    unsigned char* sIdx = reinterpret_cast<unsigned char*>(sred + 8 * 2 * C);    // [C][256]
    unsigned char* sList = sIdx + C * P1SQ;                                      // [C][256] pooled pixels with a non-zero dy1
    for (int i = threadIdx.x; i < 9 * C * C; i += MGGAN_THREADS) {
        int co = i / (9 * C), r = i - co * 9 * C, ci = r / 9, tap = r - ci * 9;
        sWT[(tap * C + co) * LDWD + ci] = __ldg(W2 + i);
    }
    if (threadIdx.x < 2 * C) {
        sPar[threadIdx.x] = __ldg(ab1 + threadIdx.x);
        sPar[2 * C + threadIdx.x] = __ldg(mean_istd1 + threadIdx.x);
        sPar[4 * C + threadIdx.x] = __ldg(ab2 + threadIdx.x);
        sPar[6 * C + threadIdx.x] = __ldg(mean_istd2 + threadIdx.x);
        sPar[8 * C + threadIdx.x] = __ldg(m12_2 + threadIdx.x);
    }
    for (int i = threadIdx.x; i < C * PPAD; i += MGGAN_THREADS) { sDX[i] = 0.f; sP[i] = 0.f; }
    for (int i = threadIdx.x; i < CIN * IMGPAD_BWD; i += MGGAN_THREADS) sImg[i] = 0.f;
    const int y = threadIdx.x >> 4, x = threadIdx.x & 15;
    const int warp = threadIdx.x >> 5, g8 = (threadIdx.x & 31) >> 2, t4 = threadIdx.x & 3;     // MMA fragment coordinates
    const int s_c = threadIdx.x / (CIN * QG), s_ci = (threadIdx.x / QG) % CIN, s_qg = threadIdx.x % QG;
    float wacc[UPW][4], sacc[9];                         // wacc[i]: C fragment of unit warp + 8 i (kept across agents)
#pragma unroll
    for (int i = 0; i < 9; ++i) sacc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < UPW; ++i) { wacc[i][0] = 0.f; wacc[i][1] = 0.f; wacc[i][2] = 0.f; wacc[i][3] = 0.f; }
    float dbs[C], st[4 * NT];                            // st: BatchNorm-1 sums of channels 8 j + 2 t + e: [2j + e], [2NT + 2j + e]
#pragma unroll
    for (int c = 0; c < C; ++c) dbs[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 4 * NT; ++c) st[c] = 0.f;

    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        const int src = rows ? rows[n] : n;
        __syncthreads();
        {
            // the crop is only read by the last stage (sparse conv1 weight gradient): cp.async (LDGSTS) it into the padded
            // layout now and wait for it there, behind the two tensor-core stages
            const float* ip = img + (size_t)src * CIN * IMG2;
            for (int i = threadIdx.x; i < CIN * IMG2; i += MGGAN_THREADS) {
                int ci = i / IMG2, p = i - ci * IMG2, yy = p / IMG, xx = p - yy * IMG;
                cp_async4(&sImg[ci * IMGPAD_BWD + (yy + 1) * LDI + xx + 1], ip + i);
            }
            cp_async_commit();
            const int win = (y >> 1) * P2 + (x >> 1), loc = (y & 1) * 2 + (x & 1);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float xv = __ldg(x2 + ((size_t)n * C + c) * P1SQ + threadIdx.x);
                float xh = (xv - sPar[6 * C + c]) * sPar[7 * C + c];
                float d = 0.f;
                if (idx2[((size_t)n * C + c) * P2SQ + win] == loc) d = __ldg(dy2 + ((size_t)n * C + c) * P2SQ + win);
                float dx = sPar[4 * C + c] * (d - sPar[8 * C + c] - xh * sPar[9 * C + c]);
                sDX[c * PPAD + (y + 1) * LDP + x + 1] = dx;
                dbs[c] += dx;
            }
            for (int i = threadIdx.x; i < C * P1SQ; i += MGGAN_THREADS) {
                int c = i >> 8, pp = i & 255;
                float e = __ldg(e1 + (size_t)n * C * P1SQ + i);
                sIdx[i] = idx1[(size_t)n * C * P1SQ + i];
                sP[c * PPAD + ((pp >> 4) + 1) * LDP + (pp & 15) + 1] = fmaxf(fmaf(sPar[c], e, sPar[C + c]), 0.f);
            }
        }
        __syncthreads();
        {   // conv2 weight gradient dW2[co][ci][tap] = sum_pix dx2[co][pix] p1[ci][pix + tap] as warp-level 3 x TF32 products:
            // M = output channels, N = 8 input channels, K = pixels (8 consecutive x per step); a warp owns the units
            // (tap, ci tile) warp, warp + 8, warp + 16 and keeps their C fragments across agents.
#pragma unroll 1
            for (int ks = 0; ks < 2 * P1; ++ks) {
                const int yy = ks >> 1, x0 = (ks & 1) * 8;
                const float* pa = sDX + g8 * PPAD + (yy + 1) * LDP + x0 + t4 + 1;
                uint32_t ah[4], al[4];
                tf32_split(pa[0], ah[0], al[0]);
                tf32_split(pa[4], ah[2], al[2]);
                if (C == 16) {
                    tf32_split(pa[8 * PPAD], ah[1], al[1]);
                    tf32_split(pa[8 * PPAD + 4], ah[3], al[3]);
                } else {
                    ah[1] = al[1] = ah[3] = al[3] = 0u;
                }
#pragma unroll
                for (int i = 0; i < UPW; ++i) {
                    const int unit = warp + 8 * i;
                    if (unit < UNITS) {
                        const int tap = unit / NT, j = unit - tap * NT;
                        const float* pb = sP + (8 * j + g8) * PPAD + (yy + tap / 3) * LDP + x0 + t4 + tap % 3;
                        mma_3xtf32(wacc[i], ah, al, pb[0], pb[4]);
                    }
                }
            }
        }
        {   // conv2 input gradient as warp-level 3 x TF32 products: m-tile = one row of 16 pixels, N = 8 input channels,
            // K = output channels of one tap (logical k = t, t + 4 <-> co = c0 + 2t, c0 + 2t + 1), then pool / ReLU / BN1 -> dy1
#pragma unroll 1
            for (int rr = 0; rr < 2; ++rr) {
                const int yy = warp * 2 + rr;
                float acc[NT][4];
#pragma unroll
                for (int j = 0; j < NT; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int ks = 0; ks < NT; ++ks) {
                        const float* pa = sDX + (ks * 8 + 2 * t4) * PPAD + (yy + 2 - tap / 3) * LDP + g8 + 2 - tap % 3;
                        uint32_t ah[4], al[4];
                        tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8], ah[1], al[1]);
                        tf32_split(pa[PPAD], ah[2], al[2]); tf32_split(pa[PPAD + 8], ah[3], al[3]);
                        const float* pb = sWT + (tap * C + ks * 8 + 2 * t4) * LDWD + g8;
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_3xtf32(acc[j], ah, al, pb[8 * j], pb[LDWD + 8 * j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int ch = 8 * j + 2 * t4 + e;
                        const float mu = sPar[2 * C + ch], is = sPar[3 * C + ch];
#pragma unroll
                        for (int hx = 0; hx < 2; ++hx) {
                            const int pix = yy * P1 + g8 + 8 * hx;
                            const float ev = __ldg(e1 + ((size_t)n * C + ch) * P1SQ + pix);
                            const float d = (sIdx[ch * P1SQ + pix] & 4) ? acc[j][2 * hx + e] : 0.f;
                            st[2 * j + e] += d;
                            st[2 * NT + 2 * j + e] = fmaf(d, (ev - mu) * is, st[2 * NT + 2 * j + e]);
                            sDY[ch * LDY + pix] = d;
                        }
                    }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        {   // sparse half of the conv1 weight gradient: thread = (c, ci, lane group).  dy1 is zero wherever the ReLU was
            // inactive (about half of the pooled pixels), so every warp first compacts the non-zero pixels of the channels
            // it owns (C = 16: channels 2 warp, 2 warp + 1; C = 8: channel warp) into a list with warp ballots, and the lane
            // groups then walk the list: every lane of every trip does useful work
            constexpr int CPW = C / 8;                       // channels per warp
            const int lane = threadIdx.x & 31;
            int nnz = 0;
#pragma unroll
            for (int cc = 0; cc < CPW; ++cc) {
                const int ch = warp * CPW + cc;
                int count = 0;
#pragma unroll 2
                for (int base = 0; base < P1SQ; base += 32) {
                    const bool on = sDY[ch * LDY + base + lane] != 0.f;
                    const unsigned m = __ballot_sync(0xffffffffu, on);
                    if (on) sList[ch * P1SQ + count + __popc(m & ((1u << lane) - 1u))] = (unsigned char)(base + lane);
                    count += __popc(m);
                }
                if (ch == s_c) nnz = count;
            }
            __syncwarp();
            const float* ipc = sImg + s_ci * IMGPAD_BWD;
#pragma unroll 2
            for (int e = s_qg; e < nnz; e += QG) {
                const int q = sList[s_c * P1SQ + e];
                const float d = sDY[s_c * LDY + q];
                const int code = sIdx[s_c * P1SQ + q] & 3;
                const float* bp = ipc + (2 * (q >> 4) + (code >> 1)) * LDI + 2 * (q & 15) + (code & 1);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) sacc[ky * 3 + kx] = fmaf(d, bp[ky * LDI + kx], sacc[ky * 3 + kx]);
            }
        }
    }
    // conv2 weight gradient: C fragments -> global (c0, c1: co = g, ci = 8 j + 2t + {0, 1}; c2, c3: co = g + 8)
#pragma unroll
    for (int i = 0; i < UPW; ++i) {
        const int unit = warp + 8 * i;
        if (unit < UNITS) {
            const int tap = unit / NT, j = unit - tap * NT;
            float* dst = dW2 + ((size_t)g8 * C + 8 * j + 2 * t4) * 9 + tap;
            atomicAdd(dst, wacc[i][0]);
            atomicAdd(dst + 9, wacc[i][1]);
            if (C == 16) {
                atomicAdd(dst + 8 * C * 9, wacc[i][2]);
                atomicAdd(dst + 8 * C * 9 + 9, wacc[i][3]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(S1 + ((size_t)s_c * CIN + s_ci) * 9 + t, sacc[t]);
    block_reduce_to_global_f<C>(dbs, dbias2, sred);
    {   // BatchNorm-1 sums: lanes that differ in g hold the same channels
        double* sd = reinterpret_cast<double*>(sred);
        __syncthreads();
        if (threadIdx.x < 2 * C) sd[threadIdx.x] = 0.0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4 * NT; ++i) {
            float v = st[i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g8 == 0) {
                const int sq = i >= 2 * NT, jj = i - (sq ? 2 * NT : 0);
                atomicAdd(sd + (sq ? C : 0) + 8 * (jj >> 1) + 2 * t4 + (jj & 1), (double)v);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * C) atomicAdd(sums1 + threadIdx.x, sd[threadIdx.x]);
    }
}

// BatchNorm-1 backward closed form (one thread per (channel, tap), double precision):
//   dW1[c][a] = a_c ( S1[c][a] - m1_c P_a - m2_c istd_c ( (R W_c)_a + b_c P_a - mu_c P_a ) )
// m1, m2: means of dy1 and dy1 * xhat1 over ALL `count` conv1 outputs (global sums when data-parallel);
// S1, R, P are the local shard's (the gradient all-reduce adds the shards).  conv1's bias gradient is
// identically zero under train-mode BatchNorm.
__global__ void scene_bn1_bwd_finalize_kernel(const double* __restrict__ sums_global, const double* __restrict__ sums_local,
                                              double count, int C, const float* __restrict__ S1,
                                              const double* __restrict__ R, const double* __restrict__ P,
                                              const float* __restrict__ W, const float* __restrict__ bias,
                                              const float* __restrict__ ab1, const float* __restrict__ mean_istd1,
                                              float* __restrict__ dW, float* __restrict__ dgamma,
                                              float* __restrict__ dbeta) {
    for (int i = threadIdx.x; i < C * NTAP; i += blockDim.x) {
        const int c = i / NTAP, a = i - c * NTAP;
        const double m1 = sums_global[c] / count, m2 = sums_global[C + c] / count;
        const float* w = W + c * NTAP;
        double rw = 0.0;
        for (int b = 0; b < NTAP; ++b) rw += R[a * NTAP + b] * (double)w[b];
        const double mu = (double)mean_istd1[c], istd = (double)mean_istd1[C + c];
        const double corr = rw + ((double)bias[c] - mu) * P[a];
        dW[i] = (float)((double)ab1[c] * ((double)S1[i] - m1 * P[a] - m2 * istd * corr));
    }
    if (threadIdx.x < C) {
        dbeta[threadIdx.x] = (float)sums_local[threadIdx.x];
        dgamma[threadIdx.x] = (float)sums_local[C + threadIdx.x];
    }
}

// ------------------------------------------------------------------------------------------
// pass C: BN2 -> ReLU -> pool -> per-position channel attention.  thread = (agent slot, position)
template <int C>
struct AttnW {
    float* sWa1;   // [AH][C]
    float* sWa2;   // [C][AH]
    float* sba1;   // [AH]
    float* sba2;   // [C]
    float* sAB;    // [2C]
};

template <int C>
__device__ __forceinline__ AttnW<C> stage_attn_weights(float* base, const float* Wa1, const float* ba1, const float* Wa2,
                                                       const float* ba2, const float* ab2) {
    AttnW<C> w;
    w.sWa1 = base; w.sWa2 = w.sWa1 + AH * C; w.sba1 = w.sWa2 + C * AH; w.sba2 = w.sba1 + AH; w.sAB = w.sba2 + C;
    for (int i = threadIdx.x; i < AH * C; i += MGGAN_THREADS) { w.sWa1[i] = __ldg(Wa1 + i); w.sWa2[i] = __ldg(Wa2 + i); }
    if (threadIdx.x < AH) w.sba1[threadIdx.x] = __ldg(ba1 + threadIdx.x);
    if (threadIdx.x < C) w.sba2[threadIdx.x] = __ldg(ba2 + threadIdx.x);
    if (threadIdx.x < 2 * C) w.sAB[threadIdx.x] = __ldg(ab2 + threadIdx.x);
    return w;
}

// v[c] = maxpool(relu(BN2(x2))) at one position; optionally the arg index / pre-BN value at the arg.
template <int C, bool WITH_ARG>
__device__ __forceinline__ void pool_block2(const float* __restrict__ x2n, const float* sAB, int pos, float (&v)[C],
                                            int (&arg)[C], float (&e)[C]) {
    const int py = pos >> 3, px = pos & 7;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float a = sAB[c], b = sAB[C + c];
        const float2* s = reinterpret_cast<const float2*>(x2n + c * P1SQ + (2 * py) * P1 + 2 * px);
        float2 t0 = __ldg(s), t1 = __ldg(s + P1 / 2);
        float v0 = fmaf(a, t0.x, b), v1 = fmaf(a, t0.y, b), v2 = fmaf(a, t1.x, b), v3 = fmaf(a, t1.y, b);
        float m = v0, ee = t0.x; int ag = 0;
        if (v1 > m) { m = v1; ee = t0.y; ag = 1; }
        if (v2 > m) { m = v2; ee = t1.x; ag = 2; }
        if (v3 > m) { m = v3; ee = t1.y; ag = 3; }
        v[c] = fmaxf(m, 0.f);
        if (WITH_ARG) { arg[c] = ag | (m > 0.f ? 4 : 0); e[c] = ee; }
    }
}

template <int C>
__device__ __forceinline__ void attn_mlp(const AttnW<C>& w, const float (&v)[C], float (&hp)[AH], float (&att)[C]) {
#pragma unroll
    for (int k = 0; k < AH; ++k) {
        float s = w.sba1[k];
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            float4 q = ld4(w.sWa1 + k * C + c);
            s = fmaf(q.x, v[c], fmaf(q.y, v[c + 1], fmaf(q.z, v[c + 2], fmaf(q.w, v[c + 3], s))));
        }
        hp[k] = s;                         // pre-activation
    }
    float mx = -3.4e38f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float s = w.sba2[c];
#pragma unroll
        for (int k = 0; k < AH; k += 4) {
            float4 q = ld4(w.sWa2 + c * AH + k);
            s = fmaf(q.x, lrelu_(hp[k], 0.01f), fmaf(q.y, lrelu_(hp[k + 1], 0.01f),
                fmaf(q.z, lrelu_(hp[k + 2], 0.01f), fmaf(q.w, lrelu_(hp[k + 3], 0.01f), s))));
        }
        att[c] = s;
        mx = fmaxf(mx, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { att[c] = __expf(att[c] - mx); sum += att[c]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < C; ++c) att[c] *= inv;
}

template <int C>
__global__ void __launch_bounds__(MGGAN_THREADS)
scene_attn_fwd_kernel(const float* __restrict__ x2, int N, const float* __restrict__ ab2, const float* __restrict__ Wa1,
                      const float* __restrict__ ba1, const float* __restrict__ Wa2, const float* __restrict__ ba2,
                      float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    AttnW<C> w = stage_attn_weights<C>(smem, Wa1, ba1, Wa2, ba2, ab2);
    __syncthreads();
    const int slot = threadIdx.x >> 6, pos = threadIdx.x & 63;
    for (int n0 = blockIdx.x * 4; n0 < N; n0 += gridDim.x * 4) {
        const int n = n0 + slot;
        if (n >= N) continue;
        float v[C], hp[AH], att[C], e[C]; int arg[C];
        pool_block2<C, false>(x2 + (size_t)n * C * P1SQ, w.sAB, pos, v, arg, e);
        attn_mlp<C>(w, v, hp, att);
        float o = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) o = fmaf(att[c], v[c], o);
        out[(size_t)n * P2SQ + pos] = o;
    }
}

// ------------------------------------------------------------------------------------------
// backward of pass C: d(out) -> attention-MLP weight grads, sparse dy2 (+ arg index), BN2 sums
template <int C>
__global__ void __launch_bounds__(MGGAN_THREADS)
scene_attn_bwd_kernel(const float* __restrict__ x2, int N, const float* __restrict__ ab2,
                      const float* __restrict__ mean_istd2, const float* __restrict__ Wa1,
                      const float* __restrict__ ba1, const float* __restrict__ Wa2, const float* __restrict__ ba2,
                      const float* __restrict__ dout, float* __restrict__ dWa1, float* __restrict__ dba1,
                      float* __restrict__ dWa2, float* __restrict__ dba2, float* __restrict__ dy2,
                      unsigned char* __restrict__ idx2, double* __restrict__ sums2) {
    constexpr int LDC = C + 4, LDA = AH + 4;
    constexpr int NBLK = 4 * C;                   // 2 * (C/4) * (AH/4) weight-gradient blocks
    constexpr int RG = MGGAN_THREADS / NBLK;      // row groups
    constexpr int RPG = MGGAN_THREADS / RG;       // rows per group
    extern __shared__ __align__(16) float smem[];
    AttnW<C> w = stage_attn_weights<C>(smem, Wa1, ba1, Wa2, ba2, ab2);
    float* sDS = w.sAB + 2 * C;                   // [256][LDC]
    float* sV = sDS + MGGAN_THREADS * LDC;        // [256][LDC]
    float* sHid = sV + MGGAN_THREADS * LDC;       // [256][LDA]
    float* sDH = sHid + MGGAN_THREADS * LDA;      // [256][LDA]
    float* sMI = sDH + MGGAN_THREADS * LDA;       // [2C]
    float* sred = sMI + 2 * C;                    // [8][2C]
    if (threadIdx.x < 2 * C) sMI[threadIdx.x] = __ldg(mean_istd2 + threadIdx.x);
    __syncthreads();
    const int slot = threadIdx.x >> 6, pos = threadIdx.x & 63;
    const int blk = threadIdx.x % NBLK, rg = threadIdx.x / NBLK;
    float wacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) wacc[i][j] = 0.f;
    float bacc = 0.f;            // dba2[c] for t < C ; dba1[k] for C <= t < C + AH
    float st[2 * C];
#pragma unroll
    for (int c = 0; c < 2 * C; ++c) st[c] = 0.f;

    for (int n0 = blockIdx.x * 4; n0 < N; n0 += gridDim.x * 4) {
        const int n = n0 + slot;
        float v[C], hp[AH], att[C], e[C], ds[C], dv[C]; int arg[C];
        float dhid[AH];
        __syncthreads();
        if (n < N) {
            pool_block2<C, true>(x2 + (size_t)n * C * P1SQ, w.sAB, pos, v, arg, e);
            attn_mlp<C>(w, v, hp, att);
            const float g = __ldg(dout + (size_t)n * P2SQ + pos);
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) dot = fmaf(att[c], g * v[c], dot);
#pragma unroll
            for (int c = 0; c < C; ++c) { ds[c] = att[c] * (g * v[c] - dot); dv[c] = g * att[c]; }
#pragma unroll
            for (int k = 0; k < AH; ++k) dhid[k] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int k = 0; k < AH; k += 4) {
                    float4 q = ld4(w.sWa2 + c * AH + k);
                    dhid[k] = fmaf(q.x, ds[c], dhid[k]); dhid[k + 1] = fmaf(q.y, ds[c], dhid[k + 1]);
                    dhid[k + 2] = fmaf(q.z, ds[c], dhid[k + 2]); dhid[k + 3] = fmaf(q.w, ds[c], dhid[k + 3]);
                }
#pragma unroll
            for (int k = 0; k < AH; ++k) {
                dhid[k] *= hp[k] > 0.f ? 1.f : 0.01f;
#pragma unroll
                for (int c = 0; c < C; c += 4) {
                    float4 q = ld4(w.sWa1 + k * C + c);
                    dv[c] = fmaf(q.x, dhid[k], dv[c]); dv[c + 1] = fmaf(q.y, dhid[k], dv[c + 1]);
                    dv[c + 2] = fmaf(q.z, dhid[k], dv[c + 2]); dv[c + 3] = fmaf(q.w, dhid[k], dv[c + 3]);
                }
            }
            // sparse gradient at the pooled arg position + BatchNorm sums
            float* dyo = dy2 + (size_t)n * C * P2SQ + pos;
            unsigned char* ixo = idx2 + (size_t)n * C * P2SQ + pos;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float d = (arg[c] & 4) ? dv[c] : 0.f;
                dyo[c * P2SQ] = d;
                ixo[c * P2SQ] = (unsigned char)(arg[c] & 3);
                float xh = (e[c] - sMI[c]) * sMI[C + c];
                st[c] += d;
                st[C + c] = fmaf(d, xh, st[C + c]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) { ds[c] = 0.f; v[c] = 0.f; }
#pragma unroll
            for (int k = 0; k < AH; ++k) { dhid[k] = 0.f; hp[k] = 0.f; }
        }
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            st4(sDS + threadIdx.x * LDC + c, make_float4(ds[c], ds[c + 1], ds[c + 2], ds[c + 3]));
            st4(sV + threadIdx.x * LDC + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
        }
#pragma unroll
        for (int k = 0; k < AH; k += 4) {
            st4(sHid + threadIdx.x * LDA + k, make_float4(lrelu_(hp[k], 0.01f), lrelu_(hp[k + 1], 0.01f),
                                                          lrelu_(hp[k + 2], 0.01f), lrelu_(hp[k + 3], 0.01f)));
            st4(sDH + threadIdx.x * LDA + k, make_float4(dhid[k], dhid[k + 1], dhid[k + 2], dhid[k + 3]));
        }
        __syncthreads();
        if (blk < NBLK / 2) {      // dWa2[c][k] += sum_r ds[r][c] hid[r][k]
            int oq = blk % (C / 4), kq = blk / (C / 4);
            tile_wgrad<RPG>(wacc, sDS + rg * RPG * LDC, LDC, oq * 4, sHid + rg * RPG * LDA, LDA, kq * 4);
        } else {                   // dWa1[k][c] += sum_r dhid[r][k] v[r][c]
            int b2 = blk - NBLK / 2;
            int oq = b2 % (AH / 4), kq = b2 / (AH / 4);
            tile_wgrad<RPG>(wacc, sDH + rg * RPG * LDA, LDA, oq * 4, sV + rg * RPG * LDC, LDC, kq * 4);
        }
        if (threadIdx.x < C) {
            for (int r = 0; r < MGGAN_THREADS; ++r) bacc += sDS[r * LDC + threadIdx.x];
        } else if (threadIdx.x < C + AH) {
            for (int r = 0; r < MGGAN_THREADS; ++r) bacc += sDH[r * LDA + threadIdx.x - C];
        }
    }
    if (blk < NBLK / 2) {
        int oq = blk % (C / 4), kq = blk / (C / 4);
        atomic_block44(dWa2, AH, oq * 4, kq * 4, wacc);
    } else {
        int b2 = blk - NBLK / 2;
        int oq = b2 % (AH / 4), kq = b2 / (AH / 4);
        atomic_block44(dWa1, C, oq * 4, kq * 4, wacc);
    }
    if (threadIdx.x < C) atomicAdd(dba2 + threadIdx.x, bacc);
    else if (threadIdx.x < C + AH) atomicAdd(dba1 + threadIdx.x - C, bacc);
    block_reduce_to_global<2 * C>(st, sums2, sred);
}


template <int C>
size_t fused_fwd_smem() { constexpr int PPAD = FwdPad::PPAD; return sizeof(float) * (2 * RAW + NTAP * C + ((C * PPAD + 3) & ~3) + 9 * C * FwdLdw2<C>::value + 2 * C + 8 * 2 * C) + 16; }
template <int C>
size_t fused_bwd_smem() {
    constexpr int PPAD = BwdPad::PPAD;
    return sizeof(float) * (C * PPAD + ((C * PPAD + 3) & ~3) + 9 * C * BwdPad::LDWD + C * (P1SQ + 4) + CIN * IMGPAD_BWD + 10 * C + 8 * 2 * C) + 2 * C * P1SQ;
}
template <int C>
size_t attn_w_floats() { return 2 * AH * C + AH + C + 2 * C; }
template <int C>
size_t attn_fwd_smem() { return sizeof(float) * attn_w_floats<C>(); }
template <int C>
size_t attn_bwd_smem() {
    return sizeof(float) * (attn_w_floats<C>() + 2 * MGGAN_THREADS * (C + 4) + 2 * MGGAN_THREADS * (AH + 4) + 2 * C + 8 * 2 * C);
}

int agent_grid(int N, int per_sm) {
    int g = sm_count() * per_sm;
    return N < g ? N : g;
}

#define SCENE_DISPATCH(C_, CALL16, CALL8)                                              \
    do {                                                                               \
        if ((C_) == 16) { CALL16; } else if ((C_) == 8) { CALL8; }                     \
        else return mggan_set_error(MGGAN_ERR_INVALID, "scene kernels built for 8 or 16 channels, got %d", (C_)); \
    } while (0)

template <int C>
int fused_fwd(const float* img, const int* rows, int N, const float* W1, const float* b1, const float* ab1, const float* W2,
              const float* b2, float* x2, double* stats2, float* e1, unsigned char* idx1, cudaStream_t s) {
    size_t sm = fused_fwd_smem<C>();
    cudaFuncSetAttribute(scene_fused12_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    scene_fused12_fwd_kernel<C><<<agent_grid(N, 2), MGGAN_THREADS, sm, s>>>(img, rows, N, W1, b1, ab1, W2, b2, x2, stats2, e1, idx1);
    return mggan_check_launch("scene_fused12_fwd");
}
template <int C>
int fused_bwd(const float* img, const int* rows, int N, const float* x2, const float* e1, const unsigned char* idx1,
              const float* ab1, const float* mi1, const float* ab2, const float* mi2, const float* m12_2, const float* W2,
              const float* dy2, const unsigned char* idx2, float* dW2, float* db2, float* S1, double* sums1, cudaStream_t s) {
    size_t sm = fused_bwd_smem<C>();
    cudaFuncSetAttribute(scene_fused12_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    scene_fused12_bwd_kernel<C><<<agent_grid(N, 2), MGGAN_THREADS, sm, s>>>(img, rows, N, x2, e1, idx1, ab1, mi1, ab2, mi2, m12_2,
                                                                          W2, dy2, idx2, dW2, db2, S1, sums1);
    return mggan_check_launch("scene_fused12_bwd");
}
template <int C>
int attn_fwd(const float* x2, int N, const float* ab2, const float* Wa1, const float* ba1, const float* Wa2, const float* ba2,
             float* out, cudaStream_t s) {
    size_t sm = attn_fwd_smem<C>();
    int g = (N + 3) / 4;
    int cap = sm_count() * 4;
    scene_attn_fwd_kernel<C><<<g < cap ? g : cap, MGGAN_THREADS, sm, s>>>(x2, N, ab2, Wa1, ba1, Wa2, ba2, out);
    return mggan_check_launch("scene_attn_fwd");
}
template <int C>
int attn_bwd(const float* x2, int N, const float* ab2, const float* mi2, const float* Wa1, const float* ba1, const float* Wa2,
             const float* ba2, const float* dout, float* dWa1, float* dba1, float* dWa2, float* dba2, float* dy2,
             unsigned char* idx2, double* sums2, cudaStream_t s) {
    size_t sm = attn_bwd_smem<C>();
    cudaFuncSetAttribute(scene_attn_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int g = (N + 3) / 4;
    int cap = sm_count() * 2;
    scene_attn_bwd_kernel<C><<<g < cap ? g : cap, MGGAN_THREADS, sm, s>>>(x2, N, ab2, mi2, Wa1, ba1, Wa2, ba2, dout, dWa1,
                                                                          dba1, dWa2, dba2, dy2, idx2, sums2);
    return mggan_check_launch("scene_attn_bwd");
}

}  // namespace

extern "C" int mggan_scene_patch_stats(const float* img, const int* rows, int N, double* R, double* P,
                                       cudaStream_t stream) {
    if (N <= 0) return MGGAN_OK;
    int slots = sm_count() * 3 / NPAIRS_CH;          // 44 agent slots x 10 channel pairs = 440 CTAs, 3 per SM
    if (slots > N) slots = N;
    scene_patch_stats_kernel<<<slots * NPAIRS_CH, PS_THREADS, 0, stream>>>(img, rows, N, slots, R, P);
    return mggan_check_launch("scene_patch_stats");
}

extern "C" int mggan_scene_bn1_from_patches(const double* R, const double* P, double count, int C, const float* W,
                                            const float* bias, const float* gamma, const float* beta,
                                            float* running_mean, float* running_var, long long* num_batches_tracked,
                                            float momentum, float eps, int training, float* ab, float* mean_istd,
                                            cudaStream_t stream) {
    MGGAN_REQUIRE(C >= 1 && C <= 32, "mggan_scene_bn1_from_patches: C=%d", C);
    scene_bn1_from_patches_kernel<<<1, MGGAN_THREADS, 0, stream>>>(R, P, count, C, W, bias, gamma, beta, running_mean, running_var,
                                                        num_batches_tracked, momentum, eps, training, ab, mean_istd);
    return mggan_check_launch("scene_bn1_from_patches");
}

extern "C" int mggan_scene_bn_finalize(const double* stats, double count, int C, const float* gamma, const float* beta,
                                       float* running_mean, float* running_var, long long* num_batches_tracked,
                                       float momentum, float eps, int training, float* ab, float* mean_istd,
                                       cudaStream_t stream) {
    MGGAN_REQUIRE(C >= 1 && C <= 32, "mggan_scene_bn_finalize: C=%d", C);
    scene_bn_finalize_kernel<<<1, 32, 0, stream>>>(stats, count, C, gamma, beta, running_mean, running_var,
                                                   num_batches_tracked, momentum, eps, training, ab, mean_istd);
    return mggan_check_launch("scene_bn_finalize");
}

extern "C" int mggan_scene_bn_bwd_finalize(const double* sums, double count, int C, float* m12, float* dgamma,
                                           float* dbeta, cudaStream_t stream) {
    MGGAN_REQUIRE(C >= 1 && C <= 32, "mggan_scene_bn_bwd_finalize: C=%d", C);
    scene_bn_bwd_finalize_kernel<<<1, 32, 0, stream>>>(sums, count, C, m12, dgamma, dbeta);
    return mggan_check_launch("scene_bn_bwd_finalize");
}

extern "C" int mggan_scene_fused12_fwd(const float* img, const int* rows, int N, int C, const float* W1, const float* b1,
                                       const float* ab1, const float* W2, const float* b2, float* x2, double* stats2,
                                       float* e1, unsigned char* idx1, cudaStream_t stream) {
    if (N <= 0) return MGGAN_OK;
    MGGAN_REQUIRE((e1 == nullptr) == (idx1 == nullptr), "mggan_scene_fused12_fwd: e1 and idx1 must both be set or both NULL");
    MGGAN_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0, "mggan_scene_fused12_fwd: img must be 16-byte aligned (bulk copy)");
    SCENE_DISPATCH(C, return fused_fwd<16>(img, rows, N, W1, b1, ab1, W2, b2, x2, stats2, e1, idx1, stream),
                   return fused_fwd<8>(img, rows, N, W1, b1, ab1, W2, b2, x2, stats2, e1, idx1, stream));
}

extern "C" int mggan_scene_attn_fwd(const float* x2, int N, int C, const float* ab2, const float* Wa1, const float* ba1,
                                    const float* Wa2, const float* ba2, float* out, cudaStream_t stream) {
    if (N <= 0) return MGGAN_OK;
    SCENE_DISPATCH(C, return attn_fwd<16>(x2, N, ab2, Wa1, ba1, Wa2, ba2, out, stream),
                   return attn_fwd<8>(x2, N, ab2, Wa1, ba1, Wa2, ba2, out, stream));
}

extern "C" int mggan_scene_attn_bwd(const float* x2, int N, int C, const float* ab2, const float* mean_istd2,
                                    const float* Wa1, const float* ba1, const float* Wa2, const float* ba2,
                                    const float* dout, float* dWa1, float* dba1, float* dWa2, float* dba2, float* dy2,
                                    unsigned char* idx2, double* sums2, cudaStream_t stream) {
    if (N <= 0) return MGGAN_OK;
    SCENE_DISPATCH(C, return attn_bwd<16>(x2, N, ab2, mean_istd2, Wa1, ba1, Wa2, ba2, dout, dWa1, dba1, dWa2, dba2, dy2, idx2, sums2, stream),
                   return attn_bwd<8>(x2, N, ab2, mean_istd2, Wa1, ba1, Wa2, ba2, dout, dWa1, dba1, dWa2, dba2, dy2, idx2, sums2, stream));
}

extern "C" int mggan_scene_fused12_bwd(const float* img, const int* rows, int N, int C, const float* x2, const float* e1,
                                       const unsigned char* idx1, const float* ab1, const float* mean_istd1,
                                       const float* ab2, const float* mean_istd2, const float* m12_2, const float* W2,
                                       const float* dy2, const unsigned char* idx2, float* dW2, float* dbias2, float* S1,
                                       double* sums1, cudaStream_t stream) {
    if (N <= 0) return MGGAN_OK;
    SCENE_DISPATCH(C, return fused_bwd<16>(img, rows, N, x2, e1, idx1, ab1, mean_istd1, ab2, mean_istd2, m12_2, W2, dy2, idx2, dW2, dbias2, S1, sums1, stream),
                   return fused_bwd<8>(img, rows, N, x2, e1, idx1, ab1, mean_istd1, ab2, mean_istd2, m12_2, W2, dy2, idx2, dW2, dbias2, S1, sums1, stream));
}

extern "C" int mggan_scene_bn1_bwd_finalize(const double* sums_global, const double* sums_local, double count, int C,
                                            const float* S1, const double* R, const double* P, const float* W,
                                            const float* bias, const float* ab1, const float* mean_istd1, float* dW,
                                            float* dgamma, float* dbeta, cudaStream_t stream) {
    MGGAN_REQUIRE(C >= 1 && C <= 32, "mggan_scene_bn1_bwd_finalize: C=%d", C);
    scene_bn1_bwd_finalize_kernel<<<1, 256, 0, stream>>>(sums_global, sums_local, count, C, S1, R, P, W, bias, ab1,
                                                         mean_istd1, dW, dgamma, dbeta);
    return mggan_check_launch("scene_bn1_bwd_finalize");
}
