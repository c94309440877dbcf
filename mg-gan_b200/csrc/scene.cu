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

// ---- bulk (TMA-engine) copies ---------------------------------------------------------------------------------------
// `cp.async.bulk` (SASS UBLKCP) moves a contiguous, 16-byte aligned chunk global -> shared without touching registers and
// signals an mbarrier when the bytes have landed; `cp.async.bulk.prefetch.L2` requests a chunk into the L2.  Used where a
// kernel consumes a contiguous per-agent block as it is: the conv2 outputs of the next trip in the attention backward
// (staged while the current trip is processed) and the L2 prefetch of the next crop's inputs in the fused backward.  The
// crops themselves are staged with 4-byte cp.async into zero-padded layouts (their 132-byte rows rule out both wider
// copies and tensor-map boxes).
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
// contiguous chunk -> L2 through the bulk-copy engine (no registers, no completion tracking); bytes % 16 == 0
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
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
// The statistics are a Gram matrix X X^T of the (tap x pixel) patch matrix X of every crop, i.e. a tensor-pipe product
// with the contraction over pixels: warp-level m16n8k8 TF32 MMAs (common.cuh) whose A and B fragments are the SAME
// loads (both index X[tap][pixel]: 8 consecutive taps x 8 consecutive pixels per group), so one k-step of 8 pixels is
// 10 conflict-free LDS.32 per lane feeding the 9 upper-triangular (16 x 8) tiles of the 40 x 40 product (36 taps, a
// constant-1 tap whose column yields P, 3 zero taps).  Crops cut from 8-bit images (-1 + u8 / 128, the reference's
// BaseTrajectories.py:278-286, and the synthetic batches) are exactly representable in TF32, so their products are
// exact with ONE MMA per tile; the kernel tests every staged crop (low 13 mantissa bits) and falls back to the
// 3 x TF32 split products for arbitrary fp32 input.  FP32 thread-per-pixel form before: 1.11 ms per batch of 16,384.
constexpr int NTAP = 36;
constexpr int G_LDR = 38;                          // padded row: the 8 taps x 4 pixels of a fragment load fall on distinct banks
constexpr int G_CHS = 35 * G_LDR;                  // padded channel [35][38], zero halo
constexpr int G_BUF = CIN * G_CHS + 24;            // + slack: masked reads of the tail k-step stay inside the buffer
constexpr int G_KSTEPS = (IMG2 + 7) / 8;           // 137 k-steps of 8 pixels (the last holds one pixel)
constexpr int G_NT = 40;                           // taps padded to 5 groups of 8

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }

// one k-step of the Gram product: SPLIT = false feeds the raw fp32 bits (exact TF32 operands), true the (hi, lo) split
template <bool SPLIT>
__device__ __forceinline__ void gram_kstep(float (&acc)[9][4], const float (&x)[5][2]) {
    uint32_t hi[5][2], lo[5][2];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (SPLIT) tf32_split(x[j][h], hi[j][h], lo[j][h]);
            else { hi[j][h] = __float_as_uint(x[j][h]); lo[j][h] = 0u; }
        }
    int tile = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // m-tile i = tap groups 2i (rows g) and 2i + 1 (rows g + 8; absent for i = 2)
        const uint32_t ah[4] = {hi[2 * i][0], i < 2 ? hi[2 * i + 1][0] : 0u, hi[2 * i][1], i < 2 ? hi[2 * i + 1][1] : 0u};
        const uint32_t al[4] = {lo[2 * i][0], i < 2 ? lo[2 * i + 1][0] : 0u, lo[2 * i][1], i < 2 ? lo[2 * i + 1][1] : 0u};
#pragma unroll
        for (int j = 2 * i; j < 5; ++j, ++tile) {
            mma_tf32_16x8x8(acc[tile], ah, hi[j][0], hi[j][1]);
            if (SPLIT) {
                mma_tf32_16x8x8(acc[tile], al, hi[j][0], hi[j][1]);
                mma_tf32_16x8x8(acc[tile], ah, lo[j][0], lo[j][1]);
            }
        }
    }
}

__global__ void __launch_bounds__(MGGAN_THREADS, 2)
scene_patch_stats_kernel(const float* __restrict__ img, const int* __restrict__ rows, int N,
                         double* __restrict__ R, double* __restrict__ P) {
    extern __shared__ __align__(16) float smem[];
    float* sImg = smem;                                                   // [2][G_BUF] double-buffered padded crop
    double* sD = reinterpret_cast<double*>(smem + 2 * G_BUF);             // [G_NT][G_NT] CTA reduction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 2 * G_BUF; i += MGGAN_THREADS) sImg[i] = 0.f;
    for (int i = threadIdx.x; i < G_NT * G_NT; i += MGGAN_THREADS) sD[i] = 0.0;
    int toff[5];                                       // tap 8 j + g of this lane: offset of its window corner
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int tap = 8 * j + g;
        const int ci = tap / 9, r = tap - ci * 9;
        toff[j] = tap < NTAP ? ci * G_CHS + (r / 3) * G_LDR + r % 3 : 0;
    }
    const bool tail_tap = g < 4;                       // group 4: taps 32..35, then the constant-1 tap, then zeros
    const float tail_const = g == 4 ? 1.f : 0.f;
    float acc[9][4];
#pragma unroll
    for (int q = 0; q < 9; ++q) { acc[q][0] = 0.f; acc[q][1] = 0.f; acc[q][2] = 0.f; acc[q][3] = 0.f; }
    __syncthreads();

    // warp w copies rows w, w + 8, ... of the crop (row = (channel, y): 33 floats, only 4-byte aligned) into the padded
    // layout: lanes = consecutive x (coalesced; a lane-per-row mapping touched 32 lines per instruction and made the
    // staging, not the products, the kernel's critical path), lane 0 also takes x = 32
    auto stage = [&](int n, int buf) {
        const int src = rows ? rows[n] : n;
        const float* ip = img + (size_t)src * RAW;
        float* base = sImg + buf * G_BUF + G_LDR + 1;
        for (int r = warp; r < CIN * IMG; r += MGGAN_THREADS / 32) {
            const int ci = r / IMG, y = r - ci * IMG;
            float* d = base + ci * G_CHS + y * G_LDR;
            cp_async4(d + lane, ip + r * IMG + lane);
            if (lane == 0) cp_async4(d + 32, ip + r * IMG + 32);
        }
        cp_async_commit();
    };
    int buf = 0;
    if ((int)blockIdx.x < N) stage(blockIdx.x, 0);
    for (int n = blockIdx.x; n < N; n += gridDim.x, buf ^= 1) {
        cp_async_wait<0>();
        __syncthreads();                               // crop n has landed; every warp is done with the other buffer
        if (n + (int)gridDim.x < N) stage(n + gridDim.x, buf ^ 1);
        const float* sI = sImg + buf * G_BUF;
        uint32_t low = 0u;                             // any low mantissa bit set -> not a TF32 number
        for (int i = threadIdx.x; i < CIN * G_CHS / 4; i += MGGAN_THREADS) {
            const float4 v = ld4(sI + 4 * i);
            low |= __float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w);
        }
        const bool split = __syncthreads_or((low & 0x1FFFu) != 0u);
        // The tensor pipe adds into its fp32 accumulator with truncation (measured: a one-sided -3.5e-6 on the all-positive
        // diagonal sums after ~100 accumulations), so an accumulator only ever holds ONE crop's 17 k-steps; the sum over
        // the CTA's crops is carried in a second register set with round-to-nearest adds.
        float cur[9][4];
#pragma unroll
        for (int q = 0; q < 9; ++q) { cur[q][0] = 0.f; cur[q][1] = 0.f; cur[q][2] = 0.f; cur[q][3] = 0.f; }
        for (int ks = warp; ks < G_KSTEPS; ks += MGGAN_THREADS / 32) {
            const int p0 = ks * 8 + t, p1 = p0 + 4;    // the lane's two pixels (k = t, t + 4)
            const int y0 = p0 / IMG, y1 = p1 / IMG;
            const int a0 = p0 + y0 * (G_LDR - IMG), a1 = p1 + y1 * (G_LDR - IMG);
            float x[5][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) { x[j][0] = sI[toff[j] + a0]; x[j][1] = sI[toff[j] + a1]; }
            x[4][0] = tail_tap ? sI[toff[4] + a0] : tail_const;
            x[4][1] = tail_tap ? sI[toff[4] + a1] : tail_const;
            if (ks == G_KSTEPS - 1) {
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    if (p0 >= IMG2) x[j][0] = 0.f;
                    if (p1 >= IMG2) x[j][1] = 0.f;
                }
            }
            if (split) gram_kstep<true>(cur, x); else gram_kstep<false>(cur, x);
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) { acc[q][0] += cur[q][0]; acc[q][1] += cur[q][1]; acc[q][2] += cur[q][2]; acc[q][3] += cur[q][3]; }
    }
    cp_async_wait<0>();
    // CTA reduction in double, one warp after the other (distinct lanes own distinct entries), then one atomic per entry
    for (int w = 0; w < MGGAN_THREADS / 32; ++w) {
        __syncthreads();
        if (warp == w) {
            int tile = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 2 * i; j < 5; ++j, ++tile) {
                    const int ra = 16 * i + g, cb = 8 * j + 2 * t;
                    sD[ra * G_NT + cb] += (double)acc[tile][0];
                    sD[ra * G_NT + cb + 1] += (double)acc[tile][1];
                    if (i < 2) {
                        sD[(ra + 8) * G_NT + cb] += (double)acc[tile][2];
                        sD[(ra + 8) * G_NT + cb + 1] += (double)acc[tile][3];
                    }
                }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NTAP * G_NT; i += MGGAN_THREADS) {
        const int a = i / G_NT, b = i - a * G_NT;
        const double v = sD[i];
        if (b == NTAP) atomicAdd(P + a, v);
        else if (b < NTAP && a <= b) {                 // upper triangle (the blocks below the diagonal were not all computed;
            atomicAdd(R + a * NTAP + b, v);            // mirroring keeps R exactly symmetric under the split products too)
            if (a < b) atomicAdd(R + b * NTAP + a, v);
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
// Both convolutions run on the tensor pipe as warp-level m16n8k8 TF32 products (common.cuh).  The FP32 form of conv1
// (thread = pooled pixel, 36 C FMAs per input pixel) made the kernel ISSUE-bound: ~35 k warp instructions per crop, 18 k
// of them FFMA; as an implicit GEMM it is ~1.3 k MMAs and as many LDS.
//   conv1   M = conv output positions, K = 36 taps (ci, ky, kx) padded to 40, N = C.  One PAIR of m-tiles covers 8 pooled
//           pixels of a row: tile 0 holds window row 0, tile 1 window row 1, fragment rows g / g + 8 are window columns
//           0 / 1 of pooled pixel px0 + g -- so lane (g, t) ends up with the whole 2 x 2 window of its pooled pixel for the
//           channels 8 q + 2 t + {0, 1} and BN1 / ReLU / max-pool / arg are thread-local.  The A operand is read straight
//           from the zero-padded crop in shared memory (im2col by address: tap offset + position); the K index is permuted
//           (TAP_PERM) so that the 8 positions x 4 taps of every fragment load fall on 32 distinct banks with row stride 35
//           and channel stride 1231.  W1 stays in registers as pre-split (hi, lo) B fragments.
//           Crops cut from 8-bit images are exact TF32 numbers (see scene_patch_stats): then the products are exact with
//           two MMAs per tile (x W1_hi + x W1_lo); any other fp32 crop takes the three split products.
//   conv2   m-tile = one output row (16 pixels), n-tile = 8 output channels, k-step = 8 input channels of one tap, 3 x TF32;
//           the pooled map and W2 are kept PRE-SPLIT as (hi, lo) planes in shared memory, so a fragment is plain LDS.32s
//           (splitting on the fly cost 5 ALU instructions per MMA).
// The next crop is staged (cp.async into the padded layout) while conv2 runs.  Optionally saves the pre-BN value at the
// pool arg (e1) and the arg index | active bit (idx1) for the backward.
constexpr int F_LDR = 35;                        // padded crop row (33 + halo)
constexpr int F_CHS = 1231;                      // padded crop channel: 35 * 35 + 6 (bank spread of the permuted taps)
constexpr int F_CROP = CIN * F_CHS + 5;          // floats per staged crop, rounded up to a multiple of 4 below
constexpr int F_CROP4 = (F_CROP + 3) & ~3;
// K order of conv1: k-step s holds taps TAP_PERM[8 s .. 8 s + 7] (k = t -> entry t, k = t + 4 -> entry 4 + t); 36.. = zero taps
__constant__ int TAP_PERM[40] = {0, 1, 2, 11, 3, 4, 5, 14, 6, 7, 8, 17, 9, 10, 19, 20, 12, 13, 22, 23,
                                 15, 16, 25, 26, 18, 21, 28, 31, 24, 27, 29, 34, 30, 32, 33, 35, 36, 37, 38, 39};

// conv1 products of NTP tile pairs (8 pooled pixels each): acc[p][window row][n-tile][c0 c1 | c2 c3].  The MMAs are issued
// product-major over the 2 NTP NT independent accumulators (back-to-back MMAs on one accumulator wait for each other:
// "wait" was the top stall reason of the first tensor-pipe version).  SPLIT = false: the crop holds TF32 numbers.
template <int NT, int NTP, bool SPLIT>
__device__ __forceinline__ void conv1_mma(const float* const (&pos)[NTP], const int (&toff)[5][2], const float* __restrict__ sW1f,
                                          int lane, float (&acc)[NTP][2][NT][4]) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            const float* wf = sW1f + (s * NT + q) * 128 + lane;
            bh[q][0] = __float_as_uint(wf[0]); bl[q][0] = __float_as_uint(wf[32]);
            bh[q][1] = __float_as_uint(wf[64]); bl[q][1] = __float_as_uint(wf[96]);
        }
        uint32_t ah[NTP][2][4], al[NTP][2][4];
#pragma unroll
        for (int p = 0; p < NTP; ++p)
#pragma unroll
            for (int wy = 0; wy < 2; ++wy) {
                const float* pa = pos[p] + wy * F_LDR;
                float a[4];
                a[0] = pa[toff[s][0]]; a[1] = pa[toff[s][0] + 1];
                if (s < 4) { a[2] = pa[toff[s][1]]; a[3] = pa[toff[s][1] + 1]; } else { a[2] = 0.f; a[3] = 0.f; }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (SPLIT) tf32_split(a[i], ah[p][wy][i], al[p][wy][i]);
                    else { ah[p][wy][i] = __float_as_uint(a[i]); al[p][wy][i] = 0u; }
                }
            }
#pragma unroll
        for (int p = 0; p < NTP; ++p)
#pragma unroll
            for (int wy = 0; wy < 2; ++wy)
#pragma unroll
                for (int q = 0; q < NT; ++q) mma_tf32_16x8x8(acc[p][wy][q], ah[p][wy], bh[q][0], bh[q][1]);
#pragma unroll
        for (int p = 0; p < NTP; ++p)
#pragma unroll
            for (int wy = 0; wy < 2; ++wy)
#pragma unroll
                for (int q = 0; q < NT; ++q) mma_tf32_16x8x8(acc[p][wy][q], ah[p][wy], bl[q][0], bl[q][1]);
        if (SPLIT) {
#pragma unroll
            for (int p = 0; p < NTP; ++p)
#pragma unroll
                for (int wy = 0; wy < 2; ++wy)
#pragma unroll
                    for (int q = 0; q < NT; ++q) mma_tf32_16x8x8(acc[p][wy][q], al[p][wy], bh[q][0], bh[q][1]);
        }
    }
}

template <int C>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
scene_fused12_fwd_kernel(const float* __restrict__ img, const int* __restrict__ rows, int N,
                         const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ ab1,
                         const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ x2,
                         double* __restrict__ stats2, float* __restrict__ e1, unsigned char* __restrict__ idx1) {
    constexpr int PPAD = FwdPad::PPAD;           // channel stride of the pooled map: 8 mod 32 (conflict-free A fragments)
    constexpr int LDW2 = FwdLdw2<C>::value;        // row stride of the conv2 weights [tap][ci][co]: 8 t + g distinct banks
    constexpr int PBUF = (C * PPAD + 3) & ~3;      // one plane of the pooled map
    constexpr int NT = C / 8;                     // n-tiles (8 output channels each) = conv2 k-steps per tap
    extern __shared__ __align__(16) float smem[];
    float* sCrop = smem;                         // [4][35][35]+   zero-padded crop
    float* sPh = sCrop + F_CROP4;                // [C][PPAD]      pooled map, TF32 hi plane (zero halo)
    float* sPl = sPh + PBUF;                     //                lo plane
    float* sW2h = sPl + PBUF;                    // [9 taps][C in][LDW2]   (C out used), hi plane
    float* sW2l = sW2h + 9 * C * LDW2;           //                lo plane
    float* sAB = sW2l + 9 * C * LDW2;            // [2C]  BatchNorm-1 as an affine map
    float* sB1 = sAB + 2 * C;                    // [C]   conv1 bias
    float* sW1f = sB1 + C;                       // [5][NT][2][2][32] conv1 B fragments
    float* sred = sW1f + 5 * NT * 2 * 64;        // [8][2C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g8 = lane >> 2, t4 = lane & 3;
    for (int i = threadIdx.x; i < 9 * C * C; i += MGGAN_THREADS) {
        int co = i / (9 * C), r = i - co * 9 * C, ci = r / 9, tap = r - ci * 9;
        uint32_t hi, lo;
        tf32_split(__ldg(W2 + i), hi, lo);
        sW2h[(tap * C + ci) * LDW2 + co] = __uint_as_float(hi);
        sW2l[(tap * C + ci) * LDW2 + co] = __uint_as_float(lo);
    }
    if (threadIdx.x < 2 * C) sAB[threadIdx.x] = __ldg(ab1 + threadIdx.x);
    for (int i = threadIdx.x; i < 2 * PBUF; i += MGGAN_THREADS) sPh[i] = 0.f;
    for (int i = threadIdx.x; i < F_CROP4; i += MGGAN_THREADS) sCrop[i] = 0.f;
    // conv1 B fragments, pre-split and stored per lane ([s][q][h][hi | lo][lane]: conflict-free LDS.32; in registers they
    // pushed the C = 16 instance over 128): b0 = W1[8 q + g][tap(s, t)], b1 = W1[8 q + g][tap(s, 4 + t)]
    for (int i = threadIdx.x; i < 5 * NT * 2 * 32; i += MGGAN_THREADS) {
        const int ln = i & 31, h = (i >> 5) & 1, q = (i >> 6) % NT, sidx = i / (64 * NT);
        const int tap = TAP_PERM[8 * sidx + 4 * h + (ln & 3)];
        const float w = tap < NTAP ? __ldg(W1 + (8 * q + (ln >> 2)) * NTAP + tap) : 0.f;
        uint32_t hi, lo;
        tf32_split(w, hi, lo);
        sW1f[((sidx * NT + q) * 2 + h) * 64 + ln] = __uint_as_float(hi);
        sW1f[((sidx * NT + q) * 2 + h) * 64 + 32 + ln] = __uint_as_float(lo);
    }
    int toff[5][2];                               // tap offsets in the padded crop (zero taps: offset 0, weight 0)
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tap = TAP_PERM[8 * s + 4 * h + t4];
            const int ci = tap / 9, r = tap - ci * 9;
            toff[s][h] = tap < NTAP ? ci * F_CHS + (r / 3) * F_LDR + r % 3 : 0;
        }
    if (threadIdx.x < C) sB1[threadIdx.x] = __ldg(b1 + threadIdx.x);
    __syncthreads();

    // warp w copies rows w, w + 8, ... of the crop (row = (channel, y)): lanes = consecutive x, lane 0 also takes x = 32
    auto stage = [&](int n) {
        const int src = rows ? rows[n] : n;
        const float* ip = img + (size_t)src * RAW;
        for (int r = warp; r < CIN * IMG; r += MGGAN_THREADS / 32) {
            const int ci = r / IMG, y = r - ci * IMG;
            float* d = sCrop + ci * F_CHS + (y + 1) * F_LDR + 1;
            cp_async4(d + lane, ip + r * IMG + lane);
            if (lane == 0) cp_async4(d + 32, ip + r * IMG + 32);
        }
        cp_async_commit();
    };
    if ((int)blockIdx.x < N) stage(blockIdx.x);
    float st[4 * NT];                             // BatchNorm-2 partial sums: channels 8 j + 2 t + {0, 1}: sum [2j + e], squares [2NT + 2j + e]
#pragma unroll
    for (int c = 0; c < 4 * NT; ++c) st[c] = 0.f;

    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        cp_async_wait<0>();
        __syncthreads();                          // the crop has landed; the previous agent's conv2 has read the pooled map
        uint32_t low = 0u;                        // any low mantissa bit set -> the crop is not made of TF32 numbers
        for (int i = threadIdx.x; i < F_CROP4 / 4; i += MGGAN_THREADS) {
            const float4 v = ld4(sCrop + 4 * i);
            low |= __float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w);
        }
        const bool split = __syncthreads_or((low & 0x1FFFu) != 0u);
        // ---- conv1 -> BN1 -> ReLU -> pool: tile pairs (pooled row py, half row) warp, warp + 8, ..., NTP at a time
        constexpr int NTP = NT == 1 ? 2 : 1;      // 4 independent accumulators per k-step either way
#pragma unroll 1
        for (int tp0 = warp; tp0 < 32; tp0 += NTP * (MGGAN_THREADS / 32)) {
            const float* pos[NTP];                // window (0, 0) of pooled pixel (py, px0 + g) of tile pair p
            float acc[NTP][2][NT][4];             // [pair][window row][n-tile][col 0: ch e | col 1: ch e]
#pragma unroll
            for (int p = 0; p < NTP; ++p) {
                const int tp = tp0 + p * (MGGAN_THREADS / 32);
                pos[p] = sCrop + (2 * (tp >> 1)) * F_LDR + 2 * ((tp & 1) * 8 + g8);
#pragma unroll
                for (int wy = 0; wy < 2; ++wy)
#pragma unroll
                    for (int q = 0; q < NT; ++q) {
                        const float2 bv = *reinterpret_cast<const float2*>(sB1 + 8 * q + 2 * t4);
                        acc[p][wy][q][0] = bv.x; acc[p][wy][q][1] = bv.y; acc[p][wy][q][2] = bv.x; acc[p][wy][q][3] = bv.y;
                    }
            }
            if (split) conv1_mma<NT, NTP, true>(pos, toff, sW1f, lane, acc);
            else conv1_mma<NT, NTP, false>(pos, toff, sW1f, lane, acc);
#pragma unroll
            for (int p = 0; p < NTP; ++p) {
                const int tp = tp0 + p * (MGGAN_THREADS / 32);
                const int py = tp >> 1, px0 = (tp & 1) * 8;
                const int pix = py * P1 + px0 + g8;
#pragma unroll
                for (int q = 0; q < NT; ++q)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * q + 2 * t4 + e;
                        const float a = sAB[c], b = sAB[C + c];
                        // window order (0,0) (0,1) (1,0) (1,1), first maximum wins (as the FP32 kernel and the oracle's max_pool2d)
                        const float x0 = acc[p][0][q][e], x1 = acc[p][0][q][2 + e], x2v = acc[p][1][q][e], x3 = acc[p][1][q][2 + e];
                        float m = fmaf(a, x0, b), ev = x0; int arg = 0;
                        float v = fmaf(a, x1, b);
                        if (v > m) { m = v; ev = x1; arg = 1; }
                        v = fmaf(a, x2v, b);
                        if (v > m) { m = v; ev = x2v; arg = 2; }
                        v = fmaf(a, x3, b);
                        if (v > m) { m = v; ev = x3; arg = 3; }
                        uint32_t hi, lo;
                        tf32_split(fmaxf(m, 0.f), hi, lo);
                        const int o = c * PPAD + (py + 1) * LDP + px0 + g8 + 1;
                        sPh[o] = __uint_as_float(hi);
                        sPl[o] = __uint_as_float(lo);
                        if (e1 != nullptr) {
                            e1[((size_t)n * C + c) * P1SQ + pix] = ev;
                            idx1[((size_t)n * C + c) * P1SQ + pix] = (unsigned char)(arg | (m > 0.f ? 4 : 0));
                        }
                    }
            }
        }
        __syncthreads();                          // pooled map complete; the crop buffer is free
        if (n + (int)gridDim.x < N) stage(n + gridDim.x);
        {   // ---- conv2: the warp's two rows, MMAs issued product-major over 4 independent accumulators (C = 8: even / odd
            // taps accumulate separately and are added at the end)
            constexpr int NSET = NT == 1 ? 2 : 1;
            const int y = warp * 2;
            float acc[NSET][2][NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const float ba = __ldg(b2 + 8 * j + 2 * t4), bb = __ldg(b2 + 8 * j + 2 * t4 + 1);
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    acc[0][rr][j][0] = ba; acc[0][rr][j][1] = bb; acc[0][rr][j][2] = ba; acc[0][rr][j][3] = bb;
                    if (NSET == 2) { acc[NSET - 1][rr][j][0] = 0.f; acc[NSET - 1][rr][j][1] = 0.f; acc[NSET - 1][rr][j][2] = 0.f; acc[NSET - 1][rr][j][3] = 0.f; }
                }
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) {
                    const int ob = (tap * C + ks * 8 + t4) * LDW2 + g8;
                    uint32_t bh0[NT], bh1[NT], bl0[NT], bl1[NT], ah[2][4], al[2][4];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        bh0[j] = __float_as_uint(sW2h[ob + 8 * j]); bh1[j] = __float_as_uint(sW2h[ob + 4 * LDW2 + 8 * j]);
                        bl0[j] = __float_as_uint(sW2l[ob + 8 * j]); bl1[j] = __float_as_uint(sW2l[ob + 4 * LDW2 + 8 * j]);
                    }
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int oa = (ks * 8 + t4) * PPAD + (y + rr + tap / 3) * LDP + g8 + tap % 3;
                        ah[rr][0] = __float_as_uint(sPh[oa]); ah[rr][1] = __float_as_uint(sPh[oa + 8]);
                        ah[rr][2] = __float_as_uint(sPh[oa + 4 * PPAD]); ah[rr][3] = __float_as_uint(sPh[oa + 4 * PPAD + 8]);
                        al[rr][0] = __float_as_uint(sPl[oa]); al[rr][1] = __float_as_uint(sPl[oa + 8]);
                        al[rr][2] = __float_as_uint(sPl[oa + 4 * PPAD]); al[rr][3] = __float_as_uint(sPl[oa + 4 * PPAD + 8]);
                    }
                    float (&ac)[2][NT][4] = acc[NSET == 2 ? (tap & 1) : 0];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_tf32_16x8x8(ac[rr][j], ah[rr], bh0[j], bh1[j]);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_tf32_16x8x8(ac[rr][j], al[rr], bh0[j], bh1[j]);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_tf32_16x8x8(ac[rr][j], ah[rr], bl0[j], bl1[j]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    float* o = x2 + ((size_t)n * C + 8 * j + 2 * t4) * P1SQ + (y + rr) * P1 + g8;
                    float a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = NSET == 2 ? acc[0][rr][j][i] + acc[NSET - 1][rr][j][i] : acc[0][rr][j][i];
                    o[0] = a[0]; o[8] = a[2]; o[P1SQ] = a[1]; o[P1SQ + 8] = a[3];
                    st[2 * j] += a[0] + a[2];
                    st[2 * j + 1] += a[1] + a[3];
                    st[2 * NT + 2 * j] = fmaf(a[0], a[0], fmaf(a[2], a[2], st[2 * NT + 2 * j]));
                    st[2 * NT + 2 * j + 1] = fmaf(a[1], a[1], fmaf(a[3], a[3], st[2 * NT + 2 * j + 1]));
                }
        }
    }
    cp_async_wait<0>();
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
    unsigned short* sOff = reinterpret_cast<unsigned short*>(sIdx + C * P1SQ);   // [C][256] crop offsets of the non-zero dy1 (compacted)
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
    float st[4 * NT];                                    // st: BatchNorm-1 sums of channels 8 j + 2 t + e: [2j + e], [2NT + 2j + e]
#pragma unroll
    for (int c = 0; c < 4 * NT; ++c) st[c] = 0.f;

    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        const int src = rows ? rows[n] : n;
        __syncthreads();
        {
            // the next agent's inputs are requested into the L2 now (this stage waits on global loads with 16 warps per SM:
            // long-scoreboard stalls were a quarter of the kernel's samples)
            if (threadIdx.x < 6 && n + (int)gridDim.x < N) {
                const size_t nn = (size_t)n + gridDim.x;
                if (threadIdx.x == 0) prefetch_l2_bulk(x2 + nn * C * P1SQ, C * P1SQ * 4);
                else if (threadIdx.x == 1) prefetch_l2_bulk(e1 + nn * C * P1SQ, C * P1SQ * 4);
                else if (threadIdx.x == 2) prefetch_l2_bulk(idx1 + nn * C * P1SQ, C * P1SQ);
                else if (threadIdx.x == 3) prefetch_l2_bulk(dy2 + nn * C * P2SQ, C * P2SQ * 4);
                else if (threadIdx.x == 4) prefetch_l2_bulk(idx2 + nn * C * P2SQ, C * P2SQ);
                else prefetch_l2_bulk(img + (size_t)(rows ? rows[nn] : (int)nn) * RAW, RAW_BYTES);
            }
            // the crop is only read by the last stage (sparse conv1 weight gradient): cp.async (LDGSTS) it into the padded
            // layout now and wait for it there, behind the two tensor-core stages.  Warp = rows (channel, y), lanes = x.
            const float* ip = img + (size_t)src * RAW;
            for (int r = warp; r < CIN * IMG; r += MGGAN_THREADS / 32) {
                const int ci = r / IMG, yy = r - ci * IMG;
                float* d = sImg + ci * IMGPAD_BWD + (yy + 1) * LDI + 1;
                cp_async4(d + (threadIdx.x & 31), ip + r * IMG + (threadIdx.x & 31));
                if ((threadIdx.x & 31) == 0) cp_async4(d + 32, ip + r * IMG + 32);
            }
            cp_async_commit();
            const int win = (y >> 1) * P2 + (x >> 1), loc = (y & 1) * 2 + (x & 1);
#pragma unroll
            for (int c0 = 0; c0 < C; c0 += 8) {       // 8 channels per batch: independent loads first, one latency per batch
                float xv[8], dv[8];
                unsigned char iv[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = c0 + cc;
                    xv[cc] = __ldg(x2 + ((size_t)n * C + c) * P1SQ + threadIdx.x);
                    iv[cc] = idx2[((size_t)n * C + c) * P2SQ + win];
                    dv[cc] = __ldg(dy2 + ((size_t)n * C + c) * P2SQ + win);
                }
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = c0 + cc;
                    const float xh = (xv[cc] - sPar[6 * C + c]) * sPar[7 * C + c];
                    const float d = iv[cc] == loc ? dv[cc] : 0.f;
                    const float dx = sPar[4 * C + c] * (d - sPar[8 * C + c] - xh * sPar[9 * C + c]);
                    sDX[c * PPAD + (y + 1) * LDP + x + 1] = dx;
                }
            }
            // e1 / idx1: 4 consecutive pooled pixels of one channel per thread and trip (float4 / 32-bit loads)
            const float4* e4 = reinterpret_cast<const float4*>(e1 + (size_t)n * C * P1SQ);
            const uint32_t* i4 = reinterpret_cast<const uint32_t*>(idx1 + (size_t)n * C * P1SQ);
            float4 ev[C / 4];
            uint32_t iw[C / 4];
#pragma unroll
            for (int k = 0; k < C / 4; ++k) {
                ev[k] = __ldg(e4 + threadIdx.x + k * MGGAN_THREADS);
                iw[k] = __ldg(i4 + threadIdx.x + k * MGGAN_THREADS);
            }
#pragma unroll
            for (int k = 0; k < C / 4; ++k) {
                const int q = threadIdx.x + k * MGGAN_THREADS, c = q >> 6, pp = (q & 63) * 4;
                reinterpret_cast<uint32_t*>(sIdx)[q] = iw[k];
                float* o = sP + c * PPAD + ((pp >> 4) + 1) * LDP + (pp & 15) + 1;
                const float a = sPar[c], b = sPar[C + c];
                o[0] = fmaxf(fmaf(a, ev[k].x, b), 0.f); o[1] = fmaxf(fmaf(a, ev[k].y, b), 0.f);
                o[2] = fmaxf(fmaf(a, ev[k].z, b), 0.f); o[3] = fmaxf(fmaf(a, ev[k].w, b), 0.f);
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
                // B fragments of the warp's units first, then the MMAs product-major over the units (three back-to-back
                // MMAs on one accumulator wait for each other)
                uint32_t bh[UPW][2], bl[UPW][2];
#pragma unroll
                for (int i = 0; i < UPW; ++i) {
                    const int unit = warp + 8 * i;
                    if (unit < UNITS) {
                        const int tap = unit / NT, j = unit - tap * NT;
                        const float* pb = sP + (8 * j + g8) * PPAD + (yy + tap / 3) * LDP + x0 + t4 + tap % 3;
                        tf32_split(pb[0], bh[i][0], bl[i][0]);
                        tf32_split(pb[4], bh[i][1], bl[i][1]);
                    } else {                       // unit UNITS: an all-ones B -> every column of the C fragment is sum_pix dx2 = dbias2
                        bh[i][0] = bh[i][1] = 0x3F800000u;
                        bl[i][0] = bl[i][1] = 0u;
                    }
                }
#pragma unroll
                for (int i = 0; i < UPW; ++i) if (warp + 8 * i <= UNITS) mma_tf32_16x8x8(wacc[i], ah, bh[i][0], bh[i][1]);
#pragma unroll
                for (int i = 0; i < UPW; ++i) if (warp + 8 * i <= UNITS) mma_tf32_16x8x8(wacc[i], al, bh[i][0], bh[i][1]);
#pragma unroll
                for (int i = 0; i < UPW; ++i) if (warp + 8 * i < UNITS) mma_tf32_16x8x8(wacc[i], ah, bl[i][0], bl[i][1]);
            }
        }
        {   // conv2 input gradient as warp-level 3 x TF32 products: m-tile = one row of 16 pixels, N = 8 input channels,
            // K = output channels of one tap (logical k = t, t + 4 <-> co = c0 + 2t, c0 + 2t + 1), then pool / ReLU / BN1 -> dy1.
            // The warp's two rows run together and the MMAs are issued product-major over independent accumulators (three
            // back-to-back MMAs on one accumulator wait for each other); C = 8: even / odd taps accumulate separately.
            constexpr int NSET = NT == 1 ? 2 : 1;
            const int y0 = warp * 2;
            float acc[NSET][2][NT][4];
#pragma unroll
            for (int q = 0; q < NSET; ++q)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int j = 0; j < NT; ++j) { acc[q][rr][j][0] = 0.f; acc[q][rr][j][1] = 0.f; acc[q][rr][j][2] = 0.f; acc[q][rr][j][3] = 0.f; }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) {
                    uint32_t ah[2][4], al[2][4], bh[NT][2], bl[NT][2];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const float* pa = sDX + (ks * 8 + 2 * t4) * PPAD + (y0 + rr + 2 - tap / 3) * LDP + g8 + 2 - tap % 3;
                        tf32_split(pa[0], ah[rr][0], al[rr][0]); tf32_split(pa[8], ah[rr][1], al[rr][1]);
                        tf32_split(pa[PPAD], ah[rr][2], al[rr][2]); tf32_split(pa[PPAD + 8], ah[rr][3], al[rr][3]);
                    }
                    const float* pb = sWT + (tap * C + ks * 8 + 2 * t4) * LDWD + g8;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        tf32_split(pb[8 * j], bh[j][0], bl[j][0]);
                        tf32_split(pb[LDWD + 8 * j], bh[j][1], bl[j][1]);
                    }
                    float (&ac)[2][NT][4] = acc[NSET == 2 ? (tap & 1) : 0];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_tf32_16x8x8(ac[rr][j], ah[rr], bh[j][0], bh[j][1]);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_tf32_16x8x8(ac[rr][j], al[rr], bh[j][0], bh[j][1]);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_tf32_16x8x8(ac[rr][j], ah[rr], bl[j][0], bl[j][1]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int yy = y0 + rr;
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
                            const float av = NSET == 2 ? acc[0][rr][j][2 * hx + e] + acc[NSET - 1][rr][j][2 * hx + e] : acc[0][rr][j][2 * hx + e];
                            const float d = (sIdx[ch * P1SQ + pix] & 4) ? av : 0.f;
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
            // inactive (about half of the pooled pixels), so every warp first compacts the non-zero entries of the channels
            // it owns (C = 16: channels 2 warp, 2 warp + 1; C = 8: channel warp) with warp ballots -- the value in place, the
            // offset of its window corner in the padded crop (from the pooled pixel and the pool arg) as 16 bits -- and the
            // lane groups then walk the list: 2 loads + 9 (load, FMA) per entry, every lane of every trip does useful work
            constexpr int CPW = C / 8;                       // channels per warp
            const int lane = threadIdx.x & 31;
            int nnz = 0;
#pragma unroll
            for (int cc = 0; cc < CPW; ++cc) {
                const int ch = warp * CPW + cc;
                int count = 0;
#pragma unroll 2
                for (int base = 0; base < P1SQ; base += 32) {
                    const int q = base + lane;
                    const float v = sDY[ch * LDY + q];
                    const int code = sIdx[ch * P1SQ + q] & 3;
                    const bool on = v != 0.f;
                    const unsigned m = __ballot_sync(0xffffffffu, on);
                    __syncwarp();                            // every lane's read of this trip is ordered before the in-place writes
                    if (on) {                                // slot <= q: the compaction never overtakes its own reads
                        const int slot = count + __popc(m & ((1u << lane) - 1u));
                        sDY[ch * LDY + slot] = v;
                        sOff[ch * P1SQ + slot] = (unsigned short)((2 * (q >> 4) + (code >> 1)) * LDI + 2 * (q & 15) + (code & 1));
                    }
                    count += __popc(m);
                }
                if (ch == s_c) nnz = count;
            }
            __syncwarp();
            const float* ipc = sImg + s_ci * IMGPAD_BWD;
#pragma unroll 2
            for (int e = s_qg; e < nnz; e += QG) {
                const float d = sDY[s_c * LDY + e];
                const float* bp = ipc + sOff[s_c * P1SQ + e];
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
        if (unit == UNITS && t4 == 0) {          // the all-ones unit: column 0 of its C fragment
            atomicAdd(dbias2 + g8, wacc[i][0]);
            if (C == 16) atomicAdd(dbias2 + g8 + 8, wacc[i][2]);
        }
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
template <int C, bool WITH_ARG, bool SHARED = false>
__device__ __forceinline__ void pool_block2(const float* __restrict__ x2n, const float* sAB, int pos, float (&v)[C],
                                            int (&arg)[C], float (&e)[C]) {
    const int py = pos >> 3, px = pos & 7;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float a = sAB[c], b = sAB[C + c];
        const float2* s = reinterpret_cast<const float2*>(x2n + c * P1SQ + (2 * py) * P1 + 2 * px);
        float2 t0, t1;
        if (SHARED) { t0 = s[0]; t1 = s[P1 / 2]; } else { t0 = __ldg(s); t1 = __ldg(s + P1 / 2); }
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
// backward of pass C: d(out) -> attention-MLP weight grads, sparse dy2 (+ arg index), BN2 sums.
// One trip = 2 agents = 128 positions.  The per-position MLP (C -> 32 -> C) and its gradients are warp-level 3 x TF32
// tensor-pipe products over the 128 rows (warp w = rows 16 w .. 16 w + 15), with the operands handed over through shared
// memory; pooling, the channel softmax and the sparse output stay thread-per-position (threads 0 .. 127).  The FP32 form
// (thread = position, the whole MLP and its 4 x 4 weight-gradient blocks in registers) needed 255 registers and 114 KB:
// one CTA of 8 warps per SM, every latency exposed (12 % occupancy, 0.36 ms per launch of 16,384 crops).
//   S0  pool -> V                                S1  Hid = lrelu(V Wa1^T + ba1),  S = Hid Wa2^T + ba2
//   S2  softmax, dS (in place of S)              S3  dHid = (dS Wa2) lrelu'(Hid)
//   S4  dWa2 += dS^T Hid, dWa1 += dHid^T V, bias gradients from all-ones B tiles (C fragments kept across trips)
//   S5  dV = dHid Wa1 (in place of dS)           S6  dy2 / idx2 / BatchNorm-2 sums
constexpr int AT_ROWS = 128;
constexpr int AT_LDH = AH + 4;                    // 36: row stride of Hid / dHid and of the [.][32] weight planes
template <int C> struct AtLd { static constexpr int V = C + 4; };     // row stride of V / dS and of the [.][C] weight planes

// acc (16 rows x 8 NTL columns) += A (rows m0 .. m0 + 15 of a row-major fp32 tile, split on the fly) . B^T with B given as
// pre-split planes stored [n][k] (row stride ldb): B fragment (k = t, n = g) = B[(8 j + g) ldb + k0 + t]
template <int KSTEPS, int NTL>
__device__ __forceinline__ void rows_mma(float (&acc)[NTL][4], const float* __restrict__ A, int lda, const float* __restrict__ Bh,
                                         const float* __restrict__ Bl, int ldb, int g8, int t4) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        const float* pa = A + g8 * lda + 8 * ks + t4;
        uint32_t ah[4], al[4], bh[NTL][2], bl[NTL][2];
        tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8 * lda], ah[1], al[1]);
        tf32_split(pa[4], ah[2], al[2]); tf32_split(pa[8 * lda + 4], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
            const int o = (8 * j + g8) * ldb + 8 * ks + t4;
            bh[j][0] = __float_as_uint(Bh[o]); bh[j][1] = __float_as_uint(Bh[o + 4]);
            bl[j][0] = __float_as_uint(Bl[o]); bl[j][1] = __float_as_uint(Bl[o + 4]);
        }
#pragma unroll
        for (int j = 0; j < NTL; ++j) mma_tf32_16x8x8(acc[j], ah, bh[j][0], bh[j][1]);
#pragma unroll
        for (int j = 0; j < NTL; ++j) mma_tf32_16x8x8(acc[j], al, bh[j][0], bh[j][1]);
#pragma unroll
        for (int j = 0; j < NTL; ++j) mma_tf32_16x8x8(acc[j], ah, bl[j][0], bl[j][1]);
    }
}

// acc (16 x 8) += A^T B over the AT_ROWS rows of a trip: A (m, k = row) = X[row ldx + m0 + m] (rows m >= MROWS are zero),
// B (k = row, n) = Y[row ldy + n0 + n]; ONES: B = 1 (column sums of X)
template <int MROWS, bool ONES>
__device__ __forceinline__ void cols_mma(float (&acc)[4], const float* __restrict__ X, int ldx, const float* __restrict__ Y,
                                         int ldy, int g8, int t4) {
#pragma unroll 4
    for (int k0 = 0; k0 < AT_ROWS; k0 += 8) {
        const float* pa = X + (k0 + t4) * ldx + g8;
        uint32_t ah[4], al[4], bh[2], bl[2];
        tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[4 * ldx], ah[2], al[2]);
        if (MROWS > 8) { tf32_split(pa[8], ah[1], al[1]); tf32_split(pa[4 * ldx + 8], ah[3], al[3]); }
        else { ah[1] = al[1] = ah[3] = al[3] = 0u; }
        if (ONES) { bh[0] = bh[1] = 0x3F800000u; bl[0] = bl[1] = 0u; }
        else {
            const float* pb = Y + (k0 + t4) * ldy + g8;
            tf32_split(pb[0], bh[0], bl[0]); tf32_split(pb[4 * ldy], bh[1], bl[1]);
        }
        mma_tf32_16x8x8(acc, ah, bh[0], bh[1]);
        mma_tf32_16x8x8(acc, al, bh[0], bh[1]);
        if (!ONES) mma_tf32_16x8x8(acc, ah, bl[0], bl[1]);
    }
}

template <int C>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
scene_attn_bwd_kernel(const float* __restrict__ x2, int N, const float* __restrict__ ab2,
                      const float* __restrict__ mean_istd2, const float* __restrict__ Wa1,
                      const float* __restrict__ ba1, const float* __restrict__ Wa2, const float* __restrict__ ba2,
                      const float* __restrict__ dout, float* __restrict__ dWa1, float* __restrict__ dba1,
                      float* __restrict__ dWa2, float* __restrict__ dba2, float* __restrict__ dy2,
                      unsigned char* __restrict__ idx2, double* __restrict__ sums2) {
    constexpr int LDV = AtLd<C>::V, LDH = AT_LDH, NTC = C / 8;
    extern __shared__ __align__(16) float smem[];
    // pre-split weight planes, each stored [n][k] for the product that uses it as B
    float* sW1h = smem;                            // [AH][LDV]   Wa1   (Hid = V Wa1^T)
    float* sW1l = sW1h + AH * LDV;
    float* sW2h = sW1l + AH * LDV;                 // [C][LDH]    Wa2   (S = Hid Wa2^T)
    float* sW2l = sW2h + C * LDH;
    float* sW2Th = sW2l + C * LDH;                 // [AH][LDV]   Wa2^T (dHid = dS Wa2)
    float* sW2Tl = sW2Th + AH * LDV;
    float* sW1Th = sW2Tl + AH * LDV;               // [C][LDH]    Wa1^T (dV = dHid Wa1)
    float* sW1Tl = sW1Th + C * LDH;
    float* sb1 = sW1Tl + C * LDH;                  // [AH]
    float* sb2 = sb1 + AH;                         // [C]
    float* sAB = sb2 + C;                          // [2C] BatchNorm-2 as an affine map
    float* sMI = sAB + 2 * C;                      // [2C] mean, 1 / std
    float* sV = sMI + 2 * C;                       // [AT_ROWS][LDV]
    float* sDS = sV + AT_ROWS * LDV;               // [AT_ROWS][LDV]  S -> dS -> dV
    float* sHid = sDS + AT_ROWS * LDV;             // [AT_ROWS][LDH]
    float* sDH = sHid + AT_ROWS * LDH;             // [AT_ROWS][LDH]
    float* sred = sDH + AT_ROWS * LDH;             // [8][2C]
    float* sX2 = sred + 8 * 2 * C;                 // [2][C][256] the trip's conv2 outputs, bulk-copied one trip ahead
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sX2 + 2 * C * P1SQ);
    if (threadIdx.x == 0) {
        sbar_init(smem_addr(sBar), 1);
        sbar_fence_init();
    }
    for (int i = threadIdx.x; i < AH * C; i += MGGAN_THREADS) {
        const int u = i / C, c = i - u * C;        // Wa1[u][c]
        uint32_t hi, lo;
        tf32_split(__ldg(Wa1 + i), hi, lo);
        sW1h[u * LDV + c] = __uint_as_float(hi); sW1l[u * LDV + c] = __uint_as_float(lo);
        sW1Th[c * LDH + u] = __uint_as_float(hi); sW1Tl[c * LDH + u] = __uint_as_float(lo);
        const int c2 = i / AH, u2 = i - c2 * AH;   // Wa2[c2][u2]
        tf32_split(__ldg(Wa2 + i), hi, lo);
        sW2h[c2 * LDH + u2] = __uint_as_float(hi); sW2l[c2 * LDH + u2] = __uint_as_float(lo);
        sW2Th[u2 * LDV + c2] = __uint_as_float(hi); sW2Tl[u2 * LDV + c2] = __uint_as_float(lo);
    }
    if (threadIdx.x < AH) sb1[threadIdx.x] = __ldg(ba1 + threadIdx.x);
    if (threadIdx.x < C) sb2[threadIdx.x] = __ldg(ba2 + threadIdx.x);
    if (threadIdx.x < 2 * C) { sAB[threadIdx.x] = __ldg(ab2 + threadIdx.x); sMI[threadIdx.x] = __ldg(mean_istd2 + threadIdx.x); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g8 = lane >> 2, t4 = lane & 3;
    const int m0 = warp * 16;
    const int slot = threadIdx.x >> 6, pos = threadIdx.x & 63;       // threads 0 .. 127: (agent of the trip, position)
    // weight-gradient tiles of the warp (S4): warps 0-3: dWa2 n-tile `warp` (+ warp 0: dba2); warps 4 ..: dWa1 tile
    // (m-tile (warp - 4) / NTC, n-tile (warp - 4) % NTC) (+ n-tile 0: dba1 of that m-tile)
    const int wt = warp - 4, wm = wt / NTC, wn = wt - wm * NTC;
    const bool has_w1 = warp >= 4 && wt < 2 * NTC;
    float wacc[4] = {0.f, 0.f, 0.f, 0.f}, bacc[4] = {0.f, 0.f, 0.f, 0.f};
    float st[2 * C];
#pragma unroll
    for (int c = 0; c < 2 * C; ++c) st[c] = 0.f;
    __syncthreads();
    // one thread arms the barrier and starts the bulk copy of a trip's conv2 outputs (its 1 or 2 agents are contiguous)
    auto stage_x2 = [&](int n0) {
        const uint32_t bytes = (uint32_t)min(2, N - n0) * C * P1SQ * 4;
        sbar_expect_tx(smem_addr(sBar), bytes);
        bulk_g2s(smem_addr(sX2), x2 + (size_t)n0 * C * P1SQ, bytes, smem_addr(sBar));
    };
    if (threadIdx.x == 0 && (int)blockIdx.x * 2 < N) stage_x2(blockIdx.x * 2);
    uint32_t phase = 0;

    for (int n0 = blockIdx.x * 2; n0 < N; n0 += gridDim.x * 2, phase ^= 1u) {
        const int n = n0 + slot;
        const bool live = threadIdx.x < AT_ROWS && n < N;
        float e[C], dvd[C], gout = 0.f;
        int arg[C];
        // ---- S0: pooled values of the position (from the staged copy)
        if (threadIdx.x < AT_ROWS) {
            float v[C];
            if (live) gout = __ldg(dout + (size_t)n * P2SQ + pos);
            sbar_wait(smem_addr(sBar), phase);
            if (live) {
                pool_block2<C, true, true>(sX2 + slot * C * P1SQ, sAB, pos, v, arg, e);
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) { v[c] = 0.f; e[c] = 0.f; arg[c] = 0; }
            }
#pragma unroll
            for (int c = 0; c < C; c += 4) st4(sV + threadIdx.x * LDV + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
        }
        __syncthreads();
        // the staged block has been consumed: the next trip's copy runs behind the six stages below
        if (threadIdx.x == 0 && n0 + (int)gridDim.x * 2 < N) stage_x2(n0 + gridDim.x * 2);
        // ---- S1: Hid = lrelu(V Wa1^T + ba1) -> S = Hid Wa2^T + ba2 (the warp's own 16 rows: warp-level hand-over)
        {
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float b0 = sb1[8 * j + 2 * t4], b1 = sb1[8 * j + 2 * t4 + 1];
                acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
            }
            rows_mma<NTC, 4>(acc, sV + m0 * LDV, LDV, sW1h, sW1l, LDV, g8, t4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float* o = sHid + (m0 + g8) * LDH + 8 * j + 2 * t4;
                *reinterpret_cast<float2*>(o) = make_float2(lrelu_(acc[j][0], 0.01f), lrelu_(acc[j][1], 0.01f));
                *reinterpret_cast<float2*>(o + 8 * LDH) = make_float2(lrelu_(acc[j][2], 0.01f), lrelu_(acc[j][3], 0.01f));
            }
            __syncwarp();
            float sc[NTC][4];
#pragma unroll
            for (int j = 0; j < NTC; ++j) {
                const float b0 = sb2[8 * j + 2 * t4], b1 = sb2[8 * j + 2 * t4 + 1];
                sc[j][0] = b0; sc[j][1] = b1; sc[j][2] = b0; sc[j][3] = b1;
            }
            rows_mma<4, NTC>(sc, sHid + m0 * LDH, LDH, sW2h, sW2l, LDH, g8, t4);
#pragma unroll
            for (int j = 0; j < NTC; ++j) {
                float* o = sDS + (m0 + g8) * LDV + 8 * j + 2 * t4;
                *reinterpret_cast<float2*>(o) = make_float2(sc[j][0], sc[j][1]);
                *reinterpret_cast<float2*>(o + 8 * LDV) = make_float2(sc[j][2], sc[j][3]);
            }
        }
        __syncthreads();
        // ---- S2: channel softmax and its backward: dS in place of S; the direct part of dV stays in registers
        if (threadIdx.x < AT_ROWS) {
            float att[C], v[C];
            float mx = -3.4e38f;
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 q = ld4(sDS + threadIdx.x * LDV + c), w = ld4(sV + threadIdx.x * LDV + c);
                att[c] = q.x; att[c + 1] = q.y; att[c + 2] = q.z; att[c + 3] = q.w;
                v[c] = w.x; v[c + 1] = w.y; v[c + 2] = w.z; v[c + 3] = w.w;
            }
#pragma unroll
            for (int c = 0; c < C; ++c) mx = fmaxf(mx, att[c]);
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) { att[c] = __expf(att[c] - mx); sum += att[c]; }
            const float inv = 1.f / sum;
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) { att[c] *= inv; dot = fmaf(att[c], gout * v[c], dot); }
            float ds[C];
#pragma unroll
            for (int c = 0; c < C; ++c) { ds[c] = live ? att[c] * (gout * v[c] - dot) : 0.f; dvd[c] = gout * att[c]; }
#pragma unroll
            for (int c = 0; c < C; c += 4) st4(sDS + threadIdx.x * LDV + c, make_float4(ds[c], ds[c + 1], ds[c + 2], ds[c + 3]));
        }
        __syncthreads();
        // ---- S3: dHid = (dS Wa2) lrelu'(Hid)
        {
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
            rows_mma<NTC, 4>(acc, sDS + m0 * LDV, LDV, sW2Th, sW2Tl, LDV, g8, t4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = (m0 + g8) * LDH + 8 * j + 2 * t4;
                const float2 ha = *reinterpret_cast<const float2*>(sHid + o), hb = *reinterpret_cast<const float2*>(sHid + o + 8 * LDH);
                *reinterpret_cast<float2*>(sDH + o) = make_float2(acc[j][0] * (ha.x > 0.f ? 1.f : 0.01f), acc[j][1] * (ha.y > 0.f ? 1.f : 0.01f));
                *reinterpret_cast<float2*>(sDH + o + 8 * LDH) = make_float2(acc[j][2] * (hb.x > 0.f ? 1.f : 0.01f), acc[j][3] * (hb.y > 0.f ? 1.f : 0.01f));
            }
        }
        __syncthreads();
        // ---- S4: weight gradients over the trip's rows
        if (warp < 4) {
            cols_mma<C, false>(wacc, sDS, LDV, sHid + 8 * warp, LDH, g8, t4);              // dWa2[c][8 warp + .]
            if (warp == 0) cols_mma<C, true>(bacc, sDS, LDV, nullptr, 0, g8, t4);           // dba2
        } else if (has_w1) {
            cols_mma<16, false>(wacc, sDH + 16 * wm, LDH, sV + 8 * wn, LDV, g8, t4);        // dWa1[16 wm + .][8 wn + .]
            if (wn == 0) cols_mma<16, true>(bacc, sDH + 16 * wm, LDH, nullptr, 0, g8, t4);  // dba1[16 wm + .]
        }
        __syncthreads();
        // ---- S5: dV (through the MLP) = dHid Wa1, in place of dS
        {
            float acc[NTC][4];
#pragma unroll
            for (int j = 0; j < NTC; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
            rows_mma<4, NTC>(acc, sDH + m0 * LDH, LDH, sW1Th, sW1Tl, LDH, g8, t4);
#pragma unroll
            for (int j = 0; j < NTC; ++j) {
                float* o = sDS + (m0 + g8) * LDV + 8 * j + 2 * t4;
                *reinterpret_cast<float2*>(o) = make_float2(acc[j][0], acc[j][1]);
                *reinterpret_cast<float2*>(o + 8 * LDV) = make_float2(acc[j][2], acc[j][3]);
            }
        }
        __syncthreads();
        // ---- S6: sparse gradient at the pooled arg position + BatchNorm sums
        if (live) {
            float* dyo = dy2 + (size_t)n * C * P2SQ + pos;
            unsigned char* ixo = idx2 + (size_t)n * C * P2SQ + pos;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float dv = dvd[c] + sDS[threadIdx.x * LDV + c];
                const float d = (arg[c] & 4) ? dv : 0.f;
                dyo[c * P2SQ] = d;
                ixo[c * P2SQ] = (unsigned char)(arg[c] & 3);
                const float xh = (e[c] - sMI[c]) * sMI[C + c];
                st[c] += d;
                st[C + c] = fmaf(d, xh, st[C + c]);
            }
        }
    }
    // C fragments: c0, c1 -> row g, columns 2t, 2t + 1; c2, c3 -> row g + 8
    if (warp < 4) {
        float* dst = dWa2 + (size_t)g8 * AH + 8 * warp + 2 * t4;
        atomicAdd(dst, wacc[0]); atomicAdd(dst + 1, wacc[1]);
        if (C > 8) { atomicAdd(dst + 8 * AH, wacc[2]); atomicAdd(dst + 8 * AH + 1, wacc[3]); }
        if (warp == 0 && t4 == 0) {
            atomicAdd(dba2 + g8, bacc[0]);
            if (C > 8) atomicAdd(dba2 + g8 + 8, bacc[2]);
        }
    } else if (has_w1) {
        float* dst = dWa1 + (size_t)(16 * wm + g8) * C + 8 * wn + 2 * t4;
        atomicAdd(dst, wacc[0]); atomicAdd(dst + 1, wacc[1]);
        atomicAdd(dst + 8 * C, wacc[2]); atomicAdd(dst + 8 * C + 1, wacc[3]);
        if (wn == 0 && t4 == 0) { atomicAdd(dba1 + 16 * wm + g8, bacc[0]); atomicAdd(dba1 + 16 * wm + g8 + 8, bacc[2]); }
    }
    block_reduce_to_global<2 * C>(st, sums2, sred);
}


template <int C>
size_t fused_fwd_smem() { constexpr int PPAD = FwdPad::PPAD; return sizeof(float) * (F_CROP4 + 2 * ((C * PPAD + 3) & ~3) + 2 * 9 * C * FwdLdw2<C>::value + 3 * C + 5 * (C / 8) * 128 + 8 * 2 * C); }
template <int C>
size_t fused_bwd_smem() {
    constexpr int PPAD = BwdPad::PPAD;
    return sizeof(float) * (C * PPAD + ((C * PPAD + 3) & ~3) + 9 * C * BwdPad::LDWD + C * (P1SQ + 4) + CIN * IMGPAD_BWD + 10 * C + 8 * 2 * C) + 3 * C * P1SQ;
}
template <int C>
size_t attn_w_floats() { return 2 * AH * C + AH + C + 2 * C; }
template <int C>
size_t attn_fwd_smem() { return sizeof(float) * attn_w_floats<C>(); }
template <int C>
size_t attn_bwd_smem() {
    return sizeof(float) * (4 * AH * AtLd<C>::V + 4 * C * AT_LDH + AH + C + 4 * C + 2 * AT_ROWS * AtLd<C>::V + 2 * AT_ROWS * AT_LDH + 8 * 2 * C + 2 * C * P1SQ) + 16;
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
    int g = (N + 1) / 2;
    int cap = sm_count() * 2;
    scene_attn_bwd_kernel<C><<<g < cap ? g : cap, MGGAN_THREADS, sm, s>>>(x2, N, ab2, mi2, Wa1, ba1, Wa2, ba2, dout, dWa1,
                                                                          dba1, dWa2, dba2, dy2, idx2, sums2);
    return mggan_check_launch("scene_attn_bwd");
}

}  // namespace

extern "C" int mggan_scene_patch_stats(const float* img, const int* rows, int N, double* R, double* P,
                                       cudaStream_t stream) {
    if (N <= 0) return MGGAN_OK;
    const size_t sm = sizeof(float) * 2 * G_BUF + sizeof(double) * G_NT * G_NT;
    cudaFuncSetAttribute(scene_patch_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    scene_patch_stats_kernel<<<agent_grid(N, 2), MGGAN_THREADS, sm, stream>>>(img, rows, N, R, P);
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
