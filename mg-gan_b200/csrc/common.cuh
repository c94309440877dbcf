// Shared device helpers for the MG-GAN training-step kernels (sm_100a, fp32).
//
// Every kernel in this directory follows the same pattern: one CTA of 256 threads owns a tile
// of independent rows (agents, sampled sequences or in-scene pairs), stages the small weight
// matrices of the layer it runs in shared memory, and expresses each dense layer as a
// register-blocked tile product over shared memory (float4 loads, broadcast-friendly lane
// mapping: a warp is 4 row-lanes x 8 column-lanes so that one LDS.128 wavefront feeds 32 FMAs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MGGAN_THREADS 256

// ---- error plumbing (api.cu) -------------------------------------------------------------
extern "C" const char* mggan_last_error(void);
int mggan_set_error(int code, const char* fmt, ...);
int mggan_check_launch(const char* what);

#define MGGAN_OK 0
#define MGGAN_ERR_INVALID 1     // bad argument / unsupported shape
#define MGGAN_ERR_CUDA 2        // CUDA runtime error at launch
#define MGGAN_ERR_DEVICE 3      // not an sm_100 device

#define MGGAN_REQUIRE(cond, ...)                                    \
    do {                                                            \
        if (!(cond)) return mggan_set_error(MGGAN_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// ---- scalar math -------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// tanh via exp; |abs err| ~1e-7, saturates cleanly for large |x|.
__device__ __forceinline__ float tanhf_(float x) {
    float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}
__device__ __forceinline__ float lrelu_(float x, float a) { return x > 0.f ? x : a * x; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- shared-memory tile products ---------------------------------------------------------
// acc[i][j] += sum_k A[(r0 + i*RS)*lda + k] * W[(o0 + j*OS)*ldw + k],  k in [0, K), K % 4 == 0.
// lda / ldw are in floats, multiples of 4 (16-byte aligned rows).
template <int TR, int TO, int K>
__device__ __forceinline__ void tile_rowdot(float (&acc)[TR][TO], const float* __restrict__ A, int lda, int r0,
                                            int RS, const float* __restrict__ W, int ldw, int o0, int OS) {
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 a[TR], w[TO];
#pragma unroll
        for (int i = 0; i < TR; ++i) a[i] = ld4(A + (r0 + i * RS) * lda + k);
#pragma unroll
        for (int j = 0; j < TO; ++j) w[j] = ld4(W + (o0 + j * OS) * ldw + k);
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TO; ++j) {
                acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
            }
    }
}

// Input gradient of a dense layer: acc[i][c] += sum_o G[(r0+i*RS)*ldg + o] * W[o*ldw + k0 + c],
// o in [0, O), O % 4 == 0, c in 0..3 (the thread owns 4 consecutive input columns k0..k0+3).
template <int TR, int O>
__device__ __forceinline__ void tile_dgrad(float (&acc)[TR][4], const float* __restrict__ G, int ldg, int r0, int RS,
                                           const float* __restrict__ W, int ldw, int k0) {
#pragma unroll 2
    for (int o = 0; o < O; o += 4) {
        float4 g[TR], w[4];
#pragma unroll
        for (int i = 0; i < TR; ++i) g[i] = ld4(G + (r0 + i * RS) * ldg + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = ld4(W + (o + j) * ldw + k0);
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            acc[i][0] = fmaf(g[i].x, w[0].x, acc[i][0]); acc[i][1] = fmaf(g[i].x, w[0].y, acc[i][1]);
            acc[i][2] = fmaf(g[i].x, w[0].z, acc[i][2]); acc[i][3] = fmaf(g[i].x, w[0].w, acc[i][3]);
            acc[i][0] = fmaf(g[i].y, w[1].x, acc[i][0]); acc[i][1] = fmaf(g[i].y, w[1].y, acc[i][1]);
            acc[i][2] = fmaf(g[i].y, w[1].z, acc[i][2]); acc[i][3] = fmaf(g[i].y, w[1].w, acc[i][3]);
            acc[i][0] = fmaf(g[i].z, w[2].x, acc[i][0]); acc[i][1] = fmaf(g[i].z, w[2].y, acc[i][1]);
            acc[i][2] = fmaf(g[i].z, w[2].z, acc[i][2]); acc[i][3] = fmaf(g[i].z, w[2].w, acc[i][3]);
            acc[i][0] = fmaf(g[i].w, w[3].x, acc[i][0]); acc[i][1] = fmaf(g[i].w, w[3].y, acc[i][1]);
            acc[i][2] = fmaf(g[i].w, w[3].z, acc[i][2]); acc[i][3] = fmaf(g[i].w, w[3].w, acc[i][3]);
        }
    }
}

// Weight gradient: acc[a][b] += sum_r G[r*ldg + o0 + a] * X[r*ldx + k0 + b], r in [0, R).
// Rows that do not exist must hold zeros in G.
template <int R>
__device__ __forceinline__ void tile_wgrad(float (&acc)[4][4], const float* __restrict__ G, int ldg, int o0,
                                           const float* __restrict__ X, int ldx, int k0) {
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
        float4 g = ld4(G + r * ldg + o0);
        float4 x = ld4(X + r * ldx + k0);
        acc[0][0] = fmaf(g.x, x.x, acc[0][0]); acc[0][1] = fmaf(g.x, x.y, acc[0][1]);
        acc[0][2] = fmaf(g.x, x.z, acc[0][2]); acc[0][3] = fmaf(g.x, x.w, acc[0][3]);
        acc[1][0] = fmaf(g.y, x.x, acc[1][0]); acc[1][1] = fmaf(g.y, x.y, acc[1][1]);
        acc[1][2] = fmaf(g.y, x.z, acc[1][2]); acc[1][3] = fmaf(g.y, x.w, acc[1][3]);
        acc[2][0] = fmaf(g.z, x.x, acc[2][0]); acc[2][1] = fmaf(g.z, x.y, acc[2][1]);
        acc[2][2] = fmaf(g.z, x.z, acc[2][2]); acc[2][3] = fmaf(g.z, x.w, acc[2][3]);
        acc[3][0] = fmaf(g.w, x.x, acc[3][0]); acc[3][1] = fmaf(g.w, x.y, acc[3][1]);
        acc[3][2] = fmaf(g.w, x.z, acc[3][2]); acc[3][3] = fmaf(g.w, x.w, acc[3][3]);
    }
}

// Copy a dense (rows x cols) fp32 matrix from global into shared with a padded leading dimension.
__device__ __forceinline__ void stage_matrix(float* __restrict__ dst, int ldd, const float* __restrict__ src, int rows,
                                             int cols) {
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        int r = i / cols, c = i - r * cols;
        dst[r * ldd + c] = __ldg(src + i);
    }
}

// atomicAdd a 4x4 register block into a dense row-major matrix.
__device__ __forceinline__ void atomic_block44(float* __restrict__ dst, int ld, int o0, int k0, const float (&acc)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) atomicAdd(dst + (o0 + a) * ld + k0 + b, acc[a][b]);
}

// ---- warp-level tensor-core tile products with fp32-level accuracy (3 x TF32) -----------------------------------
// mma.sync m16n8k8 TF32 with both operands split x = hi + lo (hi = x truncated to TF32, lo = x - hi exact):
// acc += ah.bh + al.bh + ah.bl.  Used where a kernel is bound by the shared-memory wavefronts of FP32 register-tile
// products (decoder backward): a fragment load feeds 16x8x8 MACs instead of 4, at the price of the legacy tensor path's
// modest rate (measured 0.46 mma/clk/SM on B200, tools/microbench/mma_rate.cu) -- enough once LDS is the limiter.
// Fragment layout (PTX ISA, m16n8k8 .tf32), g = lane >> 2, t = lane & 3:
//   A (16 x 8): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)      B (8 x 8): b0 (k = t, n = g)  b1 (k = t+4, n = g)
//   C (16 x 8): c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
// Rows of A / columns of B / the K index may be permuted freely as long as both operands (and the reader of C) agree.
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
    uint32_t b0h, b0l, b1h, b1l;
    tf32_split(b0, b0h, b0l);
    tf32_split(b1, b1h, b1l);
    mma_tf32_16x8x8(c, ah, b0h, b1h);
    mma_tf32_16x8x8(c, al, b0h, b1h);
    mma_tf32_16x8x8(c, ah, b0l, b1l);
}

// ---- decoder saved-activation layouts (decoder.cu, decoder_tc.cu) -----------------------------------------------
// The forward saves only the recurrent state (h_t, c_t) of every step; the backward recomputes the gates from h_{t-1}
// with one more tensor-pipe product per step (round 1 saved the six gate / cell values per unit and step: 3.5 GB written
// and 4.1 GB re-read per generator step, and the backward was bound by the latency of those loads).
// Rows are grouped in 128-row tiles and the row index is the second-fastest dimension, so that the 32 rows a warp of
// the thread-per-row tensor-core kernel writes for one unit pair are 512 contiguous bytes (4 lines per 16-byte store
// instead of 32), and a 64-row tile of one unit pair is one contiguous 1 KB chunk for the backward's float4 loads.
//   acts   (T, rows/128, 16 unit-pairs, 128 rows, 4)            4 floats = (h, c) of unit 2j, (h, c) of unit 2j+1
//   u1save (T, rows/128, 4, 128 rows, 4)                        hidden2pos.0 pre-activations m = 4 q + (0..3)
//   h0save (rows/128, 8, 128 rows, 4)                           initial hidden state, units 4 q + (0..3)
__device__ __forceinline__ size_t dec_acts_off(size_t n_super, int t, size_t row, int u) {
    return (((size_t)t * n_super + (row >> 7)) * 16 + (u >> 1)) * 512 + (row & 127) * 4 + (u & 1) * 2;
}
__device__ __forceinline__ size_t dec_u1_off(size_t n_super, int t, size_t row, int m) {
    return (((size_t)t * n_super + (row >> 7)) * 4 + (m >> 2)) * 512 + (row & 127) * 4 + (m & 3);
}
__device__ __forceinline__ size_t dec_h0_off(size_t row, int u) {
    return ((row >> 7) * 8 + (u >> 2)) * 512 + (row & 127) * 4 + (u & 3);
}
