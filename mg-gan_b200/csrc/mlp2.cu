// Fused two-layer perceptron  Y = act2(W2 act1(W1 x + b1) + b2)  over a row batch, forward and backward in one launch each.
// The dense chains of the path are all of this form and all narrow (K <= 192, H <= 96, O <= 32):
//   discriminator in_encoder_fc 64 -> 32 -> 32, pred_encoder 24 -> 64 -> 32 and the two heads 192 -> 96 -> {1, G}
//   (mggan/model/modules/discriminators.py:46-56, 76-108; LeakyReLU(0.2) inside), PM-Network 128 -> 16 -> 16 (ReLU, ReLU;
//   mggan/model/modules/standard.py:99-105, its last 16 -> G layer stays a single dense layer).
// Launched layer by layer they were 2 + 4 kernels per chain (forward; input- and weight-gradient per layer) with the hidden
// activations making a round trip through HBM, and every launch latency-bound at these widths (~25 us for 16,384 rows where
// the arithmetic needs ~3).  Here a CTA keeps both weight matrices in shared memory, walks 64-row tiles, and the hidden
// tile never leaves the SM; the backward recomputes it from X (cheaper than saving it), forms both pre-activation
// gradients in shared memory, writes dX and keeps the weight-gradient blocks in registers across its tiles (one atomicAdd
// per element per CTA at the end).  All products are the register-blocked FP32 tile products of common.cuh written for
// runtime sizes (rows of 4 outputs x 4 rows, operands read as LDS.128 along the contraction axis).
#include "common.cuh"
#include "gemm_args.cuh"

namespace {

constexpr int TM = 64;                     // rows per tile
constexpr int RQ = TM / 4;                 // row quads: micro-tile rows are r, r + 16, r + 32, r + 48
constexpr int NB1_MAX = 5;                 // dW1 4x4 blocks per thread: (96 / 4) * (192 / 4) / 256 = 4.5

struct Dims {
    int K, H, O, O4;                       // O4 = O rounded up to a multiple of 4 (zero rows / columns)
    int ldk, ldh, ldo;                     // shared-memory leading dimensions: K + 4, H + 4, O4 + 4 (= 4 mod 32 for the sizes used)
};

// acc[i][j] += sum_k A[(r0 + 16 i) lda + k] W[(o0 + j) ldw + k]
__device__ __forceinline__ void rowdot(float (&acc)[4][4], const float* __restrict__ A, int lda, int r0,
                                       const float* __restrict__ W, int ldw, int o0, int n) {
#pragma unroll 2
    for (int k = 0; k < n; k += 4) {
        float4 a[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = ld4(A + (r0 + RQ * i) * lda + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = ld4(W + (o0 + j) * ldw + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                acc[i][j] = fmaf(a[i].x, w[j].x, fmaf(a[i].y, w[j].y, fmaf(a[i].z, w[j].z, fmaf(a[i].w, w[j].w, acc[i][j]))));
    }
}

// acc[i][c] += sum_o G[(r0 + 16 i) ldg + o] W[o ldw + c0 + c]     (input gradient: contraction over the layer's outputs)
__device__ __forceinline__ void dgrad(float (&acc)[4][4], const float* __restrict__ G, int ldg, int r0,
                                      const float* __restrict__ W, int ldw, int c0, int n) {
#pragma unroll 2
    for (int o = 0; o < n; o += 4) {
        float4 g[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = ld4(G + (r0 + RQ * i) * ldg + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = ld4(W + (o + j) * ldw + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(g[i].x, w[0].x, fmaf(g[i].y, w[1].x, fmaf(g[i].z, w[2].x, fmaf(g[i].w, w[3].x, acc[i][0]))));
            acc[i][1] = fmaf(g[i].x, w[0].y, fmaf(g[i].y, w[1].y, fmaf(g[i].z, w[2].y, fmaf(g[i].w, w[3].y, acc[i][1]))));
            acc[i][2] = fmaf(g[i].x, w[0].z, fmaf(g[i].y, w[1].z, fmaf(g[i].z, w[2].z, fmaf(g[i].w, w[3].z, acc[i][2]))));
            acc[i][3] = fmaf(g[i].x, w[0].w, fmaf(g[i].y, w[1].w, fmaf(g[i].z, w[2].w, fmaf(g[i].w, w[3].w, acc[i][3]))));
        }
    }
}

// acc[a][b] += sum_{r in [r_lo, r_hi)} G[r ldg + o0 + a] X[r ldx + k0 + b]     (weight gradient: contraction over rows)
__device__ __forceinline__ void wgrad(float (&acc)[4][4], const float* __restrict__ G, int ldg, int o0,
                                      const float* __restrict__ X, int ldx, int k0, int r_lo, int r_hi) {
#pragma unroll 4
    for (int r = r_lo; r < r_hi; ++r) {
        const float4 g = ld4(G + r * ldg + o0);
        const float4 x = ld4(X + r * ldx + k0);
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

// atomicAdd of a 4x4 register block into a dense row-major matrix in global memory: one 16-byte vector reduction per row
// (sm_90+ float4 atomicAdd) when the rows are 16-byte aligned -- a quarter of the reduction operations of the scalar form,
// and every CTA flushes its whole dW1 block at the end of the launch.
__device__ __forceinline__ void atomic_block44_g(float* __restrict__ dst, int ld, int o0, int k0, const float (&acc)[4][4], bool vec) {
    if (vec) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
            atomicAdd(reinterpret_cast<float4*>(dst + (size_t)(o0 + a) * ld + k0), make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]));
    } else {
        atomic_block44(dst, ld, o0, k0, acc);
    }
}

__device__ __forceinline__ void zero44(float (&a)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// Weights into shared memory: sW1 [H][ldk], sW2 [O4][ldh] (rows >= O zero), sb1 [H], sb2 [O4] (may be absent).
__device__ __forceinline__ void stage_weights(const Dims& d, const float* __restrict__ W1, const float* __restrict__ b1,
                                              const float* __restrict__ W2, const float* __restrict__ b2, float* sW1,
                                              float* sW2, float* sb1, float* sb2) {
    for (int i = threadIdx.x; i < d.H * d.K; i += MGGAN_THREADS) {
        const int h = i / d.K, k = i - h * d.K;
        sW1[h * d.ldk + k] = __ldg(W1 + i);
    }
    for (int i = threadIdx.x; i < d.O4 * d.H; i += MGGAN_THREADS) {
        const int o = i / d.H, h = i - o * d.H;
        sW2[o * d.ldh + h] = o < d.O ? __ldg(W2 + i) : 0.f;
    }
    for (int i = threadIdx.x; i < d.H; i += MGGAN_THREADS) sb1[i] = b1 != nullptr ? __ldg(b1 + i) : 0.f;
    if (sb2 != nullptr)
        for (int i = threadIdx.x; i < d.O4; i += MGGAN_THREADS) sb2[i] = (b2 != nullptr && i < d.O) ? __ldg(b2 + i) : 0.f;
}

// X tile (rows m0 .. m0 + 63, zero beyond M) -> sX [TM][ldk]; K % 4 == 0 and X 16-byte aligned: float4 loads
__device__ __forceinline__ void load_rows(float* __restrict__ sX, int ld, const float* __restrict__ X, int cols, long long m0,
                                          long long M) {
    const int q = cols >> 2;
    for (int i = threadIdx.x; i < TM * q; i += MGGAN_THREADS) {
        const int m = i / q, c = (i - m * q) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + m < M) v = __ldg(reinterpret_cast<const float4*>(X + (m0 + m) * cols + c));
        st4(sX + m * ld + c, v);
    }
}

// sH = act1(sX W1^T + b1)
__device__ __forceinline__ void hidden_tile(const Dims& d, const float* sX, const float* sW1, const float* sb1, float* sH,
                                            int act1, float slope1) {
    for (int mt = threadIdx.x; mt < RQ * (d.H >> 2); mt += MGGAN_THREADS) {
        const int mr = mt % RQ, o0 = (mt / RQ) << 2;
        float acc[4][4];
        zero44(acc);
        rowdot(acc, sX, d.ldk, mr, sW1, d.ldk, o0, d.K);
        const float4 b = ld4(sb1 + o0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st4(sH + (mr + RQ * i) * d.ldh + o0,
                make_float4(act_fwd(acc[i][0] + b.x, act1, slope1), act_fwd(acc[i][1] + b.y, act1, slope1),
                            act_fwd(acc[i][2] + b.z, act1, slope1), act_fwd(acc[i][3] + b.w, act1, slope1)));
    }
}

__global__ void __launch_bounds__(MGGAN_THREADS, 2)
mlp2_fwd_kernel(const float* __restrict__ X, long long M, Dims d, const float* __restrict__ W1, const float* __restrict__ b1,
                int act1, float slope1, const float* __restrict__ W2, const float* __restrict__ b2, int act2, float slope2,
                float* __restrict__ Y) {
    extern __shared__ __align__(16) float smem[];
    float* sW1 = smem;                          // [H][ldk]
    float* sW2 = sW1 + d.H * d.ldk;             // [O4][ldh]
    float* sb1 = sW2 + d.O4 * d.ldh;            // [H]
    float* sb2 = sb1 + d.H;                     // [O4]
    float* sX = sb2 + d.O4;                     // [TM][ldk]
    float* sH = sX + TM * d.ldk;                // [TM][ldh]
    stage_weights(d, W1, b1, W2, b2, sW1, sW2, sb1, sb2);
    const long long n_tiles = (M + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long m0 = tile * TM;
        __syncthreads();                        // weights staged / the previous tile's sH and sX are consumed
        load_rows(sX, d.ldk, X, d.K, m0, M);
        __syncthreads();
        hidden_tile(d, sX, sW1, sb1, sH, act1, slope1);
        __syncthreads();
        for (int mt = threadIdx.x; mt < RQ * (d.O4 >> 2); mt += MGGAN_THREADS) {
            const int mr = mt % RQ, o0 = (mt / RQ) << 2;
            float acc[4][4];
            zero44(acc);
            rowdot(acc, sH, d.ldh, mr, sW2, d.ldh, o0, d.H);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long m = m0 + mr + RQ * i;
                if (m >= M) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (o0 + j < d.O) Y[m * d.O + o0 + j] = act_fwd(acc[i][j] + sb2[o0 + j], act2, slope2);
            }
        }
    }
}

// Row-group split of a weight-gradient with `nb` 4x4 blocks over the 256 threads: block b of thread t and its row range.
struct Split { int blk, r_lo, r_hi; bool active; };
__device__ __forceinline__ Split split_rows(int nb) {
    int rg = MGGAN_THREADS / nb;                // row groups (>= 1 when nb <= 256)
    rg = rg >= 16 ? 16 : rg >= 8 ? 8 : rg >= 4 ? 4 : rg >= 2 ? 2 : 1;
    const int per = TM / rg;
    Split s;
    s.active = (int)threadIdx.x < nb * rg;
    s.blk = threadIdx.x % nb;
    const int g = threadIdx.x / nb;
    s.r_lo = g * per;
    s.r_hi = s.r_lo + per;
    return s;
}

template <bool WGRAD>
__global__ void __launch_bounds__(MGGAN_THREADS, WGRAD ? 1 : 2)
mlp2_bwd_kernel(const float* __restrict__ X, long long M, Dims d, const float* __restrict__ W1, const float* __restrict__ b1,
                int act1, float slope1, const float* __restrict__ W2, int act2, float slope2, const float* __restrict__ Y,
                const float* __restrict__ dY, float* __restrict__ dX, float* __restrict__ dW1, float* __restrict__ db1,
                float* __restrict__ dW2, float* __restrict__ db2) {
    extern __shared__ __align__(16) float smem[];
    float* sW1 = smem;                          // [H][ldk]
    float* sW2 = sW1 + d.H * d.ldk;             // [O4][ldh]
    float* sb1 = sW2 + d.O4 * d.ldh;            // [H]
    float* sX = sb1 + d.H;                      // [TM][ldk]
    float* sH = sX + TM * d.ldk;                // [TM][ldh]  act1 output (recomputed)
    float* sZ1 = sH + TM * d.ldh;               // [TM][ldh]  gradient w.r.t. the layer-1 pre-activation
    float* sZ2 = sZ1 + TM * d.ldh;              // [TM][ldo]  gradient w.r.t. the layer-2 pre-activation
    stage_weights(d, W1, b1, W2, nullptr, sW1, sW2, sb1, nullptr);

    const int nb1 = (d.H >> 2) * (d.K >> 2), nb2 = (d.O4 >> 2) * (d.H >> 2);
    float acc1[NB1_MAX][4][4], acc2[4][4], bacc1 = 0.f, bacc2 = 0.f;
    if (WGRAD) {
#pragma unroll
        for (int q = 0; q < NB1_MAX; ++q) zero44(acc1[q]);
        zero44(acc2);
    }
    const Split s1 = split_rows(nb1 < MGGAN_THREADS ? nb1 : MGGAN_THREADS), s2 = split_rows(nb2 < MGGAN_THREADS ? nb2 : MGGAN_THREADS);

    const long long n_tiles = (M + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long m0 = tile * TM;
        __syncthreads();
        load_rows(sX, d.ldk, X, d.K, m0, M);
        // dz2 = dY * act2'(Y), zero for rows >= M and columns >= O
        for (int i = threadIdx.x; i < TM * d.O4; i += MGGAN_THREADS) {
            const int m = i / d.O4, o = i - m * d.O4;
            float v = 0.f;
            if (m0 + m < M && o < d.O) {
                const long long off = (m0 + m) * d.O + o;
                v = __ldg(dY + off);
                if (act2 != ACT_NONE) v *= act_bwd(__ldg(Y + off), act2, slope2);
            }
            sZ2[m * d.ldo + o] = v;
        }
        __syncthreads();
        hidden_tile(d, sX, sW1, sb1, sH, act1, slope1);
        __syncthreads();
        // dz1 = (dz2 W2) * act1'(h)
        for (int mt = threadIdx.x; mt < RQ * (d.H >> 2); mt += MGGAN_THREADS) {
            const int mr = mt % RQ, h0 = (mt / RQ) << 2;
            float acc[4][4];
            zero44(acc);
            dgrad(acc, sZ2, d.ldo, mr, sW2, d.ldh, h0, d.O4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 h = ld4(sH + (mr + RQ * i) * d.ldh + h0);
                st4(sZ1 + (mr + RQ * i) * d.ldh + h0,
                    make_float4(acc[i][0] * act_bwd(h.x, act1, slope1), acc[i][1] * act_bwd(h.y, act1, slope1),
                                acc[i][2] * act_bwd(h.z, act1, slope1), acc[i][3] * act_bwd(h.w, act1, slope1)));
            }
        }
        __syncthreads();
        if (dX != nullptr) {                    // dX = dz1 W1
            for (int mt = threadIdx.x; mt < RQ * (d.K >> 2); mt += MGGAN_THREADS) {
                const int mr = mt % RQ, k0 = (mt / RQ) << 2;
                float acc[4][4];
                zero44(acc);
                dgrad(acc, sZ1, d.ldh, mr, sW1, d.ldk, k0, d.H);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long m = m0 + mr + RQ * i;
                    if (m < M) st4(dX + m * d.K + k0, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                }
            }
        }
        if (WGRAD) {
            // dW1[h][k] += sum_m dz1[m][h] x[m][k]
            if (nb1 >= MGGAN_THREADS) {
#pragma unroll
                for (int q = 0; q < NB1_MAX; ++q) {
                    const int b = threadIdx.x + q * MGGAN_THREADS;
                    if (b < nb1) wgrad(acc1[q], sZ1, d.ldh, (b % (d.H >> 2)) << 2, sX, d.ldk, (b / (d.H >> 2)) << 2, 0, TM);
                }
            } else if (s1.active) {
                wgrad(acc1[0], sZ1, d.ldh, (s1.blk % (d.H >> 2)) << 2, sX, d.ldk, (s1.blk / (d.H >> 2)) << 2, s1.r_lo, s1.r_hi);
            }
            // dW2[o][h] += sum_m dz2[m][o] h[m][h]
            if (s2.active)
                wgrad(acc2, sZ2, d.ldo, (s2.blk % (d.O4 >> 2)) << 2, sH, d.ldh, (s2.blk / (d.O4 >> 2)) << 2, s2.r_lo, s2.r_hi);
            if ((int)threadIdx.x < d.H) {
                for (int m = 0; m < TM; ++m) bacc1 += sZ1[m * d.ldh + threadIdx.x];
            } else if ((int)threadIdx.x - 128 >= 0 && (int)threadIdx.x - 128 < d.O) {
                for (int m = 0; m < TM; ++m) bacc2 += sZ2[m * d.ldo + threadIdx.x - 128];
            }
        }
    }
    if (WGRAD) {
        const bool vec = (reinterpret_cast<uintptr_t>(dW1) & 15) == 0;        // K % 4 == 0: rows stay 16-byte aligned
        if (nb1 >= MGGAN_THREADS) {
#pragma unroll
            for (int q = 0; q < NB1_MAX; ++q) {
                const int b = threadIdx.x + q * MGGAN_THREADS;
                if (b < nb1) atomic_block44_g(dW1, d.K, (b % (d.H >> 2)) << 2, (b / (d.H >> 2)) << 2, acc1[q], vec);
            }
        } else if (s1.active) {
            atomic_block44_g(dW1, d.K, (s1.blk % (d.H >> 2)) << 2, (s1.blk / (d.H >> 2)) << 2, acc1[0], vec);
        }
        if (s2.active) {
            const int o0 = (s2.blk % (d.O4 >> 2)) << 2, h0 = (s2.blk / (d.O4 >> 2)) << 2;
#pragma unroll
            for (int a = 0; a < 4; ++a)
                if (o0 + a < d.O)
#pragma unroll
                    for (int b = 0; b < 4; ++b) atomicAdd(dW2 + (o0 + a) * d.H + h0 + b, acc2[a][b]);
        }
        if ((int)threadIdx.x < d.H) {
            if (db1 != nullptr) atomicAdd(db1 + threadIdx.x, bacc1);
        } else if ((int)threadIdx.x - 128 >= 0 && (int)threadIdx.x - 128 < d.O) {
            if (db2 != nullptr) atomicAdd(db2 + threadIdx.x - 128, bacc2);
        }
    }
}

int sm_count_() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int make_dims(int K, int H, int O, Dims& d, const char* who) {
    if (!(K >= 4 && K <= 192 && (K & 3) == 0 && H >= 4 && H <= 96 && (H & 3) == 0 && O >= 1 && O <= 32))
        return mggan_set_error(MGGAN_ERR_INVALID, "%s: sizes K=%d (4..192, %%4), H=%d (4..96, %%4), O=%d (1..32)", who, K, H, O);
    d.K = K; d.H = H; d.O = O; d.O4 = (O + 3) & ~3;
    d.ldk = K + 4; d.ldh = H + 4; d.ldo = d.O4 + 4;
    return MGGAN_OK;
}

}  // namespace

extern "C" int mggan_mlp2_fwd(const float* X, long long M, int K, const float* W1, const float* b1, int H, int act1,
                              float slope1, const float* W2, const float* b2, int O, int act2, float slope2, float* Y,
                              cudaStream_t stream) {
    Dims d;
    if (int rc = make_dims(K, H, O, d, "mggan_mlp2_fwd")) return rc;
    MGGAN_REQUIRE(M >= 0 && act1 >= 0 && act1 <= 3 && act2 >= 0 && act2 <= 3, "mggan_mlp2_fwd: bad arguments");
    MGGAN_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, "mggan_mlp2_fwd: X must be 16-byte aligned");
    if (M == 0) return MGGAN_OK;
    const size_t smem = sizeof(float) * ((size_t)d.H * d.ldk + d.O4 * d.ldh + d.H + d.O4 + TM * d.ldk + TM * d.ldh);
    cudaFuncSetAttribute(mlp2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long tiles = (M + TM - 1) / TM;
    const int cap = sm_count_() * 2;
    mlp2_fwd_kernel<<<(int)(tiles < cap ? tiles : cap), MGGAN_THREADS, smem, stream>>>(X, M, d, W1, b1, act1, slope1, W2, b2,
                                                                                     act2, slope2, Y);
    return mggan_check_launch("mlp2_fwd");
}

// dX (M x K, overwritten) may be NULL; dW1 (H x K), db1 (H), dW2 (O x H), db2 (O) are ACCUMULATED into (the caller
// zero-fills) and are either all set or all NULL (frozen weights: input gradient only).  The hidden layer is recomputed.
extern "C" int mggan_mlp2_bwd(const float* X, long long M, int K, const float* W1, const float* b1, int H, int act1,
                              float slope1, const float* W2, int O, int act2, float slope2, const float* Y, const float* dY,
                              float* dX, float* dW1, float* db1, float* dW2, float* db2, cudaStream_t stream) {
    Dims d;
    if (int rc = make_dims(K, H, O, d, "mggan_mlp2_bwd")) return rc;
    MGGAN_REQUIRE(M >= 0 && act1 >= 0 && act1 <= 3 && act2 >= 0 && act2 <= 3, "mggan_mlp2_bwd: bad arguments");
    const bool wg = dW1 != nullptr;
    MGGAN_REQUIRE((dW2 != nullptr) == wg && (wg || (db1 == nullptr && db2 == nullptr)),
                  "mggan_mlp2_bwd: weight gradients are all set or all NULL");
    MGGAN_REQUIRE(wg || dX != nullptr, "mggan_mlp2_bwd: nothing to compute");
    MGGAN_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(dX) & 15) == 0,
                  "mggan_mlp2_bwd: X and dX must be 16-byte aligned");
    MGGAN_REQUIRE((d.H >> 2) * (d.K >> 2) <= NB1_MAX * MGGAN_THREADS && d.H <= 128 && d.O <= 128,
                  "mggan_mlp2_bwd: weight-gradient block table too small");
    if (M == 0) return MGGAN_OK;
    const size_t smem = sizeof(float) * ((size_t)d.H * d.ldk + d.O4 * d.ldh + d.H + TM * d.ldk + 2 * TM * d.ldh + TM * d.ldo);
    const long long tiles = (M + TM - 1) / TM;
    if (wg) {
        cudaFuncSetAttribute(mlp2_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int cap = sm_count_();
        mlp2_bwd_kernel<true><<<(int)(tiles < cap ? tiles : cap), MGGAN_THREADS, smem, stream>>>(
            X, M, d, W1, b1, act1, slope1, W2, act2, slope2, Y, dY, dX, dW1, db1, dW2, db2);
    } else {
        cudaFuncSetAttribute(mlp2_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int cap = sm_count_() * 2;
        mlp2_bwd_kernel<false><<<(int)(tiles < cap ? tiles : cap), MGGAN_THREADS, smem, stream>>>(
            X, M, d, W1, b1, act1, slope1, W2, act2, slope2, Y, dY, dX, nullptr, nullptr, nullptr, nullptr);
    }
    return mggan_check_launch("mlp2_bwd");
}
