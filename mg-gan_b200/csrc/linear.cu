// Dense layers of the path (PM-Net, enc_h_to_dec_h, discriminator encoders and heads):
//   Y = act(X W^T + b),  X (M x K) rows = agents or sampled trajectories, W (O x K) PyTorch layout.
// References: mggan/utils.py:134-149 (make_mlp), standard.py:91-105, discriminators.py:46-56,76-108.
//
// M is large (up to k*N rows), K and O are <= 192, so one strided fp32 tile product covers the
// forward, the input gradient (dX = dZ W) and the weight gradient (dW = dZ^T X, reduction over
// the rows split across CTAs and merged with atomics).  64x64 output tile, BK = 16, 4x4
// register micro-tile, operands transposed into shared memory so the inner loop is two
// LDS.128 per 16 FMAs.
#include "common.cuh"
#include "gemm_args.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int LDT = BM + 4;

__global__ void __launch_bounds__(MGGAN_THREADS)
gemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][LDT];
    __shared__ __align__(16) float Bs[BK][LDT];
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tm = threadIdx.x >> 4, tn = threadIdx.x & 15;       // 16 x 16 threads, 4 x 4 each
    int k_begin = 0, k_end = g.K;
    if (g.splitk > 1) {
        int per = (g.K + g.splitk - 1) / g.splitk;
        per = (per + BK - 1) / BK * BK;
        k_begin = blockIdx.z * per;
        k_end = min(g.K, k_begin + per);
        if (k_begin >= k_end) return;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float csum = 0.f;
    const bool a_kfast = g.sak == 1, b_kfast = g.sbk == 1;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int it = 0; it < (BM * BK) / MGGAN_THREADS; ++it) {
            int idx = threadIdx.x + it * MGGAN_THREADS;
            int mm = a_kfast ? idx / BK : idx % BM, kk = a_kfast ? idx % BK : idx / BM;
            int m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < g.M && k < k_end) {
                long long off = m * g.sam + k * g.sak;
                v = __ldg(g.A + off);
                if (g.Ay != nullptr) v *= act_bwd(__ldg(g.Ay + off), g.act_in, g.slope);
            }
            As[kk][mm] = v;
        }
#pragma unroll
        for (int it = 0; it < (BN * BK) / MGGAN_THREADS; ++it) {
            int idx = threadIdx.x + it * MGGAN_THREADS;
            int nn = b_kfast ? idx / BK : idx % BN, kk = b_kfast ? idx % BK : idx / BN;
            int n = n0 + nn, k = k0 + kk;
            float v = 0.f;
            if (n < g.N && k < k_end) v = __ldg(g.B + n * g.sbn + k * g.sbk);
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = ld4(&As[kk][tm * 4]);
            float4 b = ld4(&Bs[kk][tn * 4]);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
            acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
            acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
            acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
            acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
        }
        if (g.colsum != nullptr && blockIdx.y == 0 && threadIdx.x < BM) {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) csum += As[kk][threadIdx.x];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + tm * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tn * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            float* dst = g.C + m * g.scm + n * g.scn;
            if (g.splitk > 1) {
                atomicAdd(dst, v);
            } else {
                if (g.bias != nullptr) v += __ldg(g.bias + n);
                *dst = act_fwd(v, g.act, g.slope);
            }
        }
    }
    if (g.colsum != nullptr && blockIdx.y == 0 && threadIdx.x < BM && m0 + threadIdx.x < g.M)
        atomicAdd(g.colsum + m0 + threadIdx.x, csum);
}

// ---- variant 2 (opt-in, mggan_set_gemm_variant(2)): same contract as gemm_kernel, (16 TM) x 64 output tile with a
// TM x 4 register micro-tile (TM = 8: three LDS.128 feed 32 FMAs instead of two feeding 16) and the next K-slab's global
// loads issued into registers before the current slab is computed (one memory latency overlapped per slab instead of
// exposed).  Written after the round's GPU budget was spent: NOT the default until tests/test_gpu_zf_gemm_variants.py has run.
template <int TM>
__global__ void __launch_bounds__(MGGAN_THREADS)
gemm_kernel_v2(GemmArgs g) {
    constexpr int BM2 = 16 * TM, LDA = BM2 + 4;
    constexpr int NA = (BM2 * BK) / MGGAN_THREADS, NB = (BN * BK) / MGGAN_THREADS;
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][LDT];
    const int m0 = blockIdx.x * BM2, n0 = blockIdx.y * BN;
    const int tm = threadIdx.x >> 4, tn = threadIdx.x & 15;       // 16 x 16 threads, TM x 4 each
    int k_begin = 0, k_end = g.K;
    if (g.splitk > 1) {
        int per = (g.K + g.splitk - 1) / g.splitk;
        per = (per + BK - 1) / BK * BK;
        k_begin = blockIdx.z * per;
        k_end = min(g.K, k_begin + per);
        if (k_begin >= k_end) return;
    }
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float csum = 0.f;
    const bool a_kfast = g.sak == 1, b_kfast = g.sbk == 1;
    float ra[NA], rb[NB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int it = 0; it < NA; ++it) {
            const int idx = threadIdx.x + it * MGGAN_THREADS;
            const int mm = a_kfast ? idx / BK : idx % BM2, kk = a_kfast ? idx % BK : idx / BM2;
            const int m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < g.M && k < k_end) {
                const long long off = m * g.sam + k * g.sak;
                v = __ldg(g.A + off);
                if (g.Ay != nullptr) v *= act_bwd(__ldg(g.Ay + off), g.act_in, g.slope);
            }
            ra[it] = v;
        }
#pragma unroll
        for (int it = 0; it < NB; ++it) {
            const int idx = threadIdx.x + it * MGGAN_THREADS;
            const int nn = b_kfast ? idx / BK : idx % BN, kk = b_kfast ? idx % BK : idx / BN;
            const int n = n0 + nn, k = k0 + kk;
            rb[it] = (n < g.N && k < k_end) ? __ldg(g.B + n * g.sbn + k * g.sbk) : 0.f;
        }
    };
    fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int it = 0; it < NA; ++it) {
            const int idx = threadIdx.x + it * MGGAN_THREADS;
            const int mm = a_kfast ? idx / BK : idx % BM2, kk = a_kfast ? idx % BK : idx / BM2;
            As[kk][mm] = ra[it];
        }
#pragma unroll
        for (int it = 0; it < NB; ++it) {
            const int idx = threadIdx.x + it * MGGAN_THREADS;
            const int nn = b_kfast ? idx / BK : idx % BN, kk = b_kfast ? idx % BK : idx / BN;
            Bs[kk][nn] = rb[it];
        }
        __syncthreads();
        if (k0 + BK < k_end) fetch(k0 + BK);          // in flight while this slab is computed
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM];
#pragma unroll
            for (int i4 = 0; i4 < TM / 4; ++i4) {
                const float4 v = ld4(&As[kk][tm * TM + i4 * 4]);
                a[i4 * 4] = v.x; a[i4 * 4 + 1] = v.y; a[i4 * 4 + 2] = v.z; a[i4 * 4 + 3] = v.w;
            }
            const float4 b = ld4(&Bs[kk][tn * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]); acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]); acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
            }
        }
        if (g.colsum != nullptr && blockIdx.y == 0 && threadIdx.x < BM2) {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) csum += As[kk][threadIdx.x];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + tm * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            float* dst = g.C + m * g.scm + n * g.scn;
            if (g.splitk > 1) {
                atomicAdd(dst, v);
            } else {
                if (g.bias != nullptr) v += __ldg(g.bias + n);
                *dst = act_fwd(v, g.act, g.slope);
            }
        }
    }
    if (g.colsum != nullptr && blockIdx.y == 0 && threadIdx.x < BM2 && m0 + threadIdx.x < g.M)
        atomicAdd(g.colsum + m0 + threadIdx.x, csum);
}

int g_gemm_variant = 1;

// rows of the output tile of the selected variant for a GEMM with M output rows
int tile_rows(int M) { return g_gemm_variant == 3 ? 128 : g_gemm_variant == 2 ? (M > 64 ? 128 : 64) : BM; }

int launch(const GemmArgs& g, cudaStream_t s) {
    if (g_gemm_variant == 3) {                       // tensor cores (linear_tc.cu); -1 = shape it does not build
        const int rc = mggan_gemm_tc_launch(g, s);
        if (rc != -1) return rc;
    }
    const int bm = g_gemm_variant == 3 ? BM : tile_rows(g.M);
    dim3 grid((g.M + bm - 1) / bm, (g.N + BN - 1) / BN, g.splitk > 1 ? g.splitk : 1);
    if (g_gemm_variant == 2) {
        if (bm == 128) gemm_kernel_v2<8><<<grid, MGGAN_THREADS, 0, s>>>(g);
        else gemm_kernel_v2<4><<<grid, MGGAN_THREADS, 0, s>>>(g);
    } else {
        gemm_kernel<<<grid, MGGAN_THREADS, 0, s>>>(g);
    }
    return mggan_check_launch("linear");
}

}  // namespace

extern "C" int mggan_linear_fwd(const float* X, int M, int K, const float* W, const float* bias, int O, int act,
                                float slope, float* Y, cudaStream_t stream) {
    MGGAN_REQUIRE(M >= 0 && K >= 1 && O >= 1 && act >= 0 && act <= 3, "mggan_linear_fwd: bad shape/act");
    if (M == 0) return MGGAN_OK;
    GemmArgs g{};
    g.A = X; g.sam = K; g.sak = 1; g.Ay = nullptr;
    g.B = W; g.sbn = K; g.sbk = 1;
    g.C = Y; g.scm = O; g.scn = 1;
    g.bias = bias; g.colsum = nullptr; g.M = M; g.N = O; g.K = K; g.act = act; g.slope = slope; g.splitk = 1;
    return launch(g, stream);
}

// dX (may be null) is overwritten; dW and db (may be null) are accumulated into (caller zero-fills).
extern "C" int mggan_linear_bwd(const float* X, int M, int K, const float* W, int O, int act, float slope,
                                const float* Y, const float* dY, float* dX, float* dW, float* db,
                                cudaStream_t stream) {
    MGGAN_REQUIRE(M >= 0 && K >= 1 && O >= 1 && act >= 0 && act <= 3, "mggan_linear_bwd: bad shape/act");
    if (M == 0) return MGGAN_OK;
    const float* Ay = act == ACT_NONE ? nullptr : Y;
    if (dX != nullptr) {          // dX(m, k) = sum_o dZ(m, o) W(o, k)
        GemmArgs g{};
        g.A = dY; g.sam = O; g.sak = 1; g.Ay = Ay;
        g.B = W; g.sbn = 1; g.sbk = K;
        g.C = dX; g.scm = K; g.scn = 1;
        g.M = M; g.N = K; g.K = O; g.act = ACT_NONE; g.act_in = act; g.slope = slope; g.splitk = 1;
        int rc = launch(g, stream);
        if (rc) return rc;
    }
    if (dW != nullptr) {          // dW(o, k) = sum_m dZ(m, o) X(m, k) ; db(o) = sum_m dZ(m, o)
        GemmArgs g{};
        g.A = dY; g.sam = 1; g.sak = O; g.Ay = Ay;
        g.B = X; g.sbn = 1; g.sbk = K;
        g.C = dW; g.scm = K; g.scn = 1;
        g.colsum = db;
        g.M = O; g.N = K; g.K = M; g.act = ACT_NONE; g.act_in = act; g.slope = slope;
        const int bm = tile_rows(O);
        int tiles = ((O + bm - 1) / bm) * ((K + BN - 1) / BN);
        int want = (148 * 4 + tiles - 1) / tiles;
        int maxsplit = (M + 4 * BK - 1) / (4 * BK);
        g.splitk = want < maxsplit ? want : maxsplit;
        if (g.splitk < 2) g.splitk = 2;        // epilogue accumulates atomically in this mode
        return launch(g, stream);
    }
    return MGGAN_OK;
}

// 1 = the 64 x 64 tile kernel (default), 2 = the 128 x 64 register-prefetch kernel, 3 = tcgen05 3 x TF32 (linear_tc.cu); 2 and
// 3 are opt-in until measured.  Process-wide; returns the previous value, or -1 for an unknown variant.
extern "C" int mggan_set_gemm_variant(int variant) {
    if (variant < 1 || variant > 3) return -1;
    const int prev = g_gemm_variant;
    g_gemm_variant = variant;
    return prev;
}
