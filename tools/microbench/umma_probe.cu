// tcgen05 feature probe for the next kernels: which operand forms of tcgen05.mma kind::tf32 work on
// this part and how they are encoded, answered by exact integer-valued products checked on the host.
//
//   ss_kmajor     A, B in shared memory, K-major canonical no-swizzle layout (the form decoder_tc.cu uses), N = 128/32/16,
//                 and the same with core matrices 144 B apart (padded LBO) for conflict-free transposed staging
//   ts_tmem_a     A read from tensor memory (written by tcgen05.st, thread r = lane r), B K-major in shared memory:
//                 the form a thread-per-row backward needs for dh = dG . W_hh without a 64 KB shared-memory operand
//   mn_major      A and/or B MN-major (the reduction index K is the slow one: [k][m] storage), both readings of the
//                 (LBO, SBO) fields: the form weight gradients dW += dG^T h need (K = rows of the tile)
//   shifted_rows  K-major A whose rows sit at a 16-byte stride ([k/4][row][4] planes, SBO = 128 B) addressed from a start
//                 address that is NOT 128-byte aligned (row shift j): the form an implicit-GEMM 3x3 convolution needs
//                 (one tap = one descriptor offset)
//   rate          back-to-back issue of 2048 MMAs: clocks per MMA for N = 16 / 32 / 64 / 128, SS and TS
//
// Every probe runs in its own process (an illegal encoding kills the context); waits are bounded (trap, not hang).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a umma_probe.cu -o umma_probe
//   for p in $(./umma_probe list); do timeout 30 ./umma_probe $p; done
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct Probe {
    int M, N, K;              // M = 128; K multiple of 8
    int a_mode, b_mode;       // 0 K-major smem, 1 MN-major smem, 2 (A only) tensor memory
    // data placement (bytes): address(r, k) of operand X
    //   K-major : (r / 8) * mn_group + (k / 4) * k_group + (r % 8) * 16 + (k % 4) * 4
    //   MN-major: (r / 4) * mn_group + (k / 8) * k_group + (k % 8) * 16 + (r % 4) * 4
    int a_mn_group, a_k_group, b_mn_group, b_k_group;
    int a_lbo, a_sbo, b_lbo, b_sbo;        // descriptor fields (bytes)
    int a_kstep, b_kstep;                  // descriptor start-address advance per K = 8 step (bytes)
    int a_shift_rows;                      // shifted_rows: A staged with M + shift rows, descriptor starts at row `shift`
    int repeat;                            // rate probe: number of times the K loop is issued
    int acc_sets;                          // rate probe: accumulators (TMEM column blocks of N) the MMAs rotate over (0 / 1 = one)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(Probe p, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, long long* clocks) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(&s_bar);
    const int a_rows = p.M + p.a_shift_rows;
    // A region first, B region after it (1024-byte aligned)
    int a_bytes = 0;
    if (p.a_mode == 0) a_bytes = ((a_rows + 7) / 8) * p.a_mn_group + (p.K / 4) * p.a_k_group;      // upper bound of every offset
    if (p.a_mode == 1) a_bytes = (p.M / 4) * p.a_mn_group + (p.K / 8) * p.a_k_group;
    a_bytes = (a_bytes + 1023) & ~1023;
    unsigned char* sA = smem;
    unsigned char* sB = smem + a_bytes;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- stage the shared-memory operands
    if (p.a_mode == 0)
        for (int i = tid; i < a_rows * p.K; i += 128) {
            const int r = i / p.K, k = i - r * p.K;
            *reinterpret_cast<float*>(sA + (r / 8) * p.a_mn_group + (k / 4) * p.a_k_group + (r % 8) * 16 + (k % 4) * 4) = A[i];
        }
    if (p.a_mode == 1)
        for (int i = tid; i < p.M * p.K; i += 128) {
            const int r = i / p.K, k = i - r * p.K;
            *reinterpret_cast<float*>(sA + (r / 4) * p.a_mn_group + (k / 8) * p.a_k_group + (k % 8) * 16 + (r % 4) * 4) = A[i];
        }
    for (int i = tid; i < p.N * p.K; i += 128) {
        const int r = i / p.K, k = i - r * p.K;
        const int off = p.b_mode == 0 ? (r / 8) * p.b_mn_group + (k / 4) * p.b_k_group + (r % 8) * 16 + (k % 4) * 4
                                      : (r / 4) * p.b_mn_group + (k / 8) * p.b_k_group + (k % 8) * 16 + (r % 4) * 4;
        *reinterpret_cast<float*>(sB + off) = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t A_COL0 = 256;               // A operand columns [256, 256 + K) when it lives in tensor memory
    if (p.a_mode == 2) {                        // thread r writes row r of A: 8 columns per store
        for (int k0 = 0; k0 < p.K; k0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[tid * p.K + k0 + j]);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(tmem_lane + A_COL0 + k0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.a_mode == 1 ? 1u : 0u) << 15) | ((p.b_mode == 1 ? 1u : 0u) << 16) |
                           ((static_cast<uint32_t>(p.N) >> 3) << 17) | ((static_cast<uint32_t>(p.M) >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        // shifted start: row j of the [k/4][row][4] planes is j * 16 bytes further (K-major, mn_group = 128)
        const uint64_t adesc = make_desc(smem_u32(sA) + p.a_shift_rows * 16, p.a_lbo, p.a_sbo);
        const uint64_t bdesc = make_desc(smem_u32(sB), p.b_lbo, p.b_sbo);
        t0 = clock64();
        for (int rep = 0; rep < p.repeat; ++rep)
            for (int kk = 0; kk < p.K / 8; ++kk) {
                const uint32_t acc = (rep > 0 || kk > 0) ? 1u : 0u;
                const uint32_t d_tmem = tmem_base + (p.acc_sets > 1 ? static_cast<uint32_t>(((rep * (p.K / 8) + kk) % p.acc_sets) * p.N) : 0u);
                const uint64_t bd = bdesc + static_cast<uint64_t>((kk * p.b_kstep) >> 4);
                if (p.a_mode == 2) {
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, q;\n\t}"
                                 ::"r"(d_tmem), "r"(tmem_base + A_COL0 + kk * 8), "l"(bd), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
                } else {
                    const uint64_t ad = adesc + static_cast<uint64_t>((kk * p.a_kstep) >> 4);
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, q;\n\t}"
                                 ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
                }
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    for (uint32_t spins = 0; !mbar_try(bar, 0); ++spins)
        if (spins > (1u << 24)) __trap();
    if (tid == 0) {
        t1 = clock64();
        clocks[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < p.N; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem_lane + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) D[tid * p.N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

struct Named { const char* name; Probe p; };

static std::vector<Named> probes() {
    std::vector<Named> v;
    auto kmaj = [](int N, int K) {
        // canonical K-major as in decoder_tc.cu: LBO = 128 (next core matrix along K), SBO = K * 32 bytes (next 8 rows)
        Probe p{};
        p.M = 128; p.N = N; p.K = K; p.a_mode = 0; p.b_mode = 0;
        p.a_mn_group = p.b_mn_group = K * 32; p.a_k_group = p.b_k_group = 128;
        p.a_lbo = p.b_lbo = 128; p.a_sbo = p.b_sbo = K * 32; p.a_kstep = p.b_kstep = 256; p.repeat = 1;
        return p;
    };
    v.push_back({"ss_kmajor_n128_k32", kmaj(128, 32)});
    v.push_back({"ss_kmajor_n32_k32", kmaj(32, 32)});
    v.push_back({"ss_kmajor_n16_k32", kmaj(16, 32)});
    v.push_back({"ss_kmajor_n32_k128", kmaj(32, 128)});
    {   // K-major with padded core-matrix spacing (LBO = 144 B): lets a thread-per-row kernel write TRANSPOSED operands
        // (4-byte stores, k fastest across lanes) without 8-way bank conflicts
        Probe p = kmaj(32, 32);
        p.a_k_group = p.b_k_group = 144; p.a_lbo = p.b_lbo = 144;
        p.a_mn_group = p.b_mn_group = (32 / 4) * 144; p.a_sbo = p.b_sbo = (32 / 4) * 144;
        p.a_kstep = p.b_kstep = 2 * 144;
        v.push_back({"ss_kmajor_lbo144", p});
    }
    for (int N : {32, 128}) {
        Probe p = kmaj(N, 128);
        p.a_mode = 2;
        v.push_back({N == 32 ? "ts_tmem_a_n32_k128" : "ts_tmem_a_n128_k128", p});
    }
    {   // MN-major, data: 16-byte MN chunks 128 B apart (one core matrix = 8 k x 16 B), 8-k groups (MN / 4) * 128 B apart
        for (int variant = 0; variant < 2; ++variant)
            for (int which = 0; which < 3; ++which) {          // 0: A and B MN-major, 1: A only, 2: B only
                Probe p = kmaj(32, 128);
                const bool am = which != 2, bm = which != 1;
                if (am) { p.a_mode = 1; p.a_mn_group = 128; p.a_k_group = (128 / 4) * 128; p.a_kstep = p.a_k_group; }
                if (bm) { p.b_mode = 1; p.b_mn_group = 128; p.b_k_group = (32 / 4) * 128; p.b_kstep = p.b_k_group; }
                // variant 0: SBO = stride between MN chunks, LBO = stride between 8-k groups (cute make_umma_desc<Major::MN>,
                // INTERLEAVE row); variant 1: the two fields swapped
                if (am) { p.a_sbo = variant == 0 ? p.a_mn_group : p.a_k_group; p.a_lbo = variant == 0 ? p.a_k_group : p.a_mn_group; }
                if (bm) { p.b_sbo = variant == 0 ? p.b_mn_group : p.b_k_group; p.b_lbo = variant == 0 ? p.b_k_group : p.b_mn_group; }
                static char names[6][48];
                snprintf(names[variant * 3 + which], 48, "mn_major_%s_v%d", which == 0 ? "ab" : which == 1 ? "a" : "b", variant);
                v.push_back({names[variant * 3 + which], p});
            }
    }
    for (int shift : {0, 1, 3, 8, 19}) {       // rows at a 16-byte stride: [k/4][row][4] planes
        Probe p = kmaj(16, 16);
        p.a_shift_rows = shift;
        const int rows_alloc = 128 + 24;        // plane stride fixed, independent of the shift
        p.a_mn_group = 128; p.a_k_group = rows_alloc * 16; p.a_sbo = 128; p.a_lbo = p.a_k_group; p.a_kstep = 2 * p.a_k_group;
        static char names[5][32];
        static int n = 0;
        snprintf(names[n], 32, "shifted_rows_%d", shift);
        v.push_back({names[n++], p});
    }
    for (int N : {16, 32, 64, 128})
        for (int ts = 0; ts < 2; ++ts) {
            Probe p = kmaj(N, 32);
            p.a_mode = ts ? 2 : 0;
            p.repeat = 512;                      // 2048 MMAs
            static char names[8][32];
            static int n = 0;
            snprintf(names[n], 32, "rate_%s_n%d", ts ? "ts" : "ss", N);
            v.push_back({names[n++], p});
        }
    // the same with the MMAs rotating over 4 independent accumulators: issue rate without the accumulate dependency
    for (int N : {16, 32, 64})
        for (int ts = 0; ts < 2; ++ts) {
            Probe p = kmaj(N, 32);
            p.a_mode = ts ? 2 : 0;
            p.repeat = 512;
            p.acc_sets = 4;
            static char names[6][32];
            static int n = 0;
            snprintf(names[n], 32, "rate4_%s_n%d", ts ? "ts" : "ss", N);
            v.push_back({names[n++], p});
        }
    return v;
}

int main(int argc, char** argv) {
    auto all = probes();
    if (argc < 2 || !strcmp(argv[1], "list")) {
        for (auto& n : all) printf("%s\n", n.name);
        return 0;
    }
    const Named* sel = nullptr;
    for (auto& n : all) if (!strcmp(n.name, argv[1])) sel = &n;
    if (!sel) { printf("unknown probe %s\n", argv[1]); return 1; }
    Probe p = sel->p;
    const int a_rows = p.M + (p.a_shift_rows ? 24 : 0);
    std::vector<float> A((size_t)a_rows * p.K), B((size_t)p.N * p.K), D((size_t)p.M * p.N, -1.f);
    srand(7);
    for (auto& x : A) x = (float)(rand() % 9 - 4);          // small integers: exact in TF32, exact fp32 sums
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    if (p.repeat > 1) { for (auto& x : A) x *= 0.f; }       // rate probes: keep the accumulator finite
    float *dA, *dB, *dD; long long* dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // the shifted-row probe stages a_rows rows; the kernel computes the row count from M + shift, so stage exactly that
    // shifted_rows: the kernel stages M + shift rows; rows [shift, shift + 128) are the logical A
    probe_kernel<<<1, 128, smem>>>(p, dA, dB, dD, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-24s ERROR %s\n", sel->name, cudaGetErrorString(e)); return 3; }
    long long clk = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&clk, dC, 8, cudaMemcpyDeviceToHost));
    if (p.repeat > 1) {
        const double n_mma = (double)p.repeat * (p.K / 8);
        printf("%-24s %8.1f clk/MMA (M128 N%d K8 tf32, %s; %.0f MMAs, %lld clk) = %.0f MAC/clk/SM\n", sel->name, clk / n_mma, p.N,
               p.a_mode == 2 ? "A in TMEM" : "A in smem", n_mma, clk, 128.0 * p.N * 8 / (clk / n_mma));
        return 0;
    }
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < p.M; ++m)
        for (int n = 0; n < p.N; ++n) {
            double ref = 0;
            for (int k = 0; k < p.K; ++k) ref += (double)A[(size_t)(m + p.a_shift_rows) * p.K + k] * B[(size_t)n * p.K + k];
            const double err = fabs(ref - D[(size_t)m * p.N + n]);
            if (err > maxerr) maxerr = err;
            bad += err > 1e-3;
        }
    printf("%-24s %s  max |err| %.3g, %d / %d wrong  (%lld clk)\n", sel->name, bad ? "FAIL" : "PASS", maxerr, bad, p.M * p.N, clk);
    return bad ? 4 : 0;
}
