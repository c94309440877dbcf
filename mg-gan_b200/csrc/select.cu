// PM-Network generator selection on the device (reference: MultiGenerator.get_samples
// mggan/model/modules/standard.py:217-225 -- Categorical(logits).sample((k,)) under no_grad --
// and get_selection_indices mggan/utils.py:234-248 + the gather of standard.py:190-214).
//
//  * mggan_gumbel_sample: k categorical draws per agent by Gumbel-max on the logits (Philox
//    counter RNG; distribution-equivalent to the reference's multinomial, SURVEY.md header).
//  * mggan_selection_build: from idx (n_act, k) builds the decoder's work list: every draw
//    becomes one sequence (agent i, generator g, noise sample m = occurrence rank of g among the
//    earlier draws of agent i, output slot (j, i)), sequences are grouped by generator in
//    deterministic (agent-major) order and each group is padded to a multiple of 128 rows (two
//    64-row tiles) so that neither a 64-row decoder tile (decoder.cu) nor a 128-row tensor-core
//    tile (decoder_tc.cu, UMMA M = 128) ever mixes generators.  The reference does this with torch.unique in a
//    Python loop over agents (one host sync per agent).
//  * mggan_selection_all: the all-generators work list used by forward_all / the PM step.
#include "common.cuh"
#include <curand_kernel.h>

namespace {

constexpr int GMAX = 32;
constexpr int TILE = 64;
constexpr int GROUP = 128;     // per-generator padding granularity (rows)

__global__ void rank_kernel(const long long* __restrict__ idx, int n, int k, int G, int* __restrict__ cnt,
                            unsigned char* __restrict__ rank, int* __restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g) c[g] = 0;
    for (int j = 0; j < k; ++j) {
        long long g = idx[(size_t)i * k + j];
        if (g < 0 || g >= G) { atomicExch(err, 1); g = 0; }
        int m = 0;
#pragma unroll
        for (int q = 0; q < GMAX; ++q)
            if (q == (int)g) { m = c[q]; c[q] = m + 1; }
        rank[(size_t)i * k + j] = (unsigned char)m;
    }
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
        if (g < G) cnt[(size_t)g * n + i] = c[g];          // generator-major: coalesced here and in the scan
}

// one CTA per generator: exclusive scan of that generator's per-agent counts (in place), total draws
__global__ void __launch_bounds__(1024)
scan_kernel(int* __restrict__ cnt, int n, int* __restrict__ totals) {
    __shared__ int swarp[32];
    __shared__ int scarry;
    int* c = cnt + (size_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) scarry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const int v = i < n ? c[i] : 0;
        const int carry = scarry;          // written before the barrier that closed the previous round
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) swarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = swarp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            swarp[lane] = wi - w;
            if (lane == 31) scarry = carry + wi;
        }
        __syncthreads();
        if (i < n) c[i] = carry + swarp[warp] + incl - v;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = scarry;
}

// group bases (each generator's rows padded to a multiple of GROUP) and the tile table
__global__ void bases_kernel(const int* __restrict__ totals, int G, int n_tiles, int* __restrict__ base_row,
                             int* __restrict__ tile_gen) {
    __shared__ int sbase[GMAX + 1];
    if (threadIdx.x == 0) {
        int row = 0;
        for (int g = 0; g < G; ++g) {
            sbase[g] = row;
            base_row[g] = row;
            row += (totals[g] + GROUP - 1) / GROUP * GROUP;
        }
        sbase[G] = row;
        base_row[G] = row;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        int r = t * TILE, gsel = -1;
        for (int g = 0; g < G; ++g)
            if (r >= sbase[g] && r < sbase[g + 1]) gsel = g;
        tile_gen[t] = gsel;
    }
}

__global__ void scatter_kernel(const long long* __restrict__ idx, const unsigned char* __restrict__ rank,
                               const int* __restrict__ off, const int* __restrict__ base_row, int n, int k, int G,
                               int* __restrict__ seq_agent, int* __restrict__ seq_noise, int* __restrict__ seq_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = 0; j < k; ++j) {
        long long g = idx[(size_t)i * k + j];
        if (g < 0 || g >= G) g = 0;
        int m = rank[(size_t)i * k + j];
        int row = base_row[g] + off[(size_t)g * n + i] + m;
        seq_agent[row] = i;
        seq_noise[row] = m * n + i;
        seq_out[row] = j * n + i;
    }
}

__global__ void all_kernel(int n, int k, int G, int tiles_per_gen, int* __restrict__ tile_gen,
                           int* __restrict__ seq_agent, int* __restrict__ seq_noise, int* __restrict__ seq_out) {
    const int rows_per_gen = tiles_per_gen * TILE;
    const int total = G * rows_per_gen;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        int g = r / rows_per_gen, q = r - g * rows_per_gen;
        if ((r % TILE) == 0) tile_gen[r / TILE] = g;
        if (q < k * n) {
            int s = q / n, i = q - s * n;
            seq_agent[r] = i;
            seq_noise[r] = s * n + i;
            seq_out[r] = (s * G + g) * n + i;
        } else {
            seq_agent[r] = -1; seq_noise[r] = 0; seq_out[r] = 0;
        }
    }
}

__global__ void gumbel_kernel(const float* __restrict__ logits, int n, int k, int G, unsigned long long seed,
                              unsigned long long offset, const unsigned long long* __restrict__ dyn_offset,
                              long long* __restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (dyn_offset != nullptr) offset += __ldg(dyn_offset);
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i, offset, &st);
    for (int j = 0; j < k; ++j) {
        float best = -INFINITY;
        int arg = 0;
        for (int g = 0; g < G; ++g) {
            float u = curand_uniform(&st);                       // (0, 1]
            float gum = -__logf(-__logf(u) + 1e-20f);
            float v = __ldg(logits + (size_t)i * G + g) + gum;
            if (v > best) { best = v; arg = g; }
        }
        idx[(size_t)i * k + j] = arg;
    }
}

}  // namespace

// upper bound on 64-row tiles (even): every generator group may add up to GROUP - 1 padding rows
extern "C" int mggan_selection_tiles(int n_seq, int G) { return ((n_seq + TILE - 1) / TILE + 2 * G + 1) & ~1; }

// scratch: cnt (n*G int32), rank (n*k bytes), base_row (G+1 int32), err (1 int32, caller zero-fills)
extern "C" int mggan_selection_build(const long long* idx, int n, int k, int G, int n_tiles, int* cnt,
                                     unsigned char* rank, int* base_row, int* err, int* totals, int* tile_gen,
                                     int* seq_agent, int* seq_noise, int* seq_out, cudaStream_t stream) {
    MGGAN_REQUIRE(G >= 1 && G <= GMAX, "mggan_selection_build: num_gens %d not in [1, %d]", G, GMAX);
    MGGAN_REQUIRE(k >= 1 && k <= 255, "mggan_selection_build: num_samples %d not in [1, 255]", k);
    MGGAN_REQUIRE(n_tiles >= mggan_selection_tiles(n * k, G), "mggan_selection_build: tile table too small");
    if (n == 0) {
        cudaMemsetAsync(tile_gen, 0xFF, sizeof(int) * n_tiles, stream);
        cudaMemsetAsync(totals, 0, sizeof(int) * G, stream);
        return mggan_check_launch("selection_build");
    }
    cudaMemsetAsync(seq_agent, 0xFF, sizeof(int) * (size_t)n_tiles * TILE, stream);
    rank_kernel<<<(n + 127) / 128, 128, 0, stream>>>(idx, n, k, G, cnt, rank, err);
    scan_kernel<<<G, 1024, 0, stream>>>(cnt, n, totals);
    bases_kernel<<<1, 256, 0, stream>>>(totals, G, n_tiles, base_row, tile_gen);
    scatter_kernel<<<(n + 127) / 128, 128, 0, stream>>>(idx, rank, cnt, base_row, n, k, G, seq_agent, seq_noise, seq_out);
    return mggan_check_launch("selection_build");
}

extern "C" int mggan_selection_all(int n, int k, int G, int* tile_gen, int* seq_agent, int* seq_noise, int* seq_out,
                                   cudaStream_t stream) {
    MGGAN_REQUIRE(G >= 1 && n >= 0 && k >= 1, "mggan_selection_all: bad arguments");
    int tiles_per_gen = (n * k + GROUP - 1) / GROUP * (GROUP / TILE);
    int total = G * tiles_per_gen * TILE;
    if (total == 0) return MGGAN_OK;
    int grid = (total + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    all_kernel<<<grid, 256, 0, stream>>>(n, k, G, tiles_per_gen, tile_gen, seq_agent, seq_noise, seq_out);
    return mggan_check_launch("selection_all");
}

extern "C" int mggan_gumbel_sample(const float* logits, int n, int k, int G, unsigned long long seed,
                                   unsigned long long offset, const unsigned long long* dyn_offset, long long* idx,
                                   cudaStream_t stream) {
    MGGAN_REQUIRE(G >= 1 && k >= 1, "mggan_gumbel_sample: bad arguments");
    if (n == 0) return MGGAN_OK;
    gumbel_kernel<<<(n + 127) / 128, 128, 0, stream>>>(logits, n, k, G, seed, offset, dyn_offset, idx);
    return mggan_check_launch("gumbel_sample");
}
