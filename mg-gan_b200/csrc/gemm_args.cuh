// Argument block and activation helpers shared by the GEMM kernels behind mggan_linear_* (linear.cu: FP32 tile kernels;
// linear_tc.cu: tcgen05 3 x TF32 kernel).
#pragma once
#include "common.cuh"

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_SIGMOID_EPS = 3 };
constexpr float D_EPS = 1e-7f;      // discriminators.py:110, :203-204

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
    switch (act) {
        case ACT_RELU: return fmaxf(z, 0.f);
        case ACT_LRELU: return z > 0.f ? z : slope * z;
        case ACT_SIGMOID_EPS: return (1.f / (1.f + expf(-z))) * (1.f - 2.f * D_EPS) + D_EPS;
        default: return z;
    }
}
// derivative expressed through the stored output y
__device__ __forceinline__ float act_bwd(float y, int act, float slope) {
    switch (act) {
        case ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case ACT_LRELU: return y > 0.f ? 1.f : slope;
        case ACT_SIGMOID_EPS: {
            float s = (y - D_EPS) / (1.f - 2.f * D_EPS);
            return (1.f - 2.f * D_EPS) * s * (1.f - s);
        }
        default: return 1.f;
    }
}

struct GemmArgs {
    const float* A; long long sam, sak;      // A(m, k) = A[m*sam + k*sak]
    const float* Ay; int act_in;             // optional: multiply A(m,k) by act_in'(Ay(m,k)) (same indexing)
    const float* B; long long sbn, sbk;      // B(n, k) = B[n*sbn + k*sbk]
    float* C; long long scm, scn;            // C(m, n)
    const float* bias;                       // per n (forward)
    float* colsum;                           // per m: sum_k A(m,k)  (db in the weight-gradient call)
    int M, N, K;
    int act; float slope;
    int splitk;                              // >1: K split over blockIdx.z, atomicAdd epilogue
};


// linear_tc.cu: the same GEMM on the tensor cores (variant 3); returns MGGAN_OK, or -1 when the shape is outside what it
// builds (N > 256), in which case the caller launches the FP32 kernel.
int mggan_gemm_tc_launch(const GemmArgs& g, cudaStream_t stream);
