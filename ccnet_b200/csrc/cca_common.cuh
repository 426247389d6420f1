// Shared declarations for the criss-cross attention kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cca_b200.h"

namespace cca {

struct Dims {
    int B, Cq, C, H, W;
};

// A "line" is one image row (row branch) or one image column (column branch) of a sample.
// Element (channel c, position j) of the line lives at  c*cs + base + j*sj  inside the sample.
struct Line {
    int L;     // positions on the line (W for a row, H for a column)
    long sj;   // stride between positions (1 for a row, W for a column)
    long base; // offset of position 0, channel 0
    long cs;   // channel stride (H*W)
};

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// host-side launchers (defined in the .cu files, called from cca_capi.cu)
cudaError_t simt_forward(const void *q, const void *k, const void *v, void *out, float *lse, void *ws,
                         Dims d, int dtype, cudaStream_t st, const char **why);
cudaError_t simt_backward(const void *dout, const void *q, const void *k, const void *v, const void *out,
                          const float *lse, void *dq, void *dk, void *dv, void *ws, Dims d, int dtype,
                          cudaStream_t st, const char **why);
bool simt_supported(Dims d, bool backward);

bool tc_forward_supported(Dims d, int dtype);
cudaError_t tc_forward(const void *q, const void *k, const void *v, void *out, float *lse, void *ws,
                       Dims d, int dtype, cudaStream_t st, const char **why);

bool tc_backward_supported(Dims d, int dtype);
cudaError_t tc_backward(const void *dout, const void *q, const void *k, const void *v, const void *out, const float *lse,
                        void *dq, void *dk, void *dv, void *ws, Dims d, int dtype, cudaStream_t st, const char **why);

void count_launch(int n = 1);
void set_tc_debug_buffer(void *p);
void set_tc_two_pass(int on);
void set_tc_bwd_debug_buffer(void *p);
// L2 eviction hints of the tensor-core kernels (CCA_B200_L2HINT = 0/1, CCA_B200_L2KEEP_MB = budget of evict_last data)
int tc_l2_hints();
double tc_l2_keep_mb();
void set_tc_l2_hints(int on, double keep_mb);
// programmatic dependent launch between the passes of one op (CCA_B200_PDL = 0/1, default 1)
int tc_pdl();
void set_tc_pdl(int on);

}  // namespace cca
