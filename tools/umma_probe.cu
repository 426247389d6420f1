// Hardware probe for the sm_100a building blocks used by cca_tc_fwd.cu: checks, against a CPU
// computation, (1) tcgen05.mma with K-major x K-major no-swizzle operands written by threads,
// (2) K-major x MN-major, (3) TMEM load mapping, (4) TMA 4-D tiled load with SWIZZLE_128B from an
// NHWC tensor (row box and column box, OOB zero fill) and TMA store.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../ccnet_b200/csrc/cca_sm100.cuh"

using namespace sm100;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// plane layout: [chunk of 8 along the contiguous dim][row][16 B]
__device__ __forceinline__ uint32_t plane_off(int row, int chunk, int rows) { return (uint32_t)(chunk * rows + row) * 16u; }

// A: [128][KA] fp32 row-major; B1: [N1][KA] fp32 row-major (K-major B);  D1 = A * B1^T   [128][N1]
// P: [128][KP] fp32; V: [KP][N2] fp32 (row = k, MN contiguous);           D2 = P * V      [128][N2]
template <int KA, int N1, int KP, int N2>
__global__ void __launch_bounds__(128) umma_probe(const float *A, const float *B1, const float *P, const float *V,
                                                  float *D1, float *D2, float *D3)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sA = smem;                       // (KA/8) planes x 128 rows x 16 B
    uint8_t *sB = sA + (KA / 8) * 128 * 16;   // (KA/8) planes x N1 rows
    uint8_t *sP = sB + (KA / 8) * N1 * 16;    // (KP/8) planes x 128 rows
    uint8_t *sV = sP + (KP / 8) * 128 * 16;   // (N2/8) planes x KP rows
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    // convert + store operands (bf16, plane layout)
    auto put = [](uint8_t *base, int row, int chunk, int rows, const float *src) {
        __nv_bfloat16 v[8];
        for (int i = 0; i < 8; ++i) v[i] = __float2bfloat16_rn(src[i]);
        *reinterpret_cast<uint4 *>(base + plane_off(row, chunk, rows)) = *reinterpret_cast<uint4 *>(v);
    };
    for (int t = tid; t < 128 * (KA / 8); t += 128) { int r = t % 128, c = t / 128; put(sA, r, c, 128, A + r * KA + c * 8); }
    for (int t = tid; t < N1 * (KA / 8); t += 128) { int r = t % N1, c = t / N1; put(sB, r, c, N1, B1 + r * KA + c * 8); }
    for (int t = tid; t < 128 * (KP / 8); t += 128) { int r = t % 128, c = t / 128; put(sP, r, c, 128, P + r * KP + c * 8); }
    for (int t = tid; t < KP * (N2 / 8); t += 128) { int r = t % KP, c = t / KP; put(sV, r, c, KP, V + r * N2 + c * 8); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        // D1: K-major A (lbo = plane stride, sbo = 128), K-major B
        constexpr uint32_t id1 = instr_desc(kFmtBF16, kFmtBF16, 128, N1, false, false);
        for (int ks = 0; ks < KA / 16; ++ks) {
            uint64_t ad = smem_desc(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
            uint64_t bd = smem_desc(smem_u32(sB) + ks * 2 * N1 * 16, N1 * 16, 128);
            mma_f16(tmem, ad, bd, id1, ks > 0);
        }
        // D2: K-major A = P, MN-major B = V (planes along MN: sbo = plane stride; k groups of 8 rows: lbo = 128)
        constexpr uint32_t id2 = instr_desc(kFmtBF16, kFmtBF16, 128, N2, false, true);
        for (int ks = 0; ks < KP / 16; ++ks) {
            uint64_t ad = smem_desc(smem_u32(sP) + ks * 2 * 128 * 16, 128 * 16, 128);
            uint64_t bd = smem_desc(smem_u32(sV) + ks * 16 * 16, 128, KP * 16);
            mma_f16(tmem + 128, ad, bd, id2, ks > 0);
        }
    }
    // D3 = P * V with A = P held in TMEM (bf16 pairs, low half = even k), B = V planes (MN-major) from smem
    {
        const uint32_t lane_b = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < KP / 2; c0 += 8) {
            uint32_t w[8];
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 v2 = __floats2bfloat162_rn(P[tid * KP + 2 * (c0 + i)], P[tid * KP + 2 * (c0 + i) + 1]);
                w[i] = *reinterpret_cast<uint32_t *>(&v2);
            }
            tmem_st8(tmem + 192 + lane_b + c0, w);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        constexpr uint32_t id3 = instr_desc(kFmtBF16, kFmtBF16, 128, N2, false, true);
        for (int ks = 0; ks < KP / 16; ++ks) {
            uint64_t bd = smem_desc(smem_u32(sV) + ks * 16 * 16, 128, KP * 16);
            mma_f16_ts(tmem + 320, tmem + 192 + ks * 8, bd, id3, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t r[16];
    for (int c0 = 0; c0 < N1; c0 += 16) {
        tmem_ld16(tmem + lane_base + c0, r);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D1[tid * N1 + c0 + i] = __uint_as_float(r[i]);
    }
    for (int c0 = 0; c0 < N2; c0 += 16) {
        tmem_ld16(tmem + 128 + lane_base + c0, r);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D2[tid * N2 + c0 + i] = __uint_as_float(r[i]);
    }
    for (int c0 = 0; c0 < N2; c0 += 16) {
        tmem_ld16(tmem + 320 + lane_base + c0, r);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D3[tid * N2 + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// SWIZZLE_128B operands exactly as a bf16 TMA tile lands them: [rows][64 bf16 = 128 B], 16-byte chunk c of row r at c ^ (r & 7).
//   D4 = A * B^T : A [128][64] and B [112][64] as K-major SW128 tiles (k-step = +32 B inside the swizzled row)
//   D5 = P * V   : P no-swizzle K-major planes (A), V [112 rows = k][64 = mn] as an MN-major SW128 tile (k-step = +2048 B)
__global__ void __launch_bounds__(128) umma_sw128_probe(const float *A, const float *B1, const float *P, const float *V,
                                                        float *D4, float *D5)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sA = smem;                 // 128 rows x 128 B
    uint8_t *sB = sA + 128 * 128;       // 112 rows x 128 B
    uint8_t *sV = sB + 112 * 128;       // 112 rows x 128 B  (+ pad to keep 1024 alignment)
    uint8_t *sP = sV + 112 * 128;       // 14 planes x 128 rows x 16 B
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    auto put_sw = [](uint8_t *tile, int row, int chunk, const float *src) {
        __nv_bfloat16 v[8];
        for (int i = 0; i < 8; ++i) v[i] = __float2bfloat16_rn(src[i]);
        *reinterpret_cast<uint4 *>(tile + row * 128 + ((chunk ^ (row & 7)) * 16)) = *reinterpret_cast<uint4 *>(v);
    };
    for (int t = tid; t < 128 * 8; t += 128) put_sw(sA, t / 8, t % 8, A + (t / 8) * 64 + (t % 8) * 8);
    for (int t = tid; t < 112 * 8; t += 128) put_sw(sB, t / 8, t % 8, B1 + (t / 8) * 64 + (t % 8) * 8);
    for (int t = tid; t < 112 * 8; t += 128) put_sw(sV, t / 8, t % 8, V + (t / 8) * 64 + (t % 8) * 8);
    for (int t = tid; t < 128 * 14; t += 128) {
        const int r = t % 128, c = t / 128;
        __nv_bfloat16 v[8];
        for (int i = 0; i < 8; ++i) v[i] = __float2bfloat16_rn(P[r * 112 + c * 8 + i]);
        *reinterpret_cast<uint4 *>(sP + plane_off(r, c, 128)) = *reinterpret_cast<uint4 *>(v);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        constexpr uint32_t id4 = instr_desc(kFmtBF16, kFmtBF16, 128, 112, false, false);
        for (int ks = 0; ks < 4; ++ks)
            mma_f16(tmem, smem_desc(smem_u32(sA) + ks * 32, 16, 1024, 2), smem_desc(smem_u32(sB) + ks * 32, 16, 1024, 2), id4, ks > 0);
        constexpr uint32_t id5 = instr_desc(kFmtBF16, kFmtBF16, 128, 64, false, true);
        for (int ks = 0; ks < 7; ++ks)
            mma_f16(tmem + 128, smem_desc(smem_u32(sP) + ks * 2 * 128 * 16, 128 * 16, 128), smem_desc(smem_u32(sV) + ks * 2048, 16, 1024, 2), id5, ks > 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t r[16];
    for (int c0 = 0; c0 < 112; c0 += 16) {
        tmem_ld16(tmem + lane_base + c0, r); tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D4[tid * 112 + c0 + i] = __uint_as_float(r[i]);
    }
    for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld16(tmem + 128 + lane_base + c0, r); tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D5[tid * 64 + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

// TMA probe: load a row box and a column box of an NHWC fp32 tensor with SWIZZLE_128B, dump raw smem, store back.
__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap map_row, const __grid_constant__ CUtensorMap map_col,
                                                 const __grid_constant__ CUtensorMap map_out, float *dump_row, float *dump_col,
                                                 int c0, int h, int w)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    float *srow = reinterpret_cast<float *>(smem);               // [104][32]
    float *scol = reinterpret_cast<float *>(smem + 104 * 128);   // [104][32]  (13312 B = 13 * 1024: still 1024-aligned)
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1); fence_mbar_init();
        mbar_expect_tx(&bar, 2 * 104 * 128);
        tma_load_4d(srow, &map_row, &bar, c0, 0, h, 0);
        tma_load_4d(scol, &map_col, &bar, c0, w, 0, 0);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = tid; i < 104 * 32; i += 128) { dump_row[i] = srow[i]; dump_col[i] = scol[i]; }
    __syncthreads();
    if (tid == 0) {
        tma_store_4d(&map_out, srow, c0, 0, h, 0);
        tma_store_commit();
        tma_store_wait_all<0>();
        tma_reduce_add_4d(&map_out, srow, c0, 0, h, 0);      // Y += tile  -> Y == 2 * X on the tile
        tma_store_commit();
        tma_store_wait_all<0>();
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    // ---------------- UMMA
    constexpr int KA = 64, N1 = 112, KP = 112, N2 = 64;
    std::vector<float> A(128 * KA), B1(N1 * KA), P(128 * KP), V(KP * N2);
    srand(1);
    auto rnd = [] { return (float)(rand() % 2001 - 1000) / 1000.f; };
    for (auto &x : A) x = rnd(); for (auto &x : B1) x = rnd(); for (auto &x : P) x = rnd(); for (auto &x : V) x = rnd();
    float *dA, *dB1, *dP, *dV, *dD1, *dD2, *dD3;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB1, B1.size() * 4)); CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dV, V.size() * 4));
    CK(cudaMalloc(&dD1, 128 * N1 * 4)); CK(cudaMalloc(&dD2, 128 * N2 * 4)); CK(cudaMalloc(&dD3, 128 * N2 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB1, B1.data(), B1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dV, V.data(), V.size() * 4, cudaMemcpyHostToDevice));
    size_t smem = (KA / 8) * 128 * 16 + (KA / 8) * N1 * 16 + (KP / 8) * 128 * 16 + (N2 / 8) * KP * 16;
    CK(cudaFuncSetAttribute(umma_probe<KA, N1, KP, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe<KA, N1, KP, N2><<<1, 128, smem>>>(dA, dB1, dP, dV, dD1, dD2, dD3);
    CK(cudaDeviceSynchronize());
    std::vector<float> D1(128 * N1), D2(128 * N2), D3(128 * N2);
    CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    auto bf = [](float x) { return __bfloat162float(__float2bfloat16_rn(x)); };
    double e1 = 0, e2 = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N1; ++n) {
        double s = 0; for (int k = 0; k < KA; ++k) s += (double)bf(A[m * KA + k]) * bf(B1[n * KA + k]);
        e1 = fmax(e1, fabs(s - D1[m * N1 + n]));
    }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N2; ++n) {
        double s = 0; for (int k = 0; k < KP; ++k) s += (double)bf(P[m * KP + k]) * bf(V[k * N2 + n]);
        e2 = fmax(e2, fabs(s - D2[m * N2 + n]));
    }
    {
        float *dD4, *dD5;
        CK(cudaMalloc(&dD4, 128 * 112 * 4)); CK(cudaMalloc(&dD5, 128 * 64 * 4));
        const size_t sm2 = 128 * 128 + 112 * 128 * 2 + 14 * 128 * 16 + 2048;
        CK(cudaFuncSetAttribute(umma_sw128_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        umma_sw128_probe<<<1, 128, sm2>>>(dA, dB1, dP, dV, dD4, dD5);
        CK(cudaDeviceSynchronize());
        std::vector<float> D4(128 * 112), D5(128 * 64);
        CK(cudaMemcpy(D4.data(), dD4, D4.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D5.data(), dD5, D5.size() * 4, cudaMemcpyDeviceToHost));
        double e4 = 0, e5 = 0;
        for (size_t i = 0; i < D4.size(); ++i) e4 = fmax(e4, fabs((double)D4[i] - D1[i]));
        for (size_t i = 0; i < D5.size(); ++i) e5 = fmax(e5, fabs((double)D5[i] - D2[i]));
        printf("UMMA SW128 K-major x K-major (TMA-style bf16 tiles) vs no-swizzle result: max abs diff %.3e  %s\n", e4, e4 < 1e-3 ? "OK" : "MISMATCH");
        printf("UMMA K-major planes x SW128 MN-major tile          vs no-swizzle result: max abs diff %.3e  %s\n", e5, e5 < 1e-3 ? "OK" : "MISMATCH");
    }
    double e3 = 0;
    for (size_t i = 0; i < D3.size(); ++i) e3 = fmax(e3, fabs((double)D3[i] - D2[i]));
    printf("UMMA A-from-TMEM (tcgen05.st bf16 pairs, low half = even k) vs SS result: max abs diff %.3e  %s\n", e3, e3 < 1e-3 ? "OK" : "MISMATCH");
    printf("UMMA K-major x K-major  (M128 N%d K%d): max abs err %.3e  %s\n", N1, KA, e1, e1 < 1e-3 ? "OK" : "MISMATCH");
    printf("UMMA K-major x MN-major (M128 N%d K%d): max abs err %.3e  %s\n", N2, KP, e2, e2 < 1e-3 ? "OK" : "MISMATCH");

    // ---------------- TMA
    const int B = 2, H = 100, W = 97, C = 64;
    std::vector<float> X((size_t)B * H * W * C);
    for (size_t i = 0; i < X.size(); ++i) X[i] = (float)(i % 100003) * 0.001f;
    float *dX, *dY, *dr, *dc;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dY, X.size() * 4)); CK(cudaMemset(dY, 0, X.size() * 4));
    CK(cudaMalloc(&dr, 104 * 32 * 4)); CK(cudaMalloc(&dc, 104 * 32 * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qr));
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    cuuint32_t box_row[4] = {32, 104, 1, 1}, box_col[4] = {32, 1, 104, 1};
    CUtensorMap mrow, mcol, mout;
    CUresult r1 = encode(&mrow, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dX, dims, strides, box_row, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&mcol, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dX, dims, strides, box_col, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r3 = encode(&mout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dY, dims, strides, box_row, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode results: %d %d %d\n", (int)r1, (int)r2, (int)r3);
    const int c0 = 32, h = 7, w = 5;
    CK(cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 104 * 128 + 1024));
    tma_probe<<<1, 128, 2 * 104 * 128 + 1024>>>(mrow, mcol, mout, dr, dc, c0, h, w);
    CK(cudaDeviceSynchronize());
    std::vector<float> hr(104 * 32), hc(104 * 32), Y(X.size());
    CK(cudaMemcpy(hr.data(), dr, hr.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc.data(), dc, hc.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost));
    int bad_r = 0, bad_c = 0, bad_s = 0;
    for (int r = 0; r < 104; ++r) for (int j = 0; j < 8; ++j) for (int e = 0; e < 4; ++e) {
        const int phys = (j ^ (r & 7)) * 4 + e;        // SWIZZLE_128B: 16-byte chunk index XOR (row % 8)
        float er = r < W ? X[(((size_t)0 * H + h) * W + r) * C + c0 + j * 4 + e] : 0.f;   // row box: rows = w
        float ec = r < H ? X[(((size_t)0 * H + r) * W + w) * C + c0 + j * 4 + e] : 0.f;   // col box: rows = h
        if (hr[r * 32 + phys] != er) ++bad_r;
        if (hc[r * 32 + phys] != ec) ++bad_c;
    }
    for (size_t i = 0; i < Y.size(); ++i) {
        size_t c = i % C, ww = (i / C) % W, hh = (i / C / W) % H, bb = i / C / W / H;
        const bool inc = c >= (size_t)c0 && c < (size_t)c0 + 32 && bb == 0;
        float e = 0.f;
        if (inc && hh == (size_t)h) e += 2.f * X[i];             // store + reduce-add of the row tile
        (void)ww;
        if (Y[i] != e) ++bad_s;
    }
    printf("TMA row box  SW128 + OOB zero fill : %s (%d mismatches)\n", bad_r ? "MISMATCH" : "OK", bad_r);
    printf("TMA col box  SW128 + OOB zero fill : %s (%d mismatches)\n", bad_c ? "MISMATCH" : "OK", bad_c);
    printf("TMA store + reduce-add (swizzled smem -> NHWC, OOB rows clipped): %s (%d mismatches)\n", bad_s ? "MISMATCH" : "OK", bad_s);
    return 0;
}
