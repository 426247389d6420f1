// Microbenchmark: throughput of TMA tensor stores vs TMA reduce-adds (cp.reduce.async.bulk.tensor .add, fp32) of the
// [112 px][64 ch] fp32 staging tile the criss-cross kernels emit, issued by one lane per CTA, 148 CTAs.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_red_bench tools/tma_red_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ccnet_b200/csrc/cca_sm100.cuh"
using namespace sm100;

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int LK = 112, TILE = LK * 128, SLOT = 2 * TILE;

// mode: 0 store, 1 reduce-add; depth: stores kept in flight before the lane waits for the oldest to be read out of smem
__global__ void __launch_bounds__(128, 1) bench(const __grid_constant__ CUtensorMap map, int B, int H, int C, int rounds, int mode, int depth,
                                                int col, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    for (int i = threadIdx.x; i < 4 * SLOT / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 1.0f;
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int lines = B * H, NCH = C / 64;
    long long t0 = clock64();
    int slot = 0;
    for (int r = 0; r < rounds; ++r)
        for (int idx = blockIdx.x; idx < lines; idx += gridDim.x) {
            const int b = idx / H, i = idx - b * H;
            for (int n = 0; n < NCH; ++n) {
                const uint8_t *src = smem + (slot & 3) * SLOT;
                ++slot;
                const int cw = col ? i : 0, ch = col ? 0 : i;
                if (mode == 0) {
                    tma_store_4d(&map, src, n * 64, cw, ch, b);
                    tma_store_4d(&map, src + TILE, n * 64 + 32, cw, ch, b);
                } else {
                    tma_reduce_add_4d(&map, src, n * 64, cw, ch, b);
                    tma_reduce_add_4d(&map, src + TILE, n * 64 + 32, cw, ch, b);
                }
                tma_store_commit();
                if (depth == 0) tma_store_wait_read<0>();
                else if (depth == 1) tma_store_wait_read<1>();
                else if (depth == 2) tma_store_wait_read<2>();
                else tma_store_wait_read<3>();
            }
        }
    tma_store_wait_all<0>();
    cycles[blockIdx.x] = clock64() - t0;
}

int main()
{
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
    EncodeFn enc = (EncodeFn)fp;
    const int H = 97, W = 97, C = 512, BMAX = 8;
    float *out;
    cudaMalloc(&out, (size_t)BMAX * H * W * C * 4);
    cudaMemset(out, 0, (size_t)BMAX * H * W * C * 4);
    long long *cyc;
    cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SLOT);
    for (int col = 0; col < 2; ++col)
        for (int B : {2, 8}) {
            CUtensorMap m;
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
            cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
            cuuint32_t box[4] = {32u, col ? 1u : (cuuint32_t)LK, col ? (cuuint32_t)LK : 1u, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            for (int mode = 0; mode < 2; ++mode)
                for (int depth : {0, 1, 3}) {
                    const int rounds = B == 2 ? 16 : 4;
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0); cudaEventCreate(&e1);
                    bench<<<148, 128, 4 * SLOT>>>(m, B, H, C, 1, mode, depth, col, cyc);   // warm-up
                    cudaEventRecord(e0);
                    bench<<<148, 128, 4 * SLOT>>>(m, B, H, C, rounds, mode, depth, col, cyc);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    const double bytes = (double)rounds * B * H * 97.0 * C * 4;   // useful bytes (97 of 112 rows land)
                    std::vector<long long> hc(148);
                    cudaMemcpy(hc.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (auto c : hc) mx = c > mx ? c : mx;
                    const double tiles_per_cta = (double)rounds * B * H * (C / 64) / 148.0;
                    printf("{\"lines\": \"%s\", \"B\": %d, \"footprint_MB\": %.0f, \"op\": \"%s\", \"in_flight\": %d, \"ms\": %.4f, \"GBps\": %.0f, "
                           "\"cycles_per_tile_per_cta\": %.0f, \"err\": \"%s\"}\n",
                           col ? "column" : "row", B, B * H * W * C * 4 / 1e6, mode ? "reduce_add" : "store", depth + 1, ms,
                           bytes / ms / 1e6, mx / tiles_per_cta, cudaGetErrorString(cudaGetLastError()));
                }
        }
    return 0;
}
