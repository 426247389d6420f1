// Microbenchmark 2: does one SM move TMA loads and TMA stores / reduce-adds concurrently, and what is the per-SM rate
// when only a few SMs are active?  One CTA per SM; warp 0 lane 0 streams loads of [112 px][64 ch] fp32 tiles (28 KB) into a
// 4-slot ring (waiting for each on an mbarrier), warp 1 lane 0 streams stores / reduce-adds of 28 KB tiles.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_mix_bench tools/tma_mix_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ccnet_b200/csrc/cca_sm100.cuh"
using namespace sm100;
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int LK = 112, TILE = LK * 128, SLOT = 2 * TILE;

// ld: 1 = loader active; st: 0 none, 1 store, 2 reduce-add
__global__ void __launch_bounds__(64, 1) bench(const __grid_constant__ CUtensorMap min, const __grid_constant__ CUtensorMap mout, int B, int H,
                                               int C, int rounds, int ld, int st, long long *cyc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[4];
    uint8_t *ldbuf = smem, *stbuf = smem + 4 * SLOT;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    for (int i = threadIdx.x; i < 2 * SLOT / 4; i += blockDim.x) reinterpret_cast<float *>(stbuf)[i] = 1.0f;
    fence_proxy_async();
    __syncthreads();
    const int lines = B * H, NCH = C / 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64(), t1 = t0;
    if (warp == 0 && lane == 0 && ld) {
        uint32_t g = 0;
        for (int r = 0; r < rounds; ++r)
            for (int idx = blockIdx.x; idx < lines; idx += gridDim.x) {
                const int b = idx / H, i = idx - b * H;
                for (int n = 0; n < NCH; ++n, ++g) {
                    const int s = g & 3;
                    if (g >= 4) mbar_wait(&bars[s], ((g >> 2) - 1) & 1);      // the load that used this slot 4 tiles ago has landed
                    mbar_expect_tx(&bars[s], SLOT);
                    tma_load_4d(ldbuf + s * SLOT, &min, &bars[s], n * 64, 0, i, b);
                    tma_load_4d(ldbuf + s * SLOT + TILE, &min, &bars[s], n * 64 + 32, 0, i, b);
                }
            }
        for (uint32_t k = g >= 4 ? g - 4 : 0; k < g; ++k) mbar_wait(&bars[k & 3], (k >> 2) & 1);
        t1 = clock64();
        cyc[blockIdx.x] = t1 - t0;
    }
    if (warp == 1 && lane == 0 && st) {
        int slot = 0;
        for (int r = 0; r < rounds; ++r)
            for (int idx = blockIdx.x; idx < lines; idx += gridDim.x) {
                const int b = idx / H, i = idx - b * H;
                for (int n = 0; n < NCH; ++n) {
                    const uint8_t *src = stbuf + (slot & 1) * SLOT;
                    ++slot;
                    if (st == 1) { tma_store_4d(&mout, src, n * 64, 0, i, b); tma_store_4d(&mout, src + TILE, n * 64 + 32, 0, i, b); }
                    else { tma_reduce_add_4d(&mout, src, n * 64, 0, i, b); tma_reduce_add_4d(&mout, src + TILE, n * 64 + 32, 0, i, b); }
                    tma_store_commit();
                    tma_store_wait_read<1>();
                }
            }
        tma_store_wait_all<0>();
        cyc[gridDim.x + blockIdx.x] = clock64() - t0;
    }
}

int main()
{
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
    EncodeFn enc = (EncodeFn)fp;
    const int H = 97, W = 97, C = 512, B = 8;
    float *in, *out;
    cudaMalloc(&in, (size_t)B * H * W * C * 4);
    cudaMalloc(&out, (size_t)B * H * W * C * 4);
    cudaMemset(in, 0, (size_t)B * H * W * C * 4);
    cudaMemset(out, 0, (size_t)B * H * W * C * 4);
    long long *cyc;
    cudaMalloc(&cyc, 2 * 148 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * SLOT);
    auto mk = [&](CUtensorMap *m, void *base, int rows) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        cuuint32_t box[4] = {32u, (cuuint32_t)rows, 1u, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap min, mout;
    mk(&min, in, LK);
    mk(&mout, out, 97);                      // exact-length store boxes, as the kernels use now
    const int grids[3] = {148, 32, 4};
    for (int gi = 0; gi < 3; ++gi)
        for (int ld = 0; ld < 2; ++ld)
            for (int st = 0; st < 3; ++st) {
                if (!ld && !st) continue;
                const int grid = grids[gi], rounds = grid == 148 ? 4 : 1;
                cudaMemset(cyc, 0, 2 * 148 * 8);
                bench<<<grid, 64, 6 * SLOT>>>(min, mout, B, H, C, 1, ld, st, cyc);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                bench<<<grid, 64, 6 * SLOT>>>(min, mout, B, H, C, rounds, ld, st, cyc);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                std::vector<long long> hc(2 * 148);
                cudaMemcpy(hc.data(), cyc, 2 * 148 * 8, cudaMemcpyDeviceToHost);
                long long ml = 0, msx = 0;
                for (int i = 0; i < grid; ++i) { ml = hc[i] > ml ? hc[i] : ml; msx = hc[grid + i] > msx ? hc[grid + i] : msx; }
                const double tiles_per_cta = (double)rounds * B * H * (C / 64) / grid;
                printf("{\"ctas\": %d, \"load\": %d, \"store\": \"%s\", \"ms\": %.4f, \"load_cycles_per_tile\": %.0f, \"store_cycles_per_tile\": %.0f, \"err\": \"%s\"}\n",
                       grid, ld, st == 0 ? "none" : (st == 1 ? "store" : "reduce_add"), ms, ml / tiles_per_cta, msx / tiles_per_cta,
                       cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
