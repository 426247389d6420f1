// Pieces shared by the tcgen05 forward and backward kernels (channels-last, fp32 I/O, bf16x3 split).
#pragma once
#include <mutex>

#include "cca_common.cuh"
#include "cca_sm100.cuh"

namespace cca {
namespace tc {
using namespace sm100;

constexpr int kNC = 64;   // channels per chunk (one ring slot = [LK px][64 ch] fp32 = two swizzled TMA tiles)
constexpr uint32_t kSw128 = 2;   // UMMA smem-descriptor layout type SWIZZLE_128B: a bf16 TMA tile [rows][128 B] is directly a
                                 // K-major operand (rows = M/N, k-step = +32 B) or an MN-major one (rows = K, k-step = +2048 B)

// Thread layout of the kernels (seven warpgroups, registers rebalanced with setmaxnreg):
//   warps 0-3   (128 thr) : epilogue group        (TMEM lane == pixel; accumulators -> staging)
//   warps 4-7   (128 thr) : softmax / P / dS group (TMEM lane == pixel)
//   warps 8-23  (512 thr) : converters -- fp32 staging tile -> bf16 hi/lo operand planes.  The conversion is ~1200 warp
//                           instructions per 28 KB slot and a slot used to take ~1000 cycles with 8 warps (two per scheduler,
//                           mostly waiting on their own dependent instructions), which paced every fp32 chunk; 16 warps
//                           halve the work per warp and double the warps a scheduler can pick from.
//   warp 24               : TMA producer (one elected lane)
//   warp 25               : MMA issuer (whole warp converged, tcgen05.mma under elect.sync)
//   warps 26, 27          : store warps (one lane each, alternating items): staging slots -> global, counters
constexpr int kThreads = 896;
constexpr int kConvThreads = 512;
constexpr int kWarpConv0 = 8, kWarpProducer = 24, kWarpMma = 25, kWarpStore = 26;
constexpr int kRegsLaunch = 72;           // 65536 / 896 rounded down to a multiple of 8
constexpr int kRegsMisc = 56;
// setmaxnreg moves registers through a per-CTA pool that only holds what the CTA itself released: the increases must be
// covered by the decreases relative to the launch allocation (kRegsLaunch per thread).
constexpr bool reg_pool_ok(int soft, int epi, int conv)
{
    return kConvThreads * (kRegsLaunch - conv) + 128 * (kRegsLaunch - kRegsMisc) >= 128 * (soft - kRegsLaunch) + 128 * (epi - kRegsLaunch);
}

template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// BF = false: fp32 I/O, every operand split into bf16 hi + lo (3 MMAs per product).
// BF = true : bf16 I/O, operands used as they are (1 MMA per product); a 64-channel chunk is ONE 128-byte-wide TMA tile.
template <int LK, bool BF = false> struct Tiles {
    static constexpr int kTile = LK * 128;             // [LK px][128 B] = 32 fp32 or 64 bf16 channels, SWIZZLE_128B
    static constexpr int kSlot = BF ? kTile : 2 * kTile; // 64 channels
    static constexpr int kPlane = LK * 16;             // operand plane: LK rows x 16 B (8 bf16)
    // fp32 I/O: a converted slot holds its 8-channel planes as  hi0 lo0 hi1 lo1 ... hi7 lo7  (each 32-channel TMA box is rewritten
    // inside its own bytes, so the two boxes of a slot are converted by two independent thread groups)
    static constexpr int kPStride = 2 * kPlane;        // hi plane p -> hi plane p+1 (and lo -> lo)
    static constexpr int kLoOff = kPlane;              // hi plane p -> lo plane p
    static constexpr int kTerms = BF ? 1 : 2;          // operand copies kept: hi (+ lo)
    static constexpr int kOp = kTerms * 8 * kPlane;    // 8 planes = 64 channels, per copy
    static constexpr int kPP = LK / 8;                 // planes of a [LK x LK] pixel-pixel matrix (P, dS)
    static constexpr int kP = kTerms * kPP * kPlane;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // low half = a, high half = b
    return r;
}
// (a, b) -> packed bf16 hi pair and lo pair, x = hi + lo up to ~2^-17 |x|
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    hi = pack_bf16(a, b);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xFFFF0000u);
    lo = pack_bf16(a - ah, b - bh);
}
__device__ __forceinline__ void split8(const float *v, uint4 &hi, uint4 &lo)
{
    split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred) : "r"(0xFFFFFFFFu));
    return pred != 0;
}

// Converter: one staged slot ([LK px][64 ch] fp32 as two swizzled 32-channel TMA boxes) -> bf16 hi/lo operand planes
// [8-channel chunk][pixel][16 B], IN PLACE.  Each box is rewritten inside its own bytes (planes hi0 lo0 .. hi3 lo3 of its 32
// channels take exactly the box's LK x 128 B), so the two boxes are handled by two independent groups of 256 threads (group =
// t >> 8; inside a group thread = (pixel row, 16-channel half)): every thread reads and splits its 64 B, the group meets on its
// own named barrier (1 or 3), then overwrites.  `mid` runs after the loads have been issued and before the group barrier.
struct ConvertNoMid { __device__ __forceinline__ void operator()() const {} };
template <int LK, typename Mid = ConvertNoMid>
__device__ __forceinline__ void convert_slot_inplace(uint8_t *slot, int t, Mid mid = Mid())
{
    using T = Tiles<LK, false>;
    const int grp = t >> 8, r = t & 127, hq = (t >> 7) & 1;
    uint8_t *box = slot + grp * T::kTile;
    float4 raw[4];                                              // the thread's 16 channels
    {
        const int rr = r < LK ? r : LK - 1;                     // idle threads re-read the last row (unconditional loads keep raw[] in registers)
        const uint8_t *src = box + rr * 128;
        const int sw = rr & 7;
#pragma unroll
        for (int j = 0; j < 4; ++j) raw[j] = *reinterpret_cast<const float4 *>(src + (((hq * 4 + j) ^ sw) * 16));
    }
    mid();
    if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
    else asm volatile("bar.sync 3, 256;" ::: "memory");
    if (r < LK) {
        uint8_t *d = box + r * 16 + hq * 2 * T::kPStride;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float4 a = raw[2 * j], b = raw[2 * j + 1];
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint4 hi, lo;
            split8(v, hi, lo);
            *reinterpret_cast<uint4 *>(d + j * T::kPStride) = hi;
            *reinterpret_cast<uint4 *>(d + j * T::kPStride + T::kLoOff) = lo;
        }
    }
}

// D (+)= A*B with the bf16x3 split: Ah*Bh + Ah*Bl + Al*Bh (TERMS = 3), or just Ah*Bh for bf16 I/O (TERMS = 1).
// Executed by the whole (converged) MMA warp.
template <int TERMS = 3>
__device__ __forceinline__ void mma_split3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl,
                                           uint32_t idesc, bool accumulate)
{
    if (elect_one()) {
        mma_f16(d, ah, bh, idesc, accumulate);
        if constexpr (TERMS == 3) {
            mma_f16(d, ah, bl, idesc, true);
            mma_f16(d, al, bh, idesc, true);
        }
    }
    __syncwarp();
}
// Same, for a whole K loop: NK k-steps, descriptors advance by (a_step, b_step) bytes per step.  One election.
template <int NK, int TERMS = 3>
__device__ __forceinline__ void mma_split3_loop(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t a_step, uint32_t a_lbo, uint32_t a_sbo,
                                                uint32_t b_hi, uint32_t b_lo, uint32_t b_step, uint32_t b_lbo, uint32_t b_sbo,
                                                uint32_t idesc, bool accumulate_first, uint32_t a_layout = 0, uint32_t b_layout = 0)
{
    if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            const uint64_t ah = smem_desc(a_hi + ks * a_step, a_lbo, a_sbo, a_layout), al = smem_desc(a_lo + ks * a_step, a_lbo, a_sbo, a_layout);
            const uint64_t bh = smem_desc(b_hi + ks * b_step, b_lbo, b_sbo, b_layout), bl = smem_desc(b_lo + ks * b_step, b_lbo, b_sbo, b_layout);
            mma_f16(d, ah, bh, idesc, accumulate_first || ks > 0);
            if constexpr (TERMS == 3) {
                mma_f16(d, ah, bl, idesc, true);
                mma_f16(d, al, bh, idesc, true);
            }
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void commit_to(uint64_t *bar)
{
    if (elect_one()) mma_commit(bar);
    __syncwarp();
}

// ---- host: TMA tensor maps over channels-last fp32 tensors -------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeFn get_encode()
{
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaDriverEntryPointQueryResult qr;
        void *p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    });
    return fn;
}
// NHWC tensor [B,H,W,C] (fp32 or bf16); box = [128 bytes of channels] x [LK pixels along W (row pass) or H (column pass)]
inline bool make_map(CUtensorMap *m, const void *base, int B, int H, int W, int C, int LK, bool col, bool bf16 = false)
{
    EncodeFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t es_bytes = bf16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * es_bytes, (cuuint64_t)W * C * es_bytes, (cuuint64_t)H * W * C * es_bytes};
    cuuint32_t box[4] = {bf16 ? 64u : 32u, col ? 1u : (cuuint32_t)LK, col ? (cuuint32_t)LK : 1u, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// LK template (padded tile length) for the longest tile of a problem
inline int lk_for(int tile) { return tile <= 80 ? 80 : (tile <= 112 ? 112 : 0); }
// Cached tensor maps (cca_tc_host.cu): encoding costs ~1 us of driver time per map and an op call needs 8-15 of them;
// a map only depends on (base, shape, box, dtype), so it stays valid for as long as that address holds such a tensor.
bool get_map(CUtensorMap *m, const void *base, int B, int H, int W, int C, int LK, bool col, bool bf16);
// SM count of the CURRENT device (cached per device id)
int sm_count();
inline bool shape_supported(Dims d, int dtype)
{
    if (dtype != CCA_F32 && dtype != CCA_BF16) return false;
    if (d.Cq % 16 != 0 || d.Cq > 64 || d.Cq < 16 || d.C % kNC != 0) return false;
    if (d.H > 112 * 8 || d.W > 112 * 8) return false;        // cca_items.cuh: at most kMaxNT tiles of kMaxTile pixels per line
    return get_encode() != nullptr;
}

}  // namespace tc
}  // namespace cca
