// tcgen05 / TMA forward kernel of criss-cross attention for sm_100a (channels-last tensors).
//
// Layout: q,k [B,H,W,Cq], v,out [B,H,W,C] (torch channels_last).  In this layout an image row and an
// image column are the same object -- L pixels with a fixed pixel stride, each pixel's channels
// contiguous -- so ONE kernel serves both branches of cc_attention/functions.py:38-47:
//   pass 1 (columns, self entry masked, functions.py:38): out <- V P_c / l_c, stats <- (m_c, l_c)
//   pass 2 (rows, functions.py:39): flash-style merge with pass 1 -> out, lse       (functions.py:40-47)
//
// One persistent CTA per SM walks over lines.  Per line (L <= LK pixels, padded to LK):
//   TMA producer (1 thread)   : 4-D tiled loads [LK px][32 ch] fp32, SWIZZLE_128B, OOB pixels zero-filled,
//                               into a 2-slot ring: Q, K of the NEXT line, then the V chunks (64 channels each)
//                               of the current one (software pipeline across lines).
//   converter warps (256 thr) : fp32 -> bf16 hi + bf16 lo split (x = hi + lo to ~2^-17), written as UMMA
//                               canonical no-swizzle operand planes [8-channel chunk][pixel][16 B].
//   MMA warp (elect.sync)     : S = Q K^T as 3 bf16 MMAs per k-step (hi*hi + hi*lo + lo*hi, fp32 accumulate
//                               in TMEM, M=128 N=LK K=16), then per V chunk O = P V (M=128 N=64, K = pixels).
//   softmax group (128 thr)   : TMEM -> registers (one query pixel per thread), exp2-based softmax, P split
//                               hi/lo into K-major operand planes, per-pixel scales / stats / lse.
//   epilogue group (128 thr)  : per V chunk TMEM -> scale/merge -> swizzled smem tile -> TMA store.
// All inter-role hand-offs are mbarriers (TMA complete_tx, tcgen05.commit, thread arrives).
#include "cca_tc_common.cuh"

namespace cca {
namespace {
using namespace tc;

constexpr int kTmemCols = 256;      // S: [0,128)  O0: [128,192)  O1: [192,256)
constexpr int kRegsSoft = 168, kRegsEpi = 128;
static_assert(reg_pool_ok(kRegsSoft, kRegsEpi), "setmaxnreg pool");

struct FwdParams {
    int B, H, W, C, Cq;
    int L;        // pixels per line (H for the column pass, W for the row pass)
    int NL;       // lines per sample
    int col;      // 1: column pass (mask self, write partial + stats), 0: row pass (merge, write out + lse)
    float2 *stats;
    float *lse;
    long long *dbg;   // optional timeline buffer (4 roles x 512 stamps), CTA 0 only; nullptr in production
};

#define CCA_STAMP(role)                                                                          \
    do {                                                                                         \
        if (p.dbg && blockIdx.x == 0 && dbg_n < 512) p.dbg[(role) * 512 + dbg_n++] = clock64();  \
    } while (0)

template <int LK> struct FwdSmem {
    using T = Tiles<LK>;
    static constexpr int off_ld = 0;                      // 2 load slots
    static constexpr int off_out = off_ld + 2 * T::kSlot; // 2 out slots
    static constexpr int off_op = off_out + 2 * T::kSlot; // 2 operand buffers
    static constexpr int off_p = off_op + 2 * T::kOp;
    static constexpr int off_tail = off_p + T::kP;        // 256 B pad: M=128 MMAs read 16 rows past LK rows
    static constexpr int off_scale = off_tail + 256;      // float2 (sa, sb) [3][128]: softmax group -> epilogue group
    static constexpr int off_bar = off_scale + 3 * 128 * 8;
    static constexpr int kBytes = off_bar + 256 + 1024;   // + alignment slack
};

enum { B_LD_FULL = 0, B_LD_EMPTY = 2, B_OP_FULL = 4, B_OP_EMPTY = 6, B_S_FULL = 8, B_S_EMPTY = 9, B_P_FULL = 10,
       B_P_EMPTY = 11, B_O_FULL = 12, B_O_EMPTY = 14, B_OUT_FULL = 16, B_COUNT = 18 };

// Software pipeline across the lines k = 0..nk-1 of this CTA: the ring carries
//     Q0 K0 | Q1 K1 V0[0..NCH) | Q2 K2 V1[0..NCH) | ...
// so S(k+1) and its softmax run while P(k) V(k) is being accumulated and stored.
template <int LK>
__global__ void __launch_bounds__(kThreads, 1)
cca_tc_fwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                  const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mo, FwdParams p)
{
    using T = Tiles<LK>;
    using S = FwdSmem<LK>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::off_bar + 8 * B_COUNT);
    float2 *scale = reinterpret_cast<float2 *>(smem + S::off_scale);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NCH = p.C / kNC;
    const int KQ = p.Cq / 16;                 // k-steps of the S MMA
    const int total_lines = p.B * p.NL;
    const int nk = (total_lines - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // lines of this CTA

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[B_LD_FULL + i], 1);            mbar_init(&bars[B_LD_EMPTY + i], kConvThreads);
            mbar_init(&bars[B_OP_FULL + i], kConvThreads); mbar_init(&bars[B_OP_EMPTY + i], 1);
            mbar_init(&bars[B_O_FULL + i], 1);             mbar_init(&bars[B_O_EMPTY + i], 128);
            mbar_init(&bars[B_OUT_FULL + i], 1);
        }
        mbar_init(&bars[B_S_FULL], 1); mbar_init(&bars[B_S_EMPTY], 128);
        mbar_init(&bars[B_P_FULL], 128); mbar_init(&bars[B_P_EMPTY], 1);
        fence_mbar_init();
        prefetch_tmap(&mq); prefetch_tmap(&mk); prefetch_tmap(&mv); prefetch_tmap(&mo);
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // line -> TMA coordinates (c, w, h, b) of pixel 0 of the line
    auto line_coords = [&](int line, int &cw, int &ch, int &cb) {
        cb = line / p.NL;
        const int i = line - cb * p.NL;
        if (p.col) { cw = i; ch = 0; } else { cw = 0; ch = i; }
    };
    auto line_of = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };

    if (warp >= kWarpProducer) {
        reg_dec<kRegsMisc>();
        if (warp == kWarpProducer) {
            // =============================== TMA producer ===============================
            if (lane == 0) {
                uint32_t g = 0;
                int dbg_n = 0;
                auto emit = [&](const CUtensorMap *m, int c0, int line) {
                    int cw, ch, cb;
                    line_coords(line, cw, ch, cb);
                    const int slot = g & 1;
                    mbar_wait(&bars[B_LD_EMPTY + slot], ((g >> 1) & 1) ^ 1);
                    CCA_STAMP(0);
                    uint8_t *dst = smem + S::off_ld + slot * T::kSlot;
                    mbar_expect_tx(&bars[B_LD_FULL + slot], T::kSlot);
                    tma_load_4d(dst, m, &bars[B_LD_FULL + slot], c0, cw, ch, cb);
                    tma_load_4d(dst + T::kTile, m, &bars[B_LD_FULL + slot], c0 + 32, cw, ch, cb);
                    ++g;
                };
                emit(&mq, 0, line_of(0));
                emit(&mk, 0, line_of(0));
                for (int k = 0; k < nk; ++k) {
                    if (k + 1 < nk) { emit(&mq, 0, line_of(k + 1)); emit(&mk, 0, line_of(k + 1)); }
                    for (int n = 0; n < NCH; ++n) emit(&mv, n * kNC, line_of(k));
                }
            }
        } else if (warp == kWarpMma) {
            // =============================== MMA issuer (whole warp, elect.sync inside) ===============================
            const uint32_t idesc_s = instr_desc(kFmtBF16, kFmtBF16, 128, LK, false, false);
            const uint32_t idesc_o = instr_desc(kFmtBF16, kFmtBF16, 128, kNC, false, true);
            const uint32_t op_base = smem_u32(smem + S::off_op), p_base = smem_u32(smem + S::off_p);
            uint32_t u = 0, oc = 0;
            int dbg_n = lane == 0 ? 0 : 512;
            auto issue_s = [&](int k) {            // S(k) = Q K^T, Q in operand buffer u&1, K in (u+1)&1
                const uint32_t qb = op_base + (u & 1) * T::kOp, kb = op_base + ((u + 1) & 1) * T::kOp;
                mbar_wait(&bars[B_OP_FULL + (u & 1)], (u >> 1) & 1);
                mbar_wait(&bars[B_OP_FULL + ((u + 1) & 1)], ((u + 1) >> 1) & 1);
                mbar_wait(&bars[B_S_EMPTY], (k & 1) ^ 1);
                tc_fence_after();
                CCA_STAMP(2);
                for (int ks = 0; ks < KQ; ++ks) {
                    const uint32_t ao = ks * 2 * T::kPlane;
                    mma_split3(tmem, smem_desc(qb + ao, T::kPlane, 128), smem_desc(qb + 8 * T::kPlane + ao, T::kPlane, 128),
                               smem_desc(kb + ao, T::kPlane, 128), smem_desc(kb + 8 * T::kPlane + ao, T::kPlane, 128),
                               idesc_s, ks > 0);
                }
                commit_to(&bars[B_S_FULL]);
                commit_to(&bars[B_OP_EMPTY + (u & 1)]);
                commit_to(&bars[B_OP_EMPTY + ((u + 1) & 1)]);
                u += 2;
            };
            issue_s(0);
            for (int k = 0; k < nk; ++k) {
                if (k + 1 < nk) issue_s(k + 1);
                CCA_STAMP(2);
                mbar_wait(&bars[B_P_FULL], k & 1);
                CCA_STAMP(2);
                for (int n = 0; n < NCH; ++n, ++u, ++oc) {
                    const uint32_t vb = op_base + (u & 1) * T::kOp;
                    mbar_wait(&bars[B_OP_FULL + (u & 1)], (u >> 1) & 1);
                    mbar_wait(&bars[B_O_EMPTY + (oc & 1)], ((oc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    CCA_STAMP(2);
                    mma_split3_loop<LK / 16>(tmem + 128 + (oc & 1) * kNC,
                                             p_base, p_base + T::kPP * T::kPlane, 2 * T::kPlane, T::kPlane, 128,
                                             vb, vb + 8 * T::kPlane, 256, 128, T::kPlane, idesc_o, false);
                    commit_to(&bars[B_O_FULL + (oc & 1)]);
                    commit_to(&bars[B_OP_EMPTY + (u & 1)]);
                    CCA_STAMP(2);
                }
                commit_to(&bars[B_P_EMPTY]);
            }
        }
    } else if (warp >= kWarpConv0) {
        // =============================== converters (256 threads) ===============================
        reg_dec<kRegsConv>();
        const int t = tid - kWarpConv0 * 32;
        const uint32_t total_items = (uint32_t)nk * (2 + NCH);
        int dbg_n = t == 0 ? 0 : 512;
        for (uint32_t g = 0; g < total_items; ++g) {
            const int slot = g & 1, ob = g & 1;
            mbar_wait(&bars[B_LD_FULL + slot], (g >> 1) & 1);
            CCA_STAMP(1);
            mbar_wait(&bars[B_OP_EMPTY + ob], ((g >> 1) & 1) ^ 1);
            CCA_STAMP(1);
            convert_slot<LK>(smem + S::off_ld + slot * T::kSlot, smem + S::off_op + ob * T::kOp, t);
            fence_proxy_async();
            mbar_arrive(&bars[B_OP_FULL + ob]);
            mbar_arrive(&bars[B_LD_EMPTY + slot]);
            CCA_STAMP(1);
        }
    } else if (warp >= 4) {
        // =============================== softmax group (128 threads, TMEM lane == query pixel) ===============================
        reg_inc<kRegsSoft>();
        const int r = tid - 128;
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        int dbg_n = r == 0 ? 0 : 512;
        for (int k = 0; k < nk; ++k) {
            int cw, ch, cb;
            line_coords(line_of(k), cw, ch, cb);
            const bool rvalid = r < p.L;
            CCA_STAMP(3);
            mbar_wait(&bars[B_S_FULL], k & 1);
            tc_fence_after();
            CCA_STAMP(3);
            float s[LK];
#pragma unroll
            for (int c0 = 0; c0 < LK; c0 += 16) tmem_ld16(tl + c0, reinterpret_cast<uint32_t *>(s + c0));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_S_EMPTY]);
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < LK; ++j) {
                const bool ok = j < p.L && !(p.col && j == r);
                s[j] = ok ? s[j] * kLog2e : -INFINITY;
                m = fmaxf(m, s[j]);
            }
            const float msub = (m == -INFINITY) ? 0.f : m;      // fully masked row (L == 1 in the column pass)
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < LK; ++j) { s[j] = exp2f(s[j] - msub); l += s[j]; }
            float sa = 0.f, sb = 0.f;
            if (rvalid) {
                const long pix = p.col ? ((long)cb * p.H + r) * p.W + cw : ((long)cb * p.H + ch) * p.W + r;
                const float mn = m * kLn2;                       // natural-log units
                if (p.col) {
                    p.stats[pix] = make_float2(mn, l);
                    sa = l > 0.f ? 1.f / l : 0.f;
                } else {
                    const float2 pc = p.stats[pix];
                    const float mm = fmaxf(mn, pc.x);
                    const float ar = exp2f((mn - mm) * kLog2e);
                    const float ac = pc.y > 0.f ? exp2f((pc.x - mm) * kLog2e) : 0.f;
                    const float lt = ar * l + ac * pc.y;
                    sa = ar / lt; sb = ac * pc.y / lt;
                    p.lse[pix] = mm + logf(lt);
                }
            }
            scale[(k % 3) * 128 + r] = make_float2(sa, sb);     // read by the epilogue group after O_FULL of line k
            // ---------------- P -> operand planes [key chunk][query pixel][16 B] (hi, lo)
            CCA_STAMP(3);
            mbar_wait(&bars[B_P_EMPTY], (k & 1) ^ 1);
            if (r < LK) {
                uint8_t *ph = smem + S::off_p + r * 16, *pl = ph + T::kPP * T::kPlane;
#pragma unroll
                for (int kc = 0; kc < T::kPP; ++kc) {
                    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
                    if (rvalid) split8(s + kc * 8, hi, lo);
                    *reinterpret_cast<uint4 *>(ph + kc * T::kPlane) = hi;
                    *reinterpret_cast<uint4 *>(pl + kc * T::kPlane) = lo;
                }
            }
            fence_proxy_async();
            mbar_arrive(&bars[B_P_FULL]);
            CCA_STAMP(3);
        }
    } else {
        // =============================== epilogue group (128 threads, TMEM lane == query pixel) ===============================
        reg_inc<kRegsEpi>();
        const int r = tid;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        const bool elected = tid == 0;
        uint32_t oc = 0;
        int dbg_n = 512;
        if (!p.col && elected) {                             // prefetch the partial of the very first chunk
            int cw, ch, cb;
            line_coords(line_of(0), cw, ch, cb);
            uint8_t *dst = smem + S::off_out;
            mbar_expect_tx(&bars[B_OUT_FULL + 0], T::kSlot);
            tma_load_4d(dst, &mo, &bars[B_OUT_FULL + 0], 0, cw, ch, cb);
            tma_load_4d(dst + T::kTile, &mo, &bars[B_OUT_FULL + 0], 32, cw, ch, cb);
        }
        for (int k = 0; k < nk; ++k) {
            int cw, ch, cb;
            line_coords(line_of(k), cw, ch, cb);
            float sa = 0.f, sb = 0.f;
            for (int n = 0; n < NCH; ++n, ++oc) {
                const int os = oc & 1;
                uint8_t *slot = smem + S::off_out + os * T::kSlot;
                if (elected) {
                    if (p.col) {
                        tma_store_wait_read<1>();              // the store that used this slot two chunks ago has drained
                        mbar_arrive(&bars[B_OUT_FULL + os]);
                    } else {
                        tma_store_wait_read<0>();              // other slot drained -> prefetch next chunk's partial into it
                        int nk2 = k, nn = n + 1;
                        if (nn == NCH) { nn = 0; nk2 = k + 1; }
                        if (nk2 < nk) {
                            int w2, h2, b2;
                            line_coords(line_of(nk2), w2, h2, b2);
                            uint8_t *dst = smem + S::off_out + (os ^ 1) * T::kSlot;
                            mbar_expect_tx(&bars[B_OUT_FULL + (os ^ 1)], T::kSlot);
                            tma_load_4d(dst, &mo, &bars[B_OUT_FULL + (os ^ 1)], nn * kNC, w2, h2, b2);
                            tma_load_4d(dst + T::kTile, &mo, &bars[B_OUT_FULL + (os ^ 1)], nn * kNC + 32, w2, h2, b2);
                        }
                    }
                }
                mbar_wait(&bars[B_OUT_FULL + os], (oc >> 1) & 1);
                mbar_wait(&bars[B_O_FULL + os], (oc >> 1) & 1);
                tc_fence_after();
                if (n == 0) { const float2 sc = scale[(k % 3) * 128 + r]; sa = sc.x; sb = sc.y; }
                float o[kNC];
#pragma unroll
                for (int c0 = 0; c0 < kNC; c0 += 16) tmem_ld16(tl + 128 + os * kNC + c0, reinterpret_cast<uint32_t *>(o + c0));
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[B_O_EMPTY + os]);
                if (r < LK) {
                    uint8_t *row = slot + r * 128;
                    const int sw = r & 7;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {               // 16 chunks of 4 channels
                        float4 *dst = reinterpret_cast<float4 *>(row + (j >> 3) * T::kTile + (((j & 7) ^ sw) * 16));
                        float4 v = make_float4(o[4 * j] * sa, o[4 * j + 1] * sa, o[4 * j + 2] * sa, o[4 * j + 3] * sa);
                        if (!p.col) {
                            const float4 q = *dst;
                            v.x = fmaf(q.x, sb, v.x); v.y = fmaf(q.y, sb, v.y); v.z = fmaf(q.z, sb, v.z); v.w = fmaf(q.w, sb, v.w);
                        }
                        *dst = v;
                    }
                }
                fence_proxy_async();
                named_bar_sync(1, 128);
                if (elected) {
                    tma_store_4d(&mo, slot, n * kNC, cw, ch, cb);
                    tma_store_4d(&mo, slot + T::kTile, n * kNC + 32, cw, ch, cb);
                    tma_store_commit();
                }
            }
        }
        if (elected) tma_store_wait_all<0>();
        (void)dbg_n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

long long *g_dbg = nullptr;   // set through cca_b200__set_debug_buffer (profiling aid, not part of the ABI)

template <int LK>
cudaError_t launch_pass(const void *q, const void *k, const void *v, void *out, float *lse, float2 *stats, Dims d,
                        bool col, cudaStream_t st, const char **why)
{
    CUtensorMap mq, mk, mv, mo;
    if (!make_map(&mq, q, d.B, d.H, d.W, d.Cq, LK, col) || !make_map(&mk, k, d.B, d.H, d.W, d.Cq, LK, col) ||
        !make_map(&mv, v, d.B, d.H, d.W, d.C, LK, col) || !make_map(&mo, out, d.B, d.H, d.W, d.C, LK, col)) {
        if (why) *why = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    FwdParams p;
    p.B = d.B; p.H = d.H; p.W = d.W; p.C = d.C; p.Cq = d.Cq;
    p.L = col ? d.H : d.W; p.NL = col ? d.W : d.H; p.col = col ? 1 : 0;
    p.stats = stats; p.lse = lse;
    p.dbg = g_dbg ? g_dbg + (col ? 0 : 2048) : nullptr;
    auto kern = cca_tc_fwd_kernel<LK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem<LK>::kBytes);
    if (e != cudaSuccess) return e;
    const int lines = d.B * p.NL;
    const int grid = lines < sm_count() ? lines : sm_count();
    kern<<<grid, kThreads, FwdSmem<LK>::kBytes, st>>>(mq, mk, mv, mo, p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

void set_tc_debug_buffer(void *p) { g_dbg = reinterpret_cast<long long *>(p); }

bool tc_forward_supported(Dims d, int dtype) { return tc::shape_supported(d, dtype); }

// q,k,v,out are channels-last (NHWC) fp32.
cudaError_t tc_forward(const void *q, const void *k, const void *v, void *out, float *lse, void *ws, Dims d, int dtype,
                       cudaStream_t st, const char **why)
{
    (void)dtype;
    float2 *stats = reinterpret_cast<float2 *>(ws);
    cudaError_t e;
    e = lk_for(d.H) == 80 ? launch_pass<80>(q, k, v, out, lse, stats, d, true, st, why)
                          : launch_pass<112>(q, k, v, out, lse, stats, d, true, st, why);
    if (e != cudaSuccess) return e;
    e = lk_for(d.W) == 80 ? launch_pass<80>(q, k, v, out, lse, stats, d, false, st, why)
                          : launch_pass<112>(q, k, v, out, lse, stats, d, false, st, why);
    return e;
}

}  // namespace cca
