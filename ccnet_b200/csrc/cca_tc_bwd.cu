// tcgen05 / TMA backward kernel of criss-cross attention for sm_100a (channels-last tensors).
//
// Closed form of SURVEY.md 8(a) row a11 (autograd of cc_attention/functions.py:38-47), flash-style: the
// attention matrix is recomputed per line from (q, k, lse), never stored.  As in the forward, a row and a
// column of a channels-last image are the same object, so one kernel runs twice:
//   pass 1 (columns, self entry masked): dq, dk, dv  = column-branch contributions
//   pass 2 (rows)                      : dq, dk, dv += row-branch contributions (TMA reduce-add, performed at L2)
// Each line CTA owns all outputs of its line: no atomics, deterministic.
//
// Per line (jq = query pixel, jk = key pixel, both < L <= LK):
//   S  = Q K^T                     (K-dim Cq)        P  = exp(S - lse[jq])            [TMEM -> planes in smem]
//   dP = dO V^T                    (K-dim C, chunked, accumulated in TMEM over the V/dO chunks)
//   dV[jk,c] = sum_jq P[jq,jk] dO[jq,c]   per chunk   (A = P planes read MN-major = P^T, B = dO chunk)
//   dS = P * (dP - delta[jq])      [planes overwrite P]
//   dQ[jq,c] = sum_jk dS[jq,jk] K[jk,c]   (A = dS planes K-major)     dK[jk,c] = sum_jq dS[jq,jk] Q[jq,c]  (A = dS^T)
// All GEMMs run as bf16x3 split MMAs (hi*hi + hi*lo + lo*hi) with fp32 accumulation in TMEM.
// The same operand planes serve several GEMMs: planes over channels are a K-major operand when the
// contraction runs over channels (S, dP) and an MN-major B operand when channels are the output (dV, dQ, dK);
// planes over key pixels are K-major A for dQ and MN-major A (= transpose) for dV / dK.
#include "cca_tc_common.cuh"

namespace cca {
namespace {
using namespace tc;

constexpr int kTmemCols = 512;      // S / dP double buffer: [0,128) [128,256)   O0: [256,320)   O1: [320,384)
constexpr int kTmemO = 256;
// converters keep a whole 128 B row live across the barrier of the in-place conversion (88 registers); the P/dS group streams
// its rows from TMEM 16 columns at a time
constexpr int kRegsSoft = 104, kRegsEpi = 128, kRegsConvB = 88;
static_assert(reg_pool_ok(kRegsSoft, kRegsEpi, kRegsConvB), "setmaxnreg pool");

struct BwdParams {
    int B, H, W, C, Cq;
    int L, NL, col;
    int sync;              // 1: the row pass runs overlapped with the tail of the column pass (programmatic dependent launch);
    unsigned int *done;    // done[b] counts the column lines of sample b whose dq/dk/dv stores are complete, and a row line
                           // only waits for the column lines of its own sample before it accumulates onto them
    int hints, keep_from;  // L2 eviction hints: the column pass keeps (evict_last) the lines of samples >= keep_from for the row
                           // pass, which walks the samples backwards; everything else streams (evict_first)
    const float *lse;
    const float *delta;
    long long *dbg;
};

#define CCA_STAMP(role)                                                                          \
    do {                                                                                         \
        if (p.dbg && blockIdx.x == 0 && dbg_n < 512) p.dbg[(role) * 512 + dbg_n++] = clock64();  \
    } while (0)

__device__ __forceinline__ void wait_count(const unsigned int *cnt, unsigned int need)
{
    unsigned int spins = 0, v;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if (v >= need) return;
        __nanosleep(64);
        if (++spins > (1u << 24)) __trap();      // a broken dependency chain traps instead of hanging the GPU
    }
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int LK, bool BF> struct BwdSmem {
    using T = Tiles<LK, BF>;
    static constexpr int kNLd = BF ? 6 : 5;                // every slot is the UMMA operand itself: a bf16 tile as loaded, an fp32
                                                           // tile once the converters have rewritten it in place (hi/lo planes)
    static constexpr int off_ld = 0;                       // kNLd load slots
    static constexpr int off_out = off_ld + kNLd * T::kSlot; // 1 out slot (the epilogue has slack; the store warp drives it)
    static constexpr int off_p = off_out + T::kSlot;       // P / dS planes (hi, lo); M=128 over-reads of a slot land in the next slot / here
    static constexpr int off_tail = off_p + T::kP + (16 - LK / 8) * T::kPlane;   // pad for the P^T over-read (16 planes of 8 key pixels)
    static constexpr int off_bar = off_tail + (128 - LK) * 16 + 256;
    static constexpr int kBytes = off_bar + 320;
    static_assert(kBytes <= 232448, "shared memory budget");
};

enum { B_LD_FULL = 0, B_LD_EMPTY = 6, B_OP_FULL = 12, B_S_FULL = 18, B_S_EMPTY = 20, B_P_FULL = 22,
       B_P_EMPTY = 23, B_O_FULL = 24, B_O_EMPTY = 26, B_OUT_FULL = 28, B_STAGED = 29, B_DP_FULL = 30, B_DS_FULL = 31,
       B_COUNT = 32 };

// Per line the ring carries  Q K (V_n dO_n)* Q K ; item g lives in load slot g % kNLd (fp32: converted in place), so the
// loads and conversions of the next chunks overlap the MMAs of the current one.  MMAs per line: S (phase A), per chunk dP += dO V^T then
// dV = P^T dO (phase B; the V buffer is released right after the dP MMAs), dS by the P/dS group (phase C),
// dQ = dS K and dK = dS^T Q (phase D).
template <int LK, bool BF>
__global__ void __launch_bounds__(kThreads, 1)
cca_tc_bwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                  const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo,
                  const __grid_constant__ CUtensorMap mdq, const __grid_constant__ CUtensorMap mdk,
                  const __grid_constant__ CUtensorMap mdv, BwdParams p)
{
    using T = Tiles<LK, BF>;
    using S = BwdSmem<LK, BF>;
    constexpr int TERMS = BF ? 1 : 3;
    constexpr int kNLd = S::kNLd;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::off_bar + 8 * B_COUNT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NCH = p.C / kNC;
    const int KQ = p.Cq / 16;
    const int NI = 4 + 2 * NCH;               // load items per line: Q K (V dO)* Q K
    const int NO = NCH + 2;                   // output items per line: dV chunks, dQ, dK
    const int total_lines = p.B * p.NL;
    const int nk = (total_lines - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < kNLd; ++i) {
            mbar_init(&bars[B_LD_FULL + i], 1); mbar_init(&bars[B_LD_EMPTY + i], 1); mbar_init(&bars[B_OP_FULL + i], kConvThreads);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars[B_O_FULL + i], 1); mbar_init(&bars[B_O_EMPTY + i], 128); }
        mbar_init(&bars[B_OUT_FULL], 1); mbar_init(&bars[B_STAGED], 128);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars[B_S_FULL + i], 1); mbar_init(&bars[B_S_EMPTY + i], 128); }
        mbar_init(&bars[B_P_FULL], 128); mbar_init(&bars[B_P_EMPTY], 1);
        mbar_init(&bars[B_DP_FULL], 1);  mbar_init(&bars[B_DS_FULL], 128);
        fence_mbar_init();
        prefetch_tmap(&mq); prefetch_tmap(&mk); prefetch_tmap(&mv); prefetch_tmap(&mdo);
        prefetch_tmap(&mdq); prefetch_tmap(&mdk); prefetch_tmap(&mdv);
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // Programmatic dependent launch: the column pass is launched behind the delta kernel and the row pass behind the column
    // pass, each without waiting for its predecessor to finish.  Only two places touch what the predecessor produces: the
    // P/dS group reads delta (column pass <- delta kernel), the store warp accumulates onto dq/dk/dv (row pass <- column pass).
    pdl_launch_dependents();

    auto line_of = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };
    auto line_coords = [&](int line, int &cw, int &ch, int &cb) {
        cb = line / p.NL;
        const int i = line - cb * p.NL;
        // second pass: samples backwards (the tail of the column pass is still in L2) -- unless it overlaps the column pass,
        // whose first samples are complete first
        if (!p.col && !p.sync) cb = p.B - 1 - cb;
        if (p.col) { cw = i; ch = 0; } else { cw = 0; ch = i; }
    };
    // output item i of a line: i < NCH -> dV chunk i; NCH -> dQ; NCH+1 -> dK
    auto out_map = [&](int i) -> const CUtensorMap * { return i < NCH ? &mdv : (i == NCH ? &mdq : &mdk); };
    auto out_c0 = [&](int i) { return i < NCH ? i * kNC : 0; };

    if (warp >= kWarpProducer) {
        reg_dec<kRegsMisc>();
        if (warp == kWarpProducer) {
            // =============================== TMA producer ===============================
            if (lane == 0) {
                uint32_t g = 0;
                int dbg_n = 0;
                const uint64_t pol_keep = p.hints ? l2_policy_evict_last() : l2_policy_evict_normal();
                const uint64_t pol_stream = p.hints ? l2_policy_evict_first() : l2_policy_evict_normal();
                auto emit = [&](const CUtensorMap *m, int c0, int line) {
                    int cw, ch, cb;
                    line_coords(line, cw, ch, cb);
                    const uint64_t pol = p.col && cb >= p.keep_from ? pol_keep : pol_stream;
                    const int slot = g % kNLd;
                    mbar_wait(&bars[B_LD_EMPTY + slot], ((g / kNLd) & 1) ^ 1);
                    CCA_STAMP(0);
                    uint8_t *dst = smem + S::off_ld + slot * T::kSlot;
                    mbar_expect_tx(&bars[B_LD_FULL + slot], T::kSlot);
                    tma_load_4d(dst, m, &bars[B_LD_FULL + slot], c0, cw, ch, cb, pol);
                    if constexpr (!BF) tma_load_4d(dst + T::kTile, m, &bars[B_LD_FULL + slot], c0 + 32, cw, ch, cb, pol);
                    ++g;
                };
                // ring:  Q0 K0 | (V dO)* Q1 K1 Q0 K0 | (V dO)* Q2 K2 Q1 K1 | ...   (S of the next line is issued while the
                // dS of the current one is being computed; Q,K of the current line come back for dQ/dK)
                emit(&mq, 0, line_of(0));
                emit(&mk, 0, line_of(0));
                for (int k = 0; k < nk; ++k) {
                    for (int n = 0; n < NCH; ++n) { emit(&mv, n * kNC, line_of(k)); emit(&mdo, n * kNC, line_of(k)); }
                    if (k + 1 < nk) { emit(&mq, 0, line_of(k + 1)); emit(&mk, 0, line_of(k + 1)); }
                    emit(&mq, 0, line_of(k));
                    emit(&mk, 0, line_of(k));
                }
            }
        } else if (warp == kWarpMma) {
            // =============================== MMA issuer ===============================
            const uint32_t id_kk_s = instr_desc(kFmtBF16, kFmtBF16, 128, LK, false, false);    // S, dP : K-major x K-major, N = LK
            const uint32_t id_mn_mn = instr_desc(kFmtBF16, kFmtBF16, 128, kNC, true, true);    // dV, dK: A^T planes x channel planes
            const uint32_t id_k_mn = instr_desc(kFmtBF16, kFmtBF16, 128, kNC, false, true);    // dQ
            const uint32_t pb = smem_u32(smem + S::off_p);
            const uint32_t LO8 = 8 * T::kPlane, LOP = T::kPP * T::kPlane;
            uint32_t u = 0, oc = 0;
            int dbg_n = lane == 0 ? 0 : 512;
            // operand of ring item g = load slot g % kNLd: fp32 -> bf16 hi/lo planes written in place (K-major / MN-major by descriptor);
            //                  bf16 -> the TMA tile itself read with SWIZZLE_128B descriptors:
            //                          K-major use: k-step = +32 B; MN-major use: k-step = +2048 B (16 pixel rows)
            const uint32_t ld_base = smem_u32(smem + S::off_ld);
            auto opb = [&](uint32_t g) { return ld_base + (g % kNLd) * T::kSlot; };
            auto wait_op = [&](uint32_t g) { mbar_wait(&bars[(BF ? B_LD_FULL : B_OP_FULL) + g % kNLd], (g / kNLd) & 1); };
            auto free_op = [&](uint32_t g) { commit_to(&bars[B_LD_EMPTY + g % kNLd]); };
            // channel-tile operand parameters: (k-step, lbo, sbo, layout) when the contraction runs over channels (kmaj) or
            // over pixels (mnmaj)
            constexpr uint32_t KS_K = BF ? 32 : 2 * T::kPlane, LBO_K = BF ? 16 : T::kPlane, SBO_K = BF ? 1024 : 128;
            constexpr uint32_t KS_MN = BF ? 2048 : 256, LBO_MN = BF ? 16 : 128, SBO_MN = BF ? 1024 : T::kPlane;
            constexpr uint32_t LAY = BF ? kSw128 : 0;
            auto issue_s = [&](int k) {                       // S(k) = Q K^T into S buffer k&1
                const uint32_t q = opb(u), kk = opb(u + 1);
                wait_op(u); wait_op(u + 1);
                mbar_wait(&bars[B_S_EMPTY + (k & 1)], ((k >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem + (k & 1) * 128;
                for (int ks = 0; ks < KQ; ++ks) {
                    const uint32_t ao = ks * KS_K;
                    mma_split3<TERMS>(d, smem_desc(q + ao, LBO_K, SBO_K, LAY), smem_desc(q + LO8 + ao, LBO_K, SBO_K, LAY),
                                      smem_desc(kk + ao, LBO_K, SBO_K, LAY), smem_desc(kk + LO8 + ao, LBO_K, SBO_K, LAY), id_kk_s, ks > 0);
                }
                commit_to(&bars[B_S_FULL + (k & 1)]);
                free_op(u); free_op(u + 1);
                u += 2;
            };
            issue_s(0);
            for (int k = 0; k < nk; ++k) {
                CCA_STAMP(2);
                const uint32_t sdp = tmem + (k & 1) * 128;        // S(k), then dP(k)
                mbar_wait(&bars[B_P_FULL], k & 1);
                CCA_STAMP(2);
                // ---- phase B: per chunk  dP += dO V^T  and  dV = P^T dO
                for (int n = 0; n < NCH; ++n, u += 2, ++oc) {
                    const uint32_t vb = opb(u), db = opb(u + 1);
                    wait_op(u); wait_op(u + 1);
                    tc_fence_after();
                    mma_split3_loop<kNC / 16, TERMS>(sdp, db, db + LO8, KS_K, LBO_K, SBO_K,
                                                     vb, vb + LO8, KS_K, LBO_K, SBO_K, id_kk_s, n > 0, LAY, LAY);
                    free_op(u);                                   // V is only needed by dP
                    mbar_wait(&bars[B_O_EMPTY + (oc & 1)], ((oc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    mma_split3_loop<LK / 16, TERMS>(tmem + kTmemO + (oc & 1) * kNC, pb, pb + LOP, 256, 128, T::kPlane,
                                                    db, db + LO8, KS_MN, LBO_MN, SBO_MN, id_mn_mn, false, 0, LAY);
                    commit_to(&bars[B_O_FULL + (oc & 1)]);
                    free_op(u + 1);
                    CCA_STAMP(2);
                }
                commit_to(&bars[B_DP_FULL]);
                if (k + 1 < nk) issue_s(k + 1);                   // overlaps the dS computation of this line
                // ---- phase D: dQ = dS K,  dK = dS^T Q      (dS in the P planes)
                mbar_wait(&bars[B_DS_FULL], k & 1);
                CCA_STAMP(2);
                {
                    const uint32_t q = opb(u), kk = opb(u + 1);
                    wait_op(u); wait_op(u + 1);
                    mbar_wait(&bars[B_O_EMPTY + (oc & 1)], ((oc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    mma_split3_loop<LK / 16, TERMS>(tmem + kTmemO + (oc & 1) * kNC, pb, pb + LOP, 2 * T::kPlane, T::kPlane, 128,
                                                    kk, kk + LO8, KS_MN, LBO_MN, SBO_MN, id_k_mn, false, 0, LAY);
                    commit_to(&bars[B_O_FULL + (oc & 1)]);
                    free_op(u + 1);
                    ++oc;
                    mbar_wait(&bars[B_O_EMPTY + (oc & 1)], ((oc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    mma_split3_loop<LK / 16, TERMS>(tmem + kTmemO + (oc & 1) * kNC, pb, pb + LOP, 256, 128, T::kPlane,
                                                    q, q + LO8, KS_MN, LBO_MN, SBO_MN, id_mn_mn, false, 0, LAY);
                    commit_to(&bars[B_O_FULL + (oc & 1)]);
                    free_op(u);
                    ++oc;
                    commit_to(&bars[B_P_EMPTY]);
                    u += 2;
                }
                CCA_STAMP(2);
            }
        } else if (warp == kWarpStore) {
            // =============================== store warp (one lane): the output staging slot <-> global ===============================
            if (lane == 0) {
                const uint32_t total = (uint32_t)nk * NO;
                uint8_t *slot = smem + S::off_out;
                const uint64_t pol_keep = p.hints ? l2_policy_evict_last() : l2_policy_evict_normal();
                const uint64_t pol_stream = p.hints ? l2_policy_evict_first() : l2_policy_evict_normal();
                for (uint32_t c = 0; c < total; ++c) {
                    const int k = c / NO, i = c - k * NO;
                    int cw, ch, cb;
                    line_coords(line_of(k), cw, ch, cb);
                    // slot is free once the previous store has been read out of shared memory
                    tma_store_wait_read<0>();
                    mbar_arrive(&bars[B_OUT_FULL]);
                    mbar_wait(&bars[B_STAGED], c & 1);
                    if (p.sync && !p.col) {
                        // overlapped row pass: all column lines of this sample must have landed before we accumulate
                        if (i == 0) { wait_count(p.done + cb, (unsigned)p.W); fence_proxy_async_global(); }
                    } else if (c == 0) {
                        pdl_wait();
                    }
                    if (p.col) {                                   // column pass defines dq/dk/dv ...
                        const uint64_t pol = cb >= p.keep_from ? pol_keep : pol_stream;
                        tma_store_4d(out_map(i), slot, out_c0(i), cw, ch, cb, pol);
                        if constexpr (!BF) tma_store_4d(out_map(i), slot + T::kTile, out_c0(i) + 32, cw, ch, cb, pol);
                    } else {                                       // ... the row pass accumulates onto them (TMA reduce-add at L2)
                        tma_reduce_add_4d(out_map(i), slot, out_c0(i), cw, ch, cb, pol_stream);
                        if constexpr (!BF) tma_reduce_add_4d(out_map(i), slot + T::kTile, out_c0(i) + 32, cw, ch, cb, pol_stream);
                    }
                    tma_store_commit();
                    if (p.sync && p.col && i == 0 && k > 0) {
                        // publish the previous line: its stores were committed at least one chunk period ago, so waiting for
                        // everything but the group just committed costs (almost) nothing
                        int pw, ph2, pb;
                        line_coords(line_of(k - 1), pw, ph2, pb);
                        tma_store_wait_all<1>();
                        fence_proxy_async_global();
                        __threadfence();
                        atomicAdd(p.done + pb, 1u);
                    }
                }
                tma_store_wait_all<0>();
                if (p.sync && p.col && nk > 0) {
                    int pw, ph2, pb;
                    line_coords(line_of(nk - 1), pw, ph2, pb);
                    fence_proxy_async_global();
                    __threadfence();
                    atomicAdd(p.done + pb, 1u);
                }
            }
        }
    } else if (warp >= kWarpConv0) {
        // =============================== converters (256 threads) ===============================
        reg_dec<kRegsConvB>();
        const int t = tid - kWarpConv0 * 32;
        const uint32_t total = BF ? 0u : (uint32_t)nk * NI;         // bf16 tiles need no conversion
        int dbg_n = t == 0 ? 0 : 512;
        for (uint32_t g = 0; g < total; ++g) {
            const int slot = g % kNLd;
            mbar_wait(&bars[B_LD_FULL + slot], (g / kNLd) & 1);
            CCA_STAMP(1);
            if constexpr (!BF) convert_slot_inplace<LK>(smem + S::off_ld + slot * T::kSlot, t);
            fence_proxy_async();
            mbar_arrive(&bars[B_OP_FULL + slot]);
            CCA_STAMP(1);
        }
    } else if (warp >= 4) {
        // =============================== P / dS group (128 threads, TMEM lane == query pixel) ===============================
        reg_inc<kRegsSoft>();
        const int r = tid - 128;
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        int dbg_n = r == 0 ? 0 : 512;
        for (int k = 0; k < nk; ++k) {
            int cw, ch, cb;
            line_coords(line_of(k), cw, ch, cb);
            const bool rvalid = r < p.L;
            float lse2 = 0.f, dl = 0.f;
            if (rvalid) {
                const long pix = p.col ? ((long)cb * p.H + r) * p.W + cw : ((long)cb * p.H + ch) * p.W + r;
                lse2 = p.lse[pix] * kLog2e;
                if (p.col) pdl_wait();                 // delta comes from the kernel right before the column pass
                dl = p.delta[pix];
            }
            uint8_t *ph = smem + S::off_p + r * 16, *pl = ph + T::kPP * T::kPlane;
            // ---------------- P = exp(S - lse)
            const uint32_t tsd = tl + (k & 1) * 128;           // S(k) / dP(k) buffer
            CCA_STAMP(3);
            mbar_wait(&bars[B_S_FULL + (k & 1)], (k >> 1) & 1);
            tc_fence_after();
            CCA_STAMP(3);
            // both phases stream the TMEM row 16 columns at a time (small loops: the instruction footprint of a fully
            // unrolled LK-element register array cost more than the extra tcgen05.wait::ld round trips)
            float nx[16];                                      // next 16 columns, in flight while the current ones are processed
            tmem_ld16(tsd, reinterpret_cast<uint32_t *>(nx));
            mbar_wait(&bars[B_P_EMPTY], (k & 1) ^ 1);
#pragma unroll
            for (int c0 = 0; c0 < LK; c0 += 16) {
                float s[16];
                tmem_ld_wait16(reinterpret_cast<uint32_t *>(nx));
#pragma unroll
                for (int e = 0; e < 16; ++e) s[e] = nx[e];
                if (c0 + 16 < LK) tmem_ld16(tsd + c0 + 16, reinterpret_cast<uint32_t *>(nx));
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int j = c0 + e;
                    const bool ok = rvalid && j < p.L && !(p.col && j == r);
                    s[e] = ok ? exp2f(s[e] * kLog2e - lse2) : 0.f;
                }
                if (r < LK) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int kc = c0 / 8 + h;
                        const float *v8 = s + h * 8;
                        if constexpr (BF) {
                            *reinterpret_cast<uint4 *>(ph + kc * T::kPlane) =
                                make_uint4(pack_bf16(v8[0], v8[1]), pack_bf16(v8[2], v8[3]), pack_bf16(v8[4], v8[5]), pack_bf16(v8[6], v8[7]));
                        } else {
                            uint4 hi, lo;
                            split8(v8, hi, lo);
                            *reinterpret_cast<uint4 *>(ph + kc * T::kPlane) = hi;
                            *reinterpret_cast<uint4 *>(pl + kc * T::kPlane) = lo;
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&bars[B_P_FULL]);
            CCA_STAMP(3);
            // ---------------- dS = P * (dP - delta)   (all MMAs that read P have completed: DP_FULL is committed after them)
            mbar_wait(&bars[B_DP_FULL], k & 1);
            tc_fence_after();
            CCA_STAMP(3);
            tmem_ld16(tsd, reinterpret_cast<uint32_t *>(nx));
#pragma unroll
            for (int c0 = 0; c0 < LK; c0 += 16) {
                float dp[16];
                tmem_ld_wait16(reinterpret_cast<uint32_t *>(nx));
#pragma unroll
                for (int e = 0; e < 16; ++e) dp[e] = nx[e];
                if (c0 + 16 < LK) tmem_ld16(tsd + c0 + 16, reinterpret_cast<uint32_t *>(nx));
                if (r < LK) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int kc = c0 / 8 + h;
                        const uint4 hi = *reinterpret_cast<const uint4 *>(ph + kc * T::kPlane);
                        uint4 lo = make_uint4(0, 0, 0, 0);
                        if constexpr (!BF) lo = *reinterpret_cast<const uint4 *>(pl + kc * T::kPlane);
                        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
                        float ds[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float p0 = bf_lo(hw[e]) + bf_lo(lw[e]), p1 = bf_hi(hw[e]) + bf_hi(lw[e]);
                            ds[2 * e] = p0 * (dp[h * 8 + 2 * e] - dl);
                            ds[2 * e + 1] = p1 * (dp[h * 8 + 2 * e + 1] - dl);
                        }
                        if constexpr (BF) {
                            *reinterpret_cast<uint4 *>(ph + kc * T::kPlane) =
                                make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
                        } else {
                            uint4 dh, dlo;
                            split8(ds, dh, dlo);
                            *reinterpret_cast<uint4 *>(ph + kc * T::kPlane) = dh;
                            *reinterpret_cast<uint4 *>(pl + kc * T::kPlane) = dlo;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars[B_S_EMPTY + (k & 1)]);
            fence_proxy_async();
            mbar_arrive(&bars[B_DS_FULL]);
            CCA_STAMP(3);
        }
    } else {
        // =============================== epilogue group (128 threads, TMEM lane == output pixel) ===============================
        reg_inc<kRegsEpi>();
        const int r = tid;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        uint8_t *slot = smem + S::off_out;
        uint32_t oc = 0;
        int dbg_n = tid == 0 ? 0 : 512;
        for (int k = 0; k < nk; ++k) {
            for (int i = 0; i < NO; ++i, ++oc) {
                const int ob = oc & 1;
                CCA_STAMP(4);
                mbar_wait(&bars[B_OUT_FULL], oc & 1);                 // staging slot free / column partial landed
                mbar_wait(&bars[B_O_FULL + ob], (oc >> 1) & 1);
                tc_fence_after();
                CCA_STAMP(4);
                float o[kNC];
#pragma unroll
                for (int c0 = 0; c0 < kNC; c0 += 16) tmem_ld16(tl + kTmemO + ob * kNC + c0, reinterpret_cast<uint32_t *>(o + c0));
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[B_O_EMPTY + ob]);
                if (r < p.L) {                                        // rows >= L are clipped by the TMA store
                    uint8_t *row = slot + r * 128;
                    const int sw = r & 7;
                    if constexpr (BF) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4 *>(row + ((j ^ sw) * 16)) =
                                make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                                           pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            *reinterpret_cast<float4 *>(row + (j >> 3) * T::kTile + (((j & 7) ^ sw) * 16)) =
                                make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&bars[B_STAGED]);
                CCA_STAMP(4);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

// delta[pix] = sum_c dout[pix][c] * out[pix][c]   (channels-last: one warp per pixel, float4 lanes)
__global__ void __launch_bounds__(256) cca_delta_nhwc_kernel(const float4 *__restrict__ dout, const float4 *__restrict__ out,
                                                             float *__restrict__ delta, long npix, int c4)
{
    // pixels are walked backwards: the column pass starts with sample 0, whose dout is then the most recent data in L2
    pdl_launch_dependents();
    const long pix = npix - 1 - ((long)blockIdx.x * 8 + (threadIdx.x >> 5));
    if (pix < 0) return;
    const int lane = threadIdx.x & 31;
    const float4 *a = dout + pix * c4, *b = out + pix * c4;
    float s = 0.f;
    for (int i = lane; i < c4; i += 32) {
        const float4 x = __ldg(a + i), y = __ldg(b + i);
        s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) delta[pix] = s;
}

// bf16 variant: 8 channels per 16-byte load
__global__ void __launch_bounds__(256) cca_delta_nhwc_bf16_kernel(const uint4 *__restrict__ dout, const uint4 *__restrict__ out,
                                                                  float *__restrict__ delta, long npix, int c8)
{
    pdl_launch_dependents();
    const long pix = npix - 1 - ((long)blockIdx.x * 8 + (threadIdx.x >> 5));
    if (pix < 0) return;
    const int lane = threadIdx.x & 31;
    const uint4 *a = dout + pix * c8, *b = out + pix * c8;
    float s = 0.f;
    for (int i = lane; i < c8; i += 32) {
        const uint4 x = __ldg(a + i), y = __ldg(b + i);
        const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) s += bf_lo(xw[e]) * bf_lo(yw[e]) + bf_hi(xw[e]) * bf_hi(yw[e]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) delta[pix] = s;
}

long long *g_bwd_dbg = nullptr;

template <int LK, bool BF>
cudaError_t launch_bwd_pass(const void *dout, const void *q, const void *k, const void *v, const float *lse, const float *delta,
                            void *dq, void *dk, void *dv, unsigned int *done, int sync, Dims d, bool col, cudaStream_t st,
                            const char **why)
{
    CUtensorMap mq, mk, mv, mdo, mdq, mdk, mdv;
    const bool ok = make_map(&mq, q, d.B, d.H, d.W, d.Cq, LK, col, BF) && make_map(&mk, k, d.B, d.H, d.W, d.Cq, LK, col, BF) &&
                    make_map(&mv, v, d.B, d.H, d.W, d.C, LK, col, BF) && make_map(&mdo, dout, d.B, d.H, d.W, d.C, LK, col, BF) &&
                    make_map(&mdq, dq, d.B, d.H, d.W, d.Cq, LK, col, BF) && make_map(&mdk, dk, d.B, d.H, d.W, d.Cq, LK, col, BF) &&
                    make_map(&mdv, dv, d.B, d.H, d.W, d.C, LK, col, BF);
    if (!ok) {
        if (why) *why = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    BwdParams p;
    p.B = d.B; p.H = d.H; p.W = d.W; p.C = d.C; p.Cq = d.Cq;
    p.L = col ? d.H : d.W; p.NL = col ? d.W : d.H; p.col = col ? 1 : 0;
    p.lse = lse; p.delta = delta;
    p.done = done; p.sync = sync;
    {   // per sample the row pass re-reads q,k,v,dout and accumulates onto dq,dk,dv
        const double per_sample = (4.0 * d.Cq + 3.0 * d.C) * d.H * d.W * (BF ? 2 : 4);
        int keep = (int)(tc_l2_keep_mb() * 1e6 / per_sample);
        if (keep > d.B) keep = d.B;
        p.hints = tc_l2_hints();
        p.keep_from = d.B - keep;
    }
    p.dbg = g_bwd_dbg ? g_bwd_dbg + (col ? 0 : 2560) : nullptr;
    auto kern = cca_tc_bwd_kernel<LK, BF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<LK, BF>::kBytes);
    if (e != cudaSuccess) return e;
    const int lines = d.B * p.NL;
    const int grid = lines < sm_count() ? lines : sm_count();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = BwdSmem<LK, BF>::kBytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tc_pdl() ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, mdo, mdq, mdk, mdv, p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

void set_tc_bwd_debug_buffer(void *p) { g_bwd_dbg = reinterpret_cast<long long *>(p); }

bool tc_backward_supported(Dims d, int dtype) { return tc::shape_supported(d, dtype); }

// all tensors channels-last (NHWC), fp32 or bf16; ws = delta [B,H,W]
cudaError_t tc_backward(const void *dout, const void *q, const void *k, const void *v, const void *out, const float *lse,
                        void *dq, void *dk, void *dv, void *ws, Dims d, int dtype, cudaStream_t st, const char **why)
{
    float *delta = reinterpret_cast<float *>(ws);
    const long npix = (long)d.B * d.H * d.W;
    const bool bf = dtype == CCA_BF16;
    const unsigned dgrid = (unsigned)((npix + 7) / 8);
    // tc_pdl() == 2: the row pass overlaps the tail of the column pass; per-sample completion counters (behind delta in ws)
    unsigned int *done = reinterpret_cast<unsigned int *>(delta + npix);
    const int sync = tc_pdl() == 2 ? 1 : 0;
    if (sync) {
        cudaError_t e0 = cudaMemsetAsync(done, 0, sizeof(unsigned int) * d.B, st);
        if (e0 != cudaSuccess) return e0;
    }
    if (bf)
        cca_delta_nhwc_bf16_kernel<<<dgrid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(dout), reinterpret_cast<const uint4 *>(out),
                                                          delta, npix, d.C / 8);
    else
        cca_delta_nhwc_kernel<<<dgrid, 256, 0, st>>>(reinterpret_cast<const float4 *>(dout), reinterpret_cast<const float4 *>(out),
                                                     delta, npix, d.C / 4);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    auto go = [&](int lk, bool col) {
        if (bf)
            return lk == 80 ? launch_bwd_pass<80, true>(dout, q, k, v, lse, delta, dq, dk, dv, done, sync, d, col, st, why)
                            : launch_bwd_pass<112, true>(dout, q, k, v, lse, delta, dq, dk, dv, done, sync, d, col, st, why);
        return lk == 80 ? launch_bwd_pass<80, false>(dout, q, k, v, lse, delta, dq, dk, dv, done, sync, d, col, st, why)
                        : launch_bwd_pass<112, false>(dout, q, k, v, lse, delta, dq, dk, dv, done, sync, d, col, st, why);
    };
    e = go(lk_for(d.H), true);
    if (e != cudaSuccess) return e;
    return go(lk_for(d.W), false);
}

}  // namespace cca
