// tcgen05 / TMA forward (values) kernel of criss-cross attention for sm_100a, fp32 I/O, "channel-major accumulator" variant.
//
// Same formulation as cca_tc_fwd.cu (cc_attention/functions.py:38-47): the statistics pre-pass has left the
// partial log-sum-exp planes, an item (cca_items.cuh) computes  P = exp(S - lse)  and adds  P V_block  onto its query pixels.
// What differs is which way round the P V product runs.  cca_tc_fwd.cu computes O[query][channel] = P V with V as the
// shared-memory operand: every V chunk is rewritten in place as bf16 hi/lo planes (28 KB read + 28 KB written), fetched three
// times by the split MMAs, and the result goes TMEM -> registers -> staging tile -> TMA.  ncu shows that kernel bound by the
// shared-memory data pipe (LSU + tensor-core + TMA wavefronts ~ 90 % of the cycles, profiles/r02_tc_ncu_summary.txt).
// Here the product is transposed:
//     O^T[channel][query] = V^T[channel][key] P^T[key][query]
//   A = V^T lives in TENSOR MEMORY (lane = channel, two bf16 keys per 32-bit column, hi block + lo block): the converter
//       warps read the landed fp32 tile with lane = channel (LDS.32, one 128-byte row per warp instruction, conflict
//       free), split in registers and write TMEM with tcgen05.st -- V is never written back to shared memory and never
//       fetched from it by the tensor core;
//   B = P^T comes from shared memory as K-major planes [8-key chunk][query][16 B] (hi, lo), written ONCE per item by
//       the softmax group (50 KB) and read by all the channel groups of the item;
//   D = O^T in TMEM with lane = channel: an epilogue warp holds 32 consecutive channels of one query pixel per column,
//       i.e. exactly one 128-byte line of the channels-last output, so it stores / reduce-adds straight from registers
//       with fully coalesced st.global / red.global.add -- no staging tile, no TMA store.
// Shared-memory wavefronts per item (LK = 112, C = 512): ~10 K instead of ~17 K.
//
// Output protocol as in cca_tc_fwd.cu: producer items (column lines, first key block) STORE, all other items of the sample ADD
// after the per-sample counter cdone[b] says every producer has finished (release: __threadfence + atomicAdd by the epilogue
// group; acquire: ld.acquire spin by one thread, then a named barrier).  Items are walked in the lagged order, a waiting item
// only ever waits for lower-indexed ones.  One tile per line => one store + one add per element => bit-reproducible.
//
// Roles (896 threads, registers rebalanced with setmaxnreg):
//   warps 0-3   epilogue group  : TMEM lane = channel of the current 128-channel group; O^T halves -> st.global / red.global
//   warps 4-7   softmax group   : TMEM lane = query pixel; P = exp2(S log2e - lse2) -> bf16 hi/lo planes in shared memory
//   warps 8-23  converters      : Q, K slots: fp32 -> bf16 hi/lo planes in place (operands of S = Q K^T);
//                                 V slots: fp32 tile -> registers -> TMEM (A operand); warps with (warp & 3) < 2 take the
//                                 even chunks (TMEM lanes 0-63), the others the odd chunks (lanes 64-127)
//   warp 24     TMA producer    : same ring as cca_tc_fwd.cu
//   warp 25     MMA issuer      : S (SS), then per 128-channel group O^T = V^T P^T as two column halves (N = 64, LK - 64) so that
//                                 the epilogue drains one half while the other is being computed
#include "cca_items.cuh"
#include "cca_tc_common.cuh"

namespace cca {
namespace {
using namespace tc;

constexpr int kGroupCh = 128;       // channels per accumulator group (= TMEM lanes)
constexpr int kNA = 48;             // columns (query pixels) of the first accumulator half (the second one has LK - 48)
constexpr int kThreadsT = 1024;     // 32 warps: the 28 of cca_tc_common.cuh plus a second epilogue group (warps 28-31)
constexpr int kRegsLaunchT = 64;    // 65536 / 1024
constexpr int kWarpEpiB = 28;
// register budget (setmaxnreg; launch allocation 64 per thread): the converters hold a 32-key column of their channel plus its
// hi/lo split across a wait and get what the epilogue, producer, MMA and idle warps give up
constexpr int kRegsSoftT = 72, kRegsEpiT = 56, kRegsConvT = 72, kRegsMiscT = 40;
static_assert(256 * (kRegsLaunchT - kRegsEpiT) + 128 * (kRegsLaunchT - kRegsMiscT) >=
              kConvThreads * (kRegsConvT - kRegsLaunchT) + 128 * (kRegsSoftT - kRegsLaunchT), "setmaxnreg pool");

struct FwdTParams {
    ItemSpace sp;
    int C, Cq;
    long npix;
    const float *parts;    // [nparts][B*H*W] partial log2-sum-exp2 (statistics pre-pass)
    float *lse;            // [B,H,W] natural-log lse (saved for backward)
    float *out;            // [B,H,W,C] fp32
    unsigned int *cdone;   // [B] producer items of sample b whose stores are visible (cleared by the statistics kernel)
    int lag;
    int hints;
};

template <int LK> struct FwdTSmem {
    using T = Tiles<LK, false>;
    static constexpr int kNLd = LK == 80 ? 8 : 4;
    static constexpr int off_ld = 0;                               // load slots (Q, K converted in place; V only read)
    static constexpr int off_p = off_ld + kNLd * T::kSlot;         // P planes, double buffered: [hi block][lo block]
    static constexpr int off_tail = off_p + 2 * T::kP;             // M=128 MMAs read (128 - LK) rows past a Q slot
    static constexpr int off_cnt = off_tail + 1024;                // unsigned [8]: warps that have read a V slot
    static constexpr int off_bar = off_cnt + 64;
    static constexpr int kBytes = off_bar + 8 * 40 + 32;
    static_assert(kBytes <= 232448, "shared memory budget");
};
// TMEM columns: S [0, LK)   V^T double buffer [128, 128 + 2 LK) (hi block LK/2 columns, lo block LK/2)   O^T [128 + 2 LK, + LK)
template <int LK> struct TmemT {
    static constexpr int kS = 0, kV = 128, kO = 128 + 2 * LK;
    static_assert(kO + LK <= 512, "tensor memory budget");
};

enum { T_LD_FULL = 0, T_LD_EMPTY = 8, T_OP_FULL = 16, T_S_FULL = 24, T_S_EMPTY = 25, T_P_FULL = 26, T_P_EMPTY = 28,
       T_V_FULL = 30, T_V_EMPTY = 32, T_O_FULL = 34, T_O_EMPTY = 36, T_COUNT = 38 };

template <int LK>
__global__ void __launch_bounds__(kThreadsT, 1)
cca_tc_fwdt_kernel(const __grid_constant__ CUtensorMap mqc, const __grid_constant__ CUtensorMap mqr,
                   const __grid_constant__ CUtensorMap mkc, const __grid_constant__ CUtensorMap mkr,
                   const __grid_constant__ CUtensorMap mvc, const __grid_constant__ CUtensorMap mvr, FwdTParams p)
{
    using T = Tiles<LK, false>;
    using S = FwdTSmem<LK>;
    using TM = TmemT<LK>;
    constexpr int kNLd = S::kNLd;
    constexpr int kNB = LK - kNA;             // columns of the second accumulator half (64 or 32)
    static_assert(LK % 16 == 0 && LK / 16 <= 8 && kNB % 16 == 0 && kNB >= 16, "tile geometry");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::off_bar + 8 * T_COUNT);
    unsigned int *rd_cnt = reinterpret_cast<unsigned int *>(smem + S::off_cnt);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NCH = p.C / kNC;                // 64-channel chunks (ring slots) per item
    const int NG = p.C / kGroupCh;            // 128-channel accumulator groups per item
    const int KQ = p.Cq / 16;
    const int nk = p.sp.total > (int)blockIdx.x ? (p.sp.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto item_of = [&](int k) { return decode_item_order(p.sp, (int)blockIdx.x + k * (int)gridDim.x, p.lag); };
    // ring order:  Q0 K0 | V0[0..qkpos) Q1 K1 V0[qkpos..NCH) | V1[0..qkpos) Q2 K2 ...
    const int qkpos = NCH >= 3 ? 2 : NCH - 1;

    if (tid == 0) {
        for (int i = 0; i < kNLd; ++i) {
            mbar_init(&bars[T_LD_FULL + i], 1); mbar_init(&bars[T_LD_EMPTY + i], 1); mbar_init(&bars[T_OP_FULL + i], kConvThreads);
        }
        mbar_init(&bars[T_S_FULL], 1); mbar_init(&bars[T_S_EMPTY], 128);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[T_P_FULL + i], 128); mbar_init(&bars[T_P_EMPTY + i], 1);
            mbar_init(&bars[T_V_FULL + i], kConvThreads); mbar_init(&bars[T_V_EMPTY + i], 1);
            mbar_init(&bars[T_O_FULL + i], 1); mbar_init(&bars[T_O_EMPTY + i], 128);
        }
        fence_mbar_init();
        prefetch_tmap(&mqc); prefetch_tmap(&mqr); prefetch_tmap(&mkc); prefetch_tmap(&mkr); prefetch_tmap(&mvc); prefetch_tmap(&mvr);
    }
    if (tid < 8) rd_cnt[tid] = 0u;
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp >= kWarpProducer && warp < kWarpEpiB) {
        reg_dec<kRegsMiscT>();
        if (warp == kWarpProducer) {
            // =============================== TMA producer ===============================
            if (lane == 0) {
                uint32_t g = 0;
                const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
                auto emit = [&](const CUtensorMap *mc, const CUtensorMap *mr, int c0, const Item &it, int start) {
                    const CUtensorMap *m = it.col ? mc : mr;
                    const int cw = it.col ? it.line : start, ch = it.col ? start : it.line;
                    const int slot = g % kNLd;
                    mbar_wait(&bars[T_LD_EMPTY + slot], ((g / kNLd) & 1) ^ 1);
                    uint8_t *dst = smem + S::off_ld + slot * T::kSlot;
                    mbar_expect_tx(&bars[T_LD_FULL + slot], T::kSlot);
                    if (p.hints == 1) {     // producers' operands are read again by the sample's consumers; theirs are not
                        const uint64_t pol = is_producer(it) ? pol_keep : pol_stream;
                        tma_load_4d(dst, m, &bars[T_LD_FULL + slot], c0, cw, ch, it.b, pol);
                        tma_load_4d(dst + T::kTile, m, &bars[T_LD_FULL + slot], c0 + 32, cw, ch, it.b, pol);
                    } else {
                        tma_load_4d(dst, m, &bars[T_LD_FULL + slot], c0, cw, ch, it.b);
                        tma_load_4d(dst + T::kTile, m, &bars[T_LD_FULL + slot], c0 + 32, cw, ch, it.b);
                    }
                    ++g;
                };
                if (nk > 0) {
                    const Item it0 = item_of(0);
                    emit(&mqc, &mqr, 0, it0, it0.q0);
                    emit(&mkc, &mkr, 0, it0, it0.k0);
                }
                for (int k = 0; k < nk; ++k) {
                    const Item it = item_of(k);
                    for (int n = 0; n < NCH; ++n) {
                        if (n == qkpos && k + 1 < nk) {
                            const Item nx = item_of(k + 1);
                            emit(&mqc, &mqr, 0, nx, nx.q0);
                            emit(&mkc, &mkr, 0, nx, nx.k0);
                        }
                        emit(&mvc, &mvr, n * kNC, it, it.k0);
                    }
                }
            }
        } else if (warp == kWarpMma) {
            // =============================== MMA issuer (whole warp, elect.sync inside) ===============================
            const uint32_t idesc_s = instr_desc(kFmtBF16, kFmtBF16, 128, LK, false, false);
            const uint32_t idesc_a = instr_desc(kFmtBF16, kFmtBF16, 128, kNA, false, false);
            const uint32_t idesc_b = instr_desc(kFmtBF16, kFmtBF16, 128, kNB, false, false);
            uint32_t u = 0, gc = 0, qkpar = 0;    // ring position, accumulator groups issued, OP_FULL parity per slot
            const uint32_t ld_base = smem_u32(smem + S::off_ld);
            auto issue_s = [&](int k) {            // S(k) = Q K^T from ring items u (Q) and u+1 (K)
                // (only Q / K slots complete an OP_FULL phase -- V slots go to TMEM -- so the parity is per slot, not per ring turn)
                const uint32_t sq = u % kNLd, sk = (u + 1) % kNLd;
                const uint32_t qb = ld_base + sq * T::kSlot, kb = ld_base + sk * T::kSlot;
                mbar_wait(&bars[T_OP_FULL + sq], (qkpar >> sq) & 1);
                mbar_wait(&bars[T_OP_FULL + sk], (qkpar >> sk) & 1);
                qkpar ^= (1u << sq) | (1u << sk);
                mbar_wait(&bars[T_S_EMPTY], (k & 1) ^ 1);
                tc_fence_after();
                for (int ks = 0; ks < KQ; ++ks) {
                    const uint32_t ao = ks * 2 * T::kPStride;
                    mma_split3<3>(tmem + TM::kS, smem_desc(qb + ao, T::kPStride, 128), smem_desc(qb + T::kLoOff + ao, T::kPStride, 128),
                                  smem_desc(kb + ao, T::kPStride, 128), smem_desc(kb + T::kLoOff + ao, T::kPStride, 128),
                                  idesc_s, ks > 0);
                }
                commit_to(&bars[T_S_FULL]);
                commit_to(&bars[T_LD_EMPTY + sq]);
                commit_to(&bars[T_LD_EMPTY + sk]);
                u += 2;
            };
            if (nk > 0) issue_s(0);
            for (int k = 0; k < nk; ++k) {
                mbar_wait(&bars[T_P_FULL + (k & 1)], (k >> 1) & 1);
                tc_fence_after();
                const uint32_t pb = smem_u32(smem + S::off_p + (k & 1) * T::kP);          // hi block; lo block at + kPP * kPlane
                constexpr uint32_t LOP = T::kPP * T::kPlane;
                for (int g = 0; g < NG; ++g, ++gc) {
                    for (int h = 0; h < 2; ++h, ++u)                                      // ring accounting: chunks 2g, 2g+1
                        if (2 * g + h == qkpos && k + 1 < nk) issue_s(k + 1);
                    const uint32_t vb = gc & 1;
                    mbar_wait(&bars[T_V_FULL + vb], (gc >> 1) & 1);
                    const uint32_t vh = tmem + TM::kV + vb * LK, vl = vh + LK / 2;
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        mbar_wait(&bars[T_O_EMPTY + hb], (gc & 1) ^ 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tmem + TM::kO + hb * kNA;
                            const uint32_t idesc = hb ? idesc_b : idesc_a;
                            const uint32_t prow = pb + hb * kNA * 16;                    // query rows [0, 48) or [48, LK) of every plane
#pragma unroll
                            for (int ks = 0; ks < LK / 16; ++ks) {
                                const uint64_t bh = smem_desc(prow + ks * 2 * T::kPlane, T::kPlane, 128);
                                const uint64_t bl = smem_desc(prow + LOP + ks * 2 * T::kPlane, T::kPlane, 128);
                                mma_f16_ts(d, vh + ks * 8, bh, idesc, ks > 0);
                                mma_f16_ts(d, vl + ks * 8, bh, idesc, true);
                                mma_f16_ts(d, vh + ks * 8, bl, idesc, true);
                            }
                        }
                        __syncwarp();
                        commit_to(&bars[T_O_FULL + hb]);
                    }
                    commit_to(&bars[T_V_EMPTY + vb]);
                }
                commit_to(&bars[T_P_EMPTY + (k & 1)]);
            }
        }
        // (warps 26, 27 have no role in this kernel)
    } else if (warp >= kWarpConv0 && warp < kWarpEpiB) {
        // =============================== converters (512 threads) ===============================
        reg_inc<kRegsConvT>();
        const int t = tid - kWarpConv0 * 32;
        const int quarter = warp & 3;                              // TMEM lane quarter this warp may access
        const int half = quarter >> 1, box = quarter & 1;          // chunk parity it serves, 32-channel TMA box inside the chunk
        const int ksplit = (warp - kWarpConv0) >> 2;               // which share of the keys: 16-key units [u0, u0 + nu)
        constexpr int NU = LK / 16;
        const int u0 = 2 * ksplit;
        const int nu = NU - u0 >= 2 ? 2 : (NU - u0 > 0 ? NU - u0 : 0);
        const uint32_t tl = tmem + ((uint32_t)(quarter * 32) << 16);
        uint32_t g = 0;
        int pend = -1;                                             // Q/K slot converted but not yet fenced / published
        auto publish = [&]() {
            if (pend >= 0) {
                fence_proxy_async();
                mbar_arrive(&bars[T_OP_FULL + pend]);
                pend = -1;
            }
        };
        auto wait_full = [&](int slot, uint32_t gg) {
            if (!mbar_try_wait(&bars[T_LD_FULL + slot], (gg / kNLd) & 1)) {
                publish();
                mbar_wait(&bars[T_LD_FULL + slot], (gg / kNLd) & 1);
            }
        };
        auto convert_qk = [&](int count) {
            for (int e = 0; e < count; ++e, ++g) {
                const int slot = g % kNLd;
                wait_full(slot, g);
                convert_slot_inplace<LK>(smem + S::off_ld + slot * T::kSlot, t, publish);
                pend = slot;
            }
        };
        // chunk n of the item, accumulator-group counter gcn (global over the items of this CTA)
        auto convert_v = [&](int n, uint32_t gcn) {
            const int slot = g % kNLd;
            if ((n & 1) == half) {
                wait_full(slot, g);
                // 16-key units of this warp (tcgen05.st column addresses stay multiples of 8): units [u0, u0 + nu).
                // Row j of the swizzled tile keeps its 16-byte chunk c at ((c ^ (j & 7)) * 16) and (u0 * 16 + i) & 7 == i & 7.
                // (bits 4-6 of `src` are exactly (lane >> 2) << 4: the chunk swizzle of row i is one XOR with a constant)
                const uint32_t src = smem_u32(smem + S::off_ld + slot * T::kSlot + box * T::kTile) + u0 * 16 * 128 + (lane & 3) * 4 +
                                     ((lane >> 2) << 4);
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    if (uu < nu) {                              // one 16-key unit at a time: 16 loads in flight, then split
                        float x[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[i]) : "r"((src ^ ((i & 7) << 4)) + (uu * 16 + i) * 128));
                        if (uu == 0) publish();
#pragma unroll
                        for (int i = 0; i < 8; ++i) split2(x[2 * i], x[2 * i + 1], hi[8 * uu + i], lo[8 * uu + i]);
                    }
                }
                // the slot goes back BEFORE the wait for the tensor-memory buffer: the registers are one more stage of prefetch
                // (the ring is only kNLd slots deep and Q', K' of the next item hold two of them for a while)
                __syncwarp();
                if (lane == 0 && atomicAdd(&rd_cnt[slot], 1u) == 7u) {       // the 8 warps of this half have read the slot
                    rd_cnt[slot] = 0u;
                    mbar_arrive(&bars[T_LD_EMPTY + slot]);
                }
                const uint32_t vb = gcn & 1;
                mbar_wait(&bars[T_V_EMPTY + vb], ((gcn >> 1) & 1) ^ 1);      // the MMAs of group gcn - 2 have read this buffer
                tc_fence_after();
                const uint32_t ch = tl + TM::kV + vb * LK + u0 * 8, cl = ch + LK / 2;
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    if (uu < nu) {
                        tmem_st8(ch + uu * 8, hi + 8 * uu);
                        tmem_st8(cl + uu * 8, lo + 8 * uu);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[T_V_FULL + vb]);
            }
            ++g;
        };
        if (nk > 0) convert_qk(2);
        uint32_t gcn = 0;
        for (int k = 0; k < nk; ++k)
            for (int n = 0; n < NCH; ++n) {
                if (n == qkpos && k + 1 < nk) convert_qk(2);
                convert_v(n, gcn + (uint32_t)(n >> 1));
                if (n == NCH - 1) gcn += (uint32_t)NG;
            }
        publish();
    } else if (warp >= 4 && warp < kWarpEpiB) {
        reg_inc<kRegsSoftT>();
        // =============================== softmax group (128 threads, TMEM lane == query pixel) ===============================
        const int r = tid - 128;
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        pdl_wait();                                                // parts come from the statistics kernel
        for (int k = 0; k < nk; ++k) {
            const Item it = item_of(k);
            const bool rvalid = r < it.lq;
            float lse2 = 0.f;
            if (rvalid) {
                const long pix = item_pixel(p.sp, it, r);
                float m = -INFINITY;
                for (int i = 0; i < p.sp.nparts; ++i) m = fmaxf(m, __ldcg(p.parts + (long)i * p.npix + pix));
                float s = 0.f;
                for (int i = 0; i < p.sp.nparts; ++i) s += exp2f(__ldcg(p.parts + (long)i * p.npix + pix) - m);
                lse2 = m + log2f(s);
                if (!it.col && it.ik == 0) p.lse[pix] = lse2 * kLn2;
            }
            const int self = it.col ? it.q0 + r - it.k0 : -1;          // masked key of this query (column branch only)
            const int sw0 = it.col ? it.q0 - it.k0 + 32 * (warp & 3) : -(1 << 20);   // see cca_tc_fwd.cu: warp-uniform mask test
            const float nlse = rvalid ? -lse2 : -INFINITY;             // rows beyond the query tile: P = exp2(-inf) = 0
            uint8_t *ph = smem + S::off_p + (k & 1) * T::kP + r * 16, *pl = ph + T::kPP * T::kPlane;
            mbar_wait(&bars[T_S_FULL], k & 1);
            mbar_wait(&bars[T_P_EMPTY + (k & 1)], ((k >> 1) & 1) ^ 1);    // the MMAs of item k-2 have finished reading these planes
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < LK; c0 += 16) {
                float s[16];
                tmem_ld16(tl + TM::kS + c0, reinterpret_cast<uint32_t *>(s));
                tmem_ld_wait();
                const bool masked = (c0 + 16 > it.lk) || (c0 + 16 > sw0 && c0 < sw0 + 32);
                if (!masked) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) s[e] = exp2f(fmaf(s[e], kLog2e, nlse));
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int j = c0 + e;
                        s[e] = (j < it.lk && j != self) ? exp2f(fmaf(s[e], kLog2e, nlse)) : 0.f;
                    }
                }
                if (r < LK) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 hi, lo;
                        split8(s + 8 * h, hi, lo);
                        *reinterpret_cast<uint4 *>(ph + (c0 / 8 + h) * T::kPlane) = hi;
                        *reinterpret_cast<uint4 *>(pl + (c0 / 8 + h) * T::kPlane) = lo;
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&bars[T_P_FULL + (k & 1)]);
            mbar_arrive(&bars[T_S_EMPTY]);
        }
    } else {
        // =============================== epilogue groups (2 x 128 threads, TMEM lane == channel of the group) ===============================
        // Group A (warps 0-3) drains accumulator half a, group B (warps 28-31) half b.  ~220 four-byte stores per thread and item,
        // each a warp-wide 128-byte line: what matters is the length of the dependent instruction chain of a warp, so the code is
        // kept lean (16 columns at a time, the next 16 already in flight from tensor memory, two independent running addresses).
        reg_dec<kRegsEpiT>();
        const int hb = warp >= kWarpEpiB ? 1 : 0;
        const int etid = (warp & 3) * 32 + lane;                   // TMEM lane = channel inside the 128-channel group
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int ncols = hb ? kNB : kNA, q0 = hb * kNA;
        const uint32_t src = tl + TM::kO + q0;
        const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first(), pol_normal = l2_policy_evict_normal();
        const bool one_tile = p.sp.col.nt == 1 && p.sp.row.nt == 1;
        uint32_t gc = 0;
        pdl_wait();                                                // statistics kernel complete: counters cleared, the output is ours
        for (int k = 0; k < nk; ++k) {
            const Item it = item_of(k);
            const bool prod = is_producer(it);
            // producers' tiles are added onto by the sample's consumers: keep them in L2; a consumer's add is the last touch
            const uint64_t pol = !p.hints ? pol_normal : (prod ? pol_keep : (one_tile ? pol_stream : pol_normal));
            const uint64_t qb = (uint64_t)(it.col ? p.sp.W : 1) * p.C * 4;     // bytes between consecutive query pixels
            // byte address of (query q0, channel etid of group 0)
            const uint64_t o0 = reinterpret_cast<uint64_t>(p.out + item_pixel(p.sp, it, 0) * (long)p.C + etid) + (uint64_t)q0 * qb;
            const int nq = it.lq - q0;                             // valid query pixels of this half (may be <= 0 or > ncols)
            bool waited = prod;
            // 16 columns (query pixels) of this thread's channel; `a` = byte address of the first one
            auto emit16 = [&](const float *o, uint64_t a, int n) {
                uint64_t a0 = a, a1 = a + qb;
                const uint64_t qb2 = 2 * qb;
                if (n >= 16) {
                    if (prod) {
#pragma unroll
                        for (int e = 0; e < 16; e += 2, a0 += qb2, a1 += qb2) { st_global_f32(a0, o[e], pol); st_global_f32(a1, o[e + 1], pol); }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e += 2, a0 += qb2, a1 += qb2) { red_global_add_f32(a0, o[e], pol); red_global_add_f32(a1, o[e + 1], pol); }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e, a0 += qb) {
                        if (e < n) {
                            if (prod) st_global_f32(a0, o[e], pol);
                            else red_global_add_f32(a0, o[e], pol);
                        }
                    }
                }
            };
            for (int g = 0; g < NG; ++g, ++gc) {
                float a[16], b[16];
                mbar_wait(&bars[T_O_FULL + hb], gc & 1);
                tc_fence_after();
                tmem_ld16(src, reinterpret_cast<uint32_t *>(a));
                if (!waited) {                                     // every producer of this sample has stored its tiles (both halves)
                    if (etid == 0) wait_count(p.cdone + it.b, 2u * (unsigned)p.sp.seg0);
                    named_bar_sync(6 + hb, 128);
                    waited = true;
                }
                uint64_t dst = o0 + (uint64_t)g * (kGroupCh * 4);
#pragma unroll 1
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    tmem_ld_wait16(reinterpret_cast<uint32_t *>(a));
                    if (c0 + 16 < ncols) {
                        tmem_ld16(src + c0 + 16, reinterpret_cast<uint32_t *>(b));
                    } else {                                       // the accumulator half is in registers: the MMAs may overwrite it
                        tc_fence_before();
                        mbar_arrive(&bars[T_O_EMPTY + hb]);
                    }
                    emit16(a, dst, nq - c0);
                    dst += 16 * qb;
                    if (c0 + 16 < ncols) {
                        tmem_ld_wait16(reinterpret_cast<uint32_t *>(b));
                        if (c0 + 32 < ncols) {
                            tmem_ld16(src + c0 + 32, reinterpret_cast<uint32_t *>(a));
                        } else {
                            tc_fence_before();
                            mbar_arrive(&bars[T_O_EMPTY + hb]);
                        }
                        emit16(b, dst, nq - c0 - 16);
                        dst += 16 * qb;
                    }
                }
            }
            if (prod) {      // publish: the group's stores happen-before the barrier, the fence of its first thread is cumulative
                named_bar_sync(6 + hb, 128);
                if (etid == 0) {
                    __threadfence();
                    atomicAdd(p.cdone + it.b, 1u);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int LK>
cudaError_t launch_fwdt(const void *q, const void *k, const void *v, void *out, float *lse, const float *parts, unsigned int *cdone,
                        Dims d, cudaStream_t st, const char **why)
{
    CUtensorMap m[6];
    const void *base[3] = {q, k, v};
    const int ch[3] = {d.Cq, d.Cq, d.C};
    for (int t = 0; t < 3; ++t)
        for (int r = 0; r < 2; ++r)
            if (!get_map(&m[2 * t + r], base[t], d.B, d.H, d.W, ch[t], LK, r == 0, false)) {
                if (why) *why = "cuTensorMapEncodeTiled failed";
                return cudaErrorInvalidValue;
            }
    FwdTParams p;
    p.sp = make_space(d.B, d.H, d.W);
    p.C = d.C; p.Cq = d.Cq;
    p.npix = (long)d.B * d.H * d.W;
    p.parts = parts; p.lse = lse; p.out = reinterpret_cast<float *>(out); p.cdone = cdone;
    p.lag = tc_lag() != 0 ? 1 : 0;
    p.hints = tc_l2_hints();
    auto kern = cca_tc_fwdt_kernel<LK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdTSmem<LK>::kBytes);
    if (e != cudaSuccess) return e;
    const int sms = sm_count();
    const int grid = p.sp.total < sms ? p.sp.total : sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreadsT); cfg.dynamicSmemBytes = FwdTSmem<LK>::kBytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tc_pdl() ? 1 : 0;     // may start ahead of the statistics kernel's completion (griddepcontrol.wait inside)
    e = cudaLaunchKernelEx(&cfg, kern, m[0], m[1], m[2], m[3], m[4], m[5], p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

// fp32, C a multiple of 128: the channel-major values kernel (statistics pre-pass already launched by the caller)
bool tc_forward_t_supported(Dims d, int dtype) { return dtype == CCA_F32 && d.C % kGroupCh == 0 && tc::shape_supported(d, dtype); }

cudaError_t tc_forward_values_t(const void *q, const void *k, const void *v, void *out, float *lse, const float *parts,
                                unsigned int *cdone, Dims d, int lk, cudaStream_t st, const char **why)
{
    return lk == 80 ? launch_fwdt<80>(q, k, v, out, lse, parts, cdone, d, st, why)
                    : launch_fwdt<112>(q, k, v, out, lse, parts, cdone, d, st, why);
}

}  // namespace cca
