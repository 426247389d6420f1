// tcgen05 / TMA forward kernel of criss-cross attention for sm_100a (channels-last tensors).
//
// Layout: q,k [B,H,W,Cq], v,out [B,H,W,C] (torch channels_last).  In this layout an image row and an
// image column are the same object -- L pixels with a fixed pixel stride, each pixel's channels
// contiguous -- so ONE kernel serves both branches of cc_attention/functions.py:38-47:
//   column items (self entry masked, functions.py:38): out <- V P_c / l_c, stats <- (m_c, l_c)
//   row items    (functions.py:39): flash-style merge with the column result -> out, lse  (functions.py:40-47)
//
// ONE persistent launch processes both kinds of items: the item list is  col(0) | col(1) row(0) | col(2) row(1) ...
// (all column lines of a sample; its row lines one block later), CTA c takes items c, c+grid, ...  A row line of
// sample b needs every column line of b: column items bump a per-sample counter after their last TMA store
// has completed; row items spin on it (all earlier items are owned by running CTAs, so this cannot deadlock).
// Scheduling a sample's rows shortly after its columns keeps q,k,v and the partial output in the 126 MB L2.
//
// Roles (warpgroups, registers rebalanced with setmaxnreg), software-pipelined across items:
//   TMA producer (1 thread)   : 4-D tiled loads [LK px][32 ch] fp32, SWIZZLE_128B, OOB pixels zero-filled,
//                               3-slot ring: Q, K of the NEXT item, then the V chunks (64 ch) of the current.
//   converter warps (256 thr) : fp32 -> bf16 hi + bf16 lo split (x = hi + lo to ~2^-17), written as UMMA
//                               canonical no-swizzle operand planes [8-channel chunk][pixel][16 B].
//   MMA warp (elect.sync)     : S = Q K^T (SS, 3 bf16 MMAs per k-step: hi*hi + hi*lo + lo*hi, fp32 in TMEM),
//                               then per V chunk O = P V with A = P read from TMEM (TS): the probabilities
//                               never touch shared memory, so the MMA is not SMEM-bandwidth bound.
//   softmax group (128 thr)   : TMEM -> registers (one query pixel per thread), exp2 softmax, P split hi/lo
//                               and written back to TMEM as packed bf16 pairs, per-pixel scales / stats / lse.
//   epilogue group (128 thr)  : per V chunk TMEM -> scale/merge -> swizzled smem tile -> TMA store
//                               (3 staging slots; the column partial is prefetched into the slot by TMA).
// All hand-offs are mbarriers (TMA complete_tx, tcgen05.commit, thread arrives).
#include <cstdlib>

#include "cca_tc_common.cuh"

namespace cca {
namespace {
using namespace tc;

constexpr int kTmemCols = 512;      // S [0,128)   P double buffer (hi/lo) [128,256) [256,384)   O ring 2 x 64 [384,512)
constexpr int kTmemP = 128, kTmemO = 384, kNOB = 2;
// the in-place conversion keeps a whole 128 B row (split into hi/lo) live across a barrier: converters get 88 registers; the
// softmax group streams its row from TMEM 16 columns at a time and needs few
constexpr int kRegsSoft = 104, kRegsEpi = 128, kRegsConvF = 88;
static_assert(reg_pool_ok(kRegsSoft, kRegsEpi, kRegsConvF), "setmaxnreg pool");
constexpr int kNOut = 3;            // staging slots
constexpr int kNLdMax = 6;          // load ring: 5 slots (fp32: 28 KB each, converted to bf16 hi/lo planes IN PLACE) or 6 (bf16: 14 KB each)

enum { MODE_FUSED = 0, MODE_DYNAMIC = 1, MODE_COL_ONLY = 2, MODE_ROW_ONLY = 3 };

struct FwdParams {
    int B, H, W, C, Cq;
    int mode;              // MODE_DYNAMIC: one launch, lines claimed from two global queues (a row line only once its
                           // sample's column lines are complete); MODE_FUSED: one launch, static interleaved order;
                           // *_ONLY: one pass per launch
    unsigned int *sched;   // [0] next column line to hand out, [1] first sample that may still have row lines (MODE_DYNAMIC)
    unsigned int *rown;    // [B] next row line of each sample (MODE_DYNAMIC)
    float2 *stats;         // [B,H,W] (m_c, l_c) of the column branch
    float *lse;            // [B,H,W]
    unsigned int *done;    // [B] column lines completed (MODE_FUSED only)
    int sync;              // two launches, the row pass overlapping the tail of the column pass (programmatic dependent launch):
                           // column lines are counted in done[] as in the fused modes and a row line waits for its own sample only
    int hints;             // L2 eviction hints on the bulk copies: column lines of samples >= keep_from are kept (evict_last:
    int keep_from;         // the row pass, which walks the samples backwards, finds them in L2), everything else streams
    long long *dbg;        // optional timeline buffer (4 roles x 512 stamps), CTA 0 only; nullptr in production
};

struct Item { int col, b, i, L; };

__device__ __forceinline__ int total_items(const FwdParams &p)
{
    return p.mode <= MODE_DYNAMIC ? p.B * (p.W + p.H) : (p.mode == MODE_COL_ONLY ? p.B * p.W : p.B * p.H);
}
__device__ __forceinline__ Item decode_item(const FwdParams &p, int idx)
{
    Item it;
    if (p.mode == MODE_FUSED) {
        // order: col(0) | col(1) row(0) | col(2) row(1) | ... | row(B-1)  -- the rows of a sample trail its columns by one
        // block of column lines, so a row line (almost) never has to wait for the column lines it depends on, while
        // the partial output and q,k,v of the sample are still in L2
        if (idx < p.W) { it.col = 1; it.b = 0; it.i = idx; }
        else {
            const int per = p.W + p.H, x = idx - p.W;
            const int j = x / per, rem = x - j * per;
            if (j < p.B - 1 && rem < p.W) { it.col = 1; it.b = j + 1; it.i = rem; }
            else if (j < p.B - 1) { it.col = 0; it.b = j; it.i = rem - p.W; }
            else { it.col = 0; it.b = p.B - 1; it.i = rem; }
        }
    } else {
        it.col = p.mode == MODE_COL_ONLY;
        const int nl = it.col ? p.W : p.H;
        it.b = idx / nl;
        it.i = idx - it.b * nl;
        // the row pass walks the samples backwards: what the column pass touched last (v, q, k and the partial output of
        // the last samples) is still in L2 when the row pass starts
        if (!it.col && !p.sync) it.b = p.B - 1 - it.b;       // (an overlapped row pass starts with the samples completed first)
    }
    it.L = it.col ? p.H : p.W;
    return it;
}

#define CCA_STAMP(role)                                                                          \
    do {                                                                                         \
        if (p.dbg && blockIdx.x == 0 && dbg_n < 512) p.dbg[(role) * 512 + dbg_n++] = clock64();  \
    } while (0)

template <int LK, bool BF> struct FwdSmem {
    using T = Tiles<LK, BF>;
    static constexpr int kNLd = BF ? 6 : 5;
    static constexpr int off_ld = 0;                          // kNLd load slots; every slot is the UMMA operand itself: a bf16 tile as
                                                              // loaded, an fp32 tile once the converters have rewritten it in place
    static constexpr int off_out = off_ld + kNLd * T::kSlot;  // kNOut out slots
    static constexpr int off_tail = off_out + kNOut * T::kSlot; // pad: an M=128 MMA reads (128 - LK) rows past the last plane of a slot
                                                              // (the rows land in the next slot / the staging slots; they only feed S rows >= LK)
    static constexpr int off_scale = off_tail + (128 - LK) * 16;          // float2 (sa, sb) [2][128]: softmax group -> epilogue group
    static constexpr int off_bar = off_scale + 2 * 128 * 8;
    static constexpr int kBytes = off_bar + 8 * 54 + 2 * 8 * 16 + 32;
    static_assert(kBytes <= 232448, "shared memory budget");
};

enum { B_LD_FULL = 0, B_LD_EMPTY = 6, B_OP_FULL = 12, B_S_FULL = 18, B_S_EMPTY = 19, B_P_FULL = 20,
       B_P_EMPTY = 22, B_O_FULL = 24, B_O_EMPTY = 26, B_OUT_FULL = 28, B_SC_EMPTY = 31, B_SC_FULL = 33, B_STAGED = 35, B_EARLY = 38,
       B_FINAL = 46, B_COUNT = 54 };

// Work distribution inside a CTA: the producer thread is the scheduler.  For the item that follows item j-1 it publishes
//   EARLY[j] when it reaches the slot where the next item's Q/K would be slipped into the ring (kind: 1 = item, 2 = no more
//            work, 0 = not decided yet -- e.g. only row lines are left and their sample's columns are not complete), and
//   FINAL[j] the definitive answer (same as EARLY unless that was 0; then after the current item's last chunk).
// Every role reads these records instead of computing a static schedule.
struct Rec { int kind, col, b, i; };
constexpr int kRecRing = 8;      // the producer never runs more than ~3 items ahead of the slowest role

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_done(const unsigned int *cnt, unsigned int need)
{
    unsigned int spins = 0;
    while (ld_acquire(cnt) < need) {
        __nanosleep(64);
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int LK, bool BF>
__global__ void __launch_bounds__(kThreads, 1)
cca_tc_fwd_kernel(const __grid_constant__ CUtensorMap mqc, const __grid_constant__ CUtensorMap mqr,
                  const __grid_constant__ CUtensorMap mkc, const __grid_constant__ CUtensorMap mkr,
                  const __grid_constant__ CUtensorMap mvc, const __grid_constant__ CUtensorMap mvr,
                  const __grid_constant__ CUtensorMap moc, const __grid_constant__ CUtensorMap mor, FwdParams p)
{
    using T = Tiles<LK, BF>;
    using S = FwdSmem<LK, BF>;
    constexpr int TERMS = BF ? 1 : 3;
    constexpr int kNLd = S::kNLd;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::off_bar + 8 * B_COUNT + 2 * kRecRing * 16);
    float2 *scale = reinterpret_cast<float2 *>(smem + S::off_scale);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NCH = p.C / kNC;
    const int KQ = p.Cq / 16;                 // k-steps of the S MMA
    const int n_items = total_items(p);
    Rec *rec_early = reinterpret_cast<Rec *>(smem + S::off_bar + 8 * B_COUNT);
    Rec *rec_final = rec_early + kRecRing;
    // ring order:  Q0 K0 | V0[0..qkpos) Q1 K1 V0[qkpos..NCH) | V1[0..qkpos) Q2 K2 ...   (Q,K of the next item are slipped in
    // after the first chunks of the current one, so neither S(k+1) nor the first P V chunk of an item waits for the other)
    const int qkpos = NCH >= 3 ? 2 : NCH - 1;

    if (tid == 0) {
        for (int i = 0; i < kNLd; ++i) {
            mbar_init(&bars[B_LD_FULL + i], 1); mbar_init(&bars[B_LD_EMPTY + i], 1); mbar_init(&bars[B_OP_FULL + i], kConvThreads);
        }
        for (int i = 0; i < kNOB; ++i) { mbar_init(&bars[B_O_FULL + i], 1); mbar_init(&bars[B_O_EMPTY + i], 128); }
        for (int i = 0; i < kNOut; ++i) { mbar_init(&bars[B_OUT_FULL + i], 1); mbar_init(&bars[B_STAGED + i], 128); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars[B_SC_EMPTY + i], 128); mbar_init(&bars[B_SC_FULL + i], 128); }
        for (int i = 0; i < kRecRing; ++i) { mbar_init(&bars[B_EARLY + i], 1); mbar_init(&bars[B_FINAL + i], 1); }
        mbar_init(&bars[B_S_FULL], 1); mbar_init(&bars[B_S_EMPTY], 128);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars[B_P_FULL + i], 128); mbar_init(&bars[B_P_EMPTY + i], 1); }
        fence_mbar_init();
        prefetch_tmap(&mqc); prefetch_tmap(&mqr); prefetch_tmap(&mkc); prefetch_tmap(&mkr);
        prefetch_tmap(&mvc); prefetch_tmap(&mvr); prefetch_tmap(&moc); prefetch_tmap(&mor);
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // Programmatic dependent launch: the row pass is launched while the column pass is still draining; its CTAs load q,k,v,
    // compute S, the row softmax and the first P V products right away and only wait (pdl_wait) where they first touch what
    // the column pass produces: the column statistics (softmax group) and the partial output (store warp).
    pdl_launch_dependents();

    // ---- item records (see Rec): blocking / non-blocking readers used by every role
    auto rec_item = [&](const Rec &r) { Item it; it.col = r.col; it.b = r.b; it.i = r.i; it.L = r.col ? p.H : p.W; return it; };
    auto early_kind = [&](int j, Item &it) {                 // blocking on the early decision only
        mbar_wait(&bars[B_EARLY + (j & (kRecRing - 1))], (j / kRecRing) & 1);
        const Rec r = rec_early[j & (kRecRing - 1)];
        if (r.kind == 1) it = rec_item(r);
        return r.kind;
    };
    auto get_item = [&](int j, Item &it) {                   // blocking on the definitive answer; false = no more work
        mbar_wait(&bars[B_FINAL + (j & (kRecRing - 1))], (j / kRecRing) & 1);
        const Rec r = rec_final[j & (kRecRing - 1)];
        if (r.kind == 1) it = rec_item(r);
        return r.kind == 1;
    };
    auto try_item = [&](int j, Item &it) {                   // non-blocking: 1 item, 2 end, -1 not known yet
        if (!mbar_try_wait(&bars[B_FINAL + (j & (kRecRing - 1))], (j / kRecRing) & 1)) return -1;
        const Rec r = rec_final[j & (kRecRing - 1)];
        if (r.kind == 1) it = rec_item(r);
        return r.kind;
    };

    if (warp >= kWarpProducer) {
        reg_dec<kRegsMisc>();
        if (warp == kWarpProducer) {
            // =============================== TMA producer ===============================
            if (lane == 0) {
                uint32_t g = 0;
                int dbg_n = 0;
                const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
                auto emit = [&](const CUtensorMap *mc, const CUtensorMap *mr, int c0, const Item &it) {
                    const CUtensorMap *m = it.col ? mc : mr;
                    const uint64_t pol = it.col && it.b >= p.keep_from ? pol_keep : pol_stream;
                    const int cw = it.col ? it.i : 0, ch = it.col ? 0 : it.i;
                    const int slot = g % kNLd;
                    mbar_wait(&bars[B_LD_EMPTY + slot], ((g / kNLd) & 1) ^ 1);
                    CCA_STAMP(0);
                    uint8_t *dst = smem + S::off_ld + slot * T::kSlot;
                    mbar_expect_tx(&bars[B_LD_FULL + slot], T::kSlot);
                    if (p.hints) {
                        tma_load_4d(dst, m, &bars[B_LD_FULL + slot], c0, cw, ch, it.b, pol);
                        if constexpr (!BF) tma_load_4d(dst + T::kTile, m, &bars[B_LD_FULL + slot], c0 + 32, cw, ch, it.b, pol);
                    } else {
                        tma_load_4d(dst, m, &bars[B_LD_FULL + slot], c0, cw, ch, it.b);
                        if constexpr (!BF) tma_load_4d(dst + T::kTile, m, &bars[B_LD_FULL + slot], c0 + 32, cw, ch, it.b);
                    }
                    ++g;
                };
                // ---- scheduler: hand out the next line of this CTA
                int static_k = 0;
                auto try_fetch = [&](Item &it) -> int {            // 1 item, 2 no more work, 0 nothing available right now
                    if (p.mode != MODE_DYNAMIC) {
                        const int idx = (int)blockIdx.x + static_k * (int)gridDim.x;
                        if (idx >= n_items) return 2;
                        it = decode_item(p, idx);
                        ++static_k;
                        return 1;
                    }
                    // dynamic: row lines first (in sample order, only of samples whose column lines are all published),
                    // else the next column line.  All claims are atomicAdd on per-queue counters: no CAS retry storms.
                    const unsigned total_cols = (unsigned)(p.B * p.W);
                    unsigned b = ld_acquire(p.sched + 1);             // first sample that may still have row lines to hand out
                    bool rows_left = false;
                    for (; b < (unsigned)p.B; ++b) {
                        if (ld_acquire(p.done + b) < (unsigned)p.W) { rows_left = true; break; }   // not ready yet (keep sample order)
                        const unsigned r = atomicAdd(p.rown + b, 1u);
                        if (r < (unsigned)p.H) { it.col = 0; it.b = (int)b; it.i = (int)r; it.L = p.W; return 1; }
                        atomicMax(p.sched + 1, b + 1);                // this sample's rows are all handed out
                    }
                    if (ld_acquire(p.sched + 0) < total_cols) {
                        const unsigned c = atomicAdd(p.sched + 0, 1u);
                        if (c < total_cols) {
                            it.col = 1; it.b = (int)(c / (unsigned)p.W); it.i = (int)(c - it.b * p.W); it.L = p.H;
                            return 1;
                        }
                    }
                    return rows_left ? 0 : 2;                          // 0: only row lines are left and they are not ready yet
                };
                auto fetch_blocking = [&](Item &it) -> int {
                    unsigned spins = 0;
                    for (;;) {
                        const int kd = try_fetch(it);
                        if (kd) return kd;
                        __nanosleep(256);
                        if (++spins > (1u << 22)) __trap();
                    }
                };
                auto publish = [&](int base, Rec *arr, int j, int kind, const Item &it) {
                    Rec r; r.kind = kind; r.col = it.col; r.b = it.b; r.i = it.i;
                    arr[j & (kRecRing - 1)] = r;
                    mbar_arrive(&bars[base + (j & (kRecRing - 1))]);       // release: the record is visible to the waiters
                };
                Item cur;
                cur.col = cur.b = cur.i = cur.L = 0;
                int kind = fetch_blocking(cur);
                publish(B_EARLY, rec_early, 0, kind, cur);
                publish(B_FINAL, rec_final, 0, kind, cur);
                if (kind == 1) { emit(&mqc, &mqr, 0, cur); emit(&mkc, &mkr, 0, cur); }
                for (int k = 0; kind == 1; ++k) {
                    Item nxt = cur;
                    int nkind = -1;                                 // -1: not decided yet
                    for (int n = 0; n < NCH; ++n) {
                        if (n == qkpos) {
                            nkind = try_fetch(nxt);
                            publish(B_EARLY, rec_early, k + 1, nkind, nxt);
                            if (nkind != 0) publish(B_FINAL, rec_final, k + 1, nkind, nxt);
                            if (nkind == 1) { emit(&mqc, &mqr, 0, nxt); emit(&mkc, &mkr, 0, nxt); }
                        }
                        emit(&mvc, &mvr, n * kNC, cur);
                    }
                    if (nkind == 0) {                               // decided late: after the last chunk of the current item
                        nkind = fetch_blocking(nxt);
                        publish(B_FINAL, rec_final, k + 1, nkind, nxt);
                        if (nkind == 1) { emit(&mqc, &mqr, 0, nxt); emit(&mkc, &mkr, 0, nxt); }
                    }
                    kind = nkind;
                    cur = nxt;
                }
            }
        } else if (warp == kWarpMma) {
            // =============================== MMA issuer (whole warp, elect.sync inside) ===============================
            const uint32_t idesc_s = instr_desc(kFmtBF16, kFmtBF16, 128, LK, false, false);
            const uint32_t idesc_o = instr_desc(kFmtBF16, kFmtBF16, 128, kNC, false, true);
            uint32_t u = 0, oc = 0;
            int dbg_n = lane == 0 ? 0 : 512;
            // operand of ring item u = load slot u % kNLd: fp32 -> bf16 hi/lo planes written in place by the converters;
            //                                              bf16 -> the TMA tile itself, read with SWIZZLE_128B descriptors
            const uint32_t ld_base = smem_u32(smem + S::off_ld);
            auto wait_item = [&](uint32_t g) {
                mbar_wait(&bars[(BF ? B_LD_FULL : B_OP_FULL) + g % kNLd], (g / kNLd) & 1);
            };
            auto free_item = [&](uint32_t g) {
                commit_to(&bars[B_LD_EMPTY + g % kNLd]);
            };
            auto item_addr = [&](uint32_t g) { return ld_base + (g % kNLd) * T::kSlot; };
            auto issue_s = [&](int k) {            // S(k) = Q K^T from items u (Q) and u+1 (K)
                const uint32_t qb = item_addr(u), kb = item_addr(u + 1);
                wait_item(u); wait_item(u + 1);
                mbar_wait(&bars[B_S_EMPTY], (k & 1) ^ 1);
                tc_fence_after();
                CCA_STAMP(2);
                for (int ks = 0; ks < KQ; ++ks) {
                    if constexpr (BF) {
                        mma_split3<1>(tmem, smem_desc(qb + ks * 32, 16, 1024, kSw128), 0, smem_desc(kb + ks * 32, 16, 1024, kSw128), 0,
                                      idesc_s, ks > 0);
                    } else {
                        const uint32_t ao = ks * 2 * T::kPlane;
                        mma_split3<3>(tmem, smem_desc(qb + ao, T::kPlane, 128), smem_desc(qb + 8 * T::kPlane + ao, T::kPlane, 128),
                                      smem_desc(kb + ao, T::kPlane, 128), smem_desc(kb + 8 * T::kPlane + ao, T::kPlane, 128),
                                      idesc_s, ks > 0);
                    }
                }
                commit_to(&bars[B_S_FULL]);
                free_item(u); free_item(u + 1);
                u += 2;
            };
            Item it_unused;
            bool have = get_item(0, it_unused);
            if (have) issue_s(0);
            for (int k = 0; have; ++k) {
                int nkind = -1;
                CCA_STAMP(2);
                mbar_wait(&bars[B_P_FULL + (k & 1)], (k >> 1) & 1);
                tc_fence_after();
                CCA_STAMP(2);
                const uint32_t pbuf = tmem + kTmemP + (k & 1) * 128;
                for (int n = 0; n < NCH; ++n, ++u, ++oc) {
                    if (n == qkpos) {                              // same decision the producer took when it filled the ring
                        nkind = early_kind(k + 1, it_unused);
                        if (nkind == 1) issue_s(k + 1);
                    }
                    const uint32_t vb = item_addr(u);
                    const uint32_t ob = oc % kNOB;
                    wait_item(u);
                    mbar_wait(&bars[B_O_EMPTY + ob], ((oc / kNOB) & 1) ^ 1);
                    tc_fence_after();
                    CCA_STAMP(2);
                    if (elect_one()) {
                        const uint32_t d = tmem + kTmemO + ob * kNC;
#pragma unroll
                        for (int ks = 0; ks < LK / 16; ++ks) {
                            const uint32_t ph = pbuf + ks * 8, pl = ph + LK / 2;
                            if constexpr (BF) {
                                mma_f16_ts(d, ph, smem_desc(vb + ks * 2048, 16, 1024, kSw128), idesc_o, ks > 0);
                            } else {
                                const uint64_t vh = smem_desc(vb + ks * 256, 128, T::kPlane);
                                const uint64_t vl = smem_desc(vb + 8 * T::kPlane + ks * 256, 128, T::kPlane);
                                mma_f16_ts(d, ph, vh, idesc_o, ks > 0);
                                mma_f16_ts(d, ph, vl, idesc_o, true);
                                mma_f16_ts(d, pl, vh, idesc_o, true);
                            }
                        }
                    }
                    __syncwarp();
                    commit_to(&bars[B_O_FULL + ob]);
                    free_item(u);
                    CCA_STAMP(2);
                }
                commit_to(&bars[B_P_EMPTY + (k & 1)]);
                if (nkind == 0) {                                  // late decision: Q,K of the next item follow the last chunk
                    have = get_item(k + 1, it_unused);
                    if (have) issue_s(k + 1);
                } else have = nkind == 1;
            }
        } else if (warp == kWarpStore) {
            // =============================== store warp (one lane): staging slots <-> global ===============================
            // Per output chunk c (slot c % kNOut): make the slot ready for the epilogue group (a row item gets its column
            // partial prefetched by TMA, two chunks ahead), and once the group has staged the merged tile, TMA-store it.
            // Column items are published (per-sample counter) after their last store has completed.
            if (lane == 0) {
                uint32_t prep = 0;                         // chunks prepared so far (chunk c belongs to item c / NCH)
                const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
                int pub_k = 0;                             // items [0, pub_k) are published / need no publishing
                // a row chunk must not wait for a column item this warp has yet to publish (static fused order only:
                // the dynamic scheduler hands out a row line only after all column lines of its sample were published)
                auto can_prepare = [&](uint32_t c) {
                    Item nx;
                    if (try_item((int)(c / NCH), nx) != 1) return false;
                    if (nx.col || p.mode != MODE_FUSED) return true;
                    for (int j = pub_k; j < (int)(c / NCH); ++j) {
                        Item pj;
                        if (try_item(j, pj) == 1 && pj.col && pj.b == nx.b) return false;
                    }
                    return true;
                };
                auto prepare = [&](uint32_t c, const Item &it) {
                    const int n = c % NCH, os = c % kNOut;
                    if (it.col) { mbar_arrive(&bars[B_OUT_FULL + os]); return; }
                    if (p.mode <= MODE_DYNAMIC || p.sync) { wait_done(p.done + it.b, (unsigned)p.W); fence_proxy_async_all(); }
                    else pdl_wait();                           // (row pass launched ahead of the column pass's completion)
                    uint8_t *dst = smem + S::off_out + os * T::kSlot;
                    mbar_expect_tx(&bars[B_OUT_FULL + os], T::kSlot);
                    if (p.hints) {                                 // the partial is read exactly once
                        tma_load_4d(dst, &mor, &bars[B_OUT_FULL + os], n * kNC, 0, it.i, it.b, pol_stream);
                        if constexpr (!BF) tma_load_4d(dst + T::kTile, &mor, &bars[B_OUT_FULL + os], n * kNC + 32, 0, it.i, it.b, pol_stream);
                    } else {
                        tma_load_4d(dst, &mor, &bars[B_OUT_FULL + os], n * kNC, 0, it.i, it.b);
                        if constexpr (!BF) tma_load_4d(dst + T::kTile, &mor, &bars[B_OUT_FULL + os], n * kNC + 32, 0, it.i, it.b);
                    }
                };
                Item it, prev;
                bool has_prev = false;
                prev.col = prev.b = prev.i = prev.L = 0;
                for (int k = 0; get_item(k, it); ++k) {
                    const CUtensorMap *mo = it.col ? &moc : &mor;
                    const int cw = it.col ? it.i : 0, ch = it.col ? 0 : it.i;
                    for (int n = 0; n < NCH; ++n) {
                        const uint32_t c = (uint32_t)k * NCH + n;
                        const int os = c % kNOut;
                        while (prep <= c) {                        // not prefetched: prepare the current chunk now
                            tma_store_wait_read<1>();
                            Item pi = it;
                            if ((int)(prep / NCH) != k) get_item((int)(prep / NCH), pi);
                            prepare(prep, pi);
                            ++prep;
                        }
                        mbar_wait(&bars[B_STAGED + os], (c / kNOut) & 1);
                        const uint8_t *slot = smem + S::off_out + os * T::kSlot;
                        if (p.hints) {
                            const uint64_t pol = it.col && it.b >= p.keep_from ? pol_keep : pol_stream;
                            tma_store_4d(mo, slot, n * kNC, cw, ch, it.b, pol);
                            if constexpr (!BF) tma_store_4d(mo, slot + T::kTile, n * kNC + 32, cw, ch, it.b, pol);
                        } else {
                            tma_store_4d(mo, slot, n * kNC, cw, ch, it.b);
                            if constexpr (!BF) tma_store_4d(mo, slot + T::kTile, n * kNC + 32, cw, ch, it.b);
                        }
                        tma_store_commit();
                        if (p.sync && n == 0 && has_prev && prev.col) {
                            // publish the previous column line: its last store was committed a chunk period ago
                            tma_store_wait_all<1>();
                            fence_proxy_async_all();
                            __threadfence();
                            atomicAdd(p.done + prev.b, 1u);
                        }
                        if (n == NCH - 1) { prev = it; has_prev = true; }
                        if (n == NCH - 1) {
                            if (it.col && p.mode <= MODE_DYNAMIC) {
                                // publish this column line: its stores (async proxy) and the stats written by the softmax group
                                tma_store_wait_all<0>();
                                fence_proxy_async_all();
                                __threadfence();
                                atomicAdd(p.done + it.b, 1u);
                            }
                            pub_k = k + 1;
                        }
                        // keep the ring of kNOut slots full: chunk c+kNOut reuses the slot of the store just issued, as soon as
                        // that store has been read out of shared memory (a few hundred cycles; this lane has nothing else to do
                        // until the next chunk is staged).  The partial of a row item is thus in flight two full chunk periods
                        // before the epilogue needs it -- one period did not cover the L2 latency of the tile.
                        while (prep <= c + kNOut && can_prepare(prep)) {
                            Item pi;
                            try_item((int)(prep / NCH), pi);
                            if (prep == c + kNOut) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
                            prepare(prep, pi);
                            ++prep;
                        }
                    }
                }
                tma_store_wait_all<0>();
                if (p.sync && has_prev && prev.col) {         // the last column line of this CTA
                    fence_proxy_async_all();
                    __threadfence();
                    atomicAdd(p.done + prev.b, 1u);
                }
            }
        }
    } else if (warp >= kWarpConv0) {
        // =============================== converters (256 threads) ===============================
        reg_dec<kRegsConvF>();
        const int t = tid - kWarpConv0 * 32;
        int dbg_n = t == 0 ? 0 : 512;
        uint32_t g = 0;
        auto convert_next = [&](int count) {                       // the next `count` ring slots, whatever they hold
            for (int e = 0; e < count; ++e, ++g) {
                const int slot = g % kNLd;
                mbar_wait(&bars[B_LD_FULL + slot], (g / kNLd) & 1);
                CCA_STAMP(1);
                convert_slot_inplace<LK>(smem + S::off_ld + slot * T::kSlot, t);
                fence_proxy_async();
                mbar_arrive(&bars[B_OP_FULL + slot]);
                CCA_STAMP(1);
            }
        };
        if constexpr (!BF) {                                       // bf16 tiles need no conversion
            Item it;
            bool have = get_item(0, it);
            if (have) convert_next(2);                             // Q, K of the first item
            for (int k = 0; have; ++k) {                           // mirrors the ring order chosen by the producer
                int nkind = -1;
                for (int n = 0; n < NCH; ++n) {
                    if (n == qkpos) {
                        nkind = early_kind(k + 1, it);
                        if (nkind == 1) convert_next(2);
                    }
                    convert_next(1);
                }
                if (nkind == 0) {
                    have = get_item(k + 1, it);
                    if (have) convert_next(2);
                } else have = nkind == 1;
            }
        }
    } else if (warp >= 4) {
        // =============================== softmax group (128 threads, TMEM lane == query pixel) ===============================
        reg_inc<kRegsSoft>();
        const int r = tid - 128;
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        int dbg_n = r == 0 ? 0 : 512;
        Item it;
        for (int k = 0; get_item(k, it); ++k) {
            const bool rvalid = r < it.L;
            CCA_STAMP(3);
            mbar_wait(&bars[B_S_FULL], k & 1);
            tc_fence_after();
            CCA_STAMP(3);
            // two streaming passes over the S row in TMEM, 16 columns at a time (small loops instead of a 112-element
            // register array: the instruction footprint matters -- the unrolled version cost ~16K cycles on its first run)
            // ---- pass 1: row max of the valid, unmasked logits (log2 units)
            float m = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < LK; c0 += 16) {
                float s[16];
                tmem_ld16(tl + c0, reinterpret_cast<uint32_t *>(s));
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int j = c0 + e;
                    const bool ok = j < it.L && !(it.col && j == r);
                    m = fmaxf(m, ok ? s[e] * kLog2e : -INFINITY);
                }
            }
            const float msub = (m == -INFINITY) ? 0.f : m;      // fully masked row (L == 1 in a column item)
            // ---- pass 2: P = exp2(s - m) -> TMEM as packed bf16 pairs (hi at [pdst, +LK/2), lo at [pdst+LK/2, +LK/2)), row sum
            CCA_STAMP(3);
            mbar_wait(&bars[B_P_EMPTY + (k & 1)], ((k >> 1) & 1) ^ 1);    // P V of item k-2 has finished reading this buffer
            tc_fence_after();
            const uint32_t pdst = tl + kTmemP + (k & 1) * 128;
            float l = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < LK; c0 += 16) {
                float s[16];
                tmem_ld16(tl + c0, reinterpret_cast<uint32_t *>(s));
                tmem_ld_wait();
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = c0 + 2 * e;
                    const bool ok0 = rvalid && j < it.L && !(it.col && j == r);
                    const bool ok1 = rvalid && j + 1 < it.L && !(it.col && j + 1 == r);
                    const float p0 = ok0 ? exp2f(s[2 * e] * kLog2e - msub) : 0.f;
                    const float p1 = ok1 ? exp2f(s[2 * e + 1] * kLog2e - msub) : 0.f;
                    l += p0 + p1;
                    if constexpr (BF) hi[e] = pack_bf16(p0, p1);
                    else split2(p0, p1, hi[e], lo[e]);
                }
                tmem_st8(pdst + c0 / 2, hi);
                if constexpr (!BF) tmem_st8(pdst + LK / 2 + c0 / 2, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_P_FULL + (k & 1)]);
            mbar_arrive(&bars[B_S_EMPTY]);
            CCA_STAMP(3);
            // ---- per-pixel statistics / merge scales (off the MMA's critical path now)
            float sa = 0.f, sb = 0.f;
            if (!it.col && (p.mode <= MODE_DYNAMIC || p.sync)) wait_done(p.done + it.b, (unsigned)p.W);   // column stats of this sample complete
            else if (!it.col) pdl_wait();                      // row pass launched ahead of the column pass's completion
            if (rvalid) {
                const long pix = it.col ? ((long)it.b * p.H + r) * p.W + it.i : ((long)it.b * p.H + it.i) * p.W + r;
                const float mn = m * kLn2;                       // natural-log units
                if (it.col) {
                    p.stats[pix] = make_float2(mn, l);
                    sa = l > 0.f ? 1.f / l : 0.f;
                } else {
                    const float2 pc = __ldcg(p.stats + pix);
                    const float mm = fmaxf(mn, pc.x);
                    const float ar = exp2f((mn - mm) * kLog2e);
                    const float ac = pc.y > 0.f ? exp2f((pc.x - mm) * kLog2e) : 0.f;
                    const float lt = ar * l + ac * pc.y;
                    sa = ar / lt; sb = ac * pc.y / lt;
                    p.lse[pix] = mm + logf(lt);
                }
            }
            mbar_wait(&bars[B_SC_EMPTY + (k & 1)], ((k >> 1) & 1) ^ 1);   // epilogue has consumed the scales of item k-2
            scale[(k & 1) * 128 + r] = make_float2(sa, sb);
            mbar_arrive(&bars[B_SC_FULL + (k & 1)]);          // release: scales and the stats / lse written above
        }
    } else {
        // =============================== epilogue group (128 threads, TMEM lane == query pixel) ===============================
        reg_inc<kRegsEpi>();
        const int r = tid;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t oc = 0;
        int dbg_n = tid == 0 ? 0 : 512;
        Item it;
        for (int k = 0; get_item(k, it); ++k) {
            float sa = 0.f, sb = 0.f;
            for (int n = 0; n < NCH; ++n, ++oc) {
                const int os = oc % kNOut;
                const uint32_t ob = oc % kNOB;
                uint8_t *slot = smem + S::off_out + os * T::kSlot;
                CCA_STAMP(4);
                mbar_wait(&bars[B_OUT_FULL + os], (oc / kNOut) & 1);     // slot free (column item) / column partial landed (row item)
                CCA_STAMP(4);
                mbar_wait(&bars[B_O_FULL + ob], (oc / kNOB) & 1);
                tc_fence_after();
                CCA_STAMP(4);
                if (n == 0) {
                    mbar_wait(&bars[B_SC_FULL + (k & 1)], (k >> 1) & 1);
                    const float2 sc = scale[(k & 1) * 128 + r];
                    sa = sc.x; sb = sc.y;
                    mbar_arrive(&bars[B_SC_EMPTY + (k & 1)]);
                }
                float o[kNC];
#pragma unroll
                for (int c0 = 0; c0 < kNC; c0 += 16) tmem_ld16(tl + kTmemO + ob * kNC + c0, reinterpret_cast<uint32_t *>(o + c0));
                tmem_ld_wait();
                CCA_STAMP(4);
                tc_fence_before();
                mbar_arrive(&bars[B_O_EMPTY + ob]);
                if (r < it.L) {                                          // rows >= L are clipped by the TMA store
                    uint8_t *row = slot + r * 128;
                    const int sw = r & 7;
                    if constexpr (BF) {
                        // bf16 staging tile: one 128-byte row = 64 channels = 8 chunks of 8 bf16
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint4 *dst = reinterpret_cast<uint4 *>(row + ((j ^ sw) * 16));
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = o[8 * j + e] * sa;
                            if (!it.col) {
                                const uint4 q = *dst;
                                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    v[2 * e] = fmaf(bf_lo(w[e]), sb, v[2 * e]);
                                    v[2 * e + 1] = fmaf(bf_hi(w[e]), sb, v[2 * e + 1]);
                                }
                            }
                            *dst = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                        }
                    } else if (it.col) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {                   // 16 chunks of 4 channels
                            float4 *dst = reinterpret_cast<float4 *>(row + (j >> 3) * T::kTile + (((j & 7) ^ sw) * 16));
                            *dst = make_float4(o[4 * j] * sa, o[4 * j + 1] * sa, o[4 * j + 2] * sa, o[4 * j + 3] * sa);
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {                    // two batches of 8: all loads of a batch before its stores
                            float4 q[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                q[j] = *reinterpret_cast<const float4 *>(row + h * T::kTile + ((j ^ sw) * 16));
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int e = 32 * h + 4 * j;
                                float4 v;
                                v.x = fmaf(q[j].x, sb, o[e] * sa);         v.y = fmaf(q[j].y, sb, o[e + 1] * sa);
                                v.z = fmaf(q[j].z, sb, o[e + 2] * sa);     v.w = fmaf(q[j].w, sb, o[e + 3] * sa);
                                *reinterpret_cast<float4 *>(row + h * T::kTile + ((j ^ sw) * 16)) = v;
                            }
                        }
                    }
                }
                CCA_STAMP(4);
                fence_proxy_async();
                mbar_arrive(&bars[B_STAGED + os]);                       // the store warp takes over
                CCA_STAMP(4);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

long long *g_dbg = nullptr;   // set through cca_b200__set_debug_buffer (profiling aid, not part of the ABI)

template <int LK, bool BF>
cudaError_t launch_fwd(const void *q, const void *k, const void *v, void *out, float *lse, float2 *stats, unsigned int *cnt,
                       Dims d, int mode, cudaStream_t st, const char **why)
{
    CUtensorMap m[8];
    const void *base[4] = {q, k, v, out};
    const int ch[4] = {d.Cq, d.Cq, d.C, d.C};
    for (int t = 0; t < 4; ++t)
        for (int r = 0; r < 2; ++r)
            if (!make_map(&m[2 * t + r], base[t], d.B, d.H, d.W, ch[t], LK, r == 0, BF)) {
                if (why) *why = "cuTensorMapEncodeTiled failed";
                return cudaErrorInvalidValue;
            }
    FwdParams p;
    p.B = d.B; p.H = d.H; p.W = d.W; p.C = d.C; p.Cq = d.Cq;
    p.mode = mode; p.stats = stats; p.lse = lse; p.sched = cnt; p.done = cnt + 2; p.rown = cnt + 2 + d.B;
    p.dbg = g_dbg ? g_dbg + (mode == MODE_ROW_ONLY ? 2560 : 0) : nullptr;
    // L2 residency: keep (evict_last) what the column pass touches for its last samples -- as many as fit the budget --
    // because the row pass starts with exactly those; everything else is marked evict_first
    const double per_sample = (2.0 * d.Cq + 2.0 * d.C) * d.H * d.W * (BF ? 2 : 4);
    int keep = (int)(tc_l2_keep_mb() * 1e6 / per_sample);
    if (keep > d.B) keep = d.B;
    p.hints = tc_l2_hints();
    p.keep_from = mode <= MODE_DYNAMIC ? 0 : d.B - keep;
    p.sync = (mode == MODE_COL_ONLY || mode == MODE_ROW_ONLY) && tc_pdl() == 2 ? 1 : 0;
    auto kern = cca_tc_fwd_kernel<LK, BF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem<LK, BF>::kBytes);
    if (e != cudaSuccess) return e;
    const int items = mode <= MODE_DYNAMIC ? d.B * (d.W + d.H) : (mode == MODE_COL_ONLY ? d.B * d.W : d.B * d.H);
    const int grid = items < sm_count() ? items : sm_count();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = FwdSmem<LK, BF>::kBytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (mode == MODE_ROW_ONLY && tc_pdl()) ? 1 : 0;     // only the row pass may start ahead of its predecessor
    e = cudaLaunchKernelEx(&cfg, kern, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// Launch policy of the forward (CCA_B200_FUSED): 0 = two launches (column pass, row pass); 1 = ONE launch with dynamic
// scheduling (lines are claimed from a column queue and a row queue; a row line only once the column lines of its sample
// are complete, so the partial output and q,k,v are re-read from L2 and no CTA ever waits on another); 2 = one launch with
// the static interleaved order (kept for comparison: it stalls on the column->row dependency).
int g_fused = -1;
int fused_mode()
{
    if (g_fused < 0) {
        const char *e = getenv("CCA_B200_FUSED");
        g_fused = e ? atoi(e) : 0;
        if (g_fused < 0 || g_fused > 2) g_fused = 0;
    }
    return g_fused;
}

}  // namespace

void set_tc_debug_buffer(void *p) { g_dbg = reinterpret_cast<long long *>(p); }
void set_tc_two_pass(int on) { g_fused = on == 1 ? 0 : (on == 0 ? 1 : 2); }   // 1: two launches, 0: dynamic fused, 2: static fused

bool tc_forward_supported(Dims d, int dtype) { return tc::shape_supported(d, dtype); }

// q,k,v,out are channels-last (NHWC), fp32 or bf16.  ws: [B*H*W] float2 stats, then [B] unsigned counters.
namespace {
template <bool BF>
cudaError_t launch_fwd_lk(int lk, const void *q, const void *k, const void *v, void *out, float *lse, float2 *stats,
                          unsigned int *cnt, Dims d, int mode, cudaStream_t st, const char **why)
{
    return lk == 80 ? launch_fwd<80, BF>(q, k, v, out, lse, stats, cnt, d, mode, st, why)
                    : launch_fwd<112, BF>(q, k, v, out, lse, stats, cnt, d, mode, st, why);
}
}  // namespace

cudaError_t tc_forward(const void *q, const void *k, const void *v, void *out, float *lse, void *ws, Dims d, int dtype,
                       cudaStream_t st, const char **why)
{
    float2 *stats = reinterpret_cast<float2 *>(ws);
    unsigned int *cnt = reinterpret_cast<unsigned int *>(stats + (size_t)d.B * d.H * d.W);   // [2] queue heads, [B] columns done, [B] rows handed out
    const int lkc = lk_for(d.H), lkr = lk_for(d.W);
    const bool bf = dtype == CCA_BF16;
    auto go = [&](int lk, int mode) {
        return bf ? launch_fwd_lk<true>(lk, q, k, v, out, lse, stats, cnt, d, mode, st, why)
                  : launch_fwd_lk<false>(lk, q, k, v, out, lse, stats, cnt, d, mode, st, why);
    };
    if (lkc == lkr && fused_mode() != 0) {
        cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * (2 * d.B + 2), st);
        if (e != cudaSuccess) return e;
        return go(lkc, fused_mode() == 1 ? MODE_DYNAMIC : MODE_FUSED);
    }
    cudaError_t e = cudaSuccess;
    if (tc_pdl() == 2) e = cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * (2 * d.B + 2), st);   // done[] of the overlapped row pass
    if (e != cudaSuccess) return e;
    e = go(lkc, MODE_COL_ONLY);
    if (e != cudaSuccess) return e;
    return go(lkr, MODE_ROW_ONLY);
}

}  // namespace cca
