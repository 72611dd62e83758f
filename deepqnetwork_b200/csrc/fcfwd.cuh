// Dense hidden layer forward with the weights streamed by TMA.
//
//   hid_pre[b][n] = sum_k act3[b][k] * W[k][n]      (K = 7680, n over both hidden streams, b <= 64 batch rows)
// At batch 32-64 this layer is a stream of the 31.5 MB weight matrix through the tensor core.  One thread issues 2-D TMA
// tensor copies of 128 (n) x 32 (k) boxes of W into a 6-stage ring (16 KB each, 96 KB in flight per CTA), and the
// converters read their column of the box from shared memory (conflict-free: consecutive lanes are consecutive n),
// split hi/lo and feed the MMA through tensor memory.  What the in-kernel timeline (tools/dense_fwd_timeline.py) taught:
//   * the activation tile must not be produced by threads with row-scattered global loads: those LDGs (8 lines per
//     instruction) share the L1tex wavefront queue with the converters' LDS and doubled the k-block time; the compiler
//     also sank the prefetch loads to their use (volatile asm keeps them early), and a select on a just-loaded value
//     stalls on the load.  PACKED mode removes the thread work altogether: conv3's epilogue (EpiBiasReluPack) writes the
//     tiles, one cp.async.bulk per k-block fetches hi+lo;
//   * with weights L2-resident (repeated predict) the kernel is no faster: it is not HBM-bound at this size but bound by
//     the serial k-loop of 13 k-blocks per CTA + prologue/epilogue, and the two towers' 2 x 144 CTAs run as two waves.
// Measured equal to the generic kernel within noise, so it stays an option (fc_tma_fwd).
//   swap-AB: MMA rows (M = 128) = hidden units, MMA columns (N) = batch rows, split-K over blockIdx.y
//   A (TMEM)  : W^T tile, thread == hidden unit
//   B (smem)  : activation tile [rows][32 k] -> dense K-major hi/lo tiles through registers (warps 4-7)
//   epilogue  : part[z][b][n] partial sums, reduced (+ bias, ReLU, heads) by fc_reduce_heads_kernel as before
#pragma once
#include "fcadam.cuh"

namespace ff {
using namespace tc;

struct WMaps { CUtensorMap w[2]; };  // per hidden stream: [flat][hid] f32, box 128 (n) x 32 (k), no swizzle
constexpr int FF_S = 6, FF_THREADS = 448;  // warps 0-7 converters / B producers / epilogue, 8-11 MMA issuers, 12 W TMA, 13 B TMA (PACKED)

// PACKED: the B tiles arrive ready-made (EpiBiasReluPack wrote them), one bulk copy per k-block; warps 4-7 become a second
// converter group instead of B producers
template <int BN, int NTERMS, bool PACKED>
__global__ void __launch_bounds__(FF_THREADS, 1) fcfwd_kernel(const float* __restrict__ act, const float* __restrict__ act_pk, int rows, int flat, int NH, int hid,
                                                              const __grid_constant__ WMaps maps, float* __restrict__ part, int w_policy,
                                                              int kb_per_split, unsigned long long* __restrict__ dbg) {
    // several MMA-issuing warps (cvgemm.cuh, cvconv_kernel): warp i takes the k-blocks j % NI == i into its own accumulator,
    // the epilogue adds them in index order
    constexpr int NI = BN == 32 ? 4 : 2;
    constexpr int S = FF_S, ACOLS = BK * (NTERMS == 3 ? 2 : 1);
    const bool cta0 = dbg && blockIdx.x == 0 && blockIdx.y == 0;
#define FF_STAMP(cond, row, col) do { if (cta0 && (cond) && (row) < 64) dbg[(row) * 8 + (col)] = clock64(); } while (0)
    FF_STAMP(threadIdx.x == 0, 63, 0);
    constexpr uint32_t W_TILE = BM * BK * 4, B_TILE = BN * 128, STAGE = W_TILE + 2 * B_TILE;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bars = sbase + S * STAGE;
    auto wfull = [&](int s) { return bars + 8u * s; };
    auto afull = [&](int s) { return bars + 8u * (S + s); };
    auto bfull = [&](int s) { return bars + 8u * (2 * S + s); };
    auto empty = [&](int s) { return bars + 8u * (3 * S + s); };
    const uint32_t accum = bars + 8u * (4 * S), tmem_slot = bars + 8u * (4 * S + 1);
    // The W half of a stage is released by the converters that have read it (wfree), not by the MMAs of the block (empty): the weight
    // stream -- the long pole, 16 KB per k-block from L2 / HBM -- turns a stage around in TMA latency + one conversion instead of
    // waiting for the block's MMAs to retire as well (2500 -> ~1300 cycles per stage, tools/dense_fwd_timeline.py); and, needing
    // nothing from the stream predecessor, it starts under the predecessor's tail (PDL).
    auto wfree = [&](int s) { return bars + 8u * (4 * S + 2 + s); };
    auto stage_w = [&](int s) { return sbase + s * STAGE; };
    auto stage_b = [&](int s) { return sbase + s * STAGE + W_TILE; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * BM, z = blockIdx.y;
    const int strm = n0 >= hid, ncol = n0 - (strm ? hid : 0);
    const int nkb_total = flat / BK;
    const int kb0 = z * kb_per_split, kb1 = min(nkb_total, kb0 + kb_per_split), nkb = kb1 - kb0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(wfull(s), 1); mbar_init(afull(s), 4); mbar_init(bfull(s), PACKED ? 1 : 4); mbar_init(empty(s), 1); mbar_init(wfree(s), 4);
        }
        mbar_init(accum, nkb < NI ? (nkb > 0 ? nkb : 1) : NI);  // the last commit of every issuing warp that has a k-block
        fence_barrier_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    FF_STAMP(threadIdx.x == 0, 63, 1);
    constexpr int DCOLS = NI * BN;
    static_assert(DCOLS + S * ACOLS <= 512, "tensor memory budget");
    auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(DCOLS + s * ACOLS); };

    if (warp == 12) {
        // ---------------- W boxes by TMA (weights were written by an earlier step: no dependency on the predecessor) ----------------
        if (lane == 0) {
            int s = 0, ph = 0;
            // w_policy: 1 = evict_last (online weights: re-read by the dgrad and the optimizer pass), 2 = evict_first (target
            // weights: read once per step), 0 = no hint
            const uint64_t pol = w_policy == 1 ? fa::l2_policy_evict_last() : fa::l2_policy_evict_first();
            for (int j = 0; j < nkb; ++j) {
                if (j >= S) mbar_wait(wfree(s), ph ^ 1);  // the converters of block j - S have read the box
                FF_STAMP(true, j, 0);
                cv::mbar_expect_tx(wfull(s), W_TILE);
                if (w_policy) fa::tma_load_2d_hint(stage_w(s), &maps.w[strm], ncol, (kb0 + j) * BK, wfull(s), pol);
                else fa::tma_load_2d(stage_w(s), &maps.w[strm], ncol, (kb0 + j) * BK, wfull(s));
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 13) {
        // ---------------- packed activation tiles by TMA (PACKED): one bulk copy per k-block ----------------
        if (PACKED && lane == 0) {
            pdl_wait();  // the tiles are the stream predecessor's output
            int s = 0, ph = 0;
            for (int j = 0; j < nkb; ++j) {
                if (j >= S) mbar_wait(empty(s), ph ^ 1);  // the MMAs of block j - S have retired
                cv::mbar_expect_tx(bfull(s), 2 * B_TILE);
                cv::tma_bulk_g2s(stage_b(s), act_pk + (size_t)(kb0 + j) * (2 * BN * 32), 2 * B_TILE, bfull(s));
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp < (PACKED ? 8 : 4)) {
        // ---------------- A converters: thread == hidden unit (column of the W box); PACKED: two groups, alternate k-blocks ----------------
        const int q4 = warp & 3, grp = warp >> 2;
        constexpr int NG = PACKED ? 2 : 1;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        const uint32_t col = (uint32_t)(q4 * 32 + lane) * 4u;
        for (int j = grp; j < nkb; j += NG) {
            const int s = j % S, ph = (j / S) & 1;
            mbar_wait(wfull(s), ph);
            FF_STAMP(tid == 0, j, 1);
            float hi[32];
            const uint32_t src = stage_w(s) + col;
#pragma unroll
            for (int k = 0; k < 32; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[k]) : "r"(src + (uint32_t)k * (BM * 4)));
            if (j >= S) { mbar_wait(empty(s), ph ^ 1); tc_fence_after(); }  // the stage's tensor-memory half: block j - S's MMAs have retired
            const uint32_t ta = tmem_a(s) + lane_base;
            if (NTERMS == 3) cv::a_stage_store3(ta, hi); else ts::tmem_st32(ta, hi);
            ts::tmem_st_wait();
            tc_fence_before();
            FF_STAMP(tid == 0, j, 2);
            __syncwarp();
            if (lane == 0) { mbar_arrive(wfree(s)); mbar_arrive(afull(s)); }  // (the box's values went through registers into the store above)
        }
    } else if (warp < 8) {
        // ---------------- B producers: act[b][k0 .. k0+31] -> dense K-major hi/lo tiles ----------------
        pdl_wait();  // the activations are the stream predecessor's output
        FF_STAMP(tid == 128, 63, 2);
        const int t = tid - 128;
        constexpr int CH = BN * 8 / 128;  // 16-byte chunks per thread
        constexpr int PF = 4;             // k-blocks of activation loads in flight per thread (L2 latency ~1 us >> a k-block)
        // chunk q of a thread: 8 consecutive lanes write 8 consecutive 16-byte rows of a core matrix (conflict-free)
        const float* srcp[CH];
        uint32_t dsto[CH], dsto16[CH];
        bool okr[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int q = t + i * 128;
            const int r8 = q & 7, c = (q >> 3) & 7, rg = q >> 6;
            const int r = rg * 8 + r8;
            okr[i] = r < rows;
            srcp[i] = act + (size_t)(okr[i] ? r : 0) * flat + (size_t)kb0 * BK + c * 4;
            dsto[i] = (uint32_t)(rg * cv::PK_SBO + c * cv::PK_LBO + r8 * 16);
            dsto16[i] = cv::pk_off16(r, c * 4);
        }
        float4 pf[PF][CH];
        auto load_blk = [&](int j, float4 (&v)[CH]) {
#pragma unroll
            for (int i = 0; i < CH; ++i)
            {
                // volatile asm: the compiler must not sink these loads to their use PF k-blocks later (it did: the timeline
                // showed one full L2 latency per k-block); out-of-range rows / blocks read the first row (finite, unused or
                // multiplied by nothing: rows >= `rows` are never stored, blocks >= nkb are never consumed)
                const float* p = srcp[i] + (size_t)(j < nkb ? j : 0) * BK;
                asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "l"(p));
                // (no select on the loaded value: it would make the thread wait for the load right here; rows == BN is
                // guaranteed by the launcher, so every row of the tile exists)
            }
        };
#pragma unroll
        for (int p = 0; p < PF; ++p) load_blk(p, pf[p]);
        int s = 0, ph = 0;
        auto put_blk = [&](int j, float4 (&v)[CH]) {
            if (j >= S) mbar_wait(empty(s), ph ^ 1);
            FF_STAMP(tid == 128, j, 3);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const uint32_t a = stage_b(s) + dsto[i];
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[i].x), "f"(v[i].y), "f"(v[i].z), "f"(v[i].w) : "memory");
                if (NTERMS == 3) {  // bf16 copies for the cross terms (cvpack.cuh layout: 8 bytes per four k)
                    const uint32_t a16 = stage_b(s) + dsto16[i];
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a16 + B_TILE), "r"(cv::bf16_pair(v[i].x, v[i].y)), "r"(cv::bf16_pair(v[i].z, v[i].w)) : "memory");
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a16 + B_TILE + B_TILE / 2), "r"(cv::bf16_pair(lo_part(v[i].x), lo_part(v[i].y))),
                                 "r"(cv::bf16_pair(lo_part(v[i].z), lo_part(v[i].w))) : "memory");
                }
            }
            // release-arrive; the generic->async proxy fence runs on the MMA thread after its acquire (a producer-side
            // fence.proxy.async costs ~1000 cycles per k-block here: fftl timeline)
            __syncwarp();
            if (lane == 0) mbar_arrive(bfull(s));
            FF_STAMP(tid == 128, j, 4);
            if (++s == S) { s = 0; ph ^= 1; }
            load_blk(j + PF, v);
        };
        for (int j = 0; j < nkb; j += PF) {
#pragma unroll
            for (int p = 0; p < PF; ++p)
                if (j + p < nkb) put_blk(j + p, pf[p]);
        }
    } else if (warp - 8 < NI) {
        // ---------------- MMA issuers (warps 8 .. 8 + NI - 1) ----------------
        const int wi = warp - 8;
        const uint32_t idesc = make_idesc(BN, false, false), idesc16 = cv::make_idesc_bf16(BN);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint32_t tmem_d = tmem_base + (uint32_t)(wi * BN);
        for (int j = wi; j < nkb; j += NI) {
            const int s = j % S, ph = (j / S) & 1;
            mbar_wait(afull(s), ph);
            mbar_wait(bfull(s), ph);  // B tile: by TMA (PACKED, async proxy) or from the producer warps
            FF_STAMP(lane == 0, j, 5);
            if (!PACKED) fence_proxy_async();  // B tile: generic -> async proxy
            FF_STAMP(lane == 0, j, 6);
            tc_fence_after();
            const uint32_t ta = tmem_a(s);
            if (leader) {
                if (NTERMS == 3) cv::mma_kblock3(tmem_d, ta, stage_b(s), BN, idesc, idesc16, j == wi);
                else {
                    const uint64_t db = make_desc(stage_b(s), cv::PK_LBO, cv::PK_SBO);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks)
                        ts::mma_tf32_ts(tmem_d, ta + ks * 8, db + ((uint64_t)(ks * 2 * cv::PK_LBO) >> 4), idesc, !(j == wi && ks == 0));
                }
                mma_commit(empty(s));
                if (j + NI >= nkb) mma_commit(accum);
            }
            FF_STAMP(lane == 0, j, 7);
            __syncwarp();
        }
        tc_fence_before();
    }
    if (warp < 8) {
        pdl_launch_dependents();
        // ---------------- epilogue: part[z][b][n0 + unit], warp (q4, half): units 32 q4 + lane, batch rows [half BN/2, +BN/2) ----------------
        const int q4 = warp & 3, half = warp >> 2;
        if (nkb > 0) { mbar_wait(accum, 0); tc_fence_after(); }
        FF_STAMP(tid == 0, 63, 3);
        constexpr int HC = BN / 2;
        float* o = part + (size_t)z * rows * NH + n0 + q4 * 32 + lane;
#pragma unroll
        for (int c = 0; c < HC; c += 16) {
            float v[16];
            if (nkb > 0) {
                const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * HC + c);
                if (nkb >= NI) {  // every issuing warp has written its accumulator: all of them in index order, one wait
                    if (NI == 4) tmem_ld16_sum4(t0, t0 + BN, t0 + 2 * BN, t0 + 3 * BN, v);
                    else tmem_ld16_sum2(t0, t0 + BN, v);
                } else {
                    tmem_ld16(t0, v);
#pragma unroll
                    for (int i = 1; i < NI; ++i)
                        if (i < nkb) {
                            float v1[16];
                            tmem_ld16(t0 + (uint32_t)(i * BN), v1);
#pragma unroll
                            for (int k = 0; k < 16; ++k) v[k] += v1[k];
                        }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int b = half * HC + c + i;
                if (b < rows) o[(size_t)b * NH] = v[i];
            }
        }
        tc_fence_before();
        FF_STAMP(tid == 0, 63, 4);
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
#undef FF_STAMP
}

static void encode_wmaps(dqn_learner* h, const float* P, WMaps& out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DQN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DQN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available");
    const NetDesc& net = h->net;
    for (int st = 0; st < 2; ++st) {
        const cuuint64_t gdim[2] = {(cuuint64_t)net.hidden, (cuuint64_t)net.flat};
        const cuuint64_t gstr[1] = {(cuuint64_t)net.hidden * 4};
        const cuuint32_t box[2] = {BM, BK}, estr[2] = {1, 1};
        const CUresult r = ((EncodeFn)fn)(&out.w[st], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(P) + net.off_fc_w[st], gdim, gstr,
                                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DQN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
    }
}

static inline bool supported(const dqn_learner* h, int rows) {
    const NetDesc& net = h->net;
    return (rows == 32 || rows == 64) && net.hidden % BM == 0 && net.flat % BK == 0 && (net.off_fc_w[0] % 4) == 0 && (net.off_fc_w[1] % 4) == 0;
}

// returns the number of split-K partials written to part[z][rows][NH]
static int launch(dqn_learner* h, const float* P, const float* act, int rows, float* part, int max_splits, const float* act_pk = nullptr) {
    const NetDesc& net = h->net;
    const int tw = P == h->online ? 0 : 1;
    if (!h->ff_maps[tw]) {
        h->ff_maps[tw] = aligned_alloc(64, sizeof(WMaps));
        try { encode_wmaps(h, P, *static_cast<WMaps*>(h->ff_maps[tw])); } catch (...) { free(h->ff_maps[tw]); h->ff_maps[tw] = nullptr; throw; }
    }
    const WMaps& maps = *static_cast<WMaps*>(h->ff_maps[tw]);
    const int mt = net.NH / BM, nkb = net.flat / BK;
    // the two towers' launches run side by side and must fit in one wave together.  A k-block costs the online tower
    // (64 batch columns per MMA) ~1.7x what it costs the target tower (32), so the SMs are shared 2 : 1 (fc_fwd_share = 0:
    // half each) -- both launches then end at about the same time instead of the online one trailing
    int share_num = 1, share_den = 2;
    if (h->fc_fwd_share && (rows == 64 || rows == 32)) { share_num = rows == 64 ? 2 : 1; share_den = 3; }
    int want = std::min(std::max(1, (h->sm_count * share_num) / (share_den * mt)), max_splits);
    const int per = (nkb + want - 1) / want;
    const int splits = (nkb + per - 1) / per;
    auto go = [&](auto kern, int bn) {
        const size_t bytes = (size_t)FF_S * (BM * BK * 4 + 2 * bn * 128) + 512 + 1024;
        // (no static "configured" flag here: both BN instantiations share one function-pointer type, hence one lambda body)
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        const int w_policy = tw == 0 ? ((h->l2_keep & 2) ? 1 : 0) : ((h->l2_keep & 4) ? 2 : 0);
        launch_kernel(h->pdl, kern, dim3(mt, splits), dim3(FF_THREADS), bytes, h->stream, act, act_pk, rows, net.flat, net.NH, net.hidden, maps,
                      part, w_policy, per, h->tc_dbg);
    };
    const bool x3 = h->cfg.math_mode == DQN_MATH_TC3X;
    if (act_pk) {
        if (rows == 64) { if (x3) go(fcfwd_kernel<64, 3, true>, 64); else go(fcfwd_kernel<64, 1, true>, 64); }
        else            { if (x3) go(fcfwd_kernel<32, 3, true>, 32); else go(fcfwd_kernel<32, 1, true>, 32); }
    } else {
        if (rows == 64) { if (x3) go(fcfwd_kernel<64, 3, false>, 64); else go(fcfwd_kernel<64, 1, false>, 64); }
        else            { if (x3) go(fcfwd_kernel<32, 3, false>, 32); else go(fcfwd_kernel<32, 1, false>, 32); }
    }
    h->launches++;
    return splits;
}
}  // namespace ff
