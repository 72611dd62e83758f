// Dense hidden layers: weight gradient on tcgen05 AND the optimizer in one pass, with the optimizer state staged by TMA.
//
//   dW[k][n] = sum_b act3[b][k] * dH[b][n]   (K = batch <= 32: ONE k-block per output tile)
// The dense kernels hold 98.6 % of the parameters, so this pass is the HBM roofline of the step: it has to move
// read(w, m, v) + write(w, m, v) = 24 bytes per parameter (189 MB at C2) and nothing else.  A GEMM-shaped kernel with a
// read-modify-write epilogue keeps too few bytes in flight (measured 2.1 TB/s), a separate Adam pass moves the gradient
// twice more (252 MB).  Here a CTA owns a 128 (k) x 32 (n) tile of W:
//   * at CTA start one thread issues three 2-D TMA tensor copies (cp.async.bulk.tensor, 128 x 32 box, 128-byte swizzle)
//     of the tile's w / m / v into shared memory -- 48 KB in flight per CTA, three CTAs per SM, no registers held
//     (row-wise 128-byte bulk copies were tried first: 768 TMA operations per CTA made the pass 5x slower) -- and only then
//   * the 128 x 32 x 32 product is formed: act3 columns -> registers -> hi/lo -> tensor memory (coalesced loads, lane ==
//     row), dH tile transposed into a K-major hi/lo shared-memory tile, 12 tcgen05.mma (3-term TF32 split);
//   * the epilogue reads the accumulator (tcgen05.ld) and the staged state, applies Adam / Nesterov momentum in place in
//     shared memory (the swizzle makes the row-per-lane accesses conflict-free), and the tile goes back with three TMA
//     tensor stores.
// Launched after the dense dgrad has read W (it overwrites W).
#pragma once
#include "cvgemm.cuh"
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

namespace fa {
using namespace tc;

constexpr int FA_BN = 32, FA_THREADS = 256;
struct StateMaps { CUtensorMap w[2], m[2], v[2]; };  // per hidden stream: [flat][hid] f32, box 32 x 128, SWIZZLE_128B

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
// the box into L2 only (no shared-memory destination, no barrier): a way to have more bytes in flight than the stages hold
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(x), "r"(y), "r"(src) : "memory");
}

// L2 residency control (createpolicy + .L2::cache_hint on the TMA copies).  The optimizer state of the two dense kernels
// (m, v: 63 MB) plus the online dense weights (31.5 MB) fit in the 126 MB L2: loaded and stored "evict_last" they stay
// resident from one step to the next, and the pass that bounds the step stops going to HBM for them; streams that are
// read once per step (the target tower's dense weights) go "evict_first" so that they do not push the state out.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, int x, int y, uint32_t src, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                 ::"l"(map), "r"(x), "r"(y), "r"(src), "l"(pol) : "memory");
}

template <int NTERMS>
__global__ void __launch_bounds__(FA_THREADS, 3) fcwg_adam_kernel(const float* __restrict__ act3, const float* __restrict__ dhid,
                                                                  int Bn, int flat, int NH, int hid,
                                                                  const __grid_constant__ StateMaps maps,
                                                                  float lr, float mom, int use_momentum,
                                                                  const dqn_learner::Scalars* __restrict__ sc) {
    constexpr int B_TILE = FA_BN * 128;  // bytes of the hi (or lo) tile: 32 n rows x 32 k
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    constexpr uint32_t ARR = BM * FA_BN * 4;                   // one state array of the tile: 128 rows x 128 B, swizzled
    const uint32_t state = sbase;                              // [3 arrays] (1024-byte aligned for SWIZZLE_128B)
    const uint32_t btile = state + 3 * ARR;                    // hi, lo
    const uint32_t bars = btile + 2 * B_TILE;                  // state_full, afull, accum, tmem slot
    const uint32_t st_full = bars, afull = bars + 8, accum = bars + 16, tmem_slot = bars + 24;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * FA_BN, m0r = blockIdx.y * BM;
    const int strm = n0 >= hid;
    const int ncol = n0 - (strm ? hid : 0);
    const int narr = use_momentum ? 2 : 3;

    if (tid == 0) {
        mbar_init(st_full, 1); mbar_init(afull, 6); mbar_init(accum, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();

    if (warp == 5) {
        // ---- optimizer state of the tile: {w, m, v} boxes of 128 rows x 32 columns, one tensor copy each ----
        if (lane == 0) {
            cv::mbar_expect_tx(st_full, (uint32_t)(narr * ARR));
            tma_load_2d(state, &maps.w[strm], ncol, m0r, st_full);
            tma_load_2d(state + ARR, &maps.m[strm], ncol, m0r, st_full);
            if (!use_momentum) tma_load_2d(state + 2 * ARR, &maps.v[strm], ncol, m0r, st_full);
        }
    } else if (warp < 4) {
        // ---- A: act3^T, thread == row (flat index), 32 coalesced loads along the batch ----
        const int r = warp * 32 + lane;
        float hi[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) hi[b] = b < Bn ? __ldg(act3 + (size_t)b * flat + m0r + r) : 0.f;
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + 32u;  // columns [32,64) hi, [64,96) lo
        ts::tmem_st32(ta, hi);
        if (NTERMS == 3) {
            float lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) lo[j] = lo_part(hi[j]);
            ts::tmem_st32(ta + 32u, lo);
        }
        ts::tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull);
    } else if (warp >= 6) {
        // ---- B: dH[b][n0 + n] -> K-major hi/lo tiles; lane = (b % 4) + 4 * (n % 8): conflict-free transposing stores ----
        const int bq = lane & 3, nq = lane >> 2;
        for (int gi = warp - 6; gi < 8 * (FA_BN / 8); gi += 2) {
            const int b4 = gi / (FA_BN / 8), n8 = gi - b4 * (FA_BN / 8);
            const int b = b4 * 4 + bq, n = n8 * 8 + nq;
            const float v = b < Bn ? __ldg(dhid + (size_t)b * NH + n0 + n) : 0.f;
            const uint32_t a = btile + (uint32_t)((n >> 3) * cv::PK_SBO + (b >> 2) * cv::PK_LBO + (n & 7) * 16 + (b & 3) * 4);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
            if (NTERMS == 3) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + B_TILE), "f"(lo_part(v)) : "memory");
        }
        fence_proxy_async();  // generic-proxy stores -> read by the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(afull);
    } else {
        // ---- warp 4: MMA issue ----
        mbar_wait(afull, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc(FA_BN, false, false);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        if (leader) {
            const uint64_t db = make_desc(btile, cv::PK_LBO, cv::PK_SBO), dbl = make_desc(btile + B_TILE, cv::PK_LBO, cv::PK_SBO);
            const uint32_t ta = tmem_base + 32u;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                const uint64_t o = (uint64_t)(ks * 2 * cv::PK_LBO) >> 4;
                if (NTERMS == 3) {
                    ts::mma_tf32_ts(tmem_base, ta + ks * 8, dbl + o, idesc, ks != 0);
                    ts::mma_tf32_ts(tmem_base, ta + 32u + ks * 8, db + o, idesc, 1);
                    ts::mma_tf32_ts(tmem_base, ta + ks * 8, db + o, idesc, 1);
                } else {
                    ts::mma_tf32_ts(tmem_base, ta + ks * 8, db + o, idesc, ks != 0);
                }
            }
            mma_commit(accum);
        }
        __syncwarp();
    }
    pdl_launch_dependents();

    // ---- epilogue (all 8 warps): warp w -> rows 32 (w % 4) + lane, columns [16 (w / 4), +16) ----
    {
        const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
        const float lr_t = use_momentum ? lr : lr * sqrtf(1.0f - sc->b2p) / (1.0f - sc->b1p);
        const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
        const int row = (warp & 3) * 32 + lane, c0 = (warp >> 2) * 16;
        mbar_wait(accum, 0);
        tc_fence_after();
        float g[16];
        tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, g);
        mbar_wait(st_full, 0);
        const uint32_t srow = state + row * 128;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            // SWIZZLE_128B: 16-byte chunk q of row r sits at chunk q ^ (r % 8)
            const uint32_t sw = srow + ((uint32_t)(((c0 + j) >> 2) ^ (row & 7)) << 4), sm = sw + ARR, sv = sw + 2 * ARR;
            float4 pp, mm, vv;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pp.x), "=f"(pp.y), "=f"(pp.z), "=f"(pp.w) : "r"(sw));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(mm.x), "=f"(mm.y), "=f"(mm.z), "=f"(mm.w) : "r"(sm));
            const float4 gg = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
            if (use_momentum) {
#define MOM1(c) mm.c = mm.c * mom + gg.c; pp.c = pp.c - (gg.c * lr + mm.c * mom * lr);
                MOM1(x) MOM1(y) MOM1(z) MOM1(w)
#undef MOM1
            } else {
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vv.x), "=f"(vv.y), "=f"(vv.z), "=f"(vv.w) : "r"(sv));
#define ADAM1(c) mm.c = mm.c + (gg.c - mm.c) * omb1; vv.c = vv.c + (gg.c * gg.c - vv.c) * omb2; \
                 pp.c = pp.c - (mm.c * lr_t) / (sqrtf(vv.c) + eps);
                ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sv), "f"(vv.x), "f"(vv.y), "f"(vv.z), "f"(vv.w) : "memory");
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sm), "f"(mm.x), "f"(mm.y), "f"(mm.z), "f"(mm.w) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sw), "f"(pp.x), "f"(pp.y), "f"(pp.z), "f"(pp.w) : "memory");
        }
        fence_proxy_async();  // the updated rows leave through the async proxy
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        if (lane == 0) {
            tma_store_2d(&maps.w[strm], ncol, m0r, state);
            tma_store_2d(&maps.m[strm], ncol, m0r, state + ARR);
            if (!use_momentum) tma_store_2d(&maps.v[strm], ncol, m0r, state + 2 * ARR);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be released once it has been read
        }
        __syncwarp();
    } else if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}


// ---- persistent variant: S-stage TMA ring of optimizer state, a CTA walks a contiguous range of tiles -----------------
// The one-tile-per-CTA kernel above keeps HBM requests in flight only during the first microseconds of a CTA's life.
// Here a CTA owns up to 32 consecutive tiles (n fastest, so at most two different act3 slices), and one thread keeps
// PS - 1 tiles of w / m / v loads (48 KB each) and one tile of stores in flight at all times while the other warps
// form the gradient tile (tcgen05) and apply the optimizer in shared memory:
//   warp 9    TMA: loads of tile i + PS - 1 are issued as soon as the stores of tile i - 1 have drained their stage
//   warps 0-3 act3^T -> tensor memory (once per m tile); epilogue columns [0, 16)
//   warps 4-7 dH tile i + 1 -> K-major hi/lo shared-memory tile (double-buffered); epilogue columns [16, 32)
//   warp 8    MMA issue into double-buffered accumulators
constexpr int PS = 4, FP_THREADS = 320;

template <int NTERMS>
__global__ void __launch_bounds__(FP_THREADS, 1) fcwg_adam_persist_kernel(const float* __restrict__ act3, const float* __restrict__ dhid,
                                                                          int Bn, int flat, int NH, int hid,
                                                                          const __grid_constant__ StateMaps maps,
                                                                          float lr, float mom, int use_momentum,
                                                                          const dqn_learner::Scalars* __restrict__ sc) {
    constexpr int B_TILE = FA_BN * 128;
    constexpr uint32_t ARR = BM * FA_BN * 4, STAGE = 3 * ARR;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t ring = sbase;                               // PS stages x {w, m, v}
    const uint32_t btiles = ring + PS * STAGE;                 // 2 x {hi, lo}
    const uint32_t bars = btiles + 2 * 2 * B_TILE;
    auto st_full = [&](int s) { return bars + 8u * s; };
    auto computed = [&](int s) { return bars + 8u * (PS + s); };
    auto b_full = [&](int b) { return bars + 8u * (2 * PS + b); };
    auto b_empty = [&](int b) { return bars + 8u * (2 * PS + 2 + b); };
    auto acc_full = [&](int b) { return bars + 8u * (2 * PS + 4 + b); };
    auto acc_empty = [&](int b) { return bars + 8u * (2 * PS + 6 + b); };
    auto a_full = [&](int b) { return bars + 8u * (2 * PS + 8 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * PS + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntn = NH / FA_BN, total = ntn * (flat / BM);
    const int t0 = (int)((long long)blockIdx.x * total / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
    const int ntiles = t1 - t0;
    const int narr = use_momentum ? 2 : 3;
    const int mt_first = t0 / ntn;

    if (tid == 0) {
        for (int s = 0; s < PS; ++s) { mbar_init(st_full(s), 1); mbar_init(computed(s), 8); }
        for (int b = 0; b < 2; ++b) { mbar_init(b_full(b), 4); mbar_init(b_empty(b), 1); mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 8); mbar_init(a_full(b), 4); }
        fence_barrier_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    auto tmem_a = [&](int ap) { return tmem_base + 64u + (uint32_t)ap * 64u; };  // accumulators at [0,32), [32,64)

    if (warp == 9) {
        if (lane == 0 && ntiles > 0) {
            auto issue_loads = [&](int i) {
                const int t = t0 + i, nt = t % ntn, mt = t / ntn;
                const int n0 = nt * FA_BN, strm = n0 >= hid, ncol = n0 - (strm ? hid : 0), m0r = mt * BM;
                const int s = i % PS;
                cv::mbar_expect_tx(st_full(s), (uint32_t)(narr * ARR));
                tma_load_2d(ring + s * STAGE, &maps.w[strm], ncol, m0r, st_full(s));
                tma_load_2d(ring + s * STAGE + ARR, &maps.m[strm], ncol, m0r, st_full(s));
                if (!use_momentum) tma_load_2d(ring + s * STAGE + 2 * ARR, &maps.v[strm], ncol, m0r, st_full(s));
            };
            for (int j = 0; j < PS - 1 && j < ntiles; ++j) issue_loads(j);
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % PS;
                mbar_wait(computed(s), (i / PS) & 1);
                const int t = t0 + i, nt = t % ntn, mt = t / ntn;
                const int n0 = nt * FA_BN, strm = n0 >= hid, ncol = n0 - (strm ? hid : 0), m0r = mt * BM;
                tma_store_2d(&maps.w[strm], ncol, m0r, ring + s * STAGE);
                tma_store_2d(&maps.m[strm], ncol, m0r, ring + s * STAGE + ARR);
                if (!use_momentum) tma_store_2d(&maps.v[strm], ncol, m0r, ring + s * STAGE + 2 * ARR);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (i + PS - 1 < ntiles) {
                    // stage (i - 1) % PS is the one to refill: its stores (previous group) must have read it
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    issue_loads(i + PS - 1);
                }
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else if (warp == 8) {
        // ---------------- MMA issue ----------------
        const uint32_t idesc = make_idesc(FA_BN, false, false);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        int a_seen = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int bb = i & 1, use = i >> 1;
            const int ap = (t0 + i) / ntn - mt_first;  // 0 or 1: which act3 slice (tensor-memory A buffer)
            if (ap >= a_seen) { mbar_wait(a_full(ap), 0); a_seen = ap + 1; }
            mbar_wait(b_full(bb), use & 1);
            if (use >= 1) mbar_wait(acc_empty(bb), (use - 1) & 1);
            tc_fence_after();
            if (leader) {
                const uint32_t bt = btiles + bb * 2 * B_TILE;
                const uint64_t db = make_desc(bt, cv::PK_LBO, cv::PK_SBO), dbl = make_desc(bt + B_TILE, cv::PK_LBO, cv::PK_SBO);
                const uint32_t ta = tmem_a(ap), td = tmem_base + (uint32_t)bb * 32u;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t o = (uint64_t)(ks * 2 * cv::PK_LBO) >> 4;
                    if (NTERMS == 3) {
                        ts::mma_tf32_ts(td, ta + ks * 8, dbl + o, idesc, ks != 0);
                        ts::mma_tf32_ts(td, ta + 32u + ks * 8, db + o, idesc, 1);
                        ts::mma_tf32_ts(td, ta + ks * 8, db + o, idesc, 1);
                    } else {
                        ts::mma_tf32_ts(td, ta + ks * 8, db + o, idesc, ks != 0);
                    }
                }
                mma_commit(b_empty(bb));
                mma_commit(acc_full(bb));
            }
            __syncwarp();
        }
        tc_fence_before();
    } else {
        // ---------------- warps 0-7: operand producers, then the optimizer epilogue of every tile ----------------
        const int q4 = warp & 3, half = warp >> 2;
        const int row = q4 * 32 + lane, c0 = half * 16;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        if (warp < 4 && ntiles > 0) {
            const int mt_last = (t1 - 1) / ntn;
            for (int ap = 0; ap <= mt_last - mt_first && ap < 2; ++ap) {
                const int m0r = (mt_first + ap) * BM;
                float hi[32];
#pragma unroll
                for (int b = 0; b < 32; ++b) hi[b] = b < Bn ? __ldg(act3 + (size_t)b * flat + m0r + row) : 0.f;
                const uint32_t ta = tmem_a(ap) + lane_base;
                ts::tmem_st32(ta, hi);
                if (NTERMS == 3) {
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) lo[j] = lo_part(hi[j]);
                    ts::tmem_st32(ta + 32u, lo);
                }
                ts::tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full(ap));
            }
        }
        auto produce_b = [&](int i) {  // warps 4-7
            const int bb = i & 1, use = i >> 1;
            if (use >= 1) mbar_wait(b_empty(bb), (use - 1) & 1);
            const int n0 = ((t0 + i) % ntn) * FA_BN;
            const uint32_t bt = btiles + bb * 2 * B_TILE;
            const int bq = lane & 3, nq = lane >> 2;
            for (int gi = warp - 4; gi < 8 * (FA_BN / 8); gi += 4) {
                const int b4 = gi / (FA_BN / 8), n8 = gi - b4 * (FA_BN / 8);
                const int b = b4 * 4 + bq, n = n8 * 8 + nq;
                const float v = b < Bn ? __ldg(dhid + (size_t)b * NH + n0 + n) : 0.f;
                const uint32_t a = bt + (uint32_t)((n >> 3) * cv::PK_SBO + (b >> 2) * cv::PK_LBO + (n & 7) * 16 + (b & 3) * 4);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
                if (NTERMS == 3) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + B_TILE), "f"(lo_part(v)) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_full(bb));
        };
        if (warp >= 4 && ntiles > 0) produce_b(0);
        const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
        const float lr_t = use_momentum ? lr : lr * sqrtf(1.0f - sc->b2p) / (1.0f - sc->b1p);
        const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
        for (int i = 0; i < ntiles; ++i) {
            if (warp >= 4 && i + 1 < ntiles) produce_b(i + 1);
            const int ab = i & 1, s = i % PS;
            mbar_wait(acc_full(ab), (i >> 1) & 1);
            tc_fence_after();
            float g[16];
            tmem_ld16(tmem_base + lane_base + (uint32_t)(ab * 32 + c0), g);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(ab));
            mbar_wait(st_full(s), (i / PS) & 1);
            const uint32_t srow = ring + s * STAGE + row * 128;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const uint32_t sw = srow + ((uint32_t)(((c0 + j) >> 2) ^ (row & 7)) << 4), sm = sw + ARR, sv = sw + 2 * ARR;
                float4 pp, mm, vv;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pp.x), "=f"(pp.y), "=f"(pp.z), "=f"(pp.w) : "r"(sw));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(mm.x), "=f"(mm.y), "=f"(mm.z), "=f"(mm.w) : "r"(sm));
                const float4 gg = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
                if (use_momentum) {
#define MOM1(c) mm.c = mm.c * mom + gg.c; pp.c = pp.c - (gg.c * lr + mm.c * mom * lr);
                    MOM1(x) MOM1(y) MOM1(z) MOM1(w)
#undef MOM1
                } else {
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vv.x), "=f"(vv.y), "=f"(vv.z), "=f"(vv.w) : "r"(sv));
#define ADAM1(c) mm.c = mm.c + (gg.c - mm.c) * omb1; vv.c = vv.c + (gg.c * gg.c - vv.c) * omb2; \
                 pp.c = pp.c - (mm.c * lr_t) / (sqrtf(vv.c) + eps);
                    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sv), "f"(vv.x), "f"(vv.y), "f"(vv.z), "f"(vv.w) : "memory");
                }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sm), "f"(mm.x), "f"(mm.y), "f"(mm.z), "f"(mm.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sw), "f"(pp.x), "f"(pp.y), "f"(pp.z), "f"(pp.w) : "memory");
            }
            fence_proxy_async();  // the updated tile leaves through the async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(computed(s));
        }
        pdl_launch_dependents();
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

// ---- row-contiguous persistent variant -------------------------------------------------------------------------------
// The tile is 16 rows x 256 columns of one hidden stream's kernel: every row segment is 1 KB of CONTIGUOUS parameters
// and consecutive tiles of a CTA cover whole 2 KB rows, so the optimizer state streams through DRAM pages in order --
// the 128 x 32 tiles above touch 128 rows x 128 bytes at a 2 KB pitch and lose most of the DRAM bandwidth to page
// misses.  The product is formed transposed:
//   D[n][kr] = sum_b dH[b][n] * act3[b][k0 + kr]     M = n (2 MMA tiles of 128 per 256-column tile), N = 16 rows
//   A (TMEM): dH^T of the CTA's hidden stream, hi/lo, all 4 M tiles of the 512 columns resident (256 TMEM columns)
//   B (smem): the 16 x 32 act3 slice of the tile, hi/lo (2 KB each), one warp per tile, double-buffered
//   24 MMAs (N = 16) per tile into double-buffered accumulators (2 x 32 columns)
// State ring: RS stages x {w, m, v} x 16 KB, one 2-D TMA box of 256 n x 16 rows per array (no swizzle: in the
// epilogue thread == n, consecutive lanes read consecutive words).  16 epilogue warps (thread = one n, 8 of the 16
// rows) apply the optimizer with approximate reciprocal / square root (2 ulp: the update term is lr * m / (sqrt(v) +
// eps), so the weight sees < 1e-10 of it) -- with IEEE division and 8 warps the epilogue, not HBM, bound the pass.
// A CTA walks a contiguous range of tiles (n-half fastest, then rows, then stream).
constexpr int RS = 4, RT_THREADS = 608, RT_ROWS = 16, RT_N = 256;  // 16 epilogue warps + MMA + TMA + B producer

struct RowMaps { CUtensorMap w[2], m[2], v[2], wout[2]; };  // box 256 (n) x 16 (k rows), SWIZZLE_NONE; wout: where the updated weights go

template <int NTERMS>
__global__ void __launch_bounds__(RT_THREADS, 1) fcwg_adam_rows_kernel(const float* __restrict__ act3, const float* __restrict__ dhid,
                                                                       int Bn, int flat, int NH, int hid,
                                                                       const __grid_constant__ RowMaps maps,
                                                                       float lr, float mom, int use_momentum,
                                                                       const dqn_learner::Scalars* sc,
                                                                       unsigned int* __restrict__ tile_ctr, int l2_keep, int pf_dist,
                                                                       unsigned int* join_ticket, float* host_mirror, int mirror_words) {
    // Tiles are claimed DYNAMICALLY (atomic counter, stream-major order, so a CTA sees at most one switch of stream): the CTAs of this low-priority
    // kernel get their SMs at different times -- whenever the conv chain of the backward pass leaves some free -- and
    // with a static partition the late starters, each still holding a full share, ended the kernel ~15 us late.
    constexpr uint32_t ARR = RT_ROWS * RT_N * 4, STAGE = 3 * ARR, B_TILE = 16 * 128;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t ring = sbase;
    const uint32_t btiles = ring + RS * STAGE;                 // 2 x {hi, lo} x 2 KB
    const uint32_t bars = btiles + 2 * 2 * B_TILE;
    auto st_full = [&](int s) { return bars + 8u * s; };
    auto computed = [&](int s) { return bars + 8u * (RS + s); };
    auto b_full = [&](int b) { return bars + 8u * (2 * RS + b); };
    auto b_empty = [&](int b) { return bars + 8u * (2 * RS + 2 + b); };
    auto acc_full = [&](int b) { return bars + 8u * (2 * RS + 4 + b); };
    auto acc_empty = [&](int b) { return bars + 8u * (2 * RS + 6 + b); };
    const uint32_t a_full = bars + 8u * (2 * RS + 8), a_empty = bars + 8u * (2 * RS + 9), tmem_slot = bars + 8u * (2 * RS + 10);
    const uint32_t tile_ids = bars + 8u * (2 * RS + 12);      // int[RS]: tile claimed for ring slot s, -1 = no more tiles
    auto id_full = [&](int s) { return bars + 8u * (2 * RS + 14 + s); };  // tile id of slot s published (at claim time, before its data lands)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int halves = hid / RT_N;                             // 2
    const int per_stream = (flat / RT_ROWS) * halves;          // tile t: stream t / per_stream, rows [16 ((t % per_stream) / halves), +16), n half t % halves
    const int ntotal = per_stream * (NH / hid);
    const int narr = use_momentum ? 2 : 3;
    auto tile_of = [&](int s) { int t; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(t) : "r"(tile_ids + 4u * s)); return t; };

    if (tid == 0) {
        for (int s = 0; s < RS; ++s) { mbar_init(st_full(s), 1); mbar_init(computed(s), 16); mbar_init(id_full(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(b_full(b), 1); mbar_init(b_empty(b), 1); mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 16); }
        mbar_init(a_full, 4); mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    // TMEM: accumulators [0,32), [32,64) (M tile j of the tile at +16 j); A of M tile mt (0..3): hi at 64 + 64 mt, lo at +32
    auto tmem_a = [&](int mt) { return tmem_base + 64u + (uint32_t)mt * 64u; };

    if (warp == 17) {
        if (lane == 0) {
            const uint64_t pol_keep = l2_policy_evict_last();
            const bool keep_mv = (l2_keep & 1) != 0, keep_w = (l2_keep & 2) != 0;
            auto ld = [&](uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar, bool keep) {
                if (keep) tma_load_2d_hint(dst, map, x, y, bar, pol_keep); else tma_load_2d(dst, map, x, y, bar);
            };
            auto stg = [&](const CUtensorMap* map, int x, int y, uint32_t src, bool keep) {
                if (keep) tma_store_2d_hint(map, x, y, src, pol_keep); else tma_store_2d(map, x, y, src);
            };
            // claim a tile for ring slot i % RS and start its loads; past the last tile publish -1 (the barrier still completes)
            auto publish = [&](int i) {
                const int s = i % RS;
                const int t = (int)atomicAdd(tile_ctr, 1u);
                const int tt = t < ntotal ? t : -1;
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_ids + 4u * s), "r"(tt) : "memory");
                mbar_arrive(id_full(s));  // release: the B producer and the MMA warp may look at the id now
                // The pass is bound by bytes in flight per SM (3 stages x 48 KB against the loaded HBM latency), not by HBM
                // bandwidth.  Tiles are claimed in order by the whole grid, so the tile `pf_dist` claims ahead will be wanted by
                // some CTA a few tile times from now: pull it into L2 (no stage needed), its loads then see L2 latency.
                if (pf_dist > 0 && t + pf_dist < ntotal) {
                    const int tp = t + pf_dist, sp = tp / per_stream, rp = tp - sp * per_stream;
                    const int rowp = (rp / halves) * RT_ROWS, nhp = rp % halves;
                    tma_prefetch_2d(&maps.w[sp], nhp * RT_N, rowp);
                    tma_prefetch_2d(&maps.m[sp], nhp * RT_N, rowp);
                    if (!use_momentum) tma_prefetch_2d(&maps.v[sp], nhp * RT_N, rowp);
                }
                if (tt < 0) { mbar_arrive(st_full(s)); return; }
                const int strm = tt / per_stream, r = tt - strm * per_stream;
                const int row0 = (r / halves) * RT_ROWS, nh = r % halves;
                const uint32_t st = ring + s * STAGE;
                cv::mbar_expect_tx(st_full(s), (uint32_t)(narr * ARR));
                ld(st, &maps.w[strm], nh * RT_N, row0, st_full(s), keep_w);
                ld(st + ARR, &maps.m[strm], nh * RT_N, row0, st_full(s), keep_mv);
                if (!use_momentum) ld(st + 2 * ARR, &maps.v[strm], nh * RT_N, row0, st_full(s), keep_mv);
            };
            for (int j = 0; j < RS - 1; ++j) publish(j);
            for (int i = 0;; ++i) {
                const int s = i % RS;
                const int tt = tile_of(s);
                if (tt < 0) break;
                mbar_wait(computed(s), (i / RS) & 1);
                fence_proxy_async();  // the epilogue warps' shared-memory writes -> visible to the TMA stores
                const int strm = tt / per_stream, r = tt - strm * per_stream;
                const int row0 = (r / halves) * RT_ROWS, nh = r % halves;
                const uint32_t st = ring + s * STAGE;
                stg(&maps.wout[strm], nh * RT_N, row0, st, keep_w);
                stg(&maps.m[strm], nh * RT_N, row0, st + ARR, keep_mv);
                if (!use_momentum) stg(&maps.v[strm], nh * RT_N, row0, st + 2 * ARR, keep_mv);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // stage (i - 1) % RS has been read by its stores
                publish(i + RS - 1);
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else if (warp == 18) {
        // ---------------- B producer: act3[b][k0 .. k0+15] -> B(kr, b) hi/lo ----------------
        for (int i = 0;; ++i) {
            const int s = i % RS;
            mbar_wait(id_full(s), (i / RS) & 1);
            const int tt = tile_of(s);
            if (tt < 0) break;
            const int k0 = ((tt % per_stream) / halves) * RT_ROWS;
            float v[16];
            const float* p = act3 + (size_t)(lane < Bn ? lane : 0) * flat + k0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4 * q]), "=f"(v[4 * q + 1]), "=f"(v[4 * q + 2]), "=f"(v[4 * q + 3]) : "l"(p + 4 * q));
            const int bb = i & 1, use = i >> 1;
            if (use >= 1) mbar_wait(b_empty(bb), (use - 1) & 1);
            const uint32_t bt = btiles + bb * 2 * B_TILE;
#pragma unroll
            for (int kr = 0; kr < 16; ++kr) {
                const float x = lane < Bn ? v[kr] : 0.f;
                const uint32_t a = bt + (uint32_t)((kr >> 3) * cv::PK_SBO + (lane >> 2) * cv::PK_LBO + (kr & 7) * 16 + (lane & 3) * 4);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
                if (NTERMS == 3) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + B_TILE), "f"(lo_part(x)) : "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(b_full(bb));
        }
    } else if (warp == 16) {
        // ---------------- MMA issue: 2 M tiles x 4 k-steps x NTERMS per tile ----------------
        const uint32_t idesc = make_idesc(16, false, false);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        int a_strm = -1, a_gen = 0;  // stream whose dH^T is in tensor memory, number of conversions seen
        for (int i = 0;; ++i) {
            const int s = i % RS;
            mbar_wait(id_full(s), (i / RS) & 1);
            const int tt = tile_of(s);
            if (tt < 0) break;
            const int strm = tt / per_stream, nh = tt % halves;
            if (strm != a_strm) {
                // first tile of a stream: (after the first) let the converters replace A once every MMA issued so far has retired
                if (a_strm >= 0 && leader) mma_commit(a_empty);
                __syncwarp();
                mbar_wait(a_full, a_gen & 1);
                ++a_gen; a_strm = strm;
            }
            const int bb = i & 1, use = i >> 1;
            mbar_wait(b_full(bb), use & 1);
            if (use >= 1) mbar_wait(acc_empty(bb), (use - 1) & 1);
            fence_proxy_async();  // B slice was written through the generic proxy
            tc_fence_after();
            if (leader) {
                const uint32_t bt = btiles + bb * 2 * B_TILE;
                const uint64_t db = make_desc(bt, cv::PK_LBO, cv::PK_SBO), dbl = make_desc(bt + B_TILE, cv::PK_LBO, cv::PK_SBO);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t ta = tmem_a(nh * 2 + j), td = tmem_base + (uint32_t)(bb * 32 + j * 16);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t o = (uint64_t)(ks * 2 * cv::PK_LBO) >> 4;
                        if (NTERMS == 3) {
                            ts::mma_tf32_ts(td, ta + ks * 8, dbl + o, idesc, ks != 0);
                            ts::mma_tf32_ts(td, ta + 32u + ks * 8, db + o, idesc, 1);
                            ts::mma_tf32_ts(td, ta + ks * 8, db + o, idesc, 1);
                        } else {
                            ts::mma_tf32_ts(td, ta + ks * 8, db + o, idesc, ks != 0);
                        }
                    }
                }
                mma_commit(b_empty(bb));
                mma_commit(acc_full(bb));
            }
            __syncwarp();
        }
        tc_fence_before();
    } else {
        // ---------------- warps 0-15 ----------------
        const int q4 = warp & 3, grp = warp >> 2;                // grp: M tile of the tile (grp & 1), row half (grp >> 1)
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        auto convert_a = [&](int strm) {  // warps 0-3: dH^T of one stream, 4 M tiles, thread == n within the M tile
            for (int mt = 0; mt < 4; ++mt) {
                const int n = strm * hid + mt * 128 + q4 * 32 + lane;
                float hi[32];
#pragma unroll
                for (int b = 0; b < 32; ++b) hi[b] = b < Bn ? __ldg(dhid + (size_t)b * NH + n) : 0.f;
                const uint32_t ta = tmem_a(mt) + lane_base;
                ts::tmem_st32(ta, hi);
                if (NTERMS == 3) {
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) lo[j] = lo_part(hi[j]);
                    ts::tmem_st32(ta + 32u, lo);
                }
            }
            ts::tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        };
        int a_strm = -1;
        const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
        const float lr_t = use_momentum ? lr : lr * sqrtf(1.0f - sc->b2p) / (1.0f - sc->b1p);
        const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
        const int mj = grp & 1, rh = grp >> 1;                   // this warp: M tile mj of the tile, rows [8 rh, 8 rh + 8)
        for (int i = 0;; ++i) {
            const int ab = i & 1, s = i % RS;
            mbar_wait(st_full(s), (i / RS) & 1);
            const int tt = tile_of(s);
            if (tt < 0) break;
            if (warp < 4 && tt / per_stream != a_strm) {
                if (a_strm >= 0) { mbar_wait(a_empty, 0); tc_fence_after(); }  // MMAs of the previous stream have retired
                a_strm = tt / per_stream;
                convert_a(a_strm);
            }
            mbar_wait(acc_full(ab), (i >> 1) & 1);
            tc_fence_after();
            float g[8];
            tmem_ld8(tmem_base + lane_base + (uint32_t)(ab * 32 + mj * 16 + rh * 8), g);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(ab));
            const uint32_t col = ring + s * STAGE + (uint32_t)((mj * 128 + q4 * 32 + lane) * 4 + rh * 8 * (RT_N * 4));
#pragma unroll
            for (int kr = 0; kr < 8; ++kr) {
                const uint32_t sw = col + kr * (RT_N * 4), sm = sw + ARR, sv = sw + 2 * ARR;
                float pp, mm, vv;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pp) : "r"(sw));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mm) : "r"(sm));
                const float gg = g[kr];
                if (use_momentum) {
                    mm = mm * mom + gg; pp = pp - (gg * lr + mm * mom * lr);
                } else {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(vv) : "r"(sv));
                    mm = mm + (gg - mm) * omb1; vv = vv + (gg * gg - vv) * omb2;
                    float rt, rc;
                    asm("sqrt.approx.f32 %0, %1;" : "=f"(rt) : "f"(vv));
                    asm("rcp.approx.f32 %0, %1;" : "=f"(rc) : "f"(rt + eps));
                    pp = pp - (mm * lr_t) * rc;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sv), "f"(vv) : "memory");
                }
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sm), "f"(mm) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sw), "f"(pp) : "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(computed(s));  // release; the generic->async proxy fence runs once, on the TMA thread
        }
        pdl_launch_dependents();
    }
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    // (option tail_join) the optimizer tail runs beside this kernel: the last CTA here arrives at the end-of-step join (learner.h);
    // every CTA has made its last tile claim by now, so the bookkeeping may reset the tile counter
    if (join_ticket != nullptr && tid == 0) {
        __threadfence();
        if (atomicAdd(join_ticket + 2, 1u) == gridDim.x - 1)
            if (step_join_arrive(join_ticket, 2u))
                step_bookkeeping(const_cast<dqn_learner::Scalars*>(sc), join_ticket, use_momentum, host_mirror, mirror_words);
    }
}

static void encode_row_maps(dqn_learner* h, RowMaps& out, float* wout_base) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DQN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DQN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available");
    const NetDesc& net = h->net;
    // wout: the in-place slab, or (w_scratch) a scratch copy of the two dense kernels [2][flat][hidden] that is copied over
    // the slab once the dense dgrad has read the old weights -- the pass can then start before that dgrad
    float* slabs[4] = {h->online, h->slot_m, h->slot_v, h->online};
    CUtensorMap* dst[4] = {out.w, out.m, out.v, out.wout};
    for (int a = 0; a < 4; ++a)
        for (int st = 0; st < 2; ++st) {
            float* base = (a == 3 && wout_base) ? wout_base + (size_t)st * net.flat * net.hidden : slabs[a] + net.off_fc_w[st];
            const cuuint64_t gdim[2] = {(cuuint64_t)net.hidden, (cuuint64_t)net.flat};
            const cuuint64_t gstr[1] = {(cuuint64_t)net.hidden * 4};
            const cuuint32_t box[2] = {RT_N, RT_ROWS}, estr[2] = {1, 1};
            const CUresult r = ((EncodeFn)fn)(&dst[a][st], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box,
                                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            DQN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
        }
}

static inline bool rows_supported(const dqn_learner* h, int rows) {
    const NetDesc& net = h->net;
    return rows <= 32 && net.hidden == 2 * RT_N && net.flat % RT_ROWS == 0 && net.dueling && (net.off_fc_w[0] % 4) == 0 && (net.off_fc_w[1] % 4) == 0;
}

static void launch_rows(dqn_learner* h, int rows, int ctas, bool join = false) {
    const NetDesc& net = h->net;
    if (!h->fa_row_maps) {
        h->fa_row_maps = aligned_alloc(64, sizeof(RowMaps));
        try { encode_row_maps(h, *static_cast<RowMaps*>(h->fa_row_maps), h->w_scratch_on ? h->w_scratch : nullptr); } catch (...) { free(h->fa_row_maps); h->fa_row_maps = nullptr; throw; }
    }
    const RowMaps& maps = *static_cast<RowMaps*>(h->fa_row_maps);
    const size_t bytes = (size_t)RS * 3 * RT_ROWS * RT_N * 4 + 2 * 2 * 16 * 128 + 256 + 1024;
    // (the tile counter is zeroed by optimizer_tail_kernel at the end of every step: no memset node in front of this launch)
    auto go = [&](auto kern) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        launch_kernel(h->pdl, kern, dim3(ctas), dim3(RT_THREADS), bytes, h->stream, (const float*)h->acts_on.act[2], (const float*)h->d_hid,
                      rows, net.flat, net.NH, net.hidden, maps, (float)h->cfg.learning_rate, (float)h->cfg.momentum, h->cfg.use_momentum,
                      (const dqn_learner::Scalars*)h->d_sc, h->fa_tile_ctr, h->l2_keep, h->fc_rows_pf * ctas,
                      join ? h->opt_ticket : (unsigned int*)nullptr, join ? h->tail_mirror : (float*)nullptr, join ? h->tail_mirror_words : 0);
    };
    if (h->cfg.math_mode == DQN_MATH_TC3X) go(fcwg_adam_rows_kernel<3>); else go(fcwg_adam_rows_kernel<1>);
    h->launches++;
}

static inline bool supported(const dqn_learner* h, int rows) {
    const NetDesc& net = h->net;
    return rows <= 32 && net.flat % BM == 0 && net.hidden % FA_BN == 0 && (net.off_fc_w[0] % 4) == 0 && (net.off_fc_w[1] % 4) == 0;
}

// tensor maps of the six state arrays (w, m, v of the two hidden streams), encoded once per handle
static void encode_maps(dqn_learner* h, StateMaps& out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DQN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DQN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available");
    const NetDesc& net = h->net;
    float* slabs[3] = {h->online, h->slot_m, h->slot_v};
    CUtensorMap* dst[3] = {out.w, out.m, out.v};
    for (int a = 0; a < 3; ++a)
        for (int st = 0; st < 2; ++st) {
            const cuuint64_t gdim[2] = {(cuuint64_t)net.hidden, (cuuint64_t)net.flat};
            const cuuint64_t gstr[1] = {(cuuint64_t)net.hidden * 4};
            const cuuint32_t box[2] = {FA_BN, BM}, estr[2] = {1, 1};
            const CUresult r = ((EncodeFn)fn)(&dst[a][st], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, slabs[a] + net.off_fc_w[st], gdim, gstr, box,
                                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            DQN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
        }
}

static void launch(dqn_learner* h, int rows) {
    const NetDesc& net = h->net;
    if (!h->fa_maps) {
        h->fa_maps = aligned_alloc(64, sizeof(StateMaps));  // released with free() in dqn_destroy
        try { encode_maps(h, *static_cast<StateMaps*>(h->fa_maps)); } catch (...) { free(h->fa_maps); h->fa_maps = nullptr; throw; }
    }
    const StateMaps& maps = *static_cast<StateMaps*>(h->fa_maps);
    const size_t bytes = 3 * BM * FA_BN * 4 + 2 * FA_BN * 128 + 64 + 1024;
    auto go = [&](auto kern) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        launch_kernel(h->pdl, kern, dim3(net.NH / FA_BN, net.flat / BM), dim3(FA_THREADS), bytes, h->stream,
                      (const float*)h->acts_on.act[2], (const float*)h->d_hid, rows, net.flat, net.NH, net.hidden, maps,
                      (float)h->cfg.learning_rate, (float)h->cfg.momentum, h->cfg.use_momentum, (const dqn_learner::Scalars*)h->d_sc);
    };
    if (h->fc_tma_ctas > 0) {
        // persistent form: fc_tma_ctas CTAs (>= 60, so that a CTA sees at most two act3 slices), one per SM
        const int total = (net.NH / FA_BN) * (net.flat / BM);
        int ctas = h->fc_tma_ctas < 60 ? 60 : h->fc_tma_ctas;
        while ((total + ctas - 1) / ctas > net.NH / FA_BN) ++ctas;
        const size_t pbytes = (size_t)PS * 3 * BM * FA_BN * 4 + 2 * 2 * FA_BN * 128 + 256 + 1024;
        auto gop = [&](auto kern) {
            DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pbytes));
            launch_kernel(h->pdl, kern, dim3(ctas), dim3(FP_THREADS), pbytes, h->stream, (const float*)h->acts_on.act[2],
                          (const float*)h->d_hid, rows, net.flat, net.NH, net.hidden, maps, (float)h->cfg.learning_rate,
                          (float)h->cfg.momentum, h->cfg.use_momentum, (const dqn_learner::Scalars*)h->d_sc);
        };
        if (h->cfg.math_mode == DQN_MATH_TC3X) gop(fcwg_adam_persist_kernel<3>); else gop(fcwg_adam_persist_kernel<1>);
        h->launches++;
        return;
    }
    if (h->cfg.math_mode == DQN_MATH_TC3X) go(fcwg_adam_kernel<3>); else go(fcwg_adam_kernel<1>);
    h->launches++;
}
}  // namespace fa
