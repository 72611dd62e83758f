// tcgen05 implicit GEMM, second generation: the 128-row operand A is fed to the tensor core from TENSOR MEMORY
// (tcgen05.mma with a TMEM A operand), only the narrow operand B goes through shared-memory descriptors.
//
// Why: with fp32 operands and the 3-term TF32 split every k-step issues three MMAs, and in the all-shared-memory
// kernel (tcgemm.cuh) each of them re-reads a 128 x 8 A slice; together with writing a hi and a lo copy of every
// tile the shared-memory port (128 B/clk) -- not the tensor pipe, not HBM -- bounded the k-loop (measured 0.6-0.9 us
// per 32-wide k-block).  Here A never makes that round trip:
//   * A contiguous along M (TF-layout dense weights, transposed activations): lane == row, so the 32 lanes of a warp
//     read 128 contiguous bytes per k; global -> registers -> tcgen05.st, no shared memory at all and no transpose.
//   * A contiguous along K (im2col rows, dgrad rows): 16-byte cp.async (zero-fill) into a raw staging tile with
//     coalesced global reads, then each thread reads ITS row back (conflict-free, one pass), splits hi/lo in
//     registers and stores both to TMEM.
// TMEM layout (512 columns): [0, BN) accumulator, then S stages of {32 columns A_hi, 32 columns A_lo}; for M = 128
// the tensor core expects A(m, k) at lane m, column k (cute tmem_frg, M_MMA == 128).
// Per k-block and CTA the shared-memory traffic drops from 2*(A+B) written + 3*(A+B) read to A written once and read
// once (K-contiguous A) or nothing (M-contiguous A), plus 2*B written + 3*B read.
//
// Thread roles (544 threads): warps 0-15 producers/converters, then the epilogue (warp w owns TMEM lanes
// 32*(w%4).., column group w/4); warp 16 allocates TMEM and one elected lane issues the MMAs.
#pragma once
#include "tcgemm.cuh"

namespace ts {
using namespace tc;

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// the mbarrier receives one (pre-counted) arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// operand functors that are contiguous along M opt in to the register path with addr1(row, m, k, z) -> &A(m, k)
template <class Op, class = void> struct has_addr1 { static constexpr bool value = false; };
template <class Op>
struct has_addr1<Op, decltype((void)static_cast<const float*>(std::declval<const Op&>().addr1(std::declval<const typename Op::Row&>(), 0, 0, 0)))> {
    static constexpr bool value = true;
};
template <class AOp> struct capable { static constexpr bool value = AOp::KCONTIG || has_addr1<AOp>::value; };

constexpr int A_PITCH = 144;  // staging row pitch: 128 data bytes + 16, so that 8 consecutive rows hit 8 distinct 16-byte bank groups

template <int BN, int NTERMS, bool A_DIRECT>
struct Plan {
    static constexpr int A_STAGE = A_DIRECT ? 0 : BM * A_PITCH;
    static constexpr int B_BYTES = Tile<BN>::BYTES;
    static constexpr int STAGE = A_STAGE + B_BYTES * (NTERMS == 3 ? 2 : 1);
    static constexpr int DCOLS = BN < 32 ? 32 : BN;
    static constexpr int ACOLS = BK * (NTERMS == 3 ? 2 : 1);
    static constexpr int S_TMEM = (512 - DCOLS) / ACOLS;
    static constexpr int S_SMEM = (196 * 1024) / STAGE;
    static constexpr int S0 = S_TMEM < S_SMEM ? S_TMEM : S_SMEM;
    static constexpr int STAGES = S0 > 6 ? 6 : S0;
    static_assert(STAGES >= 3, "pipeline too shallow");
    static constexpr int DA = STAGES - 2 > 4 ? 4 : STAGES - 2;  // cp.async issue distance (k-blocks ahead of the convert)
    static constexpr int EPI_BYTES = BM * (BN + 4) * 4;         // coalesced-epilogue transpose tile
    static constexpr int PIPE_BYTES = STAGES * STAGE;
    static constexpr int BYTES = (PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES) + 256;
    static constexpr int NEED = DCOLS + STAGES * ACOLS;
    static constexpr uint32_t TMEM_COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
};

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(THREADS, 1) tsgemm_kernel(const AOp a, const BOp b, const Epi epi, int M, int N, int K,
                                                            int kchunk, const float* __restrict__ zeros) {
    constexpr bool A_DIRECT = !AOp::KCONTIG;
    using P = Plan<BN, NTERMS, A_DIRECT>;
    constexpr int S = P::STAGES;
    using TB = Tile<BN>;
    using AA = Assign<true, BM>;  // cp.async assignment of the K-contiguous A path
    using AB = Assign<BOp::KCONTIG, BN>;
    constexpr int LA = A_DIRECT ? 1 : AA::N, LB = AB::N;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bars = sbase + P::BYTES - 256;  // full[S], empty[S], afull[S], accum, tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    auto afull_bar = [&](int s) { return bars + 8u * (2 * S + s); };
    const uint32_t accum_bar = bars + 8u * (3 * S);
    const uint32_t tmem_slot = bars + 8u * (3 * S + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = SPLIT ? z * kchunk : 0;
    const int kend = SPLIT ? min(K, kbeg + kchunk) : K;
    const int nkb = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), PRODUCERS / 32);
            mbar_init(empty_bar(s), 1);
            mbar_init(afull_bar(s), PRODUCERS);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(P::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto stage_a = [&](int s) { return sbase + s * P::STAGE; };  // raw staging tile (K-contiguous A only)
    auto stage_b = [&](int s) { return sbase + s * P::STAGE + P::A_STAGE; };
    auto stage_blo = [&](int s) { return stage_b(s) + P::B_BYTES; };
    auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(P::DCOLS + s * P::ACOLS); };

    if (warp < PRODUCERS / 32) {
        // ---------------- producers / converters ----------------
        const int grp = warp & 3, cgrp = warp >> 2;  // TMEM lane group (rows 32*grp..), column group (k 8*cgrp..)
        const int my_row = grp * 32 + lane;          // the A row this thread converts
        const uint32_t lane_base = (uint32_t)(grp * 32) << 16;
        // K-contiguous A: cp.async assignment (coalesced along K)
        int a_r[LA], a_k[LA];
        typename AOp::Row a_row[LA];
        // M-contiguous A: the thread's own row
        const bool my_row_ok = m0 + my_row < M;
        if constexpr (A_DIRECT) {
            a_r[0] = my_row; a_k[0] = cgrp * 8;
            a_row[0] = a.row(min(m0 + my_row, M - 1), z);
        } else {
#pragma unroll
            for (int i = 0; i < LA; ++i) {
                AA::rk(tid, i, a_r[i], a_k[i]);
                a_row[i] = a.row(min(m0 + max(a_r[i], 0), M - 1), z);
            }
        }
        int b_r[LB], b_k[LB];
        uint32_t b_off[LB];
        typename BOp::Row b_row[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            AB::rk(tid, i, b_r[i], b_k[i]);
            b_off[i] = b_r[i] >= 0 ? TB::off(b_r[i], b_k[i]) : 0;
            b_row[i] = b.row(min(n0 + max(b_r[i], 0), N - 1), z);
        }
        constexpr int PF = 4;  // k-blocks of register-staged loads in flight per thread
        float4 b_pf[PF][LB];
        float a_pf[PF][A_DIRECT ? 8 : 1];

        auto load_regs = [&](int kb, float4 (&bp)[LB], float (&ap)[A_DIRECT ? 8 : 1]) {
            const int k0 = kbeg + kb * BK;
            if constexpr (A_DIRECT) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = k0 + a_k[0] + j;
                    // ragged edges read zeros: the select is on the address, nothing waits on the value here
                    const float* src = (kb < nkb && my_row_ok && k < kend) ? a.addr1(a_row[0], m0 + my_row, k, z) : zeros;
                    ap[j] = __ldg(src);
                }
            }
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                const int r = n0 + b_r[i], k = k0 + b_k[i];
                const float* src = (kb < nkb && b_r[i] >= 0 && r < N && k < kend) ? b.addr4(b_row[i], r, k, z) : nullptr;
                bp[i] = __ldg(reinterpret_cast<const float4*>(src ? src : zeros));
            }
        };
        auto issue_a = [&](int kb) {  // K-contiguous A: cp.async block kb into its staging tile
            const int s = kb % S;
            const int k0 = kbeg + kb * BK;
#pragma unroll
            for (int i = 0; i < LA; ++i) {
                if (a_r[i] >= 0) {
                    const int r = m0 + a_r[i], k = k0 + a_k[i];
                    const float* src = (r < M && k < kend) ? a.addr4(a_row[i], r, k, z) : nullptr;
                    cp_async16(stage_a(s) + a_r[i] * A_PITCH + a_k[i] * 4, src ? src : zeros, src != nullptr);
                }
            }
            cp_async_arrive_noinc(afull_bar(s));
        };
        auto sts_v = [&](uint32_t addr, const float4& v) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        };
        auto sts_t = [&](uint32_t addr, const float4& v) {
            asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+16], %2;\n\tst.shared.f32 [%0+32], %3;\n\tst.shared.f32 [%0+48], %4;"
                         ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        };
        auto block = [&](int kb, float4 (&bp)[LB], float (&ap)[A_DIRECT ? 8 : 1]) {
            const int s = kb % S;
            if constexpr (!A_DIRECT) {
                // keep DA blocks of A copies in flight; stage of block kb+DA was last used by block kb+DA-S
                const int j = kb + P::DA;
                if (j < nkb) {
                    if (j >= S) mbar_wait(empty_bar(j % S), ((j / S) - 1) & 1);
                    issue_a(j);
                }
            } else {
                if (kb >= S) mbar_wait(empty_bar(s), ((kb / S) - 1) & 1);  // MMAs of block kb-S are done with stage s
            }
            // ---- A(kb) -> TMEM (hi = the raw fp32 word: the tensor core reads its top 19 bits; lo = the rest)
            float hi[8];
            if constexpr (A_DIRECT) {
#pragma unroll
                for (int j = 0; j < 8; ++j) hi[j] = ap[j];
            } else {
                mbar_wait(afull_bar(s), (kb / S) & 1);  // every producer's copies for this block have landed
                const uint32_t src = stage_a(s) + my_row * A_PITCH + cgrp * 32;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(hi[0]), "=f"(hi[1]), "=f"(hi[2]), "=f"(hi[3]) : "r"(src));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(hi[4]), "=f"(hi[5]), "=f"(hi[6]), "=f"(hi[7]) : "r"(src));
            }
            const uint32_t ta = tmem_a(s) + lane_base + (uint32_t)(cgrp * 8);
            tmem_st8(ta, hi);
            if (NTERMS == 3) {
                float lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) lo[j] = lo_part(hi[j]);
                tmem_st8(ta + BK, lo);
            }
            // ---- B(kb) -> shared memory tiles (hi / lo), K-major no-swizzle UMMA layout
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                if (b_r[i] >= 0) {
                    const float4 v = bp[i];
                    const float4 l = make_float4(lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w));
                    if (BOp::KCONTIG) { sts_v(stage_b(s) + b_off[i], v); if (NTERMS == 3) sts_v(stage_blo(s) + b_off[i], l); }
                    else { sts_t(stage_b(s) + b_off[i], v); if (NTERMS == 3) sts_t(stage_blo(s) + b_off[i], l); }
                }
            }
            // ---- publish: TMEM stores complete, then release-arrive (one per warp); the generic->async proxy fence
            // for the B tile is executed by the MMA thread after its acquire
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
            load_regs(kb + PF, bp, ap);
        };
        if constexpr (!A_DIRECT) {
#pragma unroll
            for (int j = 0; j < P::DA; ++j)
                if (j < nkb) issue_a(j);
        }
#pragma unroll
        for (int p = 0; p < PF; ++p) load_regs(p, b_pf[p], a_pf[p]);
        for (int kb = 0; kb < nkb; kb += PF) {
#pragma unroll
            for (int p = 0; p < PF; ++p)
                if (kb + p < nkb) block(kb + p, b_pf[p], a_pf[p]);
        }
        // ---------------- epilogue ----------------
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        const int sub = warp & 3, quarter = warp >> 2;
        const int m = m0 + sub * 32 + lane;
        constexpr int QCOLS = BN / 4;  // columns per warp: 8, 16 or 32
        const uint32_t trow = tmem_base + ((uint32_t)(sub * 32) << 16);
        if constexpr (wants_coalesce<Epi>::value) {
            // accumulators -> smem tile [128][BN+4] (the pipeline buffers are idle now) -> rows handed to whole warps
            constexpr int PITCH = BN + 4;
            float* tile = reinterpret_cast<float*>(smem + (sbase - smem_u32(smem)));
            const int row = sub * 32 + lane;
            if (QCOLS == 8) {
                float v[8];
                if (nkb > 0) tmem_ld8(trow + (uint32_t)(quarter * 8), v);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; i += 4)
                    *reinterpret_cast<float4*>(tile + row * PITCH + quarter * 8 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < QCOLS; c += 16) {
                    float v[16];
                    if (nkb > 0) tmem_ld16(trow + (uint32_t)(quarter * QCOLS + c), v);
                    else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(tile + row * PITCH + quarter * QCOLS + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(PRODUCERS) : "memory");
            constexpr int LANES_PER_ROW = BN / 4;            // float4 per row
            constexpr int ROWS_PER_PASS = PRODUCERS / LANES_PER_ROW;
            const int c4 = (tid % LANES_PER_ROW) * 4;
            const float step = epi.step_size();
            constexpr int RB = 4;  // rows per batch: all their read-modify-write loads are issued before any store
            for (int r0 = tid / LANES_PER_ROW; r0 < BM; r0 += ROWS_PER_PASS * RB) {
                typename Epi::State st[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int r = r0 + j * ROWS_PER_PASS;
                    if (r < BM && m0 + r < M && n0 + c4 < N) epi.load(m0 + r, n0 + c4, st[j]);
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int r = r0 + j * ROWS_PER_PASS;
                    if (r < BM && m0 + r < M && n0 + c4 < N)
                        epi.apply(m0 + r, n0 + c4, *reinterpret_cast<const float4*>(tile + r * PITCH + c4), st[j], step);
                }
            }
        } else if (QCOLS == 8) {
            const int col = quarter * 8;
            float v[8];
            if (nkb > 0) tmem_ld8(trow + (uint32_t)col, v);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            if (m < M && n0 + col < N) epi.template store<8>(m, n0 + col, v, z);
        } else {
#pragma unroll
            for (int c = 0; c < QCOLS; c += 16) {
                const int col = quarter * QCOLS + c;
                float v[16];
                if (nkb > 0) tmem_ld16(trow + (uint32_t)col, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (m < M && n0 + col < N) epi.template store<16>(m, n0 + col, v, z);
            }
        }
        tc_fence_before();
    } else {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc(BN, false, false);  // A (TMEM) and B (smem) both K-major
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint64_t db0 = TB::desc(stage_b(0), 0), dbl0 = TB::desc(stage_blo(0), 0);
        constexpr uint64_t STAGE16 = P::STAGE >> 4, KS16 = (2 * LBO) >> 4;  // descriptor address units (16 bytes)
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            mbar_wait(full_bar(s), (kb / S) & 1);   // acquire: B tile stores and A TMEM stores of this block
            fence_proxy_async();                      // B tile: generic -> async proxy
            tc_fence_after();
            const uint64_t so = (uint64_t)s * STAGE16;
            const uint32_t ta = tmem_a(s);
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                const uint64_t o = so + ks * KS16;
                const uint32_t acc = (kb | ks) != 0;
                if (leader) {
                    if (NTERMS == 3) {
                        mma_tf32_ts(tmem_base, ta + ks * 8, dbl0 + o, idesc, acc);       // A_hi * B_lo
                        mma_tf32_ts(tmem_base, ta + BK + ks * 8, db0 + o, idesc, 1);    // A_lo * B_hi
                        mma_tf32_ts(tmem_base, ta + ks * 8, db0 + o, idesc, 1);         // A_hi * B_hi
                    } else {
                        mma_tf32_ts(tmem_base, ta + ks * 8, db0 + o, idesc, acc);
                    }
                }
            }
            if (leader) {
                mma_commit(empty_bar(s));                    // frees the stage (smem B, staging A, TMEM A) when these retire
                if (kb == nkb - 1) mma_commit(accum_bar);    // accumulator complete
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == PRODUCERS / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P::TMEM_COLS));
    }
}

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
static void launch(dqn_learner* h, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, int Z, int kchunk) {
    using P = Plan<BN, NTERMS, !AOp::KCONTIG>;
    auto kern = tsgemm_kernel<BN, NTERMS, SPLIT, AOp, BOp, Epi>;
    static bool configured = false;  // one per template instantiation
    if (!configured) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::BYTES + 1024));
        configured = true;
    }
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), Z);
    kern<<<grid, THREADS, P::BYTES + 1024, h->stream>>>(a, b, e, M, N, K, kchunk, h->zeros16);
    DQN_CUDA_OK(cudaGetLastError());
    h->launches++;
}

}  // namespace ts
