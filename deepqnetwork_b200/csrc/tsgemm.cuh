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
// Thread roles (544 threads).  The k-loop of the first version was bound by the INSTRUCTIONS its 16 producer warps
// issued per k-block (every thread paid the barrier / stage / predicate bookkeeping for two 16-byte chunks), so the
// roles are split and each thread amortises that bookkeeping over a whole row or 4-8 chunks:
//   warps 0-3   A loaders (K-contiguous A): 8 cp.async chunks per thread and k-block, one shared k cursor;
//               they run ahead of the converters by the depth of the ring
//   warps 4-7   B producers: global -> registers (2-3 k-blocks ahead) -> hi/lo shared-memory tiles
//   warps 8-15  A converters, two groups taking alternate k-blocks: thread == row, 128 bytes of the staging row (or 32
//               coalesced scalar loads when A is M-contiguous; warps 0-3 then form a third group) -> hi/lo -> two
//               tcgen05.st.x32
//   warp 16     allocates TMEM; one elected lane issues the MMAs
// and all 16 then run the epilogue (warp w owns TMEM lanes 32*(w%4).., column group w/4).
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
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
template <class AOp> struct capable { static constexpr bool value = AOp::KCONTIG || has_addr1<AOp>::value; };  // M-contiguous: needs kstride()

// operand functors whose per-chunk address needs im2col arithmetic ask for 8 loader warps instead of 4
template <class Op, class = void> struct is_heavy { static constexpr bool value = false; };
template <class Op> struct is_heavy<Op, decltype((void)Op::HEAVY)> { static constexpr bool value = Op::HEAVY; };

constexpr int A_PITCH = 144;  // staging row pitch: 128 data bytes + 16, so that 8 consecutive rows hit 8 distinct 16-byte bank groups

template <int BN, int NTERMS>
struct Plan {
    static constexpr int A_STAGE = BM * A_PITCH;  // K-contiguous: [128 rows][144]; M-contiguous: [32 k][512] fits too
    static constexpr int B_BYTES = Tile<BN>::BYTES;
    static constexpr int STAGE = A_STAGE + B_BYTES * (NTERMS == 3 ? 2 : 1);
    // Two MMA-issuing warps for N <= 64 (one thread issues a tcgen05.mma every ~53 cycles, the tensor pipe needs N / 2: DESIGN.md D19):
    // warp i takes the k-blocks kb % NI == i into its own accumulator (column i * DC1), the epilogue adds them in index order.  The ring
    // depth is a multiple of NI, so that a warp waits on the barriers of "its" stages only and sees every phase of them (D20).
    static constexpr int NI = BN <= 64 ? 2 : 1;
    static constexpr int DC1 = BN < 32 ? 32 : BN;
    static constexpr int DCOLS = NI * DC1;
    static constexpr int ACOLS = BK * (NTERMS == 3 ? 2 : 1);
    static constexpr int S_TMEM = (512 - DCOLS) / ACOLS;
    static constexpr int S_SMEM = (196 * 1024) / STAGE;
    static constexpr int S0 = S_TMEM < S_SMEM ? S_TMEM : S_SMEM;
    static constexpr int STAGES = (S0 > 6 ? 6 : S0) / NI * NI;
    static_assert(STAGES >= 3, "pipeline too shallow");
    static constexpr int NG = 2;                                // converter groups (take k-blocks round-robin)
    static constexpr int EPI_BYTES = BM * (BN + 4) * 4;         // coalesced-epilogue transpose tile
    static constexpr int PIPE_BYTES = STAGES * STAGE;
    static constexpr int BYTES = (PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES) + 256;
    static constexpr int NEED = DCOLS + STAGES * ACOLS;
    static constexpr uint32_t TMEM_COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
};

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(((is_heavy<AOp>::value ? 20 : 16) + Plan<BN, NTERMS>::NI) * 32, 1) tsgemm_kernel(const AOp a, const BOp b, const Epi epi, int M, int N, int K,
                                                            int kchunk, int nclass, const float* __restrict__ zeros,
                                                            unsigned long long* __restrict__ dbg) {
    constexpr bool A_MC = !AOp::KCONTIG;  // A contiguous along M: staged as [k][m], converters read a column
    using P = Plan<BN, NTERMS>;
    constexpr int S = P::STAGES, NG = P::NG;
    constexpr int NLA = is_heavy<AOp>::value ? 8 : 4;  // loader warps; then 4 B-producer warps, 8 converter warps, 1 MMA warp
    constexpr int MMA_WARP = NLA + 12;
    using TB = Tile<BN>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bars = sbase + P::BYTES - 256;  // full[S], empty[S], afull[S], accum, tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    auto afull_bar = [&](int s) { return bars + 8u * (2 * S + s); };
    const uint32_t accum_bar = bars + 8u * (3 * S);
    const uint32_t tmem_slot = bars + 8u * (3 * S + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // blockIdx.z: split-K index, or a group id handed to the functors (dgrad parity class), or -- nclass > 0 -- both:
    // class = blockIdx.z % nclass for the operand functors, split = blockIdx.z / nclass; the epilogue sees blockIdx.z
    const int z = nclass > 0 ? blockIdx.z % nclass : blockIdx.z;
    const int zsplit = nclass > 0 ? blockIdx.z / nclass : blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = SPLIT ? zsplit * kchunk : 0;
    const int kend = SPLIT ? min(K, kbeg + kchunk) : K;
    const int nkb = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;
    const int ktot = kend - kbeg;
    const bool cta0 = dbg && (blockIdx.x + blockIdx.y + blockIdx.z == 0);  // in-kernel timeline (dqn_debug_timeline)
#define TS_STAMP(cond, kb, col) do { if (cta0 && (cond) && (kb) < 64) dbg[(kb) * 8 + (col)] = clock64(); } while (0)

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 8);      // 4 converter warps (the group that owns the block) + 4 B-producer warps
            mbar_init(empty_bar(s), 1);
            mbar_init(afull_bar(s), 32 * NLA);  // one pre-counted cp.async arrival per A-loader thread
        }
        mbar_init(accum_bar, nkb >= P::NI ? P::NI : nkb > 0 ? nkb : 1);  // the last commit of every issuing warp that has a k-block
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(P::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();  // everything above overlapped the predecessor's tail; its outputs are visible from here on

    auto stage_a = [&](int s) { return sbase + s * P::STAGE; };  // raw staging tile (K-contiguous A only)
    auto stage_b = [&](int s) { return sbase + s * P::STAGE + P::A_STAGE; };
    auto stage_blo = [&](int s) { return stage_b(s) + P::B_BYTES; };
    auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(P::DCOLS + s * P::ACOLS); };
    // number of k-blocks in which k offset `koff` (inside a block) is still below kend
    auto nvalid = [&](int koff) { return ktot > koff ? (ktot - koff + BK - 1) / BK : 0; };

    if (warp < MMA_WARP) {
        const bool is_b = warp >= NLA && warp < NLA + 4;
        const bool is_loader = warp < NLA;
        if (is_b) {
            // ================= B producers (warps 4-7) =================
            const int t = tid - 32 * NLA, w4 = warp - NLA;
            // K-contiguous: 16-byte chunk (t & 7) of rows (t >> 3) + 16 i.  N-contiguous: the transposing assignment of
            // tcgemm.cuh folded onto 4 warps: k = 4*(w4 + 4 u) + (lane >> 3) % 4, rows 32 h + 4 (lane & 7) .. +3
            constexpr int NH = BN / 32;
            constexpr int LB = BOp::KCONTIG ? BN / 16 : 2 * NH;
            const int koff0 = BOp::KCONTIG ? (t & 7) * 4 : w4 * 4 + ((lane >> 3) & 3);
            int b_r[LB];
            uint32_t b_off[LB];
            bool b_ok[LB];
            typename BOp::Row b_row[LB];
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                const int koff = BOp::KCONTIG ? koff0 : koff0 + 16 * (i / NH);
                b_r[i] = BOp::KCONTIG ? (t >> 3) + 16 * i : 32 * (i % NH) + 4 * (lane & 7);
                b_off[i] = TB::off(b_r[i], koff);
                b_ok[i] = n0 + b_r[i] < N;
                b_row[i] = b.row(min(n0 + b_r[i], N - 1), z);
            }
            const int nv0 = nvalid(koff0), nv1 = nvalid(koff0 + 16);
            typename BOp::Cur b_cur = b.cur(kbeg + koff0, z);
            constexpr int PF = LB <= 4 ? 3 : 2;  // k-blocks of loads in flight per thread
            float4 pf[PF][LB];
            int kload = 0;
            auto load_regs = [&](float4 (&bp)[LB]) {
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    const float* src = nullptr;
                    if constexpr (BOp::KCONTIG) {
                        if (b_ok[i] && kload < nv0) src = b.ptr(b_row[i], n0 + b_r[i], b_cur);
                    } else {
                        const bool second = i >= NH;
                        if (b_ok[i] && kload < (second ? nv1 : nv0)) {
                            src = b.ptr(b_row[i], n0 + b_r[i], b_cur);
                            if (second) src += (size_t)16 * b.kstride();
                        }
                    }
                    // ragged edges read 16 zero bytes: the select is on the ADDRESS, nothing waits on the value here
                    bp[i] = __ldg(reinterpret_cast<const float4*>(src ? src : zeros));
                }
                b.next(b_cur);
                ++kload;
            };
            auto sts_v = [&](uint32_t addr, const float4& v) {
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            };
            auto sts_t = [&](uint32_t addr, const float4& v) {
                asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+16], %2;\n\tst.shared.f32 [%0+32], %3;\n\tst.shared.f32 [%0+48], %4;"
                             ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            };
            int s = 0, ph = 0;
            auto block = [&](int kb, float4 (&bp)[LB]) {
                if (kb >= S) mbar_wait(empty_bar(s), ph ^ 1);  // MMAs of block kb-S are done with stage s
                TS_STAMP(t == 0, kb, 4);
                const uint32_t hi_t = stage_b(s), lo_t = stage_blo(s);
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    const float4 v = bp[i];
                    const float4 l = make_float4(lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w));
                    if (BOp::KCONTIG) { sts_v(hi_t + b_off[i], v); if (NTERMS == 3) sts_v(lo_t + b_off[i], l); }
                    else { sts_t(hi_t + b_off[i], v); if (NTERMS == 3) sts_t(lo_t + b_off[i], l); }
                }
                // release-arrive, one per warp; the generic->async proxy fence runs on the MMA thread after its acquire
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(s));
                TS_STAMP(t == 0, kb, 5);
                if (++s == S) { s = 0; ph ^= 1; }
                load_regs(bp);
            };
#pragma unroll
            for (int p = 0; p < PF; ++p) load_regs(pf[p]);
            for (int kb = 0; kb < nkb; kb += PF) {
#pragma unroll
                for (int p = 0; p < PF; ++p)
                    if (kb + p < nkb) block(kb + p, pf[p]);
            }
        } else if (is_loader) {
            // ================= A loaders (warps 0-3): cp.async ring, S-1 k-blocks (16 KB each) in flight =================
            int sj = 0, phj = 0;
            if constexpr (!A_MC) {
                // K-contiguous A: chunk c of rows r0 + RS i (8 lanes cover 128 contiguous bytes); staging [row][144]
                constexpr int RS = 4 * NLA, NCH = 32 / NLA;  // row step, chunks per thread
                const int c = tid & 7, r0 = tid >> 3;
                typename AOp::Row a_row[NCH];
                unsigned ok = 0;
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    const int r = m0 + r0 + RS * i;
                    if (r < M) ok |= 1u << i;
                    a_row[i] = a.row(min(r, M - 1), z);
                }
                const int nv = nvalid(c * 4);
                typename AOp::Cur cur = a.cur(kbeg + c * 4, z);
                const uint32_t dst0 = (uint32_t)(r0 * A_PITCH + c * 16);
                for (int j = 0; j < nkb; ++j) {
                    if (j >= S) mbar_wait(empty_bar(sj), phj ^ 1);  // stage sj was last used by block j-S
                    TS_STAMP(tid == 0, j, 0);
                    const uint32_t dst = stage_a(sj) + dst0;
                    const unsigned okj = j < nv ? ok : 0u;
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        const float* src = ((okj >> i) & 1u) ? a.ptr(a_row[i], m0 + r0 + RS * i, cur) : nullptr;
                        cp_async16(dst + i * RS * A_PITCH, src ? src : zeros, src != nullptr);
                    }
                    cp_async_arrive_noinc(afull_bar(sj));
                    TS_STAMP(tid == 0, j, 1);
                    a.next(cur);
                    if (++sj == S) { sj = 0; phj ^= 1; }
                }
            } else {
                // M-contiguous A: 16-byte chunk mq along M (a warp covers 512 contiguous bytes of one k), k rows
                // (tid >> 5) + 4 i; staging [k][512 bytes]
                static_assert(NLA == 4, "M-contiguous loader is written for 4 warps");
                const int mq = tid & 31, k0 = tid >> 5;
                const bool m_ok = m0 + 4 * mq < M;
                typename AOp::Row a_row = a.row(min(m0 + 4 * mq, M - 1), z);
                typename AOp::Cur cur = a.cur(kbeg + k0, z);
                const size_t kstep = (size_t)4 * a.kstride();
                const uint32_t dst0 = (uint32_t)(k0 * 512 + mq * 16);
                int left = ktot - k0;  // k rows of this thread still inside [kbeg, kend): i valid iff 4 i < left
                for (int j = 0; j < nkb; ++j) {
                    if (j >= S) mbar_wait(empty_bar(sj), phj ^ 1);
                    TS_STAMP(tid == 0, j, 0);
                    const uint32_t dst = stage_a(sj) + dst0;
                    const float* p = a.ptr(a_row, m0 + 4 * mq, cur);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool v = m_ok && 4 * i < left;
                        cp_async16(dst + i * 4 * 512, v ? p + i * kstep : zeros, v);
                    }
                    cp_async_arrive_noinc(afull_bar(sj));
                    TS_STAMP(tid == 0, j, 1);
                    a.next(cur);
                    left -= BK;
                    if (++sj == S) { sj = 0; phj ^= 1; }
                }
            }
        } else {
            // ================= A converters: thread == row, the two groups take alternate k-blocks =================
            const int grp = warp & 3;
            const int g = (warp - NLA - 4) >> 2;  // converter group
            const int my_row = grp * 32 + lane;
            const uint32_t lane_base = (uint32_t)(grp * 32) << 16;
            int s = g % S, ph = 0;
            for (int kb = g; kb < nkb; kb += NG) {
                if (kb >= S) mbar_wait(empty_bar(s), ph ^ 1);  // TMEM stage s is free again
                mbar_wait(afull_bar(s), ph);                   // every loader's copies for this block have landed
                TS_STAMP(grp == 0 && lane == 0, kb, 2);
                float hi[32];
                if constexpr (!A_MC) {
                    const uint32_t src = stage_a(s) + my_row * A_PITCH;  // 8 consecutive rows -> 8 distinct bank groups
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(hi[4 * c]), "=f"(hi[4 * c + 1]), "=f"(hi[4 * c + 2]), "=f"(hi[4 * c + 3]) : "r"(src + c * 16));
                } else {
                    const uint32_t src = stage_a(s) + my_row * 4;        // lanes read consecutive words of one k row
#pragma unroll
                    for (int j = 0; j < 32; ++j) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[j]) : "r"(src + j * 512));
                }
                // hi = the raw fp32 word (the tensor core reads its top 19 bits); lo = what that drops
                const uint32_t ta = tmem_a(s) + lane_base;
                tmem_st32(ta, hi);
                if (NTERMS == 3) {
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) lo[j] = lo_part(hi[j]);
                    tmem_st32(ta + BK, lo);
                }
                tmem_st_wait();
                tc_fence_before();
                TS_STAMP(grp == 0 && lane == 0, kb, 3);
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(s));
                s += NG; if (s >= S) { s -= S; ph ^= 1; }
            }
        }
        // the k-loop work of this warp is issued: let the stream successor start its prologue (it occupies a whole SM,
        // so triggering any earlier would take idle SMs away from the kernels of the other streams)
        pdl_launch_dependents();
        // ---------------- epilogue (warps 0-15) ----------------
        if (warp < PRODUCERS / 32) {
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        const int sub = warp & 3, quarter = warp >> 2;
        const int m = m0 + sub * 32 + lane;
        constexpr int QCOLS = BN / 4;  // columns per warp: 8, 16 or 32
        const uint32_t trow = tmem_base + ((uint32_t)(sub * 32) << 16);
        // columns [col, col + 8 / 16) of the tile: sum of the issuing warps' accumulators, in index order
        auto ld8s = [&](int col, float (&v)[8]) {
            tmem_ld8(trow + (uint32_t)col, v);
#pragma unroll
            for (int i = 1; i < P::NI; ++i)
                if (i < nkb) {
                    float v1[8];
                    tmem_ld8(trow + (uint32_t)(i * P::DC1 + col), v1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += v1[j];
                }
        };
        auto ld16s = [&](int col, float (&v)[16]) {
            if (P::NI == 2 && nkb >= 2) { tmem_ld16_sum2(trow + (uint32_t)col, trow + (uint32_t)(P::DC1 + col), v); return; }
            tmem_ld16(trow + (uint32_t)col, v);
        };
        if constexpr (wants_coalesce<Epi>::value) {
            // accumulators -> smem tile [128][BN+4] (the pipeline buffers are idle now) -> rows handed to whole warps
            constexpr int PITCH = BN + 4;
            float* tile = reinterpret_cast<float*>(smem + (sbase - smem_u32(smem)));
            const int row = sub * 32 + lane;
            if (QCOLS == 8) {
                float v[8];
                if (nkb > 0) ld8s(quarter * 8, v);
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
                    if (nkb > 0) ld16s(quarter * QCOLS + c, v);
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
            if (nkb > 0) ld8s(col, v);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            if (m < M && n0 + col < N) epi.template store<8>(m, n0 + col, v, (int)blockIdx.z);
        } else {
#pragma unroll
            for (int c = 0; c < QCOLS; c += 16) {
                const int col = quarter * QCOLS + c;
                float v[16];
                if (nkb > 0) ld16s(col, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (m < M && n0 + col < N) epi.template store<16>(m, n0 + col, v, (int)blockIdx.z);
            }
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
        const int wi = warp - MMA_WARP;  // issuing warp: k-blocks kb % NI == wi, accumulator wi
        const uint32_t tmem_d = tmem_base + (uint32_t)(wi * P::DC1);
        for (int kb = wi; kb < nkb; kb += P::NI) {
            const int s = kb % S, ph = (kb / S) & 1;
            mbar_wait(full_bar(s), ph);             // acquire: B tile stores and A TMEM stores of this block
            TS_STAMP(lane == 0, kb, 6);
            fence_proxy_async();                      // B tile: generic -> async proxy
            tc_fence_after();
            const uint64_t so = (uint64_t)s * STAGE16;
            const uint32_t ta = tmem_a(s);
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                const uint64_t o = so + ks * KS16;
                const uint32_t acc = (kb != wi) || ks != 0;
                if (leader) {
                    if (NTERMS == 3) {
                        mma_tf32_ts(tmem_d, ta + ks * 8, dbl0 + o, idesc, acc);       // A_hi * B_lo
                        mma_tf32_ts(tmem_d, ta + BK + ks * 8, db0 + o, idesc, 1);    // A_lo * B_hi
                        mma_tf32_ts(tmem_d, ta + ks * 8, db0 + o, idesc, 1);         // A_hi * B_hi
                    } else {
                        mma_tf32_ts(tmem_d, ta + ks * 8, db0 + o, idesc, acc);
                    }
                }
            }
            if (leader) {
                mma_commit(empty_bar(s));                         // frees the stage (smem B, staging A, TMEM A) when these retire
                if (kb + P::NI >= nkb) mma_commit(accum_bar);     // this warp's accumulator is complete
            }
            TS_STAMP(lane == 0, kb, 7);
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P::TMEM_COLS));
    }
}

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
static void launch(dqn_learner* h, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, int Z, int kchunk, int nclass = 0) {
    using P = Plan<BN, NTERMS>;
    auto kern = tsgemm_kernel<BN, NTERMS, SPLIT, AOp, BOp, Epi>;
    static bool configured = false;  // one per template instantiation
    if (!configured) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::BYTES + 1024));
        configured = true;
    }
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), Z);
    launch_kernel(h->pdl, kern, grid, dim3(((is_heavy<AOp>::value ? 20 : 16) + P::NI) * 32), (size_t)P::BYTES + 1024, h->stream, a, b, e, M, N, K, kchunk, nclass,
                  (const float*)h->zeros16, h->tc_dbg);
    h->launches++;
}

#undef TS_STAMP
}  // namespace ts
