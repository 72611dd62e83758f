// Image-resident convolution on tcgen05: one CTA = one image of the batch.
//
// Why (profiles/README.md, r01b): the generic implicit-GEMM kernel (tsgemm.cuh) gathers every 128 x 32 im2col tile
// from L2 with 16-byte cp.async, and ALL CTAs re-stage the same weight tile through registers.  At batch 32-64 the
// k-loop is then bound by the ~2300 warp instructions per k-block that the im2col loaders, the B producers and the
// hi/lo converters issue -- the tensor pipe is 16 % busy and DRAM idle.  A conv layer of this network has a much
// smaller footprint than its im2col view (conv3: 30 KB per image vs 480 KB of im2col rows), so here
//   * the zero-padded input image is copied ONCE into shared memory (16-byte cp.async, zero-fill for the SAME
//     padding), columns de-interleaved by stride phase and pixels padded to cin*4+16 bytes so that the row-per-thread
//     reads below are bank-conflict free;
//   * the weights arrive pre-split (hi = fp32 word, lo = x - tf32(x)) and pre-tiled in the canonical K-major UMMA
//     layout (pack_conv_kernel, run once per optimizer step), so one elected thread streams them with
//     cp.async.bulk (TMA) straight into the B ring: no B-producer warps;
//   * per k-block the only thread work left is the A converter: thread == output pixel reads its 128 contiguous
//     bytes (one tap x 32 channels, or one kernel row of a 4-channel input) from the resident image at
//     base(pixel) + offset(k-block), splits hi/lo and stores both to TENSOR MEMORY (tcgen05.st); the MMA takes A
//     from TMEM and B through a shared-memory descriptor, exactly as in tsgemm.cuh.
// Layers whose output has more than 128 pixels per image (conv1: 23 x 20) are processed as bands of whole output rows
// with one accumulator (BN TMEM columns) per band; the image and the TMEM allocation are shared by the bands.
//
// Thread roles (320 threads): warps 0-7 image loaders, then A converters (two groups of 4 warps take alternate
// k-blocks), then the epilogue; warp 8 allocates TMEM and one elected lane issues the MMAs; warp 9: one elected lane
// issues the weight TMA copies.
#pragma once
#include "tsgemm.cuh"
#include "cvpack.cuh"

namespace cv {
using namespace tc;

struct ImgGeom {
    int ih, iw, cin, oh, ow, cout, k, s, pt, pl;
    int hp;       // padded rows resident in shared memory: (oh-1)*s + k
    int wps;      // columns per stride phase: ceil(((ow-1)*s + k) / s)
    int wcols;    // s * wps
    int pitch;    // bytes per pixel in shared memory (cin*4, +16 when that is a multiple of 128)
    int cpp;      // 16-byte chunks per pixel
    int nkb;      // k*k*cin / 32
    int th;       // output rows per band (th * ow <= 128)
    int nbands;   // ceil(oh / th)
    int img_bytes;
    unsigned m_wcols, m_wps;  // fastdiv magics (learner.h)
};

// What the (up to 4) accumulators of one CTA cover.  A band is a set of <= 128 output pixels (a, b), a0 <= a < a0 + na,
// b0 <= b < b0 + nb: thread row m = (a - a0) * nb + (b - b0) reads the resident image at pixel (a * rs, b) (+ the tap
// of the k-block) and its result goes to output pixel (a * ym + ya, b * xm + xa) of an out_h x out_w map.
//   forward : bands = groups of whole output rows, rs = stride, identity output map, one set of packed weights
//   dgrad   : bands = stride-parity classes (ry, rx) of the input pixels, image = dY, rs = 1, ym = xm = stride,
//             ya = ry - pad, each class with its own packed (flipped, transposed) weights starting at k-block pk_kb0
struct Bands {
    int n, rs, ym, xm, out_w, out_hw;
    struct B { int a0, na, b0, nb, ya, xa, pk_kb0; } c[4];
};

static inline bool img_geom(const ConvGeom& g, ImgGeom& q) {
    q.ih = g.ih; q.iw = g.iw; q.cin = g.cin; q.oh = g.oh; q.ow = g.ow; q.cout = g.cout; q.k = g.k; q.s = g.s; q.pt = g.pt; q.pl = g.pl;
    if (!((g.cin % 32 == 0) || (g.cin == 4 && g.k == 8))) return false;  // a k-block = 128 contiguous source bytes
    if (g.ow > 128 || (g.cout != 32 && g.cout != 64)) return false;
    q.hp = (g.oh - 1) * g.s + g.k;
    q.wps = ((g.ow - 1) * g.s + g.k + g.s - 1) / g.s;
    q.wcols = g.s * q.wps;
    q.cpp = g.cin / 4;
    q.pitch = g.cin * 4 + ((g.cin * 4) % 128 == 0 ? 16 : 0);
    q.nkb = g.k * g.k * g.cin / 32;
    q.th = 128 / g.ow; if (q.th > g.oh) q.th = g.oh;
    q.nbands = (g.oh + q.th - 1) / q.th;
    q.img_bytes = ((q.hp * q.wcols * q.pitch + 127) / 128) * 128;
    if (q.nkb > 64 || q.nkb < 1) return false;
    q.m_wcols = fastdiv_magic((unsigned)q.wcols); q.m_wps = fastdiv_magic((unsigned)q.wps);
    return true;
}

static inline void forward_bands(const ImgGeom& q, Bands& bd) {
    bd.n = q.nbands; bd.rs = q.s; bd.ym = 1; bd.xm = 1; bd.out_w = q.ow; bd.out_hw = q.oh * q.ow;
    for (int b = 0; b < q.nbands && b < 4; ++b) {
        const int a0 = b * q.th;
        bd.c[b] = {a0, q.oh - a0 < q.th ? q.oh - a0 : q.th, 0, q.ow, 0, 0, 0};
    }
}

// data gradient of conv g as an image-resident convolution over dY: q describes the resident (padded) dY image and
// the k-blocks (taps (th, tw) of the k/s x k/s sub-kernel x cout), bd the stride-parity classes.
//   dX[s a + ry - pt][s b + rx - pl][ci] = sum_{th, tw, co} dYpad[a + th][b + tw][co] * W[ry + s (kt-1-th)][rx + s (kt-1-tw)][ci][co]
// with kt - 1 zero rows / columns in front of dY.
static inline bool dgrad_geom(const ConvGeom& g, ImgGeom& q, Bands& bd) {
    if (g.k % g.s != 0 || g.cout % 32 != 0 || (g.cin != 32 && g.cin != 64) || g.s * g.s > 4) return false;
    const int kt = g.k / g.s;
    int amax = 0, bmax = 0;
    bd.n = g.s * g.s; bd.rs = 1; bd.ym = g.s; bd.xm = g.s; bd.out_w = g.iw; bd.out_hw = g.ih * g.iw;
    const int nkb_class = kt * kt * g.cout / 32;
    for (int ry = 0; ry < g.s; ++ry)
        for (int rx = 0; rx < g.s; ++rx) {
            const int a0 = g.pt > ry ? (g.pt - ry + g.s - 1) / g.s : 0, b0 = g.pl > rx ? (g.pl - rx + g.s - 1) / g.s : 0;
            const int na = (g.ih - 1 - ry + g.pt) / g.s - a0 + 1, nb = (g.iw - 1 - rx + g.pl) / g.s - b0 + 1;
            if (na < 1 || nb < 1 || na * nb > 128) return false;
            bd.c[ry * g.s + rx] = {a0, na, b0, nb, ry - g.pt, rx - g.pl, (ry * g.s + rx) * nkb_class};
            if (a0 + na > amax) amax = a0 + na;
            if (b0 + nb > bmax) bmax = b0 + nb;
        }
    q.ih = g.oh; q.iw = g.ow; q.cin = g.cout; q.oh = g.ih; q.ow = g.iw; q.cout = g.cin; q.k = kt; q.s = 1;
    q.pt = kt - 1; q.pl = kt - 1;
    q.hp = amax + kt - 1; q.wps = bmax + kt - 1; q.wcols = q.wps;
    q.cpp = q.cin / 4;
    q.pitch = q.cin * 4 + ((q.cin * 4) % 128 == 0 ? 16 : 0);
    q.nkb = nkb_class;
    q.th = 0; q.nbands = bd.n;
    q.img_bytes = ((q.hp * q.wcols * q.pitch + 127) / 128) * 128;
    q.m_wcols = fastdiv_magic((unsigned)q.wcols); q.m_wps = fastdiv_magic((unsigned)q.wps);
    return q.nkb >= 1 && q.nkb <= 64;
}

static inline size_t packed_floats(const ConvGeom& g) { return (size_t)g.k * g.k * g.cin * g.cout * 2; }

// all layers of one tower in ONE launch (it runs after set_param / restore / target sync; every step's re-pack rides in the
// optimizer tail): blockIdx.y = job.  W is HWIO: W[kk][n], kk = (kh*k + kw)*cin + c.  kind 1: the dgrad operand, per stride
// class (ry, rx): Wd[(th, tw, co)][ci] = W[ry + s (kt-1-th)][rx + s (kt-1-tw)][ci][co] (N = cin rows), class c starting at
// k-block c * kt*kt*cout/32.  kind 2: the uint8 form of layer 0 (pk_store_u8).
struct PackJobs {
    int n;
    struct J { const float* W; float* out; int K, N, k, s, cin, cout, kind; } j[8];
};
__global__ void pack_all_kernel(const PackJobs jobs) {
    const PackJobs::J& q = jobs.j[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.K * q.N) return;
    float v; int kb, kin, n, rows;
    if (q.kind != 1) {
        const int kk = i / q.N;
        n = i - kk * q.N; v = q.W[i]; kb = kk >> 5; kin = kk & 31; rows = q.N;
    } else {
        const int kt = q.k / q.s, Kc = kt * kt * q.cout;
        const int ci = i % q.cin, r = i / q.cin;
        const int kk = r % Kc, cls = r / Kc;
        const int co = kk % q.cout, tap = kk / q.cout;
        const int th = tap / kt, tw = tap - th * kt;
        const int ry = cls / q.s, rx = cls - ry * q.s;
        const int kh = ry + q.s * (kt - 1 - th), kw = rx + q.s * (kt - 1 - tw);
        v = q.W[((size_t)(kh * q.k + kw) * q.cin + ci) * q.cout + co];
        kb = cls * (Kc / 32) + (kk >> 5); kin = kk & 31; n = ci; rows = q.cin;
    }
    uint8_t* blk = reinterpret_cast<uint8_t*>(q.out) + (size_t)kb * rows * 256;
    if (q.kind == 2) pk_store_u8(blk, rows, n, kin, v); else pk_store(blk, rows, n, kin, v);
}

// instruction descriptor of the bf16 cross-term MMAs: kind::f16, BF16 x BF16 -> fp32, M = 128, K-major operands
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32u(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// A tile of one k-block into its tensor-memory stage (64 columns: hi fp32 | bf16(x) pairs | bf16(lo) pairs), NTERMS == 3
__device__ __forceinline__ void a_stage_store3(uint32_t ta, const float (&hi)[32]) {
    ts::tmem_st32(ta, hi);
    uint32_t x[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        x[j] = bf16_pair(hi[2 * j], hi[2 * j + 1]);
        x[16 + j] = bf16_pair(lo_part(hi[2 * j]), lo_part(hi[2 * j + 1]));
    }
    tmem_st32u(ta + BK, x);
}
// the 8 MMAs of one k-block (one elected thread): ta = A stage, sb = B k-block in shared memory (N rows, layout above)
__device__ __forceinline__ void mma_kblock3(uint32_t tmem_d, uint32_t ta, uint32_t sb, int N, uint32_t idesc, uint32_t idesc16, bool first) {
    const uint64_t db = make_desc(sb, PK_LBO, PK_SBO);
    const uint64_t dbh = make_desc(sb + 128u * N, PK_LBO, PK_SBO16);
    const uint64_t dbl = make_desc(sb + 192u * N, PK_LBO, PK_SBO16);
    constexpr uint64_t STEP = (uint64_t)(2 * PK_LBO) >> 4;  // two core matrices along K per instruction, either kind
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        mma_bf16_ts(tmem_d, ta + BK + j * 8, dbl + j * STEP, idesc16, !(first && j == 0));   // A_hi * B_lo
        mma_bf16_ts(tmem_d, ta + BK + 16 + j * 8, dbh + j * STEP, idesc16, 1);                 // A_lo * B_hi
    }
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) ts::mma_tf32_ts(tmem_d, ta + ks * 8, db + ks * STEP, idesc, 1);  // A_hi * B_hi
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int BN, int NTERMS>
struct CvPlan {
    static constexpr int B_TILE = BN * 128;                       // bytes of the fp32 tile of a k-block (the two bf16 tiles: half each)
    static constexpr int B_STAGE = B_TILE * 2;                    // one k-block of packed weights (cvpack.cuh)
    static constexpr int ACOLS = BK * (NTERMS == 3 ? 2 : 1);
    static constexpr int MAXB = 4;                                // bands per CTA; also the number of accumulators / issuing warps
    static constexpr int DCOLS = MAXB * BN;                       // accumulator a at column a * BN
    static constexpr int S0 = (512 - DCOLS) / ACOLS;              // ring depth: B tiles in smem, A tiles in TMEM
    static constexpr int S = S0 > 6 ? 6 : S0;
    static constexpr int BAR_BYTES = 1024 + 4 * 128 * 8;          // barriers, tmem slot, k-block offsets, bias; row tables of 4 bands
};

// zero-padded, stride-phase-split copy of one NHWC image into shared memory (layout in the header comment); nthreads
// threads cooperate (cpp divides nthreads); the caller waits for the copies and synchronises
__device__ __forceinline__ void load_image_async(uint32_t img, const float* __restrict__ src_img, const ImgGeom& g, int tid,
                                                 int nthreads, const float* __restrict__ zeros) {
    const int c = tid % g.cpp, pstep = nthreads / g.cpp, npix = g.hp * g.wcols;
    for (int p = tid / g.cpp; p < npix; p += pstep) {
        const int iyp = fdiv(p, g.m_wcols, g.wcols), col = p - iyp * g.wcols;
        const int phase = fdiv(col, g.m_wps, g.wps), j = col - phase * g.wps;
        const int iy = iyp - g.pt, ix = j * g.s + phase - g.pl;
        const bool ok = (unsigned)iy < (unsigned)g.ih && (unsigned)ix < (unsigned)g.iw;
        const float* src = ok ? src_img + ((size_t)(iy * g.iw + ix) * g.cin + c * 4) : zeros;
        cp_async16(img + (uint32_t)(p * g.pitch + c * 16), src, ok);
    }
}

constexpr int CV_THREADS = 416;  // warps 0-7 converters / epilogue, warps 8-11 MMA issuers, warp 12 TMA

// Several MMA-issuing warps: one thread issues a tcgen05.mma every ~53 cycles whatever its shape, the tensor pipe needs N / 2
// (scratch/ubench/mma_issue.cu), so a single issuer leaves it 40-70 % idle at N <= 64.  The k-blocks of the CTA's bands are
// walked band-interleaved (step = kb * nb + band) and issuing warp i takes the steps with step % NI == i into ITS OWN
// accumulator (deterministic order inside each): 4 bands -> one accumulator per band, 2 bands -> two per band (even / odd
// k-blocks), 1 band -> four; the epilogue adds the accumulators of a band in index order.
template <int BN, int NTERMS, class Epi>
__global__ void __launch_bounds__(CV_THREADS, 1) cvconv_kernel(const float* __restrict__ in_a, const float* __restrict__ packed_a,
                                                               const Epi epi_a, int images_a, const float* __restrict__ in_b,
                                                               const float* __restrict__ packed_b, const Epi epi_b,
                                                               const ImgGeom g, const Bands bd, int bands_per_cta,
                                                               const float* __restrict__ zeros, unsigned long long* __restrict__ dbg) {
    using P = CvPlan<BN, NTERMS>;
    const bool cta0 = dbg && blockIdx.x == 0;  // in-kernel timeline (dqn_debug_timeline, mode >= 16)
#define CV_STAMP(cond, row, col) do { if (cta0 && (cond) && (row) < 64) dbg[(row) * 8 + (col)] = clock64(); } while (0)
    CV_STAMP(threadIdx.x == 0, 63, 0);
    constexpr int S = P::S;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t ring = sbase;                                  // S x B_STAGE (1024-aligned tiles)
    const uint32_t bars = ring + S * P::B_STAGE;                  // afull[S], bfull[S], empty[S], accum[MAXB], tmem slot
    const uint32_t kbtab = bars + 256;                            // int kboff[64]
    const uint32_t bias_s = bars + 512;                           // float bias[BN] (epilogue staging area)
    const uint32_t rowtab = bars + 1024;                          // per band and tile row: {image byte offset, output pixel or -1}
    const uint32_t img = bars + P::BAR_BYTES;
    auto afull_bar = [&](int s) { return bars + 8u * s; };
    auto bfull_bar = [&](int s) { return bars + 8u * (S + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * S + s); };
    auto accum_bar = [&](int b) { return bars + 8u * (3 * S + b); };
    const uint32_t tmem_slot = bars + 8u * (3 * S + P::MAXB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // two problem sets in one grid (online tower on [s; s'] and target tower on s'): CTAs past images_a take set b
    const bool set_b = (int)blockIdx.x >= images_a;
    const int image = set_b ? blockIdx.x - images_a : blockIdx.x;
    const float* __restrict__ in = set_b ? in_b : in_a;
    const float* __restrict__ packed = set_b ? packed_b : packed_a;
    const Epi& epi = set_b ? epi_b : epi_a;
    // a CTA owns bands [band0, band0 + nb) of its image (blockIdx.y): bands are independent accumulators, so an image's
    // bands can be spread over several CTAs (conv2 dgrad: 4 stride classes -> 2 CTAs) without any reduction
    const int band0 = blockIdx.y * bands_per_cta;
    const int nb = min(bd.n - band0, bands_per_cta);
    const int nsteps = nb * g.nkb;
    // issuing warps == accumulators; accumulator a belongs to band a % nb (nb = 1, 2 or 4).  Warp i's steps all have the parity
    // of i, i.e. they and every step that shares a stage with them (S is even) come from the SAME converter group, in order:
    // when the warp waits for a phase of a stage barrier, the phases it skipped (other warps' steps) are complete
    constexpr int NI = 4;
    const int per_band = NI / nb;      // accumulators (issuing warps) per band

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(afull_bar(s), 4);   // the 4 converter warps of the group that owns the block
            mbar_init(bfull_bar(s), 1);   // the TMA thread's expect_tx arrival (+ the bytes)
            mbar_init(empty_bar(s), 1);   // tcgen05.commit
        }
        for (int b = 0; b < P::MAXB; ++b) mbar_init(accum_bar(b), per_band);  // the last commit of every issuing warp of the band
        fence_barrier_init();
    }
    if (tid < g.nkb) {
        // k-block kb covers K elements [32 kb, 32 kb + 32) of (kh, kw, c): its 128 source bytes start at this offset
        // from the pixel a thread owns (cin >= 32: one tap, channel chunk; cin == 4: one kernel row, all kw)
        const int kk = tid * 32;
        const int tap = kk / g.cin, c = kk - tap * g.cin;
        const int kh = tap / g.k, kw = tap - kh * g.k;
        const int off = (kh * g.wcols + (kw % g.s) * g.wps + kw / g.s) * g.pitch + c * 4;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(kbtab + 4u * tid), "r"(off) : "memory");
    }
    for (int e = tid; e < nb * 128; e += CV_THREADS) {
        const int band = e >> 7, m = e & 127;
        const Bands::B& c = bd.c[band0 + band];
        const int ai = m / c.nb, bi = m - ai * c.nb;
        int off = 0, opix = -1;  // rows past the band: any finite data, never stored
        if (ai < c.na) {
            const int a = c.a0 + ai, b = c.b0 + bi;
            off = (a * bd.rs * g.wcols + b) * g.pitch;
            opix = image * bd.out_hw + (a * bd.ym + c.ya) * bd.out_w + (b * bd.xm + c.xa);
        }
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowtab + 8u * e), "r"(off), "r"(opix) : "memory");
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
    CV_STAMP(threadIdx.x == 0, 63, 1);
    auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(P::DCOLS + s * P::ACOLS); };
    auto stage_b = [&](int s) { return ring + s * P::B_STAGE; };
    auto step_kb = [&](int step) { return nb == 1 ? step : nb == 2 ? step >> 1 : step >> 2; };

    if (warp < 8) {
        pdl_wait();  // the input image is the stream predecessor's output
        CV_STAMP(tid == 0, 63, 2);
        // ---------------- image -> shared memory (once) ----------------
        {
            const float* src_img = in + (size_t)image * g.ih * g.iw * g.cin;
            epi.stage(tid, BN, bias_s);
            load_image_async(img, src_img, g, tid, 256, zeros);
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        CV_STAMP(tid == 0, 63, 3);
        // ---------------- A converters (two groups of 4 warps take alternate steps) ----------------
        const int grp = warp & 3, cg = warp >> 2;  // TMEM lane quarter, converter group
        const int m = grp * 32 + lane;             // row of the band tile
        const uint32_t lane_base = (uint32_t)(grp * 32) << 16;
        // the 8 chunks of a k-block: consecutive channels of one pixel, or (cin == 4) consecutive kw of one kernel row
        uint32_t choff[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) choff[j] = g.cin >= 32 ? 16u * j : (uint32_t)(((j % g.s) * g.wps + j / g.s) * g.pitch);
        for (int step = cg; step < nsteps; step += 2) {
            const int kb = step_kb(step), band = step - kb * nb;
            const int s = step % S, ph = (step / S) & 1;
            uint32_t roff, koff;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(roff) : "r"(rowtab + 8u * (band * 128 + m)));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(koff) : "r"(kbtab + 4u * kb));
            if (step >= S) mbar_wait(empty_bar(s), ph ^ 1);  // TMEM stage s is free again
            CV_STAMP(grp == 0 && lane == 0, step, 0);
            const uint32_t src = img + roff + koff;
            float hi[32];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(hi[4 * c]), "=f"(hi[4 * c + 1]), "=f"(hi[4 * c + 2]), "=f"(hi[4 * c + 3]) : "r"(src + choff[c]));
            const uint32_t ta = tmem_a(s) + lane_base;
            if (NTERMS == 3) a_stage_store3(ta, hi); else ts::tmem_st32(ta, hi);
            ts::tmem_st_wait();
            tc_fence_before();
            CV_STAMP(grp == 0 && lane == 0, step, 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(afull_bar(s));
        }
        CV_STAMP(tid == 0, 63, 4);
        pdl_launch_dependents();
        // ---------------- epilogue: warp (grp, cg) owns TMEM lanes 32 grp.., columns [cg BN/2, +BN/2) of every band ----------------
        constexpr int HC = BN / 2;
        for (int band = 0; band < nb; ++band) {
            mbar_wait(accum_bar(band), 0);
            tc_fence_after();
            CV_STAMP(tid == 0 && band == 0, 63, 5);
            int mg;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(mg) : "r"(rowtab + 8u * (band * 128 + m) + 4u));
            const uint32_t trow = tmem_base + lane_base + (uint32_t)(band * BN + cg * HC);
#pragma unroll
            for (int c = 0; c < HC; c += 16) {
                float v[16];  // the band's accumulators (index order), one wait for all of them
                const uint32_t t0 = trow + (uint32_t)c, ts_ = (uint32_t)(nb * BN);
                if (per_band == 4) tmem_ld16_sum4(t0, t0 + ts_, t0 + 2 * ts_, t0 + 3 * ts_, v);
                else if (per_band == 2) tmem_ld16_sum2(t0, t0 + ts_, v);
                else tmem_ld16(t0, v);
                if (mg >= 0) epi.template store_b<16>(mg, cg * HC + c, v, bias_s + 4u * (cg * HC + c));
            }
        }
        tc_fence_before();
        CV_STAMP(tid == 0, 63, 6);
    } else if (warp < 12) {
        // ---------------- MMA issuers (warps 8-11): warp i takes the steps i, i + NI, ... into accumulator i ----------------
        const int wi = warp - 8;
        if (wi < NI) {
            const uint32_t idesc = make_idesc(BN, false, false), idesc16 = make_idesc_bf16(BN);
            uint32_t leader;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
            const uint32_t tmem_d = tmem_base + (uint32_t)(wi * BN);
            for (int step = wi; step < nsteps; step += NI) {
                const int s = step % S, ph = (step / S) & 1;
                mbar_wait(afull_bar(s), ph);   // A tile of this step is in tensor memory
                mbar_wait(bfull_bar(s), ph);   // B tile has landed (TMA, async proxy: no proxy fence needed)
                CV_STAMP(lane == 0, step, 3);
                tc_fence_after();
                const uint32_t ta = tmem_a(s);
                const bool first = step == wi;
                if (leader) {
                    if (NTERMS == 3) mma_kblock3(tmem_d, ta, stage_b(s), BN, idesc, idesc16, first);
                    else {
                        const uint64_t db = make_desc(stage_b(s), PK_LBO, PK_SBO);
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks)
                            ts::mma_tf32_ts(tmem_d, ta + ks * 8, db + ((uint64_t)(ks * 2 * PK_LBO) >> 4), idesc, !(first && ks == 0));
                    }
                    mma_commit(empty_bar(s));
                    if (step + NI >= nsteps) mma_commit(accum_bar(wi % nb));
                }
                CV_STAMP(lane == 0, step, 4);
                __syncwarp();
            }
            tc_fence_before();
        }
    } else {
        // ---------------- weight TMA producer (warp 12, one lane): band-interleaved like the converters ----------------
        if (lane == 0) {
            pdl_wait();
            int s = 0, ph = 0;
            for (int step = 0; step < nsteps; ++step) {
                const int kb = step_kb(step), band = step - kb * nb;
                const float* src = packed + (size_t)(bd.c[band0 + band].pk_kb0 + kb) * (P::B_STAGE / 4);
                if (step >= S) mbar_wait(empty_bar(s), ph ^ 1);
                CV_STAMP(true, step, 5);
                mbar_expect_tx(bfull_bar(s), P::B_STAGE);
                tma_bulk_g2s(stage_b(s), src, P::B_STAGE, bfull_bar(s));
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

template <int BN, int NTERMS, class Epi>
static void launch_conv(dqn_learner* h, const float* in, const float* packed, const Epi& e, const ImgGeom& g, const Bands& bd, int images,
                        const float* in_b = nullptr, const float* packed_b = nullptr, const Epi* e_b = nullptr, int images_b = 0,
                        int bands_per_cta = 4) {
    using P = CvPlan<BN, NTERMS>;
    auto kern = cvconv_kernel<BN, NTERMS, Epi>;
    const size_t bytes = (size_t)P::S * P::B_STAGE + P::BAR_BYTES + g.img_bytes + 1024;
    DQN_REQUIRE(bytes <= 227 * 1024, "image-resident conv: shared memory budget");
    static_assert(P::S >= 3 && P::S % 2 == 0 && P::DCOLS + P::S * P::ACOLS <= 512, "image-resident conv: tensor memory budget");
    DQN_REQUIRE(bd.n <= P::MAXB && g.nkb >= 4, "image-resident conv: bands / k-blocks (every issuing warp needs a step)");
    DQN_REQUIRE(std::min(bd.n, bands_per_cta) != 3, "image-resident conv: 1, 2 or 4 bands per CTA");
    static size_t configured = 0;  // one per template instantiation
    if (bytes > configured) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    if (bands_per_cta > bd.n) bands_per_cta = bd.n;
    launch_kernel(h->pdl, kern, dim3(images + images_b, (bd.n + bands_per_cta - 1) / bands_per_cta), dim3(CV_THREADS), bytes, h->stream, in,
                  packed, e, images, in_b, packed_b, e_b ? *e_b : e, g, bd, bands_per_cta, (const float*)h->zeros16, h->tc_dbg);
    h->launches++;
}

#undef CV_STAMP

// ---- weight gradient, image-resident ---------------------------------------------------------------------------------
// dW[(kh,kw,ci)][co] = sum over pixels of X[pixel + tap][ci] * dY[pixel][co], one CTA per (image, slice of the M tiles):
// the CTA writes the image's partial sum part[image][(kh,kw,ci)][co] (+ the bias row sum_pixels dY), and
// wgrad_reduce_kernel adds the images in order.  Operands:
//   A (TMEM, rows m = (tap, ci), k = pixel): thread == row; for pixel p it reads X at pixoff[p] + tapoff(m) from the
//     resident padded image -- consecutive lanes are consecutive ci, so the 32 scalar reads of a k-block are
//     conflict-free -- splits hi/lo and stores both to tensor memory;
//   B (smem, rows n = co, k = pixel): dY of the image, transposed once into the dense K-major hi/lo tiles, resident.
// Accumulators are double-buffered (2 x 64 TMEM columns): a dedicated epilogue warpgroup drains M tile i while the
// converters and the MMA thread work on tile i+1.
// Several MMA-issuing warps (see cvconv_kernel): issuing warp i takes the k-blocks kb % NI == i of every M tile into its own
// accumulator (two issuing warps: the tensor pipe needs N / 2 cycles per MMA, one thread issues one every ~53).
constexpr int WG_THREADS = 544;  // warps 0-7 converters, 8-11 MMA issuers (8 allocates TMEM), 12-15 epilogue, 16 bias row

// U8IN (conv1): the layer input is the uint8 frame stack itself.  The image is staged as bytes (one pixel = 4 bytes, rows
// zero-padded, NOT phase-split: the 32 lanes of a warp are the (kw, ci) of one kernel row and read 32 consecutive bytes)
// and the converter applies cast / 255 (deep_q_model.py:226-227, exactly rounded) before the hi/lo split.
// bias row of one image's partial: db[co] = sum over pixels of dY (ascending pixel order), read from the resident fp32 B tiles
template <int BN>
__device__ __forceinline__ void wgrad_bias_row(float* __restrict__ part, uint32_t btiles, int b_kb, int image, int mreal, int npix, int lane, bool active) {
    if (!active) return;
    float* prow = part + ((size_t)image * (mreal + 4) + mreal) * BN;
    for (int co = lane; co < BN; co += 32) {
        float sacc = 0.f;
        for (int p = 0; p < npix; ++p) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(btiles + (uint32_t)((p >> 5) * b_kb) + pk_off32(co, p & 31)));
            sacc += v;
        }
        prow[co] = sacc;
    }
}

template <int BN, int NTERMS, bool U8IN>
__global__ void __launch_bounds__(WG_THREADS, 1) cvwgrad_kernel(const void* __restrict__ xin, const float* __restrict__ dy,
                                                                float* __restrict__ part, const ImgGeom g, int mreal,
                                                                int mtiles_per_cta, const float* __restrict__ zeros,
                                                                unsigned long long* __restrict__ dbg, int smem_bytes) {
    const bool cta0 = dbg && blockIdx.x == 0 && blockIdx.y == 0;  // in-kernel timeline (dqn_debug_timeline, modes 21-23)
#define WG_STAMP(cond, row, col) do { if (cta0 && (cond) && (row) < 128) dbg[(row) * 8 + (col)] = clock64(); } while (0)
    WG_STAMP(threadIdx.x == 0, 127, 0);
    // issuing warps (8 ..): warp i takes the steps with step % NI == i (step = tile * nkb + kb), so that it waits on the barriers of
    // the stages s % NI == i only and sees EVERY phase of them (S is a multiple of NI).  A waiter that skips phases of an
    // mbarrier can find it one phase behind and mistake the preceding phase of the same parity for the one it wants.
    constexpr int NI = BN == 32 ? 4 : 2;
    constexpr int DCOLS = 2 * NI * BN;               // double-buffered accumulators: buffer ab, issuer i at column (ab * NI + i) * BN
    constexpr int ACOLS = BK * (NTERMS == 3 ? 2 : 1), S0 = (512 - DCOLS) / ACOLS, S = (S0 > 6 ? 6 : S0) / NI * NI;
    constexpr int B_TILE = BN * 128, B_KB = B_TILE * 2;
    static_assert(S >= 3 && S % NI == 0, "tensor memory budget / stage ownership");
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const int npix = g.oh * g.ow, nkb = (npix + 31) / 32;
    const uint32_t btiles = sbase;                         // nkb x {hi, lo} tiles
    const uint32_t bars = btiles + nkb * B_KB;             // afull[S], empty[S], acc_full[2], acc_empty[2], tmem slot
    const uint32_t pixtab = bars + 256;                    // int pixoff[nkb * 32]
    const uint32_t img = pixtab + 4u * nkb * 32;
    const uint32_t stg = (sbase + (uint32_t)smem_bytes - (uint32_t)(BM * BN * 4) - 1024u) & ~127u;  // epilogue staging tile (end of the block)
    auto afull_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    auto accf_bar = [&](int b) { return bars + 8u * (2 * S + b); };
    auto acce_bar = [&](int b) { return bars + 8u * (2 * S + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * S + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int image = blockIdx.x;
    const int mtiles = (mreal + BM - 1) / BM;
    const int mt0 = blockIdx.y * mtiles_per_cta, mt1 = min(mtiles, mt0 + mtiles_per_cta);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(afull_bar(s), 4); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(accf_bar(b), NI); mbar_init(acce_bar(b), 4); }
        fence_barrier_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const int u8_row = ((g.ow - 1) * g.s + g.k) * g.cin;  // U8IN: bytes per padded image row
    for (int p = tid; p < nkb * 32; p += WG_THREADS) {
        const int oy = p / g.ow, ox = p - oy * g.ow;
        // padding pixels: any finite data (B is 0 there)
        const int off = p >= npix ? 0 : U8IN ? oy * g.s * u8_row + ox * g.s * g.cin : (oy * g.s * g.wcols + ox) * g.pitch;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(pixtab + 4u * p), "r"(off) : "memory");
    }
    WG_STAMP(threadIdx.x == 0, 127, 1);
    pdl_wait();
    WG_STAMP(threadIdx.x == 0, 127, 2);
    // ---- stage the image and dY (all threads) ----
    if constexpr (U8IN) {
        // 8-byte cp.async chunks: a padded row = [pl pixels of zeros][iw pixels][zeros]; pl * cin and iw * cin are multiples of 8
        const uint8_t* src_img = static_cast<const uint8_t*>(xin) + (size_t)image * g.ih * g.iw * g.cin;
        const int cpr = u8_row / 8, first = g.pl * g.cin / 8, ncopy = g.iw * g.cin / 8;
        for (int q = tid; q < g.hp * cpr; q += WG_THREADS) {
            const int iyp = q / cpr, c = q - iyp * cpr;
            const int iy = iyp - g.pt, cc = c - first;
            const bool ok = (unsigned)iy < (unsigned)g.ih && (unsigned)cc < (unsigned)ncopy;
            const void* src = ok ? (const void*)(src_img + (size_t)iy * g.iw * g.cin + cc * 8) : (const void*)zeros;
            const uint32_t n = ok ? 8u : 0u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(img + (uint32_t)(iyp * u8_row + c * 8)), "l"(src), "r"(n) : "memory");
        }
    } else {
        load_image_async(img, static_cast<const float*>(xin) + (size_t)image * g.ih * g.iw * g.cin, g, tid, WG_THREADS, zeros);
    }
    {
        // dY[p][co] -> B(co, p): a thread takes 4 consecutive pixels (k) of one channel -- four scalar loads, coalesced over the
        // warp's 32 consecutive channels -- and writes them as one 16-byte (fp32 tile) and two 8-byte (bf16 tiles) shared stores.
        // (One element per thread and pass, three scalar stores each, was ~60 instructions per element: 7 k cycles of issue for
        // conv1's 460 x 32 gradient, a third of that kernel's time -- tools/wgrad_timeline.py.)
        const float* dimg = dy + (size_t)image * npix * BN;
        constexpr int CG = BN / 32, NW = WG_THREADS / 32;
        for (int it = warp; it < nkb * 8 * CG; it += NW) {
            const int p4 = it / CG, co = (it - p4 * CG) * 32 + lane, p0 = p4 * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = p0 + j < npix ? __ldg(dimg + (size_t)(p0 + j) * BN + co) : 0.f;
            const int kb = p0 >> 5, kin = p0 & 31;
            const uint32_t blk = btiles + (uint32_t)(kb * B_KB);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + pk_off32(co, kin)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
            if (NTERMS == 3) {  // bf16 copies of the cross-term operands (layout of the packed weights)
                const uint32_t o = blk + pk_off16(co, kin);
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(o + 128u * BN), "r"(bf16_pair(v[0], v[1])), "r"(bf16_pair(v[2], v[3])) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(o + 192u * BN), "r"(bf16_pair(lo_part(v[0]), lo_part(v[1]))),
                             "r"(bf16_pair(lo_part(v[2]), lo_part(v[3]))) : "memory");
            }
        }
    }
    WG_STAMP(threadIdx.x == 0, 127, 3);
    asm volatile("cp.async.wait_all;" ::: "memory");
    fence_proxy_async();  // the B tiles were written through the generic proxy and are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    WG_STAMP(threadIdx.x == 0, 127, 4);
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(DCOLS + s * ACOLS); };

    if (warp < 8) {
        // ---------------- A converters: thread == row m = (tap, ci) of the current M tile ----------------
        const int grp = warp & 3, cg = warp >> 2;
        const uint32_t lane_base = (uint32_t)(grp * 32) << 16;
        int step = cg;
        for (int mt = mt0; mt < mt1; ++mt) {
            const int m = mt * BM + grp * 32 + lane;
            uint32_t base = img;  // rows past mreal: finite garbage, their outputs are never stored
            if (m < mreal) {
                const int tap = m / g.cin, ci = m - tap * g.cin;
                const int kh = tap / g.k, kw = tap - kh * g.k;
                base = img + (U8IN ? (uint32_t)(kh * u8_row + kw * g.cin + ci)
                                   : (uint32_t)((kh * g.wcols + (kw % g.s) * g.wps + kw / g.s) * g.pitch + ci * 4));
            }
            const int step_end = (mt - mt0 + 1) * nkb;
            for (; step < step_end; step += 2) {
                const int kb = step - (mt - mt0) * nkb;
                const int s = step % S, ph = (step / S) & 1;
                if (step >= S) mbar_wait(empty_bar(s), ph ^ 1);
                WG_STAMP(grp == 0 && lane == 0, step, 0);
                float hi[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t o0, o1, o2, o3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3)
                                 : "r"(pixtab + 4u * (kb * 32 + q * 4)));
                    if constexpr (U8IN) {
                        uint32_t b0, b1, b2, b3;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b0) : "r"(base + o0));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b1) : "r"(base + o1));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b2) : "r"(base + o2));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b3) : "r"(base + o3));
                        hi[4 * q] = u8_to_unit(b0); hi[4 * q + 1] = u8_to_unit(b1); hi[4 * q + 2] = u8_to_unit(b2); hi[4 * q + 3] = u8_to_unit(b3);
                    } else {
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[4 * q]) : "r"(base + o0));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[4 * q + 1]) : "r"(base + o1));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[4 * q + 2]) : "r"(base + o2));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[4 * q + 3]) : "r"(base + o3));
                    }
                }
                const uint32_t ta = tmem_a(s) + lane_base;
                if (NTERMS == 3) a_stage_store3(ta, hi); else ts::tmem_st32(ta, hi);
                ts::tmem_st_wait();
                tc_fence_before();
                WG_STAMP(grp == 0 && lane == 0, step, 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(afull_bar(s));
            }
        }
        pdl_launch_dependents();
    } else if (warp < 12) {
        // ---------------- MMA issuers (warps 8 .. 8 + NI - 1) ----------------
        const int wi = warp - 8;
        if (wi < NI) {
            const uint32_t idesc = make_idesc(BN, false, false), idesc16 = make_idesc_bf16(BN);
            uint32_t leader;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
            for (int mt = mt0; mt < mt1; ++mt) {
                const int ab = (mt - mt0) & 1, use = (mt - mt0) >> 1;
                if (use >= 1) { mbar_wait(acce_bar(ab), (use - 1) & 1); tc_fence_after(); }  // epilogue has drained this buffer
                const uint32_t tmem_d = tmem_base + (uint32_t)((ab * NI + wi) * BN);
                const int t0 = (mt - mt0) * nkb, t1 = t0 + nkb;
                for (int step = t0 + (wi - t0 % NI + NI) % NI; step < t1; step += NI) {  // this warp's steps of the tile (nkb >= NI: at least one)
                    const int kb = step - t0, s = step % S, ph = (step / S) & 1;
                    const bool first = step - NI < t0;
                    mbar_wait(afull_bar(s), ph);
                    WG_STAMP(lane == 0, step, 3);
                    tc_fence_after();
                    const uint32_t ta = tmem_a(s);
                    if (leader) {
                        if (NTERMS == 3) mma_kblock3(tmem_d, ta, btiles + kb * B_KB, BN, idesc, idesc16, first);
                        else {
                            const uint64_t db = make_desc(btiles + kb * B_KB, PK_LBO, PK_SBO);
#pragma unroll
                            for (int ks = 0; ks < BK / 8; ++ks)
                                ts::mma_tf32_ts(tmem_d, ta + ks * 8, db + ((uint64_t)(ks * 2 * PK_LBO) >> 4), idesc, !(first && ks == 0));
                        }
                        mma_commit(empty_bar(s));
                        if (step + NI >= t1) mma_commit(accf_bar(ab));
                    }
                    WG_STAMP(lane == 0, step, 4);
                    __syncwarp();
                }
            }
            tc_fence_before();
        }
    } else if (warp == 16) {
        wgrad_bias_row<BN>(part, btiles, B_KB, image, mreal, npix, lane, blockIdx.y == 0);
    } else {
        // ---------------- epilogue warpgroup (warps 12-15): accumulator of M tile i -> part[image][m][co] ----------------
        const int grp = warp & 3;
        const uint32_t lane_base = (uint32_t)(grp * 32) << 16;
        // The tile goes to global memory through a staging copy in shared memory and leaves it as whole 128-byte lines (a warp
        // instruction = 512 contiguous bytes of part[image][m][:]).  Stored row by row from registers -- 32 lanes x 16 bytes at a
        // 256-byte pitch, 32 lines per instruction -- the epilogue held the LSU for ~2000 cycles per tile and stalled the converters'
        // shared-memory loads of the next tile for as long (tools/wgrad_timeline.py).  A warp stages and reads back its own 32 rows.
        const int r = grp * 32 + lane;
        constexpr int NCH = BN / 4, RPI = 32 / NCH;  // 16-byte chunks per row; rows per read-back instruction
        for (int mt = mt0; mt < mt1; ++mt) {
            const int ab = (mt - mt0) & 1, use = (mt - mt0) >> 1;
            mbar_wait(accf_bar(ab), use & 1);
            tc_fence_after();
            WG_STAMP(warp == 12 && lane == 0, 100 + (mt - mt0), 0);
#pragma unroll
            for (int c = 0; c < BN; c += 16) {
                float v[16];  // the issuing warps' accumulators of this buffer (index order), one wait for all of them
                const uint32_t t0 = tmem_base + lane_base + (uint32_t)(ab * NI * BN + c);
                if (NI == 4) tmem_ld16_sum4(t0, t0 + BN, t0 + 2 * BN, t0 + 3 * BN, v);
                else tmem_ld16_sum2(t0, t0 + BN, v);
                if (c + 16 >= BN) {  // all columns of this buffer are in registers: hand it back to the MMA warps
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acce_bar(ab));
                }
                // the four 16-byte chunks of this column group, each lane starting at a different one: rows are 128 / 256 bytes
                // apart (the same banks), so the same chunk from 8 lanes would be an 8-way conflict; rotated it is 2-way
#pragma unroll
                for (int jq = 0; jq < 4; ++jq) {
                    const int q = (jq + r) & 3;
                    const float x0 = q == 0 ? v[0] : q == 1 ? v[4] : q == 2 ? v[8] : v[12];
                    const float x1 = q == 0 ? v[1] : q == 1 ? v[5] : q == 2 ? v[9] : v[13];
                    const float x2 = q == 0 ? v[2] : q == 1 ? v[6] : q == 2 ? v[10] : v[14];
                    const float x3 = q == 0 ? v[3] : q == 1 ? v[7] : q == 2 ? v[11] : v[15];
                    const uint32_t a = stg + (uint32_t)(r * BN * 4 + (c + 4 * q) * 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
                }
            }
            __syncwarp();
            float* o = part + ((size_t)image * (mreal + 4) + (size_t)mt * BM + grp * 32) * BN;
            const int rows_valid = mreal - mt * BM - grp * 32;  // rows of this warp's 32 that exist (all of them but in a ragged last tile)
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int row = i * RPI + lane / NCH, ch = lane % NCH;
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                             : "r"(stg + (uint32_t)((grp * 32 + row) * BN * 4 + ch * 16)));
                if (row < rows_valid) *reinterpret_cast<float4*>(o + (size_t)row * BN + ch * 4) = x;
            }
            __syncwarp();  // the staging rows are free again
        }
        tc_fence_before();
        WG_STAMP(warp == 12 && lane == 0, 127, 5);
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    WG_STAMP(threadIdx.x == 0, 127, 6);
#undef WG_STAMP
}

// true if launched: part[image][mreal + 4][N] holds the per-image partial sums (bias row at m = mreal).  u8_states: the
// layer input is the uint8 batch (conv1)
template <int BN, int NTERMS>
static bool launch_wgrad(dqn_learner* h, const void* x, bool u8_states, const float* dy, float* part, const ImgGeom& g, int images) {
    const int mreal = g.k * g.k * g.cin;
    const int npix = g.oh * g.ow, nkb = (npix + 31) / 32;
    if (g.cout != BN) return false;
    size_t img_bytes;
    if (u8_states) {
        if (g.cin != 4 || g.k != 8 || (g.pl * g.cin) % 8 != 0 || (g.iw * g.cin) % 8 != 0 || mreal % BM != 0) return false;
        const int u8_row = ((g.ow - 1) * g.s + g.k) * g.cin;
        if (u8_row % 8 != 0) return false;
        img_bytes = (size_t)g.hp * u8_row + 64;
    } else {
        if (g.cin % 32 != 0 || npix > 128) return false;
        img_bytes = g.img_bytes;
    }
    const size_t bytes = (size_t)nkb * BN * 256 + 256 + 4 * nkb * 32 + img_bytes + 1024 + 128 + (size_t)BM * BN * 4 + 1024 + 128;  // .. + staging tile
    if (bytes > 227 * 1024 || nkb < (BN == 32 ? 4 : 2)) return false;  // (every issuing warp needs a k-block in every M tile)
    const int mtiles = (mreal + BM - 1) / BM;
    const int per_cta = mtiles >= 8 ? mtiles / 2 : mtiles;  // two CTAs per image once an image carries >= 8 M tiles
    auto go = [&](auto kern) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        launch_kernel(h->pdl, kern, dim3(images, (mtiles + per_cta - 1) / per_cta), dim3(WG_THREADS), bytes, h->stream, x, dy, part, g, mreal,
                      per_cta, (const float*)h->zeros16, h->tc_dbg, (int)bytes);
    };
    if (u8_states) go(cvwgrad_kernel<BN, NTERMS, true>); else go(cvwgrad_kernel<BN, NTERMS, false>);
    h->launches++;
    return true;
}

}  // namespace cv
