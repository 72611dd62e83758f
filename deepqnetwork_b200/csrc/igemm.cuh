// fp32 implicit-GEMM on CUDA cores: C[M,N] = A[M,K] * B[K,N] with operand functors that gather
// conv windows / transposes on the fly (no im2col buffer is ever materialised) and epilogue functors
// that fuse bias, ReLU, ReLU-gating and split-K partial stores.
//
// This is the DQN_MATH_FP32 path (strict fp32, FFMA): register-blocked TMxTN micro tiles, double-buffered
// shared memory, 16-byte global loads along whichever dimension the operand is contiguous in.
#pragma once
#include "learner.h"
#include "cvpack.cuh"

template <int BM_, int BN_, int BK_, int TM_, int TN_>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
    static constexpr int THREADS = (BM / TM) * (BN / TN);
};

// ---------------- operand functors ----------------
// KCONTIG=true : load4(row, r, k, z) returns elements (r, k..k+3)
// KCONTIG=false: load4(row, r, k, z) returns elements (r..r+3, k)
// row(r, z) precomputes whatever depends only on r.

struct OpRowMajorK {  // Op(r,k) = p[r*ld + k]
    static constexpr bool KCONTIG = true;
    const float* p; int ld;
    typedef const float* Row;
    __device__ __forceinline__ Row row(int r, int) const { return p + (size_t)r * ld; }
    __device__ __forceinline__ float4 load4(Row rw, int, int k, int) const { return *reinterpret_cast<const float4*>(rw + k); }
    __device__ __forceinline__ const float* addr4(Row rw, int, int k, int) const { return rw + k; }
    // cursor form (tsgemm.cuh): k position shared by all rows of a thread, advanced by 32 per k-block
    typedef int Cur;
    __device__ __forceinline__ Cur cur(int k, int) const { return k; }
    __device__ __forceinline__ void next(Cur& c) const { c += 32; }
    __device__ __forceinline__ const float* ptr(Row rw, int, const Cur& c) const { return rw + c; }
};

struct OpColContig {  // Op(r,k) = p[k*ld + r]
    static constexpr bool KCONTIG = false;
    const float* p; int ld;
    typedef const float* Row;
    __device__ __forceinline__ Row row(int r, int) const { return p + r; }
    __device__ __forceinline__ float4 load4(Row rw, int, int k, int) const {
        return *reinterpret_cast<const float4*>(rw + (size_t)k * ld);
    }
    __device__ __forceinline__ const float* addr4(Row rw, int, int k, int) const { return rw + (size_t)k * ld; }
    __device__ __forceinline__ const float* addr1(Row rw, int, int k, int) const { return rw + (size_t)k * ld; }
    typedef size_t Cur;  // element offset k*ld
    __device__ __forceinline__ Cur cur(int k, int) const { return (size_t)k * ld; }
    __device__ __forceinline__ void next(Cur& c) const { c += (size_t)32 * ld; }
    __device__ __forceinline__ const float* ptr(Row rw, int, const Cur& c) const { return rw + c; }
    __device__ __forceinline__ int kstride() const { return ld; }
};

// hidden-layer weights of the (up to) two streams seen as one [flat, NH] matrix: B(k,n) = W[n/hid][k][n%hid]
struct OpFcWeightFwd {
    static constexpr bool KCONTIG = false;
    const float* w0; const float* w1; int hid;
    typedef const float* Row;
    __device__ __forceinline__ Row row(int n, int) const { return n < hid ? w0 + n : w1 + (n - hid); }
    __device__ __forceinline__ float4 load4(Row rw, int, int k, int) const {
        return *reinterpret_cast<const float4*>(rw + (size_t)k * hid);
    }
    __device__ __forceinline__ const float* addr4(Row rw, int, int k, int) const { return rw + (size_t)k * hid; }
    __device__ __forceinline__ const float* addr1(Row rw, int, int k, int) const { return rw + (size_t)k * hid; }
    typedef size_t Cur;
    __device__ __forceinline__ Cur cur(int k, int) const { return (size_t)k * hid; }
    __device__ __forceinline__ void next(Cur& c) const { c += (size_t)32 * hid; }
    __device__ __forceinline__ const float* ptr(Row rw, int, const Cur& c) const { return rw + c; }
    __device__ __forceinline__ int kstride() const { return hid; }
};

// transposed view for dgrad: B(k,n) = W[k/hid][n][k%hid]
struct OpFcWeightBwd {
    static constexpr bool KCONTIG = true;
    const float* w0; const float* w1; int hid;
    typedef size_t Row;
    __device__ __forceinline__ Row row(int n, int) const { return (size_t)n * hid; }
    __device__ __forceinline__ float4 load4(Row rw, int, int k, int) const {
        const float* w = k < hid ? w0 : w1;
        const int kk = k < hid ? k : k - hid;
        return *reinterpret_cast<const float4*>(w + rw + kk);
    }
    __device__ __forceinline__ const float* addr4(Row rw, int, int k, int) const {
        return (k < hid ? w0 : w1) + rw + (k < hid ? k : k - hid);
    }
    typedef int Cur;  // k (a 32-wide k-block never straddles the two streams: hid % 32 == 0)
    __device__ __forceinline__ Cur cur(int k, int) const { return k; }
    __device__ __forceinline__ void next(Cur& c) const { c += 32; }
    __device__ __forceinline__ const float* ptr(Row rw, int, const Cur& c) const {
        return (c < hid ? w0 + c : w1 + (c - hid)) + rw;
    }
};

__device__ __forceinline__ float4 load_px4(const void* in, int is_u8, size_t idx) {
    if (is_u8) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(in) + idx);
        return make_float4(u8_to_unit(w & 255u), u8_to_unit((w >> 8) & 255u), u8_to_unit((w >> 16) & 255u),
                           u8_to_unit(w >> 24));
    }
    return *reinterpret_cast<const float4*>(static_cast<const float*>(in) + idx);
}

// conv forward A: rows m=(n,oy,ox), k=(kh,kw,ci) -- the im2col view of an NHWC tensor (u8 or f32)
struct OpConvFwdA {
    static constexpr bool KCONTIG = true;
    static constexpr bool HEAVY = true;  // im2col address arithmetic per chunk (tsgemm.cuh: 8 loader warps)
    const void* in; ConvGeom g; int is_u8;
    struct Row { size_t base; int iy0, ix0; };
    __device__ __forceinline__ Row row(int m, int) const {
        const int hw = g.oh * g.ow;
        const int n = m / hw, rem = m - n * hw;
        const int oy = rem / g.ow, ox = rem - oy * g.ow;
        Row r; r.base = (size_t)n * g.ih * g.iw * g.cin; r.iy0 = oy * g.s - g.pt; r.ix0 = ox * g.s - g.pl;
        return r;
    }
    __device__ __forceinline__ float4 load4(const Row& r, int, int k, int) const {
        const int tap = k >> g.cin_log2, ci = k & (g.cin - 1);
        const int kh = fdiv(tap, g.m_k, g.k), kw = tap - kh * g.k;
        const int iy = r.iy0 + kh, ix = r.ix0 + kw;
        if ((unsigned)iy >= (unsigned)g.ih || (unsigned)ix >= (unsigned)g.iw) return make_float4(0.f, 0.f, 0.f, 0.f);
        return load_px4(in, is_u8, r.base + ((size_t)(iy * g.iw + ix) << g.cin_log2) + ci);
    }
    __device__ __forceinline__ const float* addr4(const Row& r, int, int k, int) const {  // f32 input only
        const int tap = k >> g.cin_log2, ci = k & (g.cin - 1);
        const int kh = fdiv(tap, g.m_k, g.k), kw = tap - kh * g.k;
        const int iy = r.iy0 + kh, ix = r.ix0 + kw;
        if ((unsigned)iy >= (unsigned)g.ih || (unsigned)ix >= (unsigned)g.iw) return nullptr;
        return static_cast<const float*>(in) + r.base + ((size_t)(iy * g.iw + ix) << g.cin_log2) + ci;
    }
    struct Cur { int kh, kw, ci; };
    __device__ __forceinline__ Cur cur(int k, int) const {
        const int tap = k >> g.cin_log2;
        Cur c; c.ci = k & (g.cin - 1); c.kh = fdiv(tap, g.m_k, g.k); c.kw = tap - c.kh * g.k;
        return c;
    }
    __device__ __forceinline__ void next(Cur& c) const {
        c.ci += 32;
        c.kw += c.ci >> g.cin_log2;
        c.ci &= g.cin - 1;
        while (c.kw >= g.k) { c.kw -= g.k; ++c.kh; }
    }
    __device__ __forceinline__ const float* ptr(const Row& r, int, const Cur& c) const {
        const int iy = r.iy0 + c.kh, ix = r.ix0 + c.kw;
        if ((unsigned)iy >= (unsigned)g.ih || (unsigned)ix >= (unsigned)g.iw) return nullptr;
        return static_cast<const float*>(in) + r.base + ((iy * g.iw + ix) << g.cin_log2) + c.ci;
    }
};

// conv wgrad A^T: rows m=(kh,kw,ci) (4 consecutive ci per load), k=(n,oy,ox)
struct OpConvWgradA {
    static constexpr bool KCONTIG = false;
    const void* in; ConvGeom g; int is_u8;
    struct Row { int kh, kw, ci; };
    __device__ __forceinline__ Row row(int m, int) const {
        const int tap = m >> g.cin_log2;
        Row r; r.ci = m & (g.cin - 1); r.kh = tap / g.k; r.kw = tap - r.kh * g.k;
        return r;
    }
    __device__ __forceinline__ float4 load4(const Row& r, int, int k, int) const {
        const int hw = g.oh * g.ow;
        const int n = fdiv(k, g.m_hw, hw), rem = k - n * hw;
        const int oy = fdiv(rem, g.m_ow, g.ow), ox = rem - oy * g.ow;
        const int iy = oy * g.s - g.pt + r.kh, ix = ox * g.s - g.pl + r.kw;
        if ((unsigned)iy >= (unsigned)g.ih || (unsigned)ix >= (unsigned)g.iw) return make_float4(0.f, 0.f, 0.f, 0.f);
        return load_px4(in, is_u8, (((size_t)n * g.ih + iy) * g.iw + ix) * g.cin + r.ci);
    }
    __device__ __forceinline__ const float* addr4(const Row& r, int, int k, int) const {  // f32 input only
        const int hw = g.oh * g.ow;
        const int n = fdiv(k, g.m_hw, hw), rem = k - n * hw;
        const int oy = fdiv(rem, g.m_ow, g.ow), ox = rem - oy * g.ow;
        const int iy = oy * g.s - g.pt + r.kh, ix = ox * g.s - g.pl + r.kw;
        if ((unsigned)iy >= (unsigned)g.ih || (unsigned)ix >= (unsigned)g.iw) return nullptr;
        return static_cast<const float*>(in) + (((size_t)n * g.ih + iy) * g.iw + ix) * g.cin + r.ci;
    }
};

// conv dgrad, decomposed by stride-parity class z = ry*s + rx so that only taps that actually reach an
// input pixel are multiplied: rows m=(n,a,b) with iy = y0(ry)+s*a, ix = x0(rx)+s*b; k=(jh,jw,co), kh = ry+s*jh.
struct DgradClass {
    ConvGeom g;
    __device__ __forceinline__ void cls(int z, int& ry, int& rx, int& y0, int& x0, int& cy, int& cx) const {
        ry = z / g.s; rx = z - ry * g.s;
        // smallest iy >= 0 with (iy + pt) % s == ry
        y0 = ((ry - g.pt) % g.s + g.s) % g.s; x0 = ((rx - g.pl) % g.s + g.s) % g.s;
        cy = y0 < g.ih ? (g.ih - y0 + g.s - 1) / g.s : 0;
        cx = x0 < g.iw ? (g.iw - x0 + g.s - 1) / g.s : 0;
    }
};

struct OpConvDgradA {
    static constexpr bool KCONTIG = true;
    static constexpr bool HEAVY = true;
    const float* dy; DgradClass c; int cout_log2; int nimg;
    struct Row { size_t base; int oy0, ox0; int ok; };
    __device__ __forceinline__ Row row(int m, int z) const {
        int ry, rx, y0, x0, cy, cx; c.cls(z, ry, rx, y0, x0, cy, cx);
        Row r; r.ok = 0; r.base = 0; r.oy0 = 0; r.ox0 = 0;
        const int per = cy * cx;
        if (per == 0) return r;
        const int n = m / per, rem = m - n * per;
        if (n >= nimg) return r;
        const int a = rem / cx, b = rem - a * cx;
        const int iy = y0 + c.g.s * a, ix = x0 + c.g.s * b;
        r.ok = 1; r.base = (size_t)n * c.g.oh * c.g.ow;
        r.oy0 = (iy + c.g.pt - ry) / c.g.s; r.ox0 = (ix + c.g.pl - rx) / c.g.s;
        return r;
    }
    __device__ __forceinline__ float4 load4(const Row& r, int, int k, int) const {
        const int kt = c.g.kt;  // taps per axis in one class
        const int tap = k >> cout_log2, co = k & (c.g.cout - 1);
        const int jh = fdiv(tap, c.g.m_kt, kt), jw = tap - jh * kt;
        const int oy = r.oy0 - jh, ox = r.ox0 - jw;
        if (!r.ok || (unsigned)oy >= (unsigned)c.g.oh || (unsigned)ox >= (unsigned)c.g.ow)
            return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(dy + ((r.base + (size_t)oy * c.g.ow + ox) << cout_log2) + co);
    }
    __device__ __forceinline__ const float* addr4(const Row& r, int, int k, int) const {
        const int kt = c.g.kt;
        const int tap = k >> cout_log2, co = k & (c.g.cout - 1);
        const int jh = fdiv(tap, c.g.m_kt, kt), jw = tap - jh * kt;
        const int oy = r.oy0 - jh, ox = r.ox0 - jw;
        if (!r.ok || (unsigned)oy >= (unsigned)c.g.oh || (unsigned)ox >= (unsigned)c.g.ow) return nullptr;
        return dy + ((r.base + (size_t)oy * c.g.ow + ox) << cout_log2) + co;
    }
    struct Cur { int jh, jw, co; };
    __device__ __forceinline__ Cur cur(int k, int) const {
        const int tap = k >> cout_log2;
        Cur q; q.co = k & (c.g.cout - 1); q.jh = fdiv(tap, c.g.m_kt, c.g.kt); q.jw = tap - q.jh * c.g.kt;
        return q;
    }
    __device__ __forceinline__ void next(Cur& q) const {
        q.co += 32;
        q.jw += q.co >> cout_log2;
        q.co &= c.g.cout - 1;
        while (q.jw >= c.g.kt) { q.jw -= c.g.kt; ++q.jh; }
    }
    __device__ __forceinline__ const float* ptr(const Row& r, int, const Cur& q) const {
        const int oy = r.oy0 - q.jh, ox = r.ox0 - q.jw;
        if (!r.ok || (unsigned)oy >= (unsigned)c.g.oh || (unsigned)ox >= (unsigned)c.g.ow) return nullptr;
        return dy + ((r.base + (size_t)(oy * c.g.ow + ox)) << cout_log2) + q.co;
    }
};

struct OpConvDgradB {  // B(k=(jh,jw,co), n=ci) = W[kh][kw][ci][co]
    static constexpr bool KCONTIG = true;
    const float* w; ConvGeom g; int cout_log2;
    typedef int Row;
    __device__ __forceinline__ Row row(int n, int) const { return n; }
    __device__ __forceinline__ float4 load4(Row ci, int, int k, int z) const {
        const int ry = fdiv(z, g.m_s, g.s), rx = z - ry * g.s;
        const int kt = g.kt;
        const int tap = k >> cout_log2, co = k & (g.cout - 1);
        const int jh = fdiv(tap, g.m_kt, kt), jw = tap - jh * kt;
        const int kh = ry + g.s * jh, kw = rx + g.s * jw;
        return *reinterpret_cast<const float4*>(w + ((((size_t)kh * g.k + kw) * g.cin + ci) << cout_log2) + co);
    }
    __device__ __forceinline__ const float* addr4(Row ci, int, int k, int z) const {
        const int ry = fdiv(z, g.m_s, g.s), rx = z - ry * g.s;
        const int kt = g.kt;
        const int tap = k >> cout_log2, co = k & (g.cout - 1);
        const int jh = fdiv(tap, g.m_kt, kt), jw = tap - jh * kt;
        const int kh = ry + g.s * jh, kw = rx + g.s * jw;
        return w + ((((size_t)kh * g.k + kw) * g.cin + ci) << cout_log2) + co;
    }
    struct Cur { int kh, kw0, kw, co; };  // kernel tap reached by (jh, jw) of this parity class, and the channel
    __device__ __forceinline__ Cur cur(int k, int z) const {
        const int ry = fdiv(z, g.m_s, g.s), rx = z - ry * g.s;
        const int tap = k >> cout_log2;
        const int jh = fdiv(tap, g.m_kt, g.kt), jw = tap - jh * g.kt;
        Cur q; q.co = k & (g.cout - 1); q.kh = ry + g.s * jh; q.kw0 = rx; q.kw = rx + g.s * jw;
        return q;
    }
    __device__ __forceinline__ void next(Cur& q) const {
        q.co += 32;
        q.kw += (q.co >> cout_log2) * g.s;
        q.co &= g.cout - 1;
        while (q.kw >= q.kw0 + g.kt * g.s) { q.kw -= g.kt * g.s; q.kh += g.s; }
    }
    __device__ __forceinline__ const float* ptr(Row ci, int, const Cur& q) const {
        return w + (((q.kh * g.k + q.kw) * g.cin + ci) << cout_log2) + q.co;
    }
};

// ---------------- epilogues ----------------
struct EpiBiasRelu {  // out[m][n] = relu(acc + bias[n])
    float* out; const float* bias; int ldo;
    // image-resident conv kernel (cvgemm.cuh): bias staged in shared memory, sbias = shared address of bias[n]
    __device__ __forceinline__ void stage(int tid, int bn, uint32_t sarea) const {
        if (tid < bn) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sarea + 4u * tid), "f"(bias[tid]) : "memory");
    }
    template <int TN>
    __device__ __forceinline__ void store_b(int m, int n, const float (&v)[TN], uint32_t sbias) const {
        float* o = out + (size_t)m * ldo + n;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            float4 b;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(sbias + 4u * j));
            *reinterpret_cast<float4*>(o + j) = make_float4(fmaxf(v[j] + b.x, 0.f), fmaxf(v[j + 1] + b.y, 0.f),
                                                            fmaxf(v[j + 2] + b.z, 0.f), fmaxf(v[j + 3] + b.w, 0.f));
        }
    }
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int) const {
        float* o = out + (size_t)m * ldo + n;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + n + j);
            float4 r = make_float4(fmaxf(v[j] + b.x, 0.f), fmaxf(v[j + 1] + b.y, 0.f), fmaxf(v[j + 2] + b.z, 0.f),
                                   fmaxf(v[j + 3] + b.w, 0.f));
            *reinterpret_cast<float4*>(o + j) = r;
        }
    }
};

struct EpiPartial {  // part[z][m][n] = acc   (split-K partial sums, reduced by a consumer kernel)
    float* part; int M; int ldo;
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int z) const {
        float* o = part + ((size_t)z * M + m) * ldo + n;
#pragma unroll
        for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
};

// conv3 forward feeding the TMA dense forward (fcfwd.cuh): besides the NHWC activation it writes the SAME values as
// the dense layer's B operand tiles -- [k-block of 32 flat inputs][hi, lo][pk_rows batch rows x 32 k] in the dense
// canonical K-major UMMA layout -- so that the dense kernel fetches a k-block's tile with one bulk copy and spends no
// thread work on it.  flat index = pixel * cout + channel (tf.layers.flatten of NHWC, deep_q_model.py:309).
struct EpiBiasReluPack {
    float* out; const float* bias; int ldo;
    float* pk; int pk_rows; int ohw;  // pk == nullptr: plain EpiBiasRelu
    __device__ __forceinline__ void stage(int tid, int bn, uint32_t sarea) const {
        if (tid < bn) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sarea + 4u * tid), "f"(bias[tid]) : "memory");
    }
    template <int TN>
    __device__ __forceinline__ void store_b(int m, int n, const float (&v)[TN], uint32_t sbias) const {
        float* o = out + (size_t)m * ldo + n;
        const int img = m / ohw, pix = m - img * ohw;
        const int k0 = pix * ldo + n;  // first flat index of this run of TN channels (TN = 16, n % 16 == 0)
        // the dense layer's B operand: k-block k0 / 32 of pk_rows batch rows in the packed layout of cvpack.cuh
        uint8_t* blk = pk ? reinterpret_cast<uint8_t*>(pk) + (size_t)(k0 >> 5) * pk_rows * 256 : nullptr;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            float4 b;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(sbias + 4u * j));
            const float4 r = make_float4(fmaxf(v[j] + b.x, 0.f), fmaxf(v[j + 1] + b.y, 0.f), fmaxf(v[j + 2] + b.z, 0.f), fmaxf(v[j + 3] + b.w, 0.f));
            *reinterpret_cast<float4*>(o + j) = r;
            if (blk) cv::pk_store4(blk, pk_rows, img, (k0 & 31) + j, r);
        }
    }
};

struct EpiGatePix {  // image-resident dgrad: out[pixel][n] = acc * (gate[pixel][n] > 0)
    float* out; const float* gate; int ldo;
    __device__ __forceinline__ void stage(int, int, uint32_t) const {}
    template <int TN>
    __device__ __forceinline__ void store_b(int m, int n, const float (&v)[TN], uint32_t) const {
        const size_t o = (size_t)m * ldo + n;
        float4 gt[TN / 4];
#pragma unroll
        for (int j = 0; j < TN; j += 4) gt[j / 4] = *reinterpret_cast<const float4*>(gate + o + j);
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const float4 g4 = gt[j / 4];
            *reinterpret_cast<float4*>(out + o + j) = make_float4(g4.x > 0.f ? v[j] : 0.f, g4.y > 0.f ? v[j + 1] : 0.f,
                                                                 g4.z > 0.f ? v[j + 2] : 0.f, g4.w > 0.f ? v[j + 3] : 0.f);
        }
    }
};

struct EpiGateStore {  // out[m][n] = acc * (gate[m][n] > 0)   (ReLU backward fused into dgrad)
    float* out; const float* gate; int ldo;
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int) const {
        const size_t o = (size_t)m * ldo + n;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const float4 g = *reinterpret_cast<const float4*>(gate + o + j);
            *reinterpret_cast<float4*>(out + o + j) = make_float4(g.x > 0.f ? v[j] : 0.f, g.y > 0.f ? v[j + 1] : 0.f,
                                                                 g.z > 0.f ? v[j + 2] : 0.f, g.w > 0.f ? v[j + 3] : 0.f);
        }
    }
};

struct EpiDgradGate {  // conv dgrad: row m of class z -> pixel (n,iy,ix); out = acc * (gate > 0)
    float* out; const float* gate; DgradClass c; int nimg;
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int z) const {
        int ry, rx, y0, x0, cy, cx; c.cls(z, ry, rx, y0, x0, cy, cx);
        const int per = cy * cx;
        if (per == 0) return;
        const int img = m / per, rem = m - img * per;
        if (img >= nimg) return;
        const int a = rem / cx, b = rem - a * cx;
        const int iy = y0 + c.g.s * a, ix = x0 + c.g.s * b;
        const size_t o = (((size_t)img * c.g.ih + iy) * c.g.iw + ix) * c.g.cin + n;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const float4 g = *reinterpret_cast<const float4*>(gate + o + j);
            *reinterpret_cast<float4*>(out + o + j) = make_float4(g.x > 0.f ? v[j] : 0.f, g.y > 0.f ? v[j + 1] : 0.f,
                                                                 g.z > 0.f ? v[j + 2] : 0.f, g.w > 0.f ? v[j + 3] : 0.f);
        }
    }
};

struct EpiDgradPartial {  // conv dgrad, class + split-K launch: part[split][pixel][ci] = acc (no gate; summed by the consumer)
    float* part; DgradClass c; int nimg; int nclass; size_t slab;
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int zz) const {
        const int z = zz % nclass, split = zz / nclass;
        int ry, rx, y0, x0, cy, cx; c.cls(z, ry, rx, y0, x0, cy, cx);
        const int per = cy * cx;
        if (per == 0) return;
        const int img = m / per, rem = m - img * per;
        if (img >= nimg) return;
        const int a = rem / cx, b = rem - a * cx;
        const int iy = y0 + c.g.s * a, ix = x0 + c.g.s * b;
        float* o = part + split * slab + (((size_t)img * c.g.ih + iy) * c.g.iw + ix) * c.g.cin + n;
#pragma unroll
        for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
};

struct EpiFcWgrad {  // dW of the two hidden streams: column n -> stream n/hid, grads laid out like the params
    static constexpr bool COALESCE = true;
    float* g0; float* g1; int hid;
    struct State {};
    __device__ __forceinline__ void load(int, int, State&) const {}
    __device__ __forceinline__ void apply(int m, int n, const float4& g, State&, float) const {
        *reinterpret_cast<float4*>((n < hid ? g0 + n : g1 + (n - hid)) + (size_t)m * hid) = g;
    }
    __device__ __forceinline__ float step_size() const { return 0.f; }
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int) const {
        float* o = (n < hid ? g0 + n : g1 + (n - hid)) + (size_t)m * hid;
#pragma unroll
        for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
};

// dense wgrad with the optimizer in the epilogue: dW never reaches HBM.  Adam (TF ApplyAdam) or Nesterov momentum
// (TF ApplyMomentum) on the element this thread just reduced; layout of w/m/v = the parameter slab.
struct EpiFcWgradAdam {
    static constexpr bool COALESCE = true;  // tensor-core epilogue: go through smem so lanes run along n
    float* w0; float* w1; float* m0; float* m1; float* v0; float* v1; int hid;
    const dqn_learner::Scalars* sc; float lr; float mom; int use_momentum;
    struct State { float4 w, m, v; };
    __device__ __forceinline__ size_t index(int m, int n) const { return (size_t)m * hid + (n < hid ? n : n - hid); }
    // two-phase form (used by the coalesced tensor-core epilogue to keep several rows of loads in flight)
    __device__ __forceinline__ void load(int m, int n, State& st) const {
        const size_t o = index(m, n);
        st.w = *reinterpret_cast<const float4*>((n < hid ? w0 : w1) + o);
        st.m = *reinterpret_cast<const float4*>((n < hid ? m0 : m1) + o);
        if (!use_momentum) st.v = *reinterpret_cast<const float4*>((n < hid ? v0 : v1) + o);
    }
    __device__ __forceinline__ void apply(int m, int n, const float4& g, State& st, float lr_t) const {
        const size_t o = index(m, n);
        if (use_momentum) {
#define MOM1(c) st.m.c = st.m.c * mom + g.c; st.w.c = st.w.c - (g.c * lr + st.m.c * mom * lr);
            MOM1(x) MOM1(y) MOM1(z) MOM1(w)
#undef MOM1
        } else {
            const float omb1 = 1.0f - 0.9f, omb2 = 1.0f - 0.999f, eps = 1e-8f;
#define ADAM1(c) st.m.c = st.m.c + (g.c - st.m.c) * omb1; st.v.c = st.v.c + (g.c * g.c - st.v.c) * omb2; \
                 st.w.c = st.w.c - (st.m.c * lr_t) / (sqrtf(st.v.c) + eps);
            ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
            *reinterpret_cast<float4*>((n < hid ? v0 : v1) + o) = st.v;
        }
        *reinterpret_cast<float4*>((n < hid ? m0 : m1) + o) = st.m;
        *reinterpret_cast<float4*>((n < hid ? w0 : w1) + o) = st.w;
    }
    __device__ __forceinline__ float step_size() const { return use_momentum ? lr : lr * sqrtf(1.0f - sc->b2p) / (1.0f - sc->b1p); }
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&g)[TN], int) const {
        const float lr_t = step_size();
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            State st;
            load(m, n + j, st);
            apply(m, n + j, make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]), st, lr_t);
        }
    }
};

// transposed variants for swap-AB launches (rows of the GEMM = output columns): thread holds TN consecutive n'
struct EpiPartialT {  // part[z][n][m]
    float* part; int rows; int ldo;
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int z) const {
#pragma unroll
        for (int j = 0; j < TN; ++j)
            if (n + j < rows) part[((size_t)z * rows + n + j) * ldo + m] = v[j];
    }
};

struct EpiGateStoreT {  // out[n][m] = acc * (gate[n][m] > 0)
    float* out; const float* gate; int rows; int ldo;
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int) const {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (n + j < rows) {
                const size_t o = (size_t)(n + j) * ldo + m;
                out[o] = gate[o] > 0.f ? v[j] : 0.f;
            }
        }
    }
};

struct EpiStore {  // out[m][n] = acc (self-test)
    float* out; int ldo;
    template <int TN>
    __device__ __forceinline__ void store(int m, int n, const float (&v)[TN], int) const {
#pragma unroll
        for (int j = 0; j < TN; ++j) out[(size_t)m * ldo + n + j] = v[j];
    }
};

// ---------------- the kernel ----------------
// grid = (ceil(M/BM), ceil(N/BN), Z).  SPLIT: z selects the K range [z*kchunk, min(K,(z+1)*kchunk));
// otherwise z is a group id handed to the functors (dgrad parity class) and the whole K is reduced.
template <class Cfg, bool SPLIT, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS) igemm_kernel(const AOp a, const BOp b, const Epi epi, int M, int N, int K,
                                                            int kchunk) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, TM = Cfg::TM, TN = Cfg::TN, T = Cfg::THREADS;
    constexpr int PAD = 4;
    constexpr int LA = (BM * BK / 4) / T, LB = (BN * BK / 4) / T;
    static_assert(LA * T * 4 == BM * BK && LB * T * 4 == BN * BK, "tile/thread mismatch");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile must be float4 multiples");
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = SPLIT ? z * kchunk : 0;
    const int kend = SPLIT ? min(K, kbeg + kchunk) : K;

    // static (row, k-offset) assignment of this thread's 16-byte loads
    int a_r[LA], a_k[LA], b_r[LB], b_k[LB];
    typename AOp::Row a_row[LA];
    typename BOp::Row b_row[LB];
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        const int f = tid + i * T;
        if (AOp::KCONTIG) { a_r[i] = f / (BK / 4); a_k[i] = (f % (BK / 4)) * 4; }
        else { a_k[i] = f / (BM / 4); a_r[i] = (f % (BM / 4)) * 4; }
        a_row[i] = a.row(min(m0 + a_r[i], M - 1), z);
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
        const int f = tid + i * T;
        if (BOp::KCONTIG) { b_r[i] = f / (BK / 4); b_k[i] = (f % (BK / 4)) * 4; }
        else { b_k[i] = f / (BN / 4); b_r[i] = (f % (BN / 4)) * 4; }
        b_row[i] = b.row(min(n0 + b_r[i], N - 1), z);
    }

    float4 a_reg[LA], b_reg[LB];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            const int r = m0 + a_r[i], k = k0 + a_k[i];
            a_reg[i] = (r < M && k < kend) ? a.load4(a_row[i], r, k, z) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int r = n0 + b_r[i], k = k0 + b_k[i];
            b_reg[i] = (r < N && k < kend) ? b.load4(b_row[i], r, k, z) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            if (AOp::KCONTIG) {
                As[buf][a_k[i] + 0][a_r[i]] = a_reg[i].x; As[buf][a_k[i] + 1][a_r[i]] = a_reg[i].y;
                As[buf][a_k[i] + 2][a_r[i]] = a_reg[i].z; As[buf][a_k[i] + 3][a_r[i]] = a_reg[i].w;
            } else {
                *reinterpret_cast<float4*>(&As[buf][a_k[i]][a_r[i]]) = a_reg[i];
            }
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            if (BOp::KCONTIG) {
                Bs[buf][b_k[i] + 0][b_r[i]] = b_reg[i].x; Bs[buf][b_k[i] + 1][b_r[i]] = b_reg[i].y;
                Bs[buf][b_k[i] + 2][b_r[i]] = b_reg[i].z; Bs[buf][b_k[i] + 3][b_r[i]] = b_reg[i].w;
            } else {
                *reinterpret_cast<float4*>(&Bs[buf][b_k[i]][b_r[i]]) = b_reg[i];
            }
        }
    };

    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (kbeg < kend) {
        gload(kbeg);
        sstore(0);
    }
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) gload(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) *reinterpret_cast<float4*>(&av[i]) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
#pragma unroll
            for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(&bv[j]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i, n = n0 + tx * TN;
        if (m < M && n < N) epi.store(m, n, acc[i], z);
    }
}

template <class Cfg, bool SPLIT, class AOp, class BOp, class Epi>
static void launch_igemm(dqn_learner* h, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, int Z, int kchunk) {
    dim3 grid(ceil_div(M, Cfg::BM), ceil_div(N, Cfg::BN), Z);
    launch_kernel(false, igemm_kernel<Cfg, SPLIT, AOp, BOp, Epi>, grid, dim3(Cfg::THREADS), 0, h->stream, a, b, e, M, N, K, kchunk);
    h->launches++;
}
