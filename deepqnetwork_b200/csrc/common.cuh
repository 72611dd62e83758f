// Shared device/host helpers for libdqn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#define DQN_CUDA_OK(expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            char _b[512];                                                                     \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),      \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                             \
            throw std::string(_b);                                                            \
        }                                                                                     \
    } while (0)

#define DQN_REQUIRE(cond, msg)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            char _b[512];                                                                     \
            snprintf(_b, sizeof(_b), "%s (%s:%d)", msg, __FILE__, __LINE__);                  \
            throw std::string(_b);                                                            \
        }                                                                                     \
    } while (0)

// Programmatic dependent launch: a kernel launched with the attribute may start (prologue: barrier init, TMEM
// allocation) while its stream predecessor is still running; it must execute pdl_wait() before touching anything the
// predecessor writes.  Kernels call pdl_launch_dependents() as early as possible.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch priority of the kernels being issued (CUDA stream-priority scale: lower value = dispatched first).  The
// compute work distributor hands out CTAs grid by grid within one priority level, so a multi-wave side kernel (dense
// wgrad: 480 CTAs, Adam: 1184 CTAs) would hold back every later kernel of the critical chain until its last wave is
// dispatched (measured with tools/cupti_trace: 10-16 us stalls in front of 4 us kernels).  The step orchestration sets
// this around the side branches; inside a captured graph the attribute is recorded on the kernel node.
inline int g_launch_priority = 0;

template <class... KArgs, class... Args>
static inline void launch_kernel(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    at[n].id = cudaLaunchAttributePriority; at[n].val.priority = g_launch_priority; ++n;
    if (pdl) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    cfg.attrs = at; cfg.numAttrs = n;
    DQN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// x/255 correctly rounded for integer x in [0,255] (tf.divide(cast(x), 255), rl/deep_q_model.py:226-227):
// one Newton correction of x*(1/255) -- the same quotient IEEE division produces, without the slow path.
__device__ __forceinline__ float u8_to_unit(uint32_t x) {
    const float r = 1.0f / 255.0f;
    float xf = (float)x;
    float q = xf * r;
    return fmaf(fmaf(-q, 255.0f, xf), r, q);
}

// Philox-4x32-10 counter RNG (device-side synthetic data and throughput-mode uniforms).
__host__ __device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) built like CPython's random.random(): (a>>5, b>>6) -> (a*2^26+b)/2^53
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
