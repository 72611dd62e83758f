// HBM-resident uint8 replay ring: row gather, row scatter (append) and device-side synthetic fill.
//
// Replaces (reference file:line):
//   rl/data/replay_memory.py:26-43    SoA layout (states, actions, rewards, next_states, continues, losses f16)
//   rl/data/replay_memory.py:113-122  copy(idx -> target[slot])  == gather_rows_kernel, one slot per blockIdx.y
//   rl/data/replay_memory.py:54-67,94-110  append/set            == ring copies in api.cu (dqn_replay_append)
// Pure byte movement: 16-byte vector loads/stores, rows are 16-byte aligned (H*W*C % 16 == 0 is required).
#include "learner.h"

// grid: (chunks, 2*n). slot j<n copies states[row_j] -> dst[j]; slot n+j copies next_states[row_j] -> dst[n+j]
__global__ void gather_rows_kernel(const uint8_t* __restrict__ states, const uint8_t* __restrict__ next_states,
                                   const uint8_t* __restrict__ actions, const uint8_t* __restrict__ rewards,
                                   const uint8_t* __restrict__ cont, const __half* __restrict__ losses,
                                   const long long* __restrict__ rows, int n, size_t state_bytes,
                                   uint8_t* __restrict__ b_states, uint8_t* __restrict__ b_actions,
                                   uint8_t* __restrict__ b_rewards, uint8_t* __restrict__ b_cont,
                                   __half* __restrict__ b_losses, float4* __restrict__ x_f32) {
    const int slot = blockIdx.y;
    const int j = slot < n ? slot : slot - n;
    const long long row = rows[j];
    const uint4* src = reinterpret_cast<const uint4*>((slot < n ? states : next_states) + (size_t)row * state_bytes);
    uint4* dst = reinterpret_cast<uint4*>(b_states + (size_t)slot * state_bytes);
    const int nvec = (int)(state_bytes >> 4);
    // 4 independent 16-byte loads in flight per thread
    const int per_block = blockDim.x * 4;
    const int v0 = blockIdx.x * per_block + threadIdx.x;
    uint4 r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = v0 + k * blockDim.x;
        if (v < nvec) r[k] = __ldg(src + v);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = v0 + k * blockDim.x;
        if (v < nvec) dst[v] = r[k];
    }
    if (x_f32) {
        // tensor-core path: the same bytes as f32 / 255 (tf.cast + tf.divide, rl/deep_q_model.py:226-227), written here so
        // that the step needs no separate conversion pass over the batch
        float4* xd = x_f32 + ((size_t)slot * state_bytes >> 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int v = v0 + k * blockDim.x;
            if (v < nvec) {
                const uint32_t w[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    xd[(size_t)v * 4 + q] = make_float4(u8_to_unit(w[q] & 255u), u8_to_unit((w[q] >> 8) & 255u),
                                                        u8_to_unit((w[q] >> 16) & 255u), u8_to_unit(w[q] >> 24));
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && slot < n) {
        b_actions[j] = actions[row];
        b_rewards[j] = rewards[row];
        b_cont[j] = cont[row];
        b_losses[j] = losses[row];
    }
}

void launch_gather(dqn_learner* h, const long long* d_rows, int n, bool with_f32) {
    const int threads = 256;
    const int nvec = (int)(h->state_bytes >> 4);
    dim3 grid(ceil_div(nvec, threads * 4), 2 * n);
    // the batch buffer keeps s in rows [0,n) and s' in rows [n,2n): one contiguous [2n,S] input for the online tower
    launch_kernel(false, gather_rows_kernel, grid, dim3(threads), 0, h->stream, h->r_states, h->r_next, h->r_actions, h->r_rewards,
                  h->r_cont, h->r_losses, d_rows, n, h->state_bytes, h->b_states, h->b_actions, h->b_rewards, h->b_cont, h->b_losses,
                  with_f32 && h->cfg.math_mode != DQN_MATH_FP32 ? reinterpret_cast<float4*>(h->x_f32) : (float4*)nullptr);
    h->launches++;
    h->xf_from_gather = with_f32 && h->cfg.math_mode != DQN_MATH_FP32;
}

// ---- synthetic fill: states/next_states ~ U{0..255}; actions ~ U{0..A-1};
// rewards in {0,1,4,7} w.p. {.90,.06,.03,.01}; continues ~ Bernoulli(.99);
// priority p = min(|N(0,0.3)| + .01, 1)^.6 as f32 (f16 copy in `losses`, f32 handed to the tree).
__global__ void fill_states_kernel(uint4* __restrict__ dst, size_t nvec, size_t vec0, uint64_t seed, uint32_t stream_id) {
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        const size_t c = vec0 + v;  // counter = global 16-byte index, independent of how the fill is chunked
        philox4x32((uint32_t)c, (uint32_t)(c >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        dst[v] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

__global__ void fill_scalars_kernel(uint8_t* actions, uint8_t* rewards, uint8_t* cont, __half* losses,
                                    float* prio_out, long long start, long long n, int num_actions, uint64_t seed) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const long long row = start + j;
    uint32_t r[4], g[4];
    philox4x32((uint32_t)row, (uint32_t)(row >> 32), 7u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    philox4x32((uint32_t)row, (uint32_t)(row >> 32), 8u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), g);
    actions[row] = (uint8_t)(r[0] % (uint32_t)num_actions);
    const float ur = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
    rewards[row] = ur < 0.90f ? 0 : (ur < 0.96f ? 1 : (ur < 0.99f ? 4 : 7));
    cont[row] = ((float)(r[2] >> 8) * (1.0f / 16777216.0f)) < 0.99f ? 1 : 0;
    // Box-Muller
    const float u1 = ((float)(g[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = (float)(g[1] >> 8) * (1.0f / 16777216.0f);
    const float z = sqrtf(-2.0f * logf(u1)) * cosf(6.2831853071795864f * u2);
    const float p = powf(fminf(fabsf(0.3f * z) + 0.01f, 1.0f), 0.6f);
    losses[row] = __float2half_rn(p);
    prio_out[j] = p;
}

void launch_fill_synthetic(dqn_learner* h, int64_t start, int64_t n, uint64_t seed, float* d_prio) {
    const size_t nvec = (size_t)n * (h->state_bytes >> 4);
    const int blocks = h->sm_count * 16;
    const size_t vec0 = (size_t)start * (h->state_bytes >> 4);
    fill_states_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(h->r_states + (size_t)start * h->state_bytes),
                                                      nvec, vec0, seed, 1u);
    fill_states_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(h->r_next + (size_t)start * h->state_bytes),
                                                      nvec, vec0, seed, 2u);
    fill_scalars_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, h->stream>>>(h->r_actions, h->r_rewards, h->r_cont,
                                                                             h->r_losses, d_prio, start, n,
                                                                             h->cfg.num_actions, seed);
    DQN_CUDA_OK(cudaGetLastError());
    h->launches += 3;
}
