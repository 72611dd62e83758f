// HBM-resident uint8 replay ring: row gather, row scatter (append) and device-side synthetic fill.
//
// Replaces (reference file:line):
//   rl/data/replay_memory.py:26-43    SoA layout (states, actions, rewards, next_states, continues, losses f16)
//   rl/data/replay_memory.py:113-122  copy(idx -> target[slot])  == gather_rows_kernel, one slot per blockIdx.y
//   rl/data/replay_memory.py:54-67,94-110  append/set            == ring copies in api.cu (dqn_replay_append)
// Pure byte movement: 16-byte vector loads/stores, rows are 16-byte aligned (H*W*C % 16 == 0 is required).
#include "learner.h"
#include <algorithm>

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

// Frame-de-duplicated layout: state[h][w][c] = frame[slot[c]][h][w].  A thread assembles 4 pixels (16 output bytes): one
// 4-byte load from each of the 4 frames of the stack, a 4 x 4 byte transpose (prmt), one 16-byte store.  Reads are 4-byte
// coalesced per frame (a warp reads 128 contiguous bytes of each), writes 16-byte coalesced.
__global__ void gather_frames_kernel(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ fidx,
                                     const uint8_t* __restrict__ actions, const uint8_t* __restrict__ rewards,
                                     const uint8_t* __restrict__ cont, const __half* __restrict__ losses,
                                     const long long* __restrict__ rows, int n, size_t frame_bytes, int only_next,
                                     uint8_t* __restrict__ b_states, uint8_t* __restrict__ b_actions,
                                     uint8_t* __restrict__ b_rewards, uint8_t* __restrict__ b_cont,
                                     __half* __restrict__ b_losses, float4* __restrict__ x_f32) {
    // only_next < 0: slots [0,n) = s, [n,2n) = s' (the train batch); only_next = 0 / 1: n slots of s / s' (read-back)
    const int slot = blockIdx.y;
    const int j = only_next < 0 ? (slot < n ? slot : slot - n) : slot;
    const int half = only_next < 0 ? (slot < n ? 0 : 1) : only_next;
    const long long row = rows[j];
    const uint32_t* fi = fidx + (size_t)row * 8 + half * 4;
    const uint32_t* f0 = reinterpret_cast<const uint32_t*>(frames + (size_t)fi[0] * frame_bytes);
    const uint32_t* f1 = reinterpret_cast<const uint32_t*>(frames + (size_t)fi[1] * frame_bytes);
    const uint32_t* f2 = reinterpret_cast<const uint32_t*>(frames + (size_t)fi[2] * frame_bytes);
    const uint32_t* f3 = reinterpret_cast<const uint32_t*>(frames + (size_t)fi[3] * frame_bytes);
    uint4* dst = reinterpret_cast<uint4*>(b_states + (size_t)slot * frame_bytes * 4);
    const int nq = (int)(frame_bytes >> 2);  // groups of 4 pixels
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const uint32_t a = __ldg(f0 + q), b = __ldg(f1 + q), c = __ldg(f2 + q), d = __ldg(f3 + q);
        // pixel p of the group: bytes {a.p, b.p, c.p, d.p}
        const uint32_t ab_lo = __byte_perm(a, b, 0x5140), ab_hi = __byte_perm(a, b, 0x7362);  // a0 b0 a1 b1 | a2 b2 a3 b3
        const uint32_t cd_lo = __byte_perm(c, d, 0x5140), cd_hi = __byte_perm(c, d, 0x7362);
        uint4 o;
        o.x = __byte_perm(ab_lo, cd_lo, 0x5410); o.y = __byte_perm(ab_lo, cd_lo, 0x7632);
        o.z = __byte_perm(ab_hi, cd_hi, 0x5410); o.w = __byte_perm(ab_hi, cd_hi, 0x7632);
        dst[q] = o;
        if (x_f32) {
            float4* xd = x_f32 + ((size_t)slot * frame_bytes) + (size_t)q * 4;
            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                xd[k] = make_float4(u8_to_unit(w[k] & 255u), u8_to_unit((w[k] >> 8) & 255u), u8_to_unit((w[k] >> 16) & 255u), u8_to_unit(w[k] >> 24));
        }
    }
    if (only_next < 0 && blockIdx.x == 0 && threadIdx.x == 0 && slot < n) {
        b_actions[j] = actions[row];
        b_rewards[j] = rewards[row];
        b_cont[j] = cont[row];
        b_losses[j] = losses[row];
    }
}

void launch_gather_states(dqn_learner* h, const long long* d_rows, int n, bool next, uint8_t* dst) {
    if (h->dedup) {
        launch_kernel(false, gather_frames_kernel, dim3(ceil_div((int)(h->frame_bytes >> 2), 256), n), dim3(256), 0, h->stream,
                      (const uint8_t*)h->r_frames, (const uint32_t*)h->r_fidx, (const uint8_t*)nullptr, (const uint8_t*)nullptr,
                      (const uint8_t*)nullptr, (const __half*)nullptr, d_rows, n, h->frame_bytes, next ? 1 : 0, dst, (uint8_t*)nullptr,
                      (uint8_t*)nullptr, (uint8_t*)nullptr, (__half*)nullptr, (float4*)nullptr);
        h->launches++;
        return;
    }
    throw std::string("launch_gather_states is the frame_dedup read-back path");
}

void launch_gather(dqn_learner* h, const long long* d_rows, int n, bool with_f32) {
    if (h->dedup) {
        const bool f32 = with_f32 && h->cfg.math_mode != DQN_MATH_FP32 && !tower_fused_ready(h);
        launch_kernel(false, gather_frames_kernel, dim3(ceil_div((int)(h->frame_bytes >> 2), 256), 2 * n), dim3(256), 0, h->stream,
                      (const uint8_t*)h->r_frames, (const uint32_t*)h->r_fidx, (const uint8_t*)h->r_actions, (const uint8_t*)h->r_rewards,
                      (const uint8_t*)h->r_cont, (const __half*)h->r_losses, d_rows, n, h->frame_bytes, -1, h->b_states, h->b_actions,
                      h->b_rewards, h->b_cont, h->b_losses, f32 ? reinterpret_cast<float4*>(h->x_f32) : (float4*)nullptr);
        h->launches++;
        h->xf_from_gather = f32;
        return;
    }
    const int threads = 256;
    const int nvec = (int)(h->state_bytes >> 4);
    dim3 grid(ceil_div(nvec, threads * 4), 2 * n);
    // the batch buffer keeps s in rows [0,n) and s' in rows [n,2n): one contiguous [2n,S] input for the online tower
    // (the fused conv tower reads the uint8 rows themselves: no f32 copy, the gather moves 2 x 2 x B x S bytes and nothing else)
    const bool f32 = with_f32 && h->cfg.math_mode != DQN_MATH_FP32 && !tower_fused_ready(h);
    launch_kernel(false, gather_rows_kernel, grid, dim3(threads), 0, h->stream, h->r_states, h->r_next, h->r_actions, h->r_rewards,
                  h->r_cont, h->r_losses, d_rows, n, h->state_bytes, h->b_states, h->b_actions, h->b_rewards, h->b_cont, h->b_losses,
                  f32 ? reinterpret_cast<float4*>(h->x_f32) : (float4*)nullptr);
    h->launches++;
    h->xf_from_gather = f32;
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

// frame_dedup synthetic fill: transition t holds frames {t .. t+3} (s) and {t+1 .. t+4} (s'), slots modulo the frame ring
__global__ void fill_fidx_kernel(uint32_t* __restrict__ fidx, long long start, long long n, long long fcap) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n * 8) return;
    const long long t = start + (i >> 3);
    const int c = (int)(i & 7);
    fidx[(size_t)t * 8 + c] = (uint32_t)((t + (c & 3) + (c >> 2)) % fcap);
}

void launch_fill_synthetic(dqn_learner* h, int64_t start, int64_t n, uint64_t seed, float* d_prio) {
    if (h->dedup) {
        // frames [start, start + n) of the stream (+ the 4 trailing ones once, with the first chunk's successor frames)
        const int blocks = h->sm_count * 16;
        const int64_t f0 = start == 0 ? 0 : start + 4, f1 = start + n + 4;  // transition t needs frames up to t + 4
        DQN_REQUIRE(f1 <= h->fcap, "synthetic fill: frame ring too small (needs replay rows + 4)");
        const size_t nvec = (size_t)(f1 - f0) * (h->frame_bytes >> 4), vec0 = (size_t)f0 * (h->frame_bytes >> 4);
        fill_states_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(h->r_frames + (size_t)f0 * h->frame_bytes), nvec, vec0, seed, 1u);
        fill_fidx_kernel<<<(unsigned)ceil_div64(n * 8, 256), 256, 0, h->stream>>>(h->r_fidx, start, n, h->fcap);
        fill_scalars_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, h->stream>>>(h->r_actions, h->r_rewards, h->r_cont, h->r_losses, d_prio,
                                                                                 start, n, h->cfg.num_actions, seed);
        DQN_CUDA_OK(cudaGetLastError());
        if (f1 > h->f_written) h->f_written = f1;
        h->launches += 3;
        return;
    }
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


// ---- frame preprocessing (actor side; rl/game_environment.py:110-116,134-140,154-160) -----------------------------------
//   obs[top : raw_h - bottom : stride, ::stride] -> np.dot(rgb, [0.299, 0.587, 0.144]) -> astype(uint8)
// in float64 with the contraction the reference's np.dot performs on the [H, W, 3] view (pinned by executing it over all
// 2^24 colours, oracle/preprocess_ref.py): fl(r * .299), then two fused multiply-adds; the C cast truncates and wraps
// (white = 262.65 -> 6).  One thread per output pixel; a warp reads 32 pixels x 3 bytes at a 2-pixel stride.
__global__ void preprocess_kernel(const uint8_t* __restrict__ obs, int raw_h, int raw_w, int top, int stride, int oh, int ow,
                                  long long n, uint8_t* __restrict__ out) {
    const long long total = n * oh * ow;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long img = i / (oh * ow);
        const int p = (int)(i - img * oh * ow), y = p / ow, x = p - y * ow;
        const uint8_t* px = obs + ((size_t)img * raw_h * raw_w + (size_t)(top + y * stride) * raw_w + (size_t)x * stride) * 3;
        const double r = (double)px[0], g = (double)px[1], b = (double)px[2];
        const double v = fma(b, 0.144, fma(g, 0.587, __dmul_rn(r, 0.299)));
        out[i] = (uint8_t)((long long)v & 255);
    }
}

void launch_preprocess(dqn_learner* h, const uint8_t* d_obs, int raw_h, int raw_w, int top, int stride, int oh, int ow, int64_t n,
                       uint8_t* d_out) {
    const long long total = (long long)n * oh * ow;
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
    preprocess_kernel<<<blocks, 256, 0, h->stream>>>(d_obs, raw_h, raw_w, top, stride, oh, ow, (long long)n, d_out);
    DQN_CUDA_OK(cudaGetLastError());
    h->launches++;
}
