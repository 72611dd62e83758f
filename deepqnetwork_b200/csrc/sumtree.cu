// GPU sum tree: proportional stratified sampling, order-preserving batched update, leaf statistics.
//
// Replaces (reference file:line):
//   rl/data/sum_tree.py:113-129  _retrieve            -> tree_descend (warp-cooperative, 5 levels / round trip)
//   rl/data/sum_tree.py:76-91    get_with_info        -> tree_sample_kernel outputs
//   rl/data/sum_tree.py:45-52,101-110 update_score/_propagate -> tree_update_kernel (bit-exact, sample order)
//   rl/data/sum_tree.py:23-42    add                  -> tree_add (data store + update at the write pointer)
//   rl/data/sum_tree.py:139-165  total/min/max/avg    -> tree_stats kernels
//   rl/data/replay_sampler_priority.py:30-35 stratified probe; rl/deep_q_agent.py:506-519 IS weights
//
// The heap array has the reference's exact layout (2N-1 f32 nodes, leaves at N-1..2N-2, any N) and every
// f32 operation is done with explicit round-to-nearest intrinsics in the reference's order, so node values
// (which are history dependent: delta propagation drifts) stay bit-identical to the NumPy run.
#include "learner.h"

#define FULL 0xffffffffu

struct PerParams {
    double per_b_start, per_b_end, per_anneal_inv;  // (b_end-b_start)/anneal_steps
    int use_anneal;
    double per_b_override;  // >= 0: use this beta
};

__device__ __forceinline__ double current_per_b(const PerParams& pp, long long step) {
    if (pp.per_b_override >= 0.0) return pp.per_b_override;
    if (!pp.use_anneal) return pp.per_b_start;
    // deep_q_agent.py:449-450: min(b_end, b_start + step * ((b_end - b_start) / anneal_steps))
    double b = pp.per_b_start + (double)step * pp.per_anneal_inv;
    return b < pp.per_b_end ? b : pp.per_b_end;
}

// Walk from the root to a leaf.  All 32 lanes hold the same (score) and return the same leaf.
// Each round the warp fetches the left children of the 31 nodes of a 5-level subtree (lane q-1 holds
// tree[left(rel node q)]) and walks them with shuffles: ceil(depth/5) dependent memory round trips.
__device__ __forceinline__ long long tree_descend(const float* __restrict__ tree, long long n_nodes, float score,
                                                  int lane) {
    long long r = 0;
    while (true) {
        const int q = lane + 1;
        const int t = 31 - __clz(q);  // relative level 0..4 (lane 31: level 5, idle)
        const long long g = ((r + 1) << t) - 1 + (q - (1 << t));
        const long long left = 2 * g + 1;
        float lv = 0.f;
        if (q < 32 && left < n_nodes) lv = __ldcg(tree + left);
        int cur = 1;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int tl = 31 - __clz(cur);
            const long long gc = ((r + 1) << tl) - 1 + (cur - (1 << tl));
            if (2 * gc + 1 >= n_nodes) return gc;  // sum_tree.py:119 leaf test
            const float l = __shfl_sync(FULL, lv, cur - 1);
            if (score <= l) {  // sum_tree.py:126
                cur = 2 * cur;
            } else {
                score = __fsub_rn(score, l);  // sum_tree.py:129
                cur = 2 * cur + 1;
            }
        }
        r = ((r + 1) << 5) - 1 + (cur - 32);
    }
}

__global__ void tree_sample_kernel(const float* __restrict__ tree, const uint32_t* __restrict__ data, long long N,
                                   const long long* __restrict__ p_tree_size, int batch,
                                   const double* __restrict__ u_in, int device_rng, uint64_t seed,
                                   dqn_learner::Scalars* sc, PerParams pp, int use_per_weights,
                                   long long* __restrict__ o_tree_idx, long long* __restrict__ o_data_idx,
                                   long long* __restrict__ o_mem_idx, double* __restrict__ o_prio,
                                   double* __restrict__ o_isw64, float* __restrict__ o_isw, double* __restrict__ o_u,
                                   int ahead, long long* host_out, int host_stride, int raw_probe) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= batch) return;
    const long long n_nodes = 2 * N - 1;
    const long long size = *p_tree_size;
    double u;
    if (device_rng) {
        uint32_t r[4];
        unsigned long long ctr = sc->rng_counter + (unsigned long long)ahead;  // look-ahead sample: counters not bumped yet
        philox4x32((uint32_t)i, 0u, (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
        u = u53(r[0], r[1]);
    } else {
        u = u_in[i];
    }
    // replay_sampler_priority.py:30,33 under NEP 50: every operand is rounded to f32 first
    const float total = __ldcg(tree);
    const float seg = __fdiv_rn(total, (float)batch);
    // raw_probe: u already is the probe (SumTree.get_with_info(s) called directly, sum_tree.py:76-83)
    const float s = raw_probe ? __double2float_rn(u) : __fadd_rn(__fmul_rn(__double2float_rn(u), seg), __fmul_rn((float)i, seg));
    long long leaf = tree_descend(tree, n_nodes, s, lane);
    long long d = leaf - (N - 1);
    const long long over = d - size + 1;  // sum_tree.py:121-123 "float rounding bug" clamp
    if (over > 0) { leaf -= over; d -= over; }
    if (lane == 0) {
        const float p = __ldcg(tree + leaf);
        if (o_tree_idx) o_tree_idx[i] = leaf;
        if (o_data_idx) o_data_idx[i] = d;
        if (o_mem_idx) o_mem_idx[i] = (long long)data[d];
        if (o_prio) o_prio[i] = (double)p;
        if (o_u) o_u[i] = u;
        if (use_per_weights) {
            // deep_q_agent.py:518-519 (float64 array maths on f32 tree values)
            const double per_b = current_per_b(pp, sc->step + ahead);
            const double prob = (double)p / (double)total;
            const double w = pow((double)batch * prob, -per_b) / (double)sc->last_max_weight;
            if (o_isw64) o_isw64[i] = w;
            if (o_isw) o_isw[i] = (float)w;
            if (i == 0) sc->per_b = (float)per_b;
            if (host_out) reinterpret_cast<double*>(host_out)[2 * host_stride + i] = w;
        }
        // seam-3 callers: {tree_idx | priority | is_weight} columns straight into the pinned host block (no D2H copy op)
        if (host_out) {
            host_out[i] = leaf;
            reinterpret_cast<double*>(host_out)[host_stride + i] = (double)p;
        }
    }
}

// One CTA; warp w owns tree depth w, so every node is read-modified-written by exactly one warp and the
// per-node additions happen in update order j = 0..n-1 exactly as `tree[parent] += change` does in the
// reference loop.  32 updates are handled per round; lanes that hit the same node are folded in lane order.
__global__ void __launch_bounds__(1024) tree_update_kernel(float* tree, const uint32_t* __restrict__ data,
                                                           __half* losses, long long N,
                                                           const long long* __restrict__ tree_idx,
                                                           const float* __restrict__ scores, long long n,
                                                           int store_losses, long long gen_write0) {
    __shared__ float s_change[32];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += 32) {
        const long long j = base + lane;
        const bool valid = j < n;
        long long leaf = -1 - lane;  // distinct sentinels never match each other
        float p = 0.f;
        if (valid) {
            // gen_write0 >= 0: SumTree.add semantics, leaf = write pointer (sum_tree.py:28)
            leaf = gen_write0 >= 0 ? ((gen_write0 + j) % N) + N - 1 : tree_idx[j];
            p = scores[j];
        }
        if (w == 0) {
            float old = valid ? tree[leaf] : 0.f;
            const unsigned grp = __match_any_sync(FULL, leaf);
            const unsigned lower = grp & ((1u << lane) - 1u);
            const int prev = lower ? 31 - __clz(lower) : lane;
            const float pprev = __shfl_sync(FULL, p, prev);
            if (lower) old = pprev;  // an earlier update in this round already set this leaf
            s_change[lane] = __fsub_rn(p, old);  // sum_tree.py:49
            const bool last = (grp >> lane) == 1u;  // highest lane of the group: its value survives
            if (valid && last) {
                tree[leaf] = p;  // sum_tree.py:51
                if (store_losses) losses[data[leaf - (N - 1)]] = __float2half_rn(p);  // replay_sampler_priority.py:48-50
            }
        }
        __syncthreads();
        {
            const float c = s_change[lane];
            const int leafdepth = valid ? 63 - __clzll(leaf + 1) : -1;
            const bool part = valid && (w < leafdepth);
            const long long node = part ? (((leaf + 1) >> (leafdepth - w)) - 1) : (long long)(-1 - lane);
            const unsigned grp = __match_any_sync(FULL, node);
            const int leader = __ffs(grp) - 1;
            const int cnt = __popc(grp);
            const bool lead = part && lane == leader;
            float v = 0.f;
            if (lead) v = tree[node];
            const int maxcnt = __reduce_max_sync(FULL, part ? cnt : 0);
            unsigned rem = grp;  // group members in ascending lane (= update) order, one per round
            for (int r = 0; r < maxcnt; ++r) {
                const int src = rem ? __ffs(rem) - 1 : lane;
                rem &= rem - 1u;
                const float cr = __shfl_sync(FULL, c, src);
                if (lead && r < cnt) v = __fadd_rn(v, cr);  // sum_tree.py:107
            }
            if (lead) tree[node] = v;
        }
        __syncthreads();
    }
}

__global__ void tree_set_data_kernel(uint32_t* data, long long N, long long write0, long long n,
                                     const uint32_t* __restrict__ vals) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j < n) data[(write0 + j) % N] = vals ? vals[j] : (uint32_t)((write0 + j) % N);
}

// ---- leaf statistics -------------------------------------------------------
__global__ void tree_stats_partial(const float* __restrict__ leaves, long long n, float* pmin, float* pmax,
                                   double* psum) {
    float mn = INFINITY, mx = -INFINITY;
    double sm = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = leaves[i];
        mn = fminf(mn, v); mx = fmaxf(mx, v); sm += (double)v;
    }
    __shared__ float s_mn[32], s_mx[32];
    __shared__ double s_sm[32];
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        sm += __shfl_xor_sync(FULL, sm, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_mn[w] = mn; s_mx[w] = mx; s_sm[w] = sm; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        mn = lane < nw ? s_mn[lane] : INFINITY; mx = lane < nw ? s_mx[lane] : -INFINITY; sm = lane < nw ? s_sm[lane] : 0.0;
        for (int o = 16; o; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
            sm += __shfl_xor_sync(FULL, sm, o);
        }
        if (lane == 0) { pmin[blockIdx.x] = mn; pmax[blockIdx.x] = mx; psum[blockIdx.x] = sm; }
    }
}

__global__ void tree_stats_final(const float* pmin, const float* pmax, const double* psum, int parts,
                                 const float* __restrict__ tree, const long long* p_size, int batch, PerParams pp,
                                 dqn_learner::Scalars* sc, float* out4) {
    float mn = INFINITY, mx = -INFINITY;
    double sm = 0.0;
    for (int i = threadIdx.x; i < parts; i += 32) { mn = fminf(mn, pmin[i]); mx = fmaxf(mx, pmax[i]); sm += psum[i]; }
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        sm += __shfl_xor_sync(FULL, sm, o);
    }
    if (threadIdx.x == 0) {
        const long long n = *p_size;
        const float total = tree[0];
        const float avg = n > 0 ? (float)(sm / (double)n) : 0.f;
        if (out4) {
            // read-only getters (SumTree.get_min/get_max/get_average/total): the handle's PER state keeps the cadence of
            // DeepQAgent._sample_memories (step % per_calculate_steps == 0, deep_q_agent.py:505)
            out4[0] = mn; out4[1] = mx; out4[2] = avg; out4[3] = total;
            return;
        }
        sc->stat_min = mn; sc->stat_max = mx; sc->stat_total = total;
        sc->stat_avg = avg;
        // deep_q_agent.py:508-510: last_min = max(min, 0.01); last_max_weight = pow(B*(last_min/total), -per_b)
        const float last_min = fmaxf(mn, 0.01f);
        const float x = __fmul_rn((float)batch, __fdiv_rn(last_min, total));
        sc->last_max_weight = (float)pow((double)x, -current_per_b(pp, sc->step));
        sc->has_max_weight = 1;
    }
}

// ---- launchers ---------------------------------------------------------------
static PerParams make_pp(const dqn_learner* h, double override_b) {
    PerParams pp;
    pp.per_b_start = h->cfg.per_b_start;
    pp.per_b_end = h->cfg.per_b_end;
    pp.per_anneal_inv = (h->cfg.per_b_end - h->cfg.per_b_start) / (double)h->cfg.per_anneal_steps;
    pp.use_anneal = h->cfg.use_per_annealing;
    pp.per_b_override = override_b;
    return pp;
}

long long* tree_size_ptr(dqn_learner* h);  // api.cu: device mirror of tree_size

void launch_tree_sample(dqn_learner* h, int batch, int device_rng, double per_b, const double* u_src, void* host_out, bool raw_probe) {
    const int wpb = 8;
    launch_kernel(false, tree_sample_kernel, dim3(ceil_div(batch, wpb)), dim3(wpb * 32), 0, h->stream,
        h->tree, h->tree_data, h->cap, tree_size_ptr(h), batch, u_src ? u_src : (const double*)h->b_u, device_rng, h->cfg.seed, h->d_sc,
        make_pp(h, per_b), h->cfg.use_per, h->b_tree_idx, h->b_data_idx, h->b_mem_idx, h->b_prio, h->b_isw64,
        h->b_isw, device_rng ? h->b_u : (double*)nullptr, h->sample_ahead, (long long*)host_out, h->max_rows, raw_probe ? 1 : 0);
    h->launches++;
}

void launch_tree_update(dqn_learner* h, int n, const long long* d_tree_idx, const float* d_scores, int store_losses) {
    launch_kernel(false, tree_update_kernel, dim3(1), dim3(1024), 0, h->stream, h->tree, h->tree_data, h->r_losses, h->cap,
                  d_tree_idx, d_scores, n, store_losses, -1);
    h->launches++;
}

void launch_tree_add(dqn_learner* h, int64_t n, const float* d_scores, const uint32_t* d_data, int64_t write0) {
    tree_set_data_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, h->stream>>>(h->tree_data, h->cap, write0, n, d_data);
    tree_update_kernel<<<1, 1024, 0, h->stream>>>(h->tree, h->tree_data, h->r_losses, h->cap, nullptr, d_scores, n, 0,
                                                   write0);
    DQN_CUDA_OK(cudaGetLastError());
    h->launches += 2;
}

void launch_tree_stats(dqn_learner* h, double per_b_override, float* d_out4) {
    const int parts = 296;
    float* pmin = (float*)h->stats_scratch;
    float* pmax = pmin + parts;
    double* psum = (double*)(pmax + parts);
    const long long n = h->tree_size;
    tree_stats_partial<<<parts, 256, 0, h->stream>>>(h->tree + (h->cap - 1), n, pmin, pmax, psum);
    tree_stats_final<<<1, 32, 0, h->stream>>>(pmin, pmax, psum, parts, h->tree, tree_size_ptr(h), h->B,
                                               make_pp(h, per_b_override), h->d_sc, d_out4);
    DQN_CUDA_OK(cudaGetLastError());
    h->launches += 2;
}
