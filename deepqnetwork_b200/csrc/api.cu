// C ABI of libdqn_b200 (see include/dqn_b200.h): handle lifecycle, host<->device plumbing and the
// orchestration of one training step (DeepQAgent._train_run_train, rl/deep_q_agent.py:326-349).
#include "learner.h"
#include <string.h>
#include <algorithm>

static std::string g_create_error;

// launchers living in other translation units
void debug_conv_layer(dqn_learner* h, int layer, int rows);
void launch_fill_synthetic(dqn_learner* h, int64_t start, int64_t n, uint64_t seed, float* d_prio);

long long* tree_size_ptr(dqn_learner* h) { return &h->d_dev->tree_size; }
void select_set(dqn_learner* h, int k);  // point the b_* batch aliases at buffer set k

// every entry point except dqn_train_steps drops the batch that was sampled ahead (it may change the tree, the ring,
// the counters or the batch buffers); the next fused step then samples eagerly -- same draws, same result
#define API_BEGIN_KEEP(h)                   \
    if (!(h)) return 2;                     \
    try {                                   \
        DQN_CUDA_OK(cudaSetDevice((h)->device));
#define API_BEGIN(h)                        \
    API_BEGIN_KEEP(h)                       \
        (h)->pf_valid = false;
#define API_END(h)                          \
    } catch (const std::string& e) {        \
        (h)->err = e;                       \
        return 1;                           \
    }                                       \
    return 0;

template <class T>
static void dmalloc(T*& p, size_t count) {
    void* q = nullptr;
    DQN_CUDA_OK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    p = static_cast<T*>(q);
}

__global__ void set_ll_kernel(long long* p, long long v) { *p = v; }
__global__ void set_f2_kernel(float* a, float va, float* b, float vb) { *a = va; *b = vb; }
static void set_ll(dqn_learner* h, long long* p, long long v) {
    set_ll_kernel<<<1, 1, 0, h->stream>>>(p, v);
    DQN_CUDA_OK(cudaGetLastError());
}
static void push_sizes(dqn_learner* h) {
    set_ll(h, &h->d_dev->tree_size, h->tree_size);
    set_ll(h, &h->d_dev->ring_len, h->ring_size);
}
constexpr size_t SC_BLOCK = 128;  // Scalars, padded; the losses of the last step follow (device block and pinned mirror)

static void read_scalars(dqn_learner* h) {
    DQN_CUDA_OK(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(dqn_learner::Scalars), cudaMemcpyDeviceToHost, h->stream));
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
}

static void alloc_tower(dqn_learner* h, TowerActs& t, int rows) {
    const NetDesc& net = h->net;
    for (int l = 0; l < 3; ++l) dmalloc(t.act[l], (size_t)rows * net.conv[l].oh * net.conv[l].ow * net.conv[l].cout);
    dmalloc(t.fc_part, (size_t)h->fc_splits * rows * net.NH);
    dmalloc(t.hid, (size_t)rows * net.NH);
    dmalloc(t.q, (size_t)rows * net.A);
}

extern "C" int dqn_create(const dqn_config* cfg, dqn_learner** out) {
    if (!cfg || !out) { g_create_error = "null argument"; return 2; }
    dqn_learner* h = nullptr;
    try {
        DQN_REQUIRE(cfg->struct_size == (int32_t)sizeof(dqn_config), "dqn_config size mismatch (ABI drift)");
        DQN_REQUIRE(cfg->num_actions >= 1 && cfg->num_actions <= 31, "num_actions must be in [1,31]");
        DQN_REQUIRE(cfg->batch_size >= 1 && cfg->batch_size <= 4096, "batch_size out of range");
        DQN_REQUIRE(cfg->replay_capacity >= 2, "replay_capacity must be >= 2");
        DQN_REQUIRE(cfg->math_mode >= DQN_MATH_FP32 && cfg->math_mode <= DQN_MATH_TC1X, "bad math_mode");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) throw std::string("no CUDA device: libdqn_b200 has no CPU fallback");
        DQN_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
        DQN_CUDA_OK(cudaSetDevice(cfg->device));
        h = new dqn_learner();
        h->cfg = *cfg;
        h->device = cfg->device;
        cudaDeviceProp prop;
        DQN_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
        h->sm_count = prop.multiProcessorCount;
        {   // the main stream carries the critical path of a step (forward, dgrad chain); the side streams only hold
            // work that has slack (weight gradients, optimizer, look-ahead sampling): give the main stream priority
            int lo = 0, hi = 0;
            DQN_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            h->prio_lo = lo; h->prio_hi = hi;
            DQN_CUDA_OK(cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, hi));
        }
        h->stream = h->own_stream;
        DQN_CUDA_OK(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        DQN_CUDA_OK(cudaStreamCreateWithFlags(&h->side2_stream, cudaStreamNonBlocking));
        DQN_CUDA_OK(cudaStreamCreateWithFlags(&h->side3_stream, cudaStreamNonBlocking));
        DQN_CUDA_OK(cudaStreamCreateWithFlags(&h->side4_stream, cudaStreamNonBlocking));
        DQN_CUDA_OK(cudaStreamCreateWithFlags(&h->side5_stream, cudaStreamNonBlocking));
        DQN_CUDA_OK(cudaStreamCreateWithFlags(&h->side6_stream, cudaStreamNonBlocking));
        for (auto& e : h->fork_ev) DQN_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : h->phase_ev) DQN_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        net_describe(h->net, h->cfg);
        const NetDesc& net = h->net;
        h->state_bytes = (size_t)net.H * net.W * net.C;
        DQN_REQUIRE(h->state_bytes % 16 == 0, "H*W*C must be a multiple of 16 bytes");
        DQN_REQUIRE(net.C == 4, "input_channels must be 4 (NUM_FRAMES_PER_STATE)");
        h->B = cfg->batch_size;
        h->max_rows = cfg->max_rows > 0 ? cfg->max_rows : std::max(2 * h->B, 256);
        if (h->max_rows < h->B) h->max_rows = h->B;
        h->fc_splits = std::max(ceil_div(net.flat, 256), 40);

        // parameters + optimizer slots + gradient slab
        const size_t P = (size_t)net.slab;
        {
            // the five parameter-sized slabs come from ONE allocation at a fixed relative placement
            const size_t stride = (((size_t)P * 4 + (2u << 20) - 1) / (2u << 20)) * (2u << 20);  // bytes
            float* base = nullptr;
            dmalloc(base, 5 * stride / 4);
            h->slab_base = base;
            h->online = base; h->target = base + stride / 4; h->slot_m = base + 2 * (stride / 4); h->slot_v = base + 3 * (stride / 4);
            h->grads = base + 4 * (stride / 4);
        }
        for (float* p : {h->online, h->target, h->slot_m, h->slot_v, h->grads}) DQN_CUDA_OK(cudaMemsetAsync(p, 0, P * 4, h->stream));
        // scalars and the per-sample losses of a step share one block, so that dqn_train_batch reads both with one copy
        static_assert(sizeof(dqn_learner::Scalars) <= SC_BLOCK, "scalar block");
        {
            void* blk = nullptr;
            DQN_CUDA_OK(cudaMalloc(&blk, SC_BLOCK + 4 * (size_t)h->max_rows));
            h->d_sc = static_cast<dqn_learner::Scalars*>(blk);
            h->b_losses_out = reinterpret_cast<float*>(static_cast<char*>(blk) + SC_BLOCK);
        }
        dmalloc(h->d_dev, 1);
        DQN_CUDA_OK(cudaMemsetAsync(h->d_sc, 0, sizeof(dqn_learner::Scalars), h->stream));
        DQN_CUDA_OK(cudaMemsetAsync(h->d_dev, 0, sizeof(dqn_learner::DevSizes), h->stream));
        DQN_CUDA_OK(cudaHostAlloc((void**)&h->h_sc, SC_BLOCK + 4 * (size_t)h->max_rows, cudaHostAllocMapped));
        DQN_CUDA_OK(cudaHostAlloc((void**)&h->h_sc_alt, SC_BLOCK + 4 * (size_t)h->max_rows, cudaHostAllocMapped));
        for (int k = 0; k < 2; ++k) DQN_CUDA_OK(cudaEventCreateWithFlags(&h->ustep_ev[k], cudaEventDisableTiming));
        set_f2_kernel<<<1, 1, 0, h->stream>>>(&h->d_sc->b1p, 0.9f, &h->d_sc->b2p, 0.999f);

        // replay ring + tree
        h->cap = cfg->replay_capacity;
        const size_t N = (size_t)h->cap;
        h->dedup = cfg->frame_dedup != 0;
        h->frame_bytes = (size_t)net.H * net.W;
        if (h->dedup) {
            DQN_REQUIRE(h->frame_bytes % 16 == 0, "frame_dedup: H*W must be a multiple of 16 bytes");
            h->fcap = cfg->frame_capacity > 0 ? cfg->frame_capacity : h->cap + h->cap / 16 + 64;
            DQN_REQUIRE(h->fcap >= 8 && h->fcap < (int64_t)1 << 32, "frame_capacity out of range");
            dmalloc(h->r_frames, (size_t)h->fcap * h->frame_bytes);
            dmalloc(h->r_fidx, N * 8);
            DQN_CUDA_OK(cudaMemsetAsync(h->r_fidx, 0, N * 8 * sizeof(uint32_t), h->stream));
        } else {
            dmalloc(h->r_states, N * h->state_bytes); dmalloc(h->r_next, N * h->state_bytes);
        }
        dmalloc(h->r_actions, N); dmalloc(h->r_rewards, N); dmalloc(h->r_cont, N); dmalloc(h->r_losses, N);
        DQN_CUDA_OK(cudaMemsetAsync(h->r_losses, 0, N * sizeof(__half), h->stream));
        dmalloc(h->tree, 2 * N - 1); dmalloc(h->tree_data, N);
        DQN_CUDA_OK(cudaMemsetAsync(h->tree, 0, (2 * N - 1) * 4, h->stream));
        DQN_CUDA_OK(cudaMemsetAsync(h->tree_data, 0, N * 4, h->stream));

        // batches
        const size_t R = (size_t)h->max_rows;
        dmalloc(h->b_states, 2 * R * h->state_bytes); dmalloc(h->in_states, 2 * R * h->state_bytes);
        dmalloc(h->b_actions, R); dmalloc(h->b_rewards, R); dmalloc(h->b_cont, R); dmalloc(h->b_losses, R);
        {   // the small per-sample inputs of a host-fed batch share one device block and one pinned staging block, so
            // that a batch upload is the two state copies plus ONE small copy
            const size_t Rp = (R + 15) / 16 * 16;
            h->in_small_bytes = 3 * Rp + 4 * R;
            uint8_t* blk = nullptr;
            dmalloc(blk, h->in_small_bytes);
            h->in_actions = blk; h->in_rewards = blk + Rp; h->in_cont = blk + 2 * Rp; h->in_isw = reinterpret_cast<float*>(blk + 3 * Rp);
            DQN_CUDA_OK(cudaHostAlloc((void**)&h->h_in_small, std::max<size_t>(h->in_small_bytes, 16 * R), cudaHostAllocMapped));
            DQN_CUDA_OK(cudaEventCreateWithFlags(&h->stage_ev, cudaEventDisableTiming));
        }
        auto carve = [&](dqn_learner::BatchSet& b) {
            dmalloc(b.blk, 6 * R);
            b.tree_idx = b.blk; b.prio = reinterpret_cast<double*>(b.blk + R); b.isw64 = reinterpret_cast<double*>(b.blk + 2 * R);
            b.data_idx = b.blk + 3 * R; b.mem_idx = b.blk + 4 * R; b.u = reinterpret_cast<double*>(b.blk + 5 * R);
        };
        carve(h->sets[0]); carve(h->sets[1]);
        h->b_tree_idx = h->sets[0].tree_idx; h->b_data_idx = h->sets[0].data_idx; h->b_mem_idx = h->sets[0].mem_idx;
        h->b_prio = h->sets[0].prio; h->b_isw64 = h->sets[0].isw64; h->b_u = h->sets[0].u;
        dmalloc(h->b_isw, R);
        DQN_CUDA_OK(cudaHostAlloc((void**)&h->h_stage2, 64 * R, cudaHostAllocMapped));
        DQN_CUDA_OK(cudaHostAlloc((void**)&h->h_stage2_alt, 64 * R, cudaHostAllocMapped));
        DQN_CUDA_OK(cudaEventCreateWithFlags(&h->stage2_ev, cudaEventDisableTiming));
        dmalloc(h->b_y, R); dmalloc(h->b_maxq, R); dmalloc(h->b_dq, R); dmalloc(h->b_newprio, R);
        dmalloc(h->b_rows, R);

        // activations
        alloc_tower(h, h->acts_on, 2 * h->max_rows);
        alloc_tower(h, h->acts_tg, h->max_rows);
        dmalloc(h->d_hid, R * net.NH);
        for (int l = 0; l < 3; ++l) dmalloc(h->d_act[l], R * net.conv[l].oh * net.conv[l].ow * net.conv[l].cout);
        h->wg_part_floats = 0;
        int64_t wg_total = 0;
        for (int l = 0; l < 3; ++l) {
            const ConvGeom& g = net.conv[l];
            const int64_t need = (int64_t)64 * (g.k * g.k * g.cin + 4) * g.cout;
            h->wg_part_floats = std::max<int64_t>(h->wg_part_floats, need);
            h->wg_off[l] = wg_total;
            wg_total += need;
        }
        dmalloc(h->wg_part, (size_t)wg_total);
        dmalloc(h->zeros16, 4); dmalloc(h->ones_row, 4); dmalloc(h->opt_ticket, 4); h->fa_tile_ctr = h->opt_ticket + 1;
        DQN_CUDA_OK(cudaMemsetAsync(h->zeros16, 0, 16, h->stream));
        DQN_CUDA_OK(cudaMemsetAsync(h->opt_ticket, 0, 16, h->stream));
        {
            const float one[4] = {1.f, 0.f, 0.f, 0.f};
            DQN_CUDA_OK(cudaMemcpyAsync(h->ones_row, one, 16, cudaMemcpyHostToDevice, h->stream));
            DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        }
        dmalloc(h->x_f32, 2 * R * h->state_bytes);
        {
            dqn_learner::BatchSet& a = h->sets[0];
            a.states = h->b_states; a.actions = h->b_actions; a.rewards = h->b_rewards; a.cont = h->b_cont; a.losses = h->b_losses;
            a.tree_idx = h->b_tree_idx; a.data_idx = h->b_data_idx; a.mem_idx = h->b_mem_idx; a.rows = h->b_rows;
            a.prio = h->b_prio; a.isw64 = h->b_isw64; a.u = h->b_u; a.isw = h->b_isw; a.x_f32 = h->x_f32;
            dqn_learner::BatchSet& b = h->sets[1];
            dmalloc(b.states, 2 * R * h->state_bytes); dmalloc(b.actions, R); dmalloc(b.rewards, R); dmalloc(b.cont, R);
            dmalloc(b.losses, R); dmalloc(b.rows, R);
            dmalloc(b.isw, R); dmalloc(b.x_f32, 2 * R * h->state_bytes);
        }
        alloc_packed_conv(h);
        launch_pack_conv(h, 0); launch_pack_conv(h, 1);
        if (net.dueling) dmalloc(h->w_scratch, (size_t)2 * net.flat * net.hidden);
        h->dev_stage_bytes = 8u << 20;
        dmalloc(h->dev_stage, h->dev_stage_bytes);
        dmalloc(h->stats_scratch, 16384);
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    } catch (const std::string& e) {
        g_create_error = e;
        if (h) dqn_destroy(h);  // frees whatever was allocated before the failure (every pointer starts out null)
        return 1;
    }
    *out = h;
    return 0;
}

extern "C" int dqn_destroy(dqn_learner* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->sets[0].states) select_set(h, 0);  // the b_* aliases freed below must be set 0 (set 1 is freed separately)
    if (h->step_graph) cudaGraphExecDestroy(h->step_graph);
    if (h->prof_graph) cudaGraphExecDestroy(h->prof_graph);
    void* ptrs[] = {h->slab_base, h->d_sc, h->d_dev, h->r_states, h->r_next, h->r_frames, h->r_fidx,
                    h->r_actions, h->r_rewards, h->r_cont, h->r_losses, h->tree, h->tree_data, h->b_states, h->in_states,
                    h->b_actions, h->b_rewards, h->b_cont, h->b_losses, h->in_actions,
                    h->sets[0].blk, h->sets[1].blk, h->b_isw, h->b_y, h->b_maxq,
                    h->b_dq, h->b_newprio, h->b_rows, h->d_hid, h->d_act[0], h->d_act[1], h->d_act[2],
                    h->wg_part, h->dev_stage, h->stats_scratch, h->w_scratch, h->zeros16, h->ones_row, h->x_f32, h->opt_ticket, h->packed[0], h->packed[1], h->packed_dg, h->act3_pk[0], h->act3_pk[1]};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (TowerActs* t : {&h->acts_on, &h->acts_tg}) {
        for (int l = 0; l < 3; ++l) cudaFree(t->act[l]);
        cudaFree(t->fc_part); cudaFree(t->hid); cudaFree(t->q);
    }
    if (h->fa_maps) free(h->fa_maps);
    if (h->fa_row_maps) free(h->fa_row_maps);
    for (int t = 0; t < 2; ++t) if (h->ff_maps[t]) free(h->ff_maps[t]);
    if (h->h_sc) cudaFreeHost(h->h_sc);
    if (h->h_sc_alt) cudaFreeHost(h->h_sc_alt);
    if (h->h_stage2_alt) cudaFreeHost(h->h_stage2_alt);
    for (int k = 0; k < 2; ++k) if (h->ustep_ev[k]) cudaEventDestroy(h->ustep_ev[k]);
    if (h->ustep_graph_alt) cudaGraphExecDestroy(h->ustep_graph_alt);
    if (h->h_in_small) cudaFreeHost(h->h_in_small);
    if (h->h_stage2) cudaFreeHost(h->h_stage2);
    if (h->stage2_ev) cudaEventDestroy(h->stage2_ev);
    if (h->stage_ev) cudaEventDestroy(h->stage_ev);
    for (auto& ev : h->prof_events) cudaEventDestroy(ev);
    for (auto& e : h->fork_ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->phase_ev) if (e) cudaEventDestroy(e);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->side2_stream) cudaStreamDestroy(h->side2_stream);
    if (h->side3_stream) cudaStreamDestroy(h->side3_stream);
    if (h->side4_stream) cudaStreamDestroy(h->side4_stream);
    if (h->side5_stream) cudaStreamDestroy(h->side5_stream);
    if (h->side6_stream) cudaStreamDestroy(h->side6_stream);
    for (int k = 0; k < 2; ++k) for (int v = 0; v < 2; ++v) if (h->pipe_graph[k][v]) cudaGraphExecDestroy(h->pipe_graph[k][v]);
    {
        dqn_learner::BatchSet& b = h->sets[1];
        void* p2[] = {b.states, b.actions, b.rewards, b.cont, b.losses, b.rows, b.isw, b.x_f32};
        for (void* p : p2) if (p) cudaFree(p);
    }
    if (h->batch_graph) cudaGraphExecDestroy(h->batch_graph);
    if (h->ustep_graph) cudaGraphExecDestroy(h->ustep_graph);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return 0;
}

extern "C" const char* dqn_last_error(const dqn_learner* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int dqn_sync(dqn_learner* h) {
    API_BEGIN(h)
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_set_stream(dqn_learner* h, void* s) {
    API_BEGIN(h)
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    h->graph_valid = false; h->graph_valid_batch = false; h->graph_valid_u = false; h->graph_valid_u_alt = false;
    if (h->prof_graph) { cudaGraphExecDestroy(h->prof_graph); h->prof_graph = nullptr; }
    API_END(h)
}

// ---- parameters ---------------------------------------------------------------------------------------
extern "C" int dqn_num_params(const dqn_learner* h, int32_t* n_tensors, int64_t* n_floats) {
    if (!h) return 2;
    int64_t tot = 0;
    for (auto& t : h->net.tensors) tot += t.count;
    if (n_tensors) *n_tensors = (int32_t)h->net.tensors.size();
    if (n_floats) *n_floats = tot;
    return 0;
}

extern "C" int dqn_param_info(const dqn_learner* h, int32_t i, const char** name, int64_t* count, int64_t* offset) {
    if (!h || i < 0 || i >= (int)h->net.tensors.size()) return 2;
    const TensorInfo& t = h->net.tensors[i];
    if (name) *name = t.name.c_str();
    if (count) *count = t.count;
    if (offset) *offset = t.offset;
    return 0;
}

static float* slab_of(dqn_learner* h, int tower, int slot) {
    if (slot == 0) return tower == 0 ? h->online : h->target;
    DQN_REQUIRE(tower == 0, "optimizer slots exist for the online tower only");
    if (slot == 1) return h->slot_m;
    if (slot == 2) return h->slot_v;
    if (slot == 3) return h->grads;
    throw std::string("bad slot");
}
static const TensorInfo& find_tensor(dqn_learner* h, const char* name) {
    for (auto& t : h->net.tensors) if (t.name == name) return t;
    throw std::string("unknown parameter name: ") + name;
}

extern "C" int dqn_get_param(dqn_learner* h, int32_t tower, int32_t slot, const char* name, float* out, int64_t count) {
    API_BEGIN(h)
    const TensorInfo& t = find_tensor(h, name);
    DQN_REQUIRE(count == t.count, "parameter element count mismatch");
    DQN_CUDA_OK(cudaMemcpyAsync(out, slab_of(h, tower, slot) + t.offset, count * 4, cudaMemcpyDeviceToHost, h->stream));
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_set_param(dqn_learner* h, int32_t tower, int32_t slot, const char* name, const float* in, int64_t count) {
    API_BEGIN(h)
    const TensorInfo& t = find_tensor(h, name);
    DQN_REQUIRE(count == t.count, "parameter element count mismatch");
    DQN_CUDA_OK(cudaMemcpyAsync(slab_of(h, tower, slot) + t.offset, in, count * 4, cudaMemcpyHostToDevice, h->stream));
    if (slot == 0 && t.name.find("conv") != std::string::npos) launch_pack_conv(h, tower);
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_device_ptr(dqn_learner* h, const char* what, void** ptr, int64_t* bytes) {
    API_BEGIN(h)
    const std::string w = what;
    const int64_t P4 = h->net.slab * 4;
    void* p = nullptr; int64_t b = 0;
    if (w == "grads") { p = h->grads; b = P4; }
    else if (w == "online") { p = h->online; b = P4; }
    else if (w == "target") { p = h->target; b = P4; }
    else if (w == "adam_m") { p = h->slot_m; b = P4; }
    else if (w == "adam_v") { p = h->slot_v; b = P4; }
    else if (w == "tree") { p = h->tree; b = (2 * h->cap - 1) * 4; }
    else if (w == "replay_frames") { p = h->r_frames; b = h->fcap * (int64_t)h->frame_bytes; }
    else if (w == "replay_states") { p = h->r_states; b = h->cap * (int64_t)h->state_bytes; }
    else if (w == "replay_next_states") { p = h->r_next; b = h->cap * (int64_t)h->state_bytes; }
    else if (w == "batch_states") { p = h->b_states; b = 2 * (int64_t)h->max_rows * h->state_bytes; }
    else if (w == "scalars") { p = h->d_sc; b = sizeof(dqn_learner::Scalars); }
    else if (w == "stream") { p = (void*)h->stream; b = 0; }  // the cudaStream_t the handle issues its work on (CUDA-event timing by the caller)
    else throw std::string("unknown device buffer: ") + w;
    if (ptr) *ptr = p;
    if (bytes) *bytes = b;
    API_END(h)
}

extern "C" int dqn_get_step(dqn_learner* h, int64_t* step) {
    API_BEGIN(h)
    read_scalars(h);
    h->host_step = h->h_sc->step;
    if (step) *step = h->h_sc->step;
    API_END(h)
}
extern "C" int dqn_set_step(dqn_learner* h, int64_t step) {
    API_BEGIN(h)
    set_ll(h, &h->d_sc->step, step);
    h->host_step = step;
    API_END(h)
}
extern "C" int dqn_get_game_count(dqn_learner* h, int64_t* n) {
    API_BEGIN(h)
    read_scalars(h);
    if (n) *n = h->h_sc->game_count;
    API_END(h)
}
extern "C" int dqn_set_game_count(dqn_learner* h, int64_t n) {
    API_BEGIN(h)
    set_ll(h, &h->d_sc->game_count, n);
    API_END(h)
}
extern "C" int dqn_get_beta_powers(dqn_learner* h, float* b1p, float* b2p) {
    API_BEGIN(h)
    read_scalars(h);
    if (b1p) *b1p = h->h_sc->b1p;
    if (b2p) *b2p = h->h_sc->b2p;
    API_END(h)
}
extern "C" int dqn_set_beta_powers(dqn_learner* h, float b1p, float b2p) {
    API_BEGIN(h)
    set_f2_kernel<<<1, 1, 0, h->stream>>>(&h->d_sc->b1p, b1p, &h->d_sc->b2p, b2p);
    DQN_CUDA_OK(cudaGetLastError());
    API_END(h)
}
extern "C" int dqn_copy_network(dqn_learner* h) {
    API_BEGIN(h)
    launch_copy_network(h);
    API_END(h)
}
extern "C" int dqn_get_stats(dqn_learner* h, float* out8) {
    API_BEGIN(h)
    read_scalars(h);
    const dqn_learner::Scalars& s = *h->h_sc;
    const float v[8] = {s.stat_min, s.stat_max, s.stat_avg, s.stat_total, s.last_max_weight, s.per_b, s.loss, s.loss_sum};
    memcpy(out8, v, sizeof(v));
    API_END(h)
}

// ---- replay ring ------------------------------------------------------------------------------------------
__global__ void store_losses_kernel(__half* losses, long long cap, long long cur, long long n, const float* __restrict__ src) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j < n) losses[(cur + j) % cap] = __float2half_rn(src[j]);  // replay_memory.py:109-110: f16 store
}

static void ring_copy(dqn_learner* h, uint8_t* ring, const uint8_t* src, int64_t cur, int64_t n, size_t elt) {
    const int64_t first = std::min<int64_t>(n, h->cap - cur);
    DQN_CUDA_OK(cudaMemcpyAsync(ring + (size_t)cur * elt, src, (size_t)first * elt, cudaMemcpyHostToDevice, h->stream));
    if (n > first)
        DQN_CUDA_OK(cudaMemcpyAsync(ring, src + (size_t)first * elt, (size_t)(n - first) * elt, cudaMemcpyHostToDevice, h->stream));
}

extern "C" int dqn_replay_append(dqn_learner* h, int64_t n, const uint8_t* st, const uint8_t* ac, const uint8_t* rw,
                                 const uint8_t* ns, const uint8_t* ct, const float* losses) {
    API_BEGIN(h)
    DQN_REQUIRE(!(h->cfg.use_per && !losses), "use_per: losses (priorities) are required on append");
    DQN_REQUIRE(!h->dedup, "frame_dedup layout: append frames with dqn_frames_append and rows with dqn_replay_append_indexed");
    int64_t done = 0;
    const int64_t max_chunk = std::min<int64_t>(h->cap, (int64_t)(h->dev_stage_bytes / 4));
    while (done < n) {
        const int64_t m = std::min(n - done, max_chunk);
        const int64_t cur = h->ring_cur;
        if (losses) {
            float* d_sc = (float*)h->dev_stage;
            DQN_CUDA_OK(cudaMemcpyAsync(d_sc, losses + done, m * 4, cudaMemcpyHostToDevice, h->stream));
            if (h->cfg.use_per) {
                // ReplaySamplerPriority.append: sum_tree.add(loss, replay_memory.cur_idx) then replay append
                // (replay_sampler_priority.py:22-24); the tree write pointer moves in lock-step with the ring
                DQN_REQUIRE(h->tree_write == cur, "sum tree write pointer out of step with the ring");
                launch_tree_add(h, m, d_sc, nullptr, h->tree_write);
                if (h->tree_write + m >= h->cap) h->tree_max_reached = true;
                h->tree_size = h->tree_max_reached ? h->cap : h->tree_write + m;
                h->tree_write = (h->tree_write + m) % h->cap;
            }
            store_losses_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, h->stream>>>(h->r_losses, h->cap, cur, m, d_sc);
            DQN_CUDA_OK(cudaGetLastError());
        }
        ring_copy(h, h->r_states, st + (size_t)done * h->state_bytes, cur, m, h->state_bytes);
        ring_copy(h, h->r_next, ns + (size_t)done * h->state_bytes, cur, m, h->state_bytes);
        ring_copy(h, h->r_actions, ac + done, cur, m, 1);
        ring_copy(h, h->r_rewards, rw + done, cur, m, 1);
        ring_copy(h, h->r_cont, ct + done, cur, m, 1);
        if (cur + m >= h->cap) h->ring_full = true;
        h->ring_cur = (cur + m) % h->cap;
        h->ring_size = h->ring_full ? h->cap : h->ring_cur;
        done += m;
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));  // staging buffer and caller memory are reusable on return
    }
    push_sizes(h);
    API_END(h)
}

extern "C" int dqn_frames_append(dqn_learner* h, int64_t n, const uint8_t* frames, int64_t* first_slot) {
    API_BEGIN(h)
    DQN_REQUIRE(h->dedup, "handle was created without frame_dedup");
    DQN_REQUIRE(n >= 0 && n <= h->fcap, "more frames than the frame ring holds");
    const int64_t slot0 = h->f_written % h->fcap;
    if (first_slot) *first_slot = slot0;
    const int64_t first = std::min<int64_t>(n, h->fcap - slot0);
    DQN_CUDA_OK(cudaMemcpyAsync(h->r_frames + (size_t)slot0 * h->frame_bytes, frames, (size_t)first * h->frame_bytes, cudaMemcpyHostToDevice, h->stream));
    if (n > first)
        DQN_CUDA_OK(cudaMemcpyAsync(h->r_frames, frames + (size_t)first * h->frame_bytes, (size_t)(n - first) * h->frame_bytes, cudaMemcpyHostToDevice, h->stream));
    h->f_written += n;
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));  // caller memory is reusable on return
    API_END(h)
}

extern "C" int dqn_frames_info(dqn_learner* h, int64_t* frame_capacity, int64_t* frames_written) {
    if (!h) return 2;
    if (frame_capacity) *frame_capacity = h->fcap;
    if (frames_written) *frames_written = h->f_written;
    return 0;
}

extern "C" int dqn_replay_append_indexed(dqn_learner* h, int64_t n, const uint32_t* slots, const uint8_t* ac, const uint8_t* rw,
                                         const uint8_t* ct, const float* losses) {
    API_BEGIN(h)
    DQN_REQUIRE(h->dedup, "handle was created without frame_dedup");
    DQN_REQUIRE(!(h->cfg.use_per && !losses), "use_per: losses (priorities) are required on append");
    for (int64_t i = 0; i < n * 8; ++i) DQN_REQUIRE((int64_t)slots[i] < h->fcap, "frame slot out of range");
    int64_t done = 0;
    const int64_t max_chunk = std::min<int64_t>(h->cap, (int64_t)(h->dev_stage_bytes / 4));
    while (done < n) {
        const int64_t m = std::min(n - done, max_chunk);
        const int64_t cur = h->ring_cur;
        if (losses) {
            float* d_sc = (float*)h->dev_stage;
            DQN_CUDA_OK(cudaMemcpyAsync(d_sc, losses + done, m * 4, cudaMemcpyHostToDevice, h->stream));
            if (h->cfg.use_per) {  // ReplaySamplerPriority.append (replay_sampler_priority.py:22-24)
                DQN_REQUIRE(h->tree_write == cur, "sum tree write pointer out of step with the ring");
                launch_tree_add(h, m, d_sc, nullptr, h->tree_write);
                if (h->tree_write + m >= h->cap) h->tree_max_reached = true;
                h->tree_size = h->tree_max_reached ? h->cap : h->tree_write + m;
                h->tree_write = (h->tree_write + m) % h->cap;
            }
            store_losses_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, h->stream>>>(h->r_losses, h->cap, cur, m, d_sc);
            DQN_CUDA_OK(cudaGetLastError());
        }
        ring_copy(h, reinterpret_cast<uint8_t*>(h->r_fidx), reinterpret_cast<const uint8_t*>(slots + (size_t)done * 8), cur, m, 32);
        ring_copy(h, h->r_actions, ac + done, cur, m, 1);
        ring_copy(h, h->r_rewards, rw + done, cur, m, 1);
        ring_copy(h, h->r_cont, ct + done, cur, m, 1);
        if (cur + m >= h->cap) h->ring_full = true;
        h->ring_cur = (cur + m) % h->cap;
        h->ring_size = h->ring_full ? h->cap : h->ring_cur;
        done += m;
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    push_sizes(h);
    API_END(h)
}

// GameEnvironment.preprocess_observation for n raw observations (actor side): crop, take every stride-th pixel, "grey",
// uint8.  to_frame_ring != 0 (frame_dedup handles): the frames are appended to the frame ring on the device and
// *first_slot receives the slot of the first one; h_frames (may be NULL then) also receives them.
extern "C" int dqn_preprocess(dqn_learner* h, int64_t n, const uint8_t* h_obs, int32_t raw_h, int32_t raw_w, int32_t top,
                              int32_t bottom, int32_t stride, uint8_t* h_frames, int32_t to_frame_ring, int64_t* first_slot) {
    API_BEGIN(h)
    DQN_REQUIRE(n >= 1 && stride >= 1 && top >= 0 && bottom >= 0 && raw_h - bottom > top, "dqn_preprocess: bad geometry");
    const int oh = (raw_h - bottom - top + stride - 1) / stride, ow = (raw_w + stride - 1) / stride;
    const size_t in_b = (size_t)raw_h * raw_w * 3, out_b = (size_t)oh * ow;
    if (to_frame_ring) {
        DQN_REQUIRE(h->dedup, "to_frame_ring needs a frame_dedup handle");
        DQN_REQUIRE(oh == h->net.H && ow == h->net.W, "preprocessed frame size differs from the model input");
    }
    const int64_t max_chunk = std::max<int64_t>(1, (int64_t)(h->dev_stage_bytes / (in_b + out_b)));
    for (int64_t done = 0; done < n;) {
        const int64_t m = std::min(n - done, max_chunk);
        uint8_t* d_in = h->dev_stage;
        uint8_t* d_out = h->dev_stage + (size_t)m * in_b;
        DQN_CUDA_OK(cudaMemcpyAsync(d_in, h_obs + (size_t)done * in_b, (size_t)m * in_b, cudaMemcpyHostToDevice, h->stream));
        launch_preprocess(h, d_in, raw_h, raw_w, top, stride, oh, ow, m, d_out);
        if (to_frame_ring) {
            for (int64_t j = 0; j < m; ++j) {  // ring slots may wrap: one copy per frame
                const int64_t slot = (h->f_written + j) % h->fcap;
                if (done == 0 && j == 0 && first_slot) *first_slot = slot;
                DQN_CUDA_OK(cudaMemcpyAsync(h->r_frames + (size_t)slot * h->frame_bytes, d_out + (size_t)j * out_b, out_b, cudaMemcpyDeviceToDevice, h->stream));
            }
            h->f_written += m;
        }
        if (h_frames) DQN_CUDA_OK(cudaMemcpyAsync(h_frames + (size_t)done * out_b, d_out, (size_t)m * out_b, cudaMemcpyDeviceToHost, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        done += m;
    }
    API_END(h)
}

extern "C" int dqn_replay_read(dqn_learner* h, int64_t n, const int64_t* rows, uint8_t* st, uint8_t* ac, uint8_t* rw,
                               uint8_t* ns, uint8_t* ct, uint16_t* ls) {
    API_BEGIN(h)
    if (h->dedup && (st || ns)) {
        // stacked states are assembled from the frame ring on the device (the gather the training step uses), max_rows at a time
        const size_t S = h->state_bytes;
        for (int64_t done = 0; done < n;) {
            const int m = (int)std::min<int64_t>(n - done, h->max_rows);
            for (int j = 0; j < m; ++j) DQN_REQUIRE(rows[done + j] >= 0 && rows[done + j] < h->cap, "row out of range");
            DQN_CUDA_OK(cudaMemcpyAsync(h->b_rows, rows + done, (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
            if (st) {
                launch_gather_states(h, h->b_rows, m, false, h->in_states);
                DQN_CUDA_OK(cudaMemcpyAsync(st + (size_t)done * S, h->in_states, (size_t)m * S, cudaMemcpyDeviceToHost, h->stream));
            }
            if (ns) {
                launch_gather_states(h, h->b_rows, m, true, h->in_states + (size_t)h->max_rows * S);
                DQN_CUDA_OK(cudaMemcpyAsync(ns + (size_t)done * S, h->in_states + (size_t)h->max_rows * S, (size_t)m * S, cudaMemcpyDeviceToHost, h->stream));
            }
            DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
            done += m;
        }
        st = nullptr; ns = nullptr;
    }
    for (int64_t j = 0; j < n; ++j) {
        const int64_t r = rows[j];
        DQN_REQUIRE(r >= 0 && r < h->cap, "row out of range");
        const size_t S = h->state_bytes;
        if (st) DQN_CUDA_OK(cudaMemcpyAsync(st + j * S, h->r_states + r * S, S, cudaMemcpyDeviceToHost, h->stream));
        if (ns) DQN_CUDA_OK(cudaMemcpyAsync(ns + j * S, h->r_next + r * S, S, cudaMemcpyDeviceToHost, h->stream));
        if (ac) DQN_CUDA_OK(cudaMemcpyAsync(ac + j, h->r_actions + r, 1, cudaMemcpyDeviceToHost, h->stream));
        if (rw) DQN_CUDA_OK(cudaMemcpyAsync(rw + j, h->r_rewards + r, 1, cudaMemcpyDeviceToHost, h->stream));
        if (ct) DQN_CUDA_OK(cudaMemcpyAsync(ct + j, h->r_cont + r, 1, cudaMemcpyDeviceToHost, h->stream));
        if (ls) DQN_CUDA_OK(cudaMemcpyAsync(ls + j, h->r_losses + r, 2, cudaMemcpyDeviceToHost, h->stream));
    }
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_replay_info(dqn_learner* h, int64_t* size, int64_t* cur, int32_t* full) {
    if (!h) return 2;
    if (size) *size = h->ring_size;
    if (cur) *cur = h->ring_cur;
    if (full) *full = h->ring_full ? 1 : 0;
    return 0;
}

extern "C" int dqn_replay_clear(dqn_learner* h) {
    API_BEGIN(h)
    h->ring_cur = 0; h->ring_full = false; h->ring_size = 0;
    push_sizes(h);  // (the frame ring of the frame_dedup layout keeps running: slots are reused by age, not by row)
    API_END(h)
}

extern "C" int dqn_replay_fill_synthetic(dqn_learner* h, int64_t n, uint64_t seed) {
    API_BEGIN(h)
    DQN_REQUIRE(n >= 1 && n <= h->cap && h->ring_size == 0, "synthetic fill needs an empty ring and n <= capacity");
    const int64_t chunk = (int64_t)(h->dev_stage_bytes / 4);
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
        const int64_t m = std::min(chunk, n - s0);
        float* d_prio = (float*)h->dev_stage;
        launch_fill_synthetic(h, s0, m, seed, d_prio);
        if (h->cfg.use_per) {
            launch_tree_add(h, m, d_prio, nullptr, s0);
            h->tree_size = s0 + m;
            h->tree_write = (s0 + m) % h->cap;
            if (s0 + m >= h->cap) h->tree_max_reached = true;
        }
    }
    h->ring_cur = n % h->cap; h->ring_full = n >= h->cap; h->ring_size = n;
    push_sizes(h);
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

// ---- sum tree ------------------------------------------------------------------------------------------------
extern "C" int dqn_tree_add(dqn_learner* h, int64_t n, const float* scores, const uint32_t* data) {
    API_BEGIN(h)
    const int64_t max_chunk = std::min<int64_t>(h->cap, (int64_t)(h->dev_stage_bytes / 8));
    for (int64_t done = 0; done < n;) {
        const int64_t m = std::min(n - done, max_chunk);
        float* d_s = (float*)h->dev_stage;
        uint32_t* d_d = (uint32_t*)(h->dev_stage + h->dev_stage_bytes / 2);
        DQN_CUDA_OK(cudaMemcpyAsync(d_s, scores + done, m * 4, cudaMemcpyHostToDevice, h->stream));
        if (data) DQN_CUDA_OK(cudaMemcpyAsync(d_d, data + done, m * 4, cudaMemcpyHostToDevice, h->stream));
        launch_tree_add(h, m, d_s, data ? d_d : nullptr, h->tree_write);
        if (h->tree_write + m >= h->cap) h->tree_max_reached = true;
        h->tree_size = h->tree_max_reached ? h->cap : h->tree_write + m;
        h->tree_write = (h->tree_write + m) % h->cap;
        done += m;
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    push_sizes(h);
    API_END(h)
}

extern "C" int dqn_tree_update(dqn_learner* h, int64_t n, const int64_t* idx, const float* scores, int32_t store_losses) {
    API_BEGIN(h)
    for (int64_t j = 0; j < n; ++j) DQN_REQUIRE(idx[j] >= h->cap - 1 && idx[j] < 2 * h->cap - 1, "tree_idx is not a leaf");
    if (n <= h->max_rows) {
        // the per-step write-back (B rows): indices and scores ride one pinned staging block, no host sync -- the
        // update is ordered on the handle's stream like every later call
        DQN_CUDA_OK(cudaEventSynchronize(h->stage_ev));
        uint8_t* hs = h->h_in_small;
        memcpy(hs, idx, (size_t)n * 8);
        memcpy(hs + (size_t)n * 8, scores, (size_t)n * 4);
        // the kernel reads both columns from the pinned, device-mapped block over PCIe (no copy operation in front of it)
        launch_tree_update(h, (int)n, reinterpret_cast<const long long*>(hs), reinterpret_cast<const float*>(hs + (size_t)n * 8),
                           store_losses);
        DQN_CUDA_OK(cudaEventRecord(h->stage_ev, h->stream));
        return 0;
    }
    const int64_t max_chunk = (int64_t)(h->dev_stage_bytes / 16);
    for (int64_t done = 0; done < n;) {
        const int64_t m = std::min(n - done, max_chunk);
        long long* d_i = (long long*)h->dev_stage;
        float* d_s = (float*)(h->dev_stage + h->dev_stage_bytes / 2);
        DQN_CUDA_OK(cudaMemcpyAsync(d_i, idx + done, m * 8, cudaMemcpyHostToDevice, h->stream));
        DQN_CUDA_OK(cudaMemcpyAsync(d_s, scores + done, m * 4, cudaMemcpyHostToDevice, h->stream));
        launch_tree_update(h, (int)m, d_i, d_s, store_losses);
        done += m;
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    API_END(h)
}

template <class T>
static void d2h(dqn_learner* h, T* dst, const T* src, size_t n) {
    if (dst) DQN_CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
}

static void tree_probe(dqn_learner* h, int32_t batch, const double* u, bool raw, int64_t* t_idx, int64_t* d_idx, double* prio,
                       int64_t* m_idx) {
    DQN_REQUIRE(batch >= 1 && batch <= h->max_rows, "batch exceeds max_rows");
    DQN_REQUIRE(u != nullptr, "draws are required");
    DQN_REQUIRE(h->tree_size >= 1, "sum tree is empty");
    DQN_CUDA_OK(cudaMemcpyAsync(h->b_u, u, batch * 8, cudaMemcpyHostToDevice, h->stream));
    const int saved = h->cfg.use_per;
    h->cfg.use_per = 0;  // indices only: no IS weights
    try { launch_tree_sample(h, batch, 0, -1.0, nullptr, nullptr, raw); } catch (...) { h->cfg.use_per = saved; throw; }
    h->cfg.use_per = saved;
    d2h(h, (long long*)t_idx, h->b_tree_idx, batch); d2h(h, (long long*)d_idx, h->b_data_idx, batch);
    d2h(h, prio, h->b_prio, batch); d2h(h, (long long*)m_idx, h->b_mem_idx, batch);
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
}

extern "C" int dqn_tree_sample(dqn_learner* h, int32_t batch, const double* u, int64_t* t_idx, int64_t* d_idx,
                               double* prio, int64_t* m_idx) {
    API_BEGIN(h)
    tree_probe(h, batch, u, false, t_idx, d_idx, prio, m_idx);
    API_END(h)
}

extern "C" int dqn_tree_retrieve(dqn_learner* h, int32_t n, const double* scores, int64_t* t_idx, int64_t* d_idx,
                                 double* prio, int64_t* m_idx) {
    API_BEGIN(h)
    tree_probe(h, n, scores, true, t_idx, d_idx, prio, m_idx);
    API_END(h)
}

extern "C" int dqn_tree_stats(dqn_learner* h, float* mn, float* mx, float* avg, float* total) {
    API_BEGIN(h)
    // read-only: results land in a scratch block, the PER normaliser / report statistics keep their own cadence
    float* d_out = reinterpret_cast<float*>(h->stats_scratch + 8192);  // behind the 296 partials
    launch_tree_stats(h, -1.0, d_out);
    float v[4];
    DQN_CUDA_OK(cudaMemcpyAsync(v, d_out, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (mn) *mn = v[0];
    if (mx) *mx = v[1];
    if (avg) *avg = v[2];
    if (total) *total = v[3];
    API_END(h)
}

extern "C" int dqn_tree_read(dqn_learner* h, float* tree, uint32_t* data, int64_t* size, int64_t* write) {
    API_BEGIN(h)
    d2h(h, tree, h->tree, (size_t)(2 * h->cap - 1));
    d2h(h, data, h->tree_data, (size_t)h->cap);
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (size) *size = h->tree_size;
    if (write) *write = h->tree_write;
    API_END(h)
}

extern "C" int dqn_tree_info(dqn_learner* h, int64_t* size, int64_t* write, float* total) {
    API_BEGIN(h)
    if (size) *size = h->tree_size;
    if (write) *write = h->tree_write;
    if (total) {
        DQN_CUDA_OK(cudaMemcpyAsync(total, h->tree, 4, cudaMemcpyDeviceToHost, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    API_END(h)
}

extern "C" int dqn_tree_get(dqn_learner* h, int64_t n, const int64_t* idx, float* scores, uint32_t* data) {
    API_BEGIN(h)
    for (int64_t j = 0; j < n; ++j) {
        DQN_REQUIRE(idx[j] >= 0 && idx[j] < 2 * h->cap - 1, "tree_idx out of range");
        if (scores) DQN_CUDA_OK(cudaMemcpyAsync(scores + j, h->tree + idx[j], 4, cudaMemcpyDeviceToHost, h->stream));
        if (data) {
            const int64_t d = idx[j] - (h->cap - 1);
            if (d >= 0) DQN_CUDA_OK(cudaMemcpyAsync(data + j, h->tree_data + d, 4, cudaMemcpyDeviceToHost, h->stream));
            else data[j] = 0;
        }
    }
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

// ---- samplers --------------------------------------------------------------------------------------------------
extern "C" int dqn_per_sample(dqn_learner* h, const double* u, int32_t refresh, double per_b, int64_t* t_idx,
                              double* prio, double* isw) {
    API_BEGIN(h)
    DQN_REQUIRE(h->cfg.use_per, "handle was created without use_per");
    DQN_REQUIRE(u != nullptr, "uniform draws are required");
    if (refresh) launch_tree_stats(h, per_b);
    // uniforms in and {tree_idx, prio, is_w} out ride one pinned, device-mapped staging block: the sampling kernel reads
    // the draws from it and writes the three columns into it over PCIe -- no copy operations on the stream
    const size_t R = (size_t)h->max_rows, B = (size_t)h->B;
    uint8_t* hs = h->h_stage2;
    memcpy(hs, u, B * 8);
    launch_tree_sample(h, h->B, 0, per_b, reinterpret_cast<const double*>(hs), hs + 8 * R);
    DQN_CUDA_OK(cudaEventRecord(h->stage2_ev, h->stream));
    launch_gather(h, h->b_mem_idx, h->B);  // stream-ordered behind the read-back: the host does not wait for it
    DQN_CUDA_OK(cudaEventSynchronize(h->stage2_ev));
    if (t_idx) memcpy(t_idx, hs + 8 * R, B * 8);
    if (prio) memcpy(prio, hs + 16 * R, B * 8);
    if (isw) memcpy(isw, hs + 24 * R, B * 8);
    API_END(h)
}

extern "C" int dqn_uniform_sample(dqn_learner* h, const int64_t* rows) {
    API_BEGIN(h)
    for (int i = 0; i < h->B; ++i) DQN_REQUIRE(rows[i] >= 0 && rows[i] < h->cap, "row out of range");
    DQN_CUDA_OK(cudaMemcpyAsync(h->b_rows, rows, h->B * 8, cudaMemcpyHostToDevice, h->stream));
    launch_gather(h, h->b_rows, h->B);
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_batch_read(dqn_learner* h, uint8_t* st, uint8_t* ac, uint8_t* rw, uint8_t* ns, uint8_t* ct, uint16_t* ls) {
    API_BEGIN(h)
    const size_t S = h->state_bytes, B = h->B;
    d2h(h, st, h->b_states, B * S); d2h(h, ns, h->b_states + B * S, B * S);
    d2h(h, ac, h->b_actions, B); d2h(h, rw, h->b_rewards, B); d2h(h, ct, h->b_cont, B);
    d2h(h, (__half*)ls, h->b_losses, B);
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

// ---- the training step ---------------------------------------------------------------------------------------
bool overlap_enabled(const dqn_learner* h) { return h->overlap && !h->profile; }

static cudaEvent_t next_fork_event(dqn_learner* h) {
    DQN_REQUIRE(h->fork_used < 21, "fork event pool exhausted");  // [21]: dense dgrad done (nn.cu); [22], [23] belong to upload_batch
    return h->fork_ev[h->fork_used++];
}

static cudaStream_t side_of(const dqn_learner* h, int which) {
    switch (which) {
        case 0: return h->side_stream;
        case 1: return h->side2_stream;
        case 2: return h->side3_stream;
        case 3: return h->side4_stream;
        case 4: return h->side5_stream;
        default: return h->side6_stream;
    }
}

static int prio_of(const dqn_learner* h, int prio_class) {
    if (!h->use_prio) return h->prio_lo;
    const int p = prio_class >= 2 ? h->prio_hi : prio_class == 1 ? h->prio_hi + 1 : h->prio_lo;
    return p > h->prio_lo ? h->prio_lo : p;
}

void side_begin(dqn_learner* h, cudaStream_t& saved, int which, int prio_class) {
    saved = h->stream;
    h->saved_prio = g_launch_priority;
    if (!overlap_enabled(h)) return;
    g_launch_priority = prio_of(h, prio_class);
    cudaStream_t side = side_of(h, which);
    cudaEvent_t e = next_fork_event(h);
    DQN_CUDA_OK(cudaEventRecord(e, saved));
    DQN_CUDA_OK(cudaStreamWaitEvent(side, e, 0));
    h->stream = side;
}

void side_end(dqn_learner* h, cudaStream_t saved) { h->stream = saved; g_launch_priority = h->saved_prio; }

cudaEvent_t fork_event_here(dqn_learner* h) {
    if (!overlap_enabled(h)) return nullptr;
    cudaEvent_t e = next_fork_event(h);
    DQN_CUDA_OK(cudaEventRecord(e, h->stream));
    return e;
}

void side_begin_at(dqn_learner* h, cudaStream_t& saved, int which, int prio_class, cudaEvent_t after) {
    if (!after) { side_begin(h, saved, which, prio_class); return; }
    saved = h->stream;
    h->saved_prio = g_launch_priority;
    if (!overlap_enabled(h)) return;
    g_launch_priority = prio_of(h, prio_class);
    cudaStream_t side = side_of(h, which);
    DQN_CUDA_OK(cudaStreamWaitEvent(side, after, 0));
    h->stream = side;
}

void main_wait_side(dqn_learner* h, int which) {
    if (!overlap_enabled(h)) return;
    cudaEvent_t e = next_fork_event(h);
    DQN_CUDA_OK(cudaEventRecord(e, side_of(h, which)));
    DQN_CUDA_OK(cudaStreamWaitEvent(h->stream, e, 0));
}

void prof_mark(dqn_learner* h, const char* name) {
    if (!h->profile) return;
    if (h->prof_used >= h->prof_events.size()) {
        cudaEvent_t e; DQN_CUDA_OK(cudaEventCreate(&e));
        h->prof_events.push_back(e);
    }
    // inside the profiled graph the marks become event-record nodes, so the intervals are device-side times
    // between kernels with no host launch latency in them
    DQN_CUDA_OK(cudaEventRecordWithFlags(h->prof_events[h->prof_used], h->stream,
                                         h->prof_capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    h->prof_marks.push_back(name);
    h->prof_used++;
}

void select_set(dqn_learner* h, int k) {
    const dqn_learner::BatchSet& a = h->sets[k];
    h->b_states = a.states; h->b_actions = a.actions; h->b_rewards = a.rewards; h->b_cont = a.cont; h->b_losses = a.losses;
    h->b_tree_idx = a.tree_idx; h->b_data_idx = a.data_idx; h->b_mem_idx = a.mem_idx; h->b_rows = a.rows;
    h->b_prio = a.prio; h->b_isw64 = a.isw64; h->b_u = a.u; h->b_isw = a.isw; h->x_f32 = a.x_f32;
    h->cur_set = k;
}
static void sample_into_batch(dqn_learner* h, const double* u, const int64_t* rows);

// forward both towers, TD epilogue and (train) backward on a [2n,S] batch laid out as rows [0,n)=s, [n,2n)=s'
static void step_core(dqn_learner* h, int n, const uint8_t* states2, const uint8_t* actions, const uint8_t* rewards,
                      const uint8_t* cont, const float* isw, int train, double grad_scale, int per_update = 0) {
    const size_t S = h->state_bytes;
    const int A = h->net.A;
    const float* xf = nullptr;
    g_launch_priority = prio_of(h, 2);  // everything issued on the main stream of a step is the critical chain
    // the fused conv tower (cvtower.cuh) and the image-resident conv1 wgrad read the uint8 batch themselves: no f32 copy
    const bool fused_tower = tower_fused_ready(h);
    if (h->cfg.math_mode != DQN_MATH_FP32) {
        bool have = h->xf_from_gather && states2 == h->b_states;  // sampled batch: the gather may have written it
        if (!have && !fused_tower) {                               // host-fed batch: convert here
            prof_mark(h, "u8_to_f32");
            launch_u8_to_f32(h, states2, h->x_f32, 2 * (size_t)n * S);
            have = true;
        }
        h->xf_from_gather = false;
        h->xf_valid = have;
        xf = h->x_f32;  // contents valid only if xf_valid (tower_backward converts lazily if a fallback kernel needs them)
    }
    h->fork_used = 0;
    if (h->defer_start) {
        // deferred from the previous step: optimizer pass over the two dense kernels (their gradients are in the slab,
        // the beta powers of that step in b?p_prev), under this step's conv forward; joined before the dense forward
        cudaStream_t sv;
        side_begin(h, sv, 2, 0); launch_optimizer_fc_stream(h, 0, true); side_end(h, sv);
        side_begin(h, sv, 5, 0); launch_optimizer_fc_stream(h, 1, true); side_end(h, sv);
    }
    if (fused_tower) launch_dense_prefetch(h, true);  // dense kernels -> L2 while the conv towers run
    // the two towers are independent until the TD epilogue: target tower on the side stream
    cudaStream_t saved;
    const int rows_on = h->cfg.use_double ? 2 * n : n;  // double DQN: s and s' share one pass of the online tower
    if ((fused_tower && towers_conv_forward_fused(h, h->online, states2, rows_on, h->acts_on, states2 + (size_t)n * S, n, train ? n : -1)) ||
        (xf && h->xf_valid && towers_conv_forward_merged(h, xf, rows_on, xf + (size_t)n * S, n))) {
        side_begin(h, saved, 0, 2);
        h->prof_tag = "tg_";
        tower_fc_forward(h, h->target, n, h->acts_tg);
        side_end(h, saved);
        h->prof_tag = "on_";
        tower_fc_forward(h, h->online, rows_on, h->acts_on);
    } else {
        if (xf && !h->xf_valid) { launch_u8_to_f32(h, states2, h->x_f32, 2 * (size_t)n * S); h->xf_valid = true; }
        side_begin(h, saved, 0, 2);
        h->prof_tag = "tg_";
        tower_forward(h, h->target, states2 + (size_t)n * S, xf ? xf + (size_t)n * S : nullptr, n, h->acts_tg);
        side_end(h, saved);
        h->prof_tag = "on_";
        tower_forward(h, h->online, states2, xf, rows_on, h->acts_on);
    }
    main_wait_side(h);
    if (fused_tower && h->fc_prefetch) main_wait_side(h, 6);  // (join of the prefetch branch; it ended long ago)
    const bool fuse_td = train && h->fuse_td;
    TdFuse tf{h->acts_on.q, h->acts_on.q + (size_t)n * A, h->acts_tg.q, rewards, cont, isw, grad_scale};
    if (fuse_td) {
        // the TD epilogue rides in the head backward kernel: one launch less on the critical chain
        prof_mark(h, "head_backward");
        launch_head_backward(h, n, actions, &tf);
    } else {
        prof_mark(h, "td_epilogue");
        launch_td_epilogue(h, n, h->acts_on.q, h->acts_on.q + (size_t)n * A, h->acts_tg.q, actions, rewards, cont, isw, train,
                           grad_scale);
    }
    if (train && (per_update || h->pf_capture)) {
        // replay_sampler.update_sum_tree(tree_idxes, _make_priority(losses)) (deep_q_agent.py:341-347): it only needs
        // the TD errors, so it runs on a side stream underneath the whole backward pass ...
        cudaStream_t sv;
        const int which = h->pf_capture ? 2 : 0;
        side_begin(h, sv, which, 1);
        if (per_update) {
            prof_mark(h, "tree_update");
            launch_tree_update(h, n, h->b_tree_idx, h->b_newprio, 1);
        }
        if (h->pf_capture) {
            // ... and the NEXT step's batch is sampled from the updated tree and gathered into the other buffer set
            // right behind it (counters read with +1: the optimizer kernel bumps them at the end of this step)
            const int cur = h->cur_set;
            select_set(h, cur ^ 1);
            h->sample_ahead = 1;
            try { sample_into_batch(h, nullptr, nullptr); } catch (...) { h->sample_ahead = 0; select_set(h, cur); throw; }
            h->sample_ahead = 0;
            select_set(h, cur);
        }
        side_end(h, sv);
    }
    if (train) {
        if (!fuse_td) {
            prof_mark(h, "head_backward");
            launch_head_backward(h, n, actions);
        }
        tower_backward(h, states2, xf, n);
        if (h->pf_capture) main_wait_side(h, 2);  // the look-ahead batch is complete before the counters are bumped
    }
}

static void finish_step(dqn_learner* h, int per_update) {
    prof_mark(h, "optimizer");
    const bool joined = h->rows_joined;  // (tail_join) the fused dense kernel is still running on side stream 0: the tail starts beside it ...
    launch_optimizer(h, h->fuse_fc_adam || h->fc_adam_done);
    h->fc_adam_done = false;
    if (joined) main_wait_side(h, 0);    // ... and the step ends when both are done
    prof_mark(h, "post_step");
    launch_post_step(h, per_update);
    prof_mark(h, "end");
    h->host_step++;
}

static bool maybe_refresh_stats(dqn_learner* h);
static void maybe_refresh_stats_noprof(dqn_learner* h) {
    const int p = h->profile; h->profile = 0;
    maybe_refresh_stats(h);
    h->profile = p;
}
static bool maybe_refresh_stats(dqn_learner* h) {
    // deep_q_agent.py:505: step % per_calculate_steps == 0 or last_max_weight is None
    // (external_normaliser: data-parallel ranks install the maximum over all shards themselves, dqn_set_per_normaliser)
    if (h->cfg.use_per && !h->external_normaliser && (!h->has_max_weight || h->host_step % h->cfg.per_calculate_steps == 0)) {
        launch_tree_stats(h, -1.0);
        h->has_max_weight = true;
        return true;
    }
    return false;
}

__global__ void uniform_rows_kernel(int batch, const long long* p_len, uint64_t seed, dqn_learner::Scalars* sc,
                                    long long* rows, int ahead) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    // replay_sampler.py:27-30: size = len/B ; randint(int(i*size), int((i+1)*size - 1)) inclusive
    const double size = (double)(*p_len) / (double)batch;
    const long long lo = (long long)((double)i * size), hi = (long long)((double)(i + 1) * size - 1.0);
    uint32_t r[4];
    const unsigned long long ctr = sc->rng_counter + (unsigned long long)ahead;
    philox4x32((uint32_t)i, 1u, (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
    long long span = hi - lo + 1;
    if (span < 1) span = 1;
    long long off = (long long)(u53(r[0], r[1]) * (double)span);
    if (off >= span) off = span - 1;
    rows[i] = lo + off;
}

// sample + gather into the device batch. u / rows: host-provided draws for parity, NULL -> device Philox.
static void sample_into_batch(dqn_learner* h, const double* u, const int64_t* rows) {
    prof_mark(h, "sample");
    if (h->cfg.use_per) {
        if (u) DQN_CUDA_OK(cudaMemcpyAsync(h->b_u, u, h->B * 8, cudaMemcpyHostToDevice, h->stream));
        launch_tree_sample(h, h->B, u ? 0 : 1, -1.0);
        prof_mark(h, "gather");
        launch_gather(h, h->b_mem_idx, h->B, true);
    } else {
        if (rows) DQN_CUDA_OK(cudaMemcpyAsync(h->b_rows, rows, h->B * 8, cudaMemcpyHostToDevice, h->stream));
        else {
            launch_kernel(false, uniform_rows_kernel, dim3(ceil_div(h->B, 128)), dim3(128), 0, h->stream, h->B, (const long long*)&h->d_dev->ring_len, (uint64_t)h->cfg.seed, h->d_sc, h->b_rows, h->sample_ahead);
            h->launches++;
        }
        prof_mark(h, "gather");
        launch_gather(h, h->b_rows, h->B, true);
    }
}

static void fused_step_launches(dqn_learner* h, const double* u, const int64_t* rows, double grad_scale, int phases) {
    // whole step in one call: the dense wgrad carries the optimizer (no gradient round trip through HBM);
    // the split (data-parallel) form keeps the gradient slab for the external all-reduce
    h->fuse_fc_adam = (phases == 3) && h->fuse_enabled;
    h->early_fc_adam = (phases == 3) && !h->fuse_fc_adam;
    if (phases & 1) {
        if (!(phases & 4)) sample_into_batch(h, u, rows);  // bit 2: train on the batch already in the device buffers
        step_core(h, h->B, h->b_states, h->b_actions, h->b_rewards, h->b_cont, h->cfg.use_per ? h->b_isw : nullptr, 1, grad_scale,
                  h->cfg.use_per);
    }
    if (phases & 2) finish_step(h, 0);
    h->fuse_fc_adam = false; h->early_fc_adam = false;
}

// apply the optimizer pass over the dense kernels that the last pipelined step left pending
static void flush_fc_adam(dqn_learner* h) {
    if (!h->fc_adam_pending) return;
    launch_optimizer_fc_stream(h, 0, true);
    launch_optimizer_fc_stream(h, 1, true);
    h->fc_adam_pending = false;
}

static void after_step_host(dqn_learner* h) {
    // deep_q_agent.py:358-360: copy online -> target every copy_network_steps (step is the post-increment value)
    if (h->cfg.copy_network_steps > 0 && h->host_step % h->cfg.copy_network_steps == 0) launch_copy_network(h);
}


// The drop-in agent's call (DeepQAgent._train_run_train) with HOST draws: ~35 launches replayed as ONE graph.  The draws are
// copied into a pinned staging block with a fixed address (first node of the graph: H2D of that block); scalars and per-sample
// losses are written into their pinned mirror by the step's last kernel, so a caller that wants them pays one event wait and no
// copy operation.  Two slots (staging block, mirror, graph, event each): dqn_train_step_async alternates them, so that the host
// can queue step i+1 while step i runs and read step i's results afterwards (dqn_train_step_result).
static void launch_ustep(dqn_learner* h, int slot, const double* u, const int64_t* rows) {
    const size_t readback = SC_BLOCK + (size_t)h->B * 4;
    uint8_t* stage = slot ? h->h_stage2_alt : h->h_stage2;
    dqn_learner::Scalars* mirror = slot ? h->h_sc_alt : h->h_sc;
    cudaGraphExec_t& gx = slot ? h->ustep_graph_alt : h->ustep_graph;
    bool& valid = slot ? h->graph_valid_u_alt : h->graph_valid_u;
    DQN_CUDA_OK(cudaEventSynchronize(h->ustep_ev[slot]));  // the slot's previous launch is complete: block and mirror are free
    memcpy(stage, u ? (const void*)u : (const void*)rows, (size_t)h->B * 8);
    select_set(h, 0);
    if (!gx || !valid) {
        cudaGraph_t g = nullptr;
        const int64_t l0 = h->launches, s0 = h->host_step;
        DQN_CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->tail_mirror = reinterpret_cast<float*>(mirror); h->tail_mirror_words = (int)(readback / 4);
        try {
            fused_step_launches(h, u ? reinterpret_cast<const double*>(stage) : nullptr,
                                rows ? reinterpret_cast<const int64_t*>(stage) : nullptr, 1.0, 3);
            h->tail_mirror = nullptr; h->tail_mirror_words = 0;
        } catch (...) {
            h->tail_mirror = nullptr; h->tail_mirror_words = 0;
            cudaStreamEndCapture(h->stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        DQN_CUDA_OK(cudaStreamEndCapture(h->stream, &g));
        h->ustep_graph_launches = h->launches - l0;
        h->launches = l0; h->host_step = s0;
        if (gx) { cudaGraphExecDestroy(gx); gx = nullptr; }
        DQN_CUDA_OK(cudaGraphInstantiateWithFlags(&gx, g, cudaGraphInstantiateFlagUseNodePriority));
        DQN_CUDA_OK(cudaGraphDestroy(g));
        valid = true;
    }
    DQN_CUDA_OK(cudaGraphLaunch(gx, h->stream));
    DQN_CUDA_OK(cudaEventRecord(h->ustep_ev[slot], h->stream));
    h->launches += h->ustep_graph_launches;
    h->host_step++;
    after_step_host(h);
}

static void drain_ustep(dqn_learner* h) { h->ustep_pending = 0; h->ustep_slot = 0; h->ustep_read = 0; }

extern "C" int dqn_train_step(dqn_learner* h, const double* u, const int64_t* rows, int64_t* step, float* losses, float* loss) {
    API_BEGIN(h)
    DQN_REQUIRE(h->ring_size >= 1, "replay is empty");
    // the draws must match the sampler this handle was created with (a mismatched argument used to be ignored silently)
    DQN_REQUIRE(!(u && !h->cfg.use_per), "uniform draws were given but the handle samples uniformly (pass h_rows)");
    DQN_REQUIRE(!(rows && h->cfg.use_per), "rows were given but the handle samples by priority (pass h_u)");
    if (u) for (int i = 0; i < h->B; ++i) DQN_REQUIRE(u[i] >= 0.0 && u[i] < 1.0, "uniform draw outside [0,1)");
    if (rows) for (int i = 0; i < h->B; ++i) DQN_REQUIRE(rows[i] >= 0 && rows[i] < h->ring_size, "row out of range");
    h->prof_used = 0; h->prof_marks.clear();
    maybe_refresh_stats(h);
    const bool host_draws = u != nullptr || rows != nullptr;
    if (host_draws && !h->profile && h->batch_graph_enabled) {
        drain_ustep(h);  // (results of asynchronous steps nobody read are dropped)
        launch_ustep(h, 0, u, rows);
        if (step || losses || loss) {
            DQN_CUDA_OK(cudaEventSynchronize(h->ustep_ev[0]));
            if (losses) memcpy(losses, reinterpret_cast<const char*>(h->h_sc) + SC_BLOCK, (size_t)h->B * 4);
            if (step) *step = h->h_sc->step;
            if (loss) *loss = h->h_sc->loss;
        }
        return 0;
    }
    fused_step_launches(h, u, rows, 1.0, 3);
    after_step_host(h);
    if (step || losses || loss) {
        d2h(h, losses, h->b_losses_out, h->B);
        read_scalars(h);
        if (step) *step = h->h_sc->step;
        if (loss) *loss = h->h_sc->loss;
    }
    API_END(h)
}

extern "C" int dqn_train_step_async(dqn_learner* h, const double* u, const int64_t* rows) {
    API_BEGIN(h)
    DQN_REQUIRE(h->ring_size >= 1, "replay is empty");
    DQN_REQUIRE((u != nullptr) != (rows != nullptr), "dqn_train_step_async: host draws -- uniforms (PER) or rows (uniform sampler)");
    DQN_REQUIRE((u != nullptr) == (h->cfg.use_per != 0), "dqn_train_step_async: uniforms with use_per, rows without");
    DQN_REQUIRE(!h->profile && h->batch_graph_enabled, "dqn_train_step_async needs the graph path (no profile mode)");
    DQN_REQUIRE(h->ustep_pending < 2, "dqn_train_step_async: two steps are in flight, read a result first");
    if (rows) for (int i = 0; i < h->B; ++i) DQN_REQUIRE(rows[i] >= 0 && rows[i] < h->ring_size, "row index out of range");
    maybe_refresh_stats(h);
    launch_ustep(h, h->ustep_slot, u, rows);
    h->ustep_slot ^= 1;
    h->ustep_pending++;
    API_END(h)
}

extern "C" int dqn_train_step_result(dqn_learner* h, int64_t* step, float* losses, float* loss) {
    API_BEGIN_KEEP(h)
    DQN_REQUIRE(h->ustep_pending > 0, "dqn_train_step_result: no asynchronous step in flight");
    const int slot = h->ustep_read;
    DQN_CUDA_OK(cudaEventSynchronize(h->ustep_ev[slot]));
    const dqn_learner::Scalars* mirror = slot ? h->h_sc_alt : h->h_sc;
    if (losses) memcpy(losses, reinterpret_cast<const char*>(mirror) + SC_BLOCK, (size_t)h->B * 4);
    if (step) *step = mirror->step;
    if (loss) *loss = mirror->loss;
    h->ustep_read ^= 1;
    h->ustep_pending--;
    API_END(h)
}

extern "C" int dqn_train_steps(dqn_learner* h, int64_t n) {
    API_BEGIN_KEEP(h)
    DQN_REQUIRE(h->ring_size >= 1, "replay is empty");
    if (h->profile || !h->prefetch) h->pf_valid = false;
    if (h->profile) {
        if (!h->prof_graph) {
            cudaGraph_t g = nullptr;
            const int64_t l0 = h->launches, s0 = h->host_step;
            h->prof_used = 0; h->prof_marks.clear();
            h->prof_capturing = true;
            DQN_CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            try {
                fused_step_launches(h, nullptr, nullptr, 1.0, 3);
            } catch (...) {
                h->prof_capturing = false;
                cudaStreamEndCapture(h->stream, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            h->prof_capturing = false;
            DQN_CUDA_OK(cudaStreamEndCapture(h->stream, &g));
            h->graph_launches = h->launches - l0;
            h->launches = l0; h->host_step = s0;
            DQN_CUDA_OK(cudaGraphInstantiate(&h->prof_graph, g, 0));
            DQN_CUDA_OK(cudaGraphDestroy(g));
        }
        for (int64_t i = 0; i < n; ++i) {
            maybe_refresh_stats_noprof(h);
            DQN_CUDA_OK(cudaGraphLaunch(h->prof_graph, h->stream));
            h->launches += h->graph_launches;
            h->host_step++;
            after_step_host(h);
            DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
            for (size_t s = 0; s + 1 < h->prof_used; ++s) {
                float ms = 0.f;
                DQN_CUDA_OK(cudaEventElapsedTime(&ms, h->prof_events[s], h->prof_events[s + 1]));
                bool found = false;
                for (auto& kv : h->prof_acc) if (kv.first == h->prof_marks[s]) { kv.second += ms; found = true; }
                if (!found) h->prof_acc.push_back({h->prof_marks[s], ms});
            }
            h->prof_steps++;
        }
        return 0;
    }
    if (h->prefetch) {
        // Pipelined replay: graph k consumes the batch in sets[k] and, behind its priority write-back, samples and
        // gathers the next batch into sets[1-k].  Same draws, same tree states, same results as the serial order.
        const bool can_defer = h->defer_adam && h->cfg.math_mode != DQN_MATH_FP32 && h->net.dueling && h->fc_split_streams &&
                               !h->fuse_enabled && !h->fc_ffma_adam && !h->fc_tma_adam;
        if (!h->graph_valid) {
            for (int kv = 0; kv < (can_defer ? 4 : 2); ++kv) {
                const int k = kv & 1, v = kv >> 1;
                cudaGraph_t g = nullptr;
                const int64_t l0 = h->launches, s0 = h->host_step;
                select_set(h, k);
                DQN_CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                h->pf_capture = true;
                h->defer_capture = can_defer; h->defer_start = v == 1;
                try {
                    h->fuse_fc_adam = h->fuse_enabled; h->early_fc_adam = !h->fuse_fc_adam;
                    h->xf_from_gather = !tower_fused_ready(h);  // the look-ahead gather of the previous step wrote x_f32 of this set
                    step_core(h, h->B, h->b_states, h->b_actions, h->b_rewards, h->b_cont, h->cfg.use_per ? h->b_isw : nullptr, 1,
                              1.0, h->cfg.use_per);
                    finish_step(h, 0);
                } catch (...) {
                    h->pf_capture = false; h->fuse_fc_adam = false; h->early_fc_adam = false;
                    h->defer_capture = false; h->defer_start = false;
                    cudaStreamEndCapture(h->stream, &g);
                    if (g) cudaGraphDestroy(g);
                    throw;
                }
                h->pf_capture = false; h->fuse_fc_adam = false; h->early_fc_adam = false;
                h->defer_capture = false; h->defer_start = false;
                DQN_CUDA_OK(cudaStreamEndCapture(h->stream, &g));
                if (v == 0) h->graph_launches = h->launches - l0;  // per step: [k][1] has two launches more, [k][0]/flush make up for them
                h->launches = l0; h->host_step = s0;
                if (h->pipe_graph[k][v]) { cudaGraphExecDestroy(h->pipe_graph[k][v]); h->pipe_graph[k][v] = nullptr; }
                DQN_CUDA_OK(cudaGraphInstantiateWithFlags(&h->pipe_graph[k][v], g, cudaGraphInstantiateFlagUseNodePriority));
                DQN_CUDA_OK(cudaGraphDestroy(g));
            }
            h->graph_valid = true;
            h->pf_valid = false;
        }
        for (int64_t i = 0; i < n; ++i) {
            const bool refreshed = maybe_refresh_stats(h);
            if (!h->pf_valid || refreshed) {
                // first step after anything else touched the handle, or new PER normalisers (deep_q_agent.py:505-510):
                // sample eagerly with the counters as they are (the look-ahead sample used the same draws)
                select_set(h, h->pf_set);
                sample_into_batch(h, nullptr, nullptr);
                h->pf_valid = true;
            }
            select_set(h, h->pf_set);  // the batch this step trains on (what dqn_batch_read shows afterwards)
            const int v = can_defer && h->fc_adam_pending ? 1 : 0;
            DQN_CUDA_OK(cudaGraphLaunch(h->pipe_graph[h->pf_set][v], h->stream));
            h->pf_set ^= 1;
            h->launches += h->graph_launches + (v ? 2 : 0);
            if (can_defer) h->fc_adam_pending = true;
            h->host_step++;
            // the target sync copies the online weights: they must be complete (deep_q_agent.py:358-360)
            if (h->cfg.copy_network_steps > 0 && h->host_step % h->cfg.copy_network_steps == 0) flush_fc_adam(h);
            after_step_host(h);
        }
        flush_fc_adam(h);  // the handle never leaves dqn_train_steps with a pending optimizer pass
        return 0;
    }
    if (!h->graph_valid) {
        // capture one fused step (device RNG) into a graph: ~30 short kernels replayed with a single launch
        cudaGraph_t g = nullptr;
        const int64_t l0 = h->launches, s0 = h->host_step;
        DQN_CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        try {
            fused_step_launches(h, nullptr, nullptr, 1.0, 3);
        } catch (...) {
            cudaStreamEndCapture(h->stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        DQN_CUDA_OK(cudaStreamEndCapture(h->stream, &g));
        h->graph_launches = h->launches - l0;
        h->launches = l0; h->host_step = s0;
        if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
        DQN_CUDA_OK(cudaGraphInstantiateWithFlags(&h->step_graph, g, cudaGraphInstantiateFlagUseNodePriority));
        DQN_CUDA_OK(cudaGraphDestroy(g));
        h->graph_valid = true;
    }
    for (int64_t i = 0; i < n; ++i) {
        maybe_refresh_stats(h);
        DQN_CUDA_OK(cudaGraphLaunch(h->step_graph, h->stream));
        h->launches += h->graph_launches;
        h->host_step++;
        after_step_host(h);
    }
    API_END(h)
}

extern "C" int dqn_train_step_phase(dqn_learner* h, int32_t phase, double grad_scale) {
    API_BEGIN(h)
    h->prof_used = 0; h->prof_marks.clear();
    if (phase == 0 || phase == 4) {
        DQN_REQUIRE(h->ring_size >= 1, "replay is empty");
        if (phase == 0) maybe_refresh_stats(h);
        h->phase_ev_armed = true;
        try {
            fused_step_launches(h, nullptr, nullptr, grad_scale, phase == 0 ? 1 : 5);
        } catch (...) { h->phase_ev_armed = false; throw; }
        h->phase_ev_armed = false;
        DQN_CUDA_OK(cudaEventRecord(h->phase_ev[1], h->stream));
    } else if (phase == 1) {
        // (the priority write-back of the local samples ran inside phase 0: it only needs the TD errors)
        fused_step_launches(h, nullptr, nullptr, grad_scale, 2);
        after_step_host(h);
    } else {
        throw std::string("phase must be 0 (sample + forward + backward), 4 (the same on the current device batch) or 1 (optimizer)");
    }
    API_END(h)
}

extern "C" int dqn_phase_wait(dqn_learner* h, const char* what, void* cuda_stream) {
    API_BEGIN_KEEP(h)
    const std::string w = what;
    DQN_REQUIRE(w == "fc_grads" || w == "grads", "dqn_phase_wait: what must be \"fc_grads\" or \"grads\"");
    DQN_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, h->phase_ev[w == "grads" ? 1 : 0], 0));
    API_END(h)
}

extern "C" int dqn_set_per_normaliser(dqn_learner* h, float last_max_weight) {
    API_BEGIN(h)
    DQN_REQUIRE(h->cfg.use_per && last_max_weight > 0.f, "dqn_set_per_normaliser needs use_per and a positive weight");
    set_f2_kernel<<<1, 1, 0, h->stream>>>(&h->d_sc->last_max_weight, last_max_weight, &h->d_sc->last_max_weight, last_max_weight);
    DQN_CUDA_OK(cudaGetLastError());
    h->has_max_weight = true;
    API_END(h)
}

static void upload_small(dqn_learner* h) {
    DQN_CUDA_OK(cudaMemcpyAsync(h->in_actions, h->h_in_small, h->in_small_bytes, cudaMemcpyHostToDevice, h->stream));
}

// copy_small = false: the staging block is filled but its H2D copy is left to the caller (a node of the replayed graph)
static void upload_batch(dqn_learner* h, int n, const uint8_t* st, const uint8_t* ac, const uint8_t* rw, const uint8_t* ct,
                         const uint8_t* ns, const float* isw, bool copy_small = true) {
    const size_t S = h->state_bytes;
    // s on the main stream, s' on a side stream: two copy engines share the upload
    DQN_CUDA_OK(cudaEventRecord(h->fork_ev[22], h->stream));
    DQN_CUDA_OK(cudaStreamWaitEvent(h->side_stream, h->fork_ev[22], 0));
    DQN_CUDA_OK(cudaMemcpyAsync(h->in_states + n * S, ns, n * S, cudaMemcpyHostToDevice, h->side_stream));
    DQN_CUDA_OK(cudaEventRecord(h->fork_ev[23], h->side_stream));
    DQN_CUDA_OK(cudaMemcpyAsync(h->in_states, st, n * S, cudaMemcpyHostToDevice, h->stream));
    DQN_CUDA_OK(cudaStreamWaitEvent(h->stream, h->fork_ev[23], 0));
    // actions / rewards / continues / is_weights: packed through the pinned staging block (mirrors the device block)
    DQN_CUDA_OK(cudaEventSynchronize(h->stage_ev));  // the previous upload has left the staging block
    uint8_t* hs = h->h_in_small;
    memcpy(hs + (h->in_actions - h->in_actions), ac, n);
    memcpy(hs + (h->in_rewards - h->in_actions), rw, n);
    memcpy(hs + (h->in_cont - h->in_actions), ct, n);
    if (isw) memcpy(hs + (reinterpret_cast<uint8_t*>(h->in_isw) - h->in_actions), isw, (size_t)n * 4);
    if (copy_small) upload_small(h);
    DQN_CUDA_OK(cudaEventRecord(h->stage_ev, h->stream));  // (graph path: the caller synchronises the stream before returning)
}

extern "C" int dqn_train_batch(dqn_learner* h, int32_t n, const uint8_t* st, const uint8_t* ac, const uint8_t* rw,
                               const uint8_t* ct, const uint8_t* ns, const float* isw, int64_t* step, float* losses, float* loss) {
    API_BEGIN(h)
    DQN_REQUIRE(n >= 1 && n <= h->max_rows, "n exceeds max_rows");
    DQN_REQUIRE(!h->cfg.use_per || isw, "use_per: is_weights are required");
    h->prof_used = 0; h->prof_marks.clear();
    const bool use_graph = !h->profile && h->batch_graph_enabled;
    upload_batch(h, n, st, ac, rw, ct, ns, h->cfg.use_per ? isw : nullptr, !use_graph);
    const size_t readback = SC_BLOCK + (size_t)n * 4;
    auto launches = [&]() {
        h->early_fc_adam = true; h->fc_grads_needed = h->keep_grads;  // option: keep the full gradient slab readable
        try {
            step_core(h, n, h->in_states, h->in_actions, h->in_rewards, h->in_cont, h->cfg.use_per ? h->in_isw : nullptr, 1, 1.0);
        } catch (...) { h->early_fc_adam = false; h->fc_grads_needed = false; throw; }
        h->early_fc_adam = false; h->fc_grads_needed = false;
        finish_step(h, 0);  // priorities are written back by the caller (sampler.update_sum_tree), as in the reference
    };
    if (!use_graph) {
        launches();
        DQN_CUDA_OK(cudaMemcpyAsync(h->h_sc, h->d_sc, readback, cudaMemcpyDeviceToHost, h->stream));
    } else {
        // the ~28 launches of one host-fed step replayed as one graph (inputs/outputs sit in fixed device buffers);
        // re-captured when the batch size or an option changes
        if (!h->batch_graph || h->batch_graph_n != n || !h->graph_valid_batch) {
            cudaGraph_t g = nullptr;
            const int64_t l0 = h->launches, s0 = h->host_step;
            DQN_CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            // the small-input upload has a fixed pinned address: it rides the graph as its first node; the scalar/loss
            // read-back is done by the step's last kernel (optimizer tail) into the pinned, device-mapped mirror
            h->tail_mirror = reinterpret_cast<float*>(h->h_sc); h->tail_mirror_words = (int)(readback / 4);
            try {
                upload_small(h);
                launches();
                h->tail_mirror = nullptr; h->tail_mirror_words = 0;
            } catch (...) {
                h->tail_mirror = nullptr; h->tail_mirror_words = 0;
                cudaStreamEndCapture(h->stream, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            DQN_CUDA_OK(cudaStreamEndCapture(h->stream, &g));
            h->batch_graph_launches = h->launches - l0;
            h->launches = l0; h->host_step = s0;
            if (h->batch_graph) { cudaGraphExecDestroy(h->batch_graph); h->batch_graph = nullptr; }
            DQN_CUDA_OK(cudaGraphInstantiateWithFlags(&h->batch_graph, g, cudaGraphInstantiateFlagUseNodePriority));
            DQN_CUDA_OK(cudaGraphDestroy(g));
            h->batch_graph_n = n; h->graph_valid_batch = true;
        }
        DQN_CUDA_OK(cudaGraphLaunch(h->batch_graph, h->stream));
        h->launches += h->batch_graph_launches;
        h->host_step++;
    }
    // scalars and per-sample losses sit in one device block with a pinned mirror: one copy, one synchronisation (a copy
    // into the caller's pageable memory would be staged by the driver and cost a second one)
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (losses) memcpy(losses, reinterpret_cast<const char*>(h->h_sc) + SC_BLOCK, (size_t)n * 4);
    if (step) *step = h->h_sc->step;
    if (loss) *loss = h->h_sc->loss;
    API_END(h)
}

extern "C" int dqn_get_losses(dqn_learner* h, int32_t n, const uint8_t* st, const uint8_t* ac, const uint8_t* rw,
                              const uint8_t* ct, const uint8_t* ns, float* losses) {
    API_BEGIN(h)
    DQN_REQUIRE(n >= 1 && n <= h->max_rows, "n exceeds max_rows");
    upload_batch(h, n, st, ac, rw, ct, ns, nullptr);
    // the losses tensor does not depend on is_weights (deep_q_model.py:150-157 feeds none)
    const int saved_profile = h->profile; h->profile = 0;
    static_cast<void>(saved_profile);
    if (h->cfg.use_per) {
        DQN_CUDA_OK(cudaMemsetAsync(h->in_isw, 0, n * 4, h->stream));
    }
    step_core(h, n, h->in_states, h->in_actions, h->in_rewards, h->in_cont, h->cfg.use_per ? h->in_isw : nullptr, 0, 1.0);
    h->profile = saved_profile;
    d2h(h, losses, h->b_losses_out, n);
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_predict(dqn_learner* h, int32_t n, const uint8_t* st, int32_t use_target, float* q) {
    API_BEGIN(h)
    const size_t S = h->state_bytes;
    const int A = h->net.A;
    const int saved_profile = h->profile; h->profile = 0;
    for (int done = 0; done < n;) {
        const int m = std::min(n - done, h->max_rows);
        DQN_CUDA_OK(cudaMemcpyAsync(h->in_states, st + (size_t)done * S, m * S, cudaMemcpyHostToDevice, h->stream));
        TowerActs& acts = use_target ? h->acts_tg : h->acts_on;
        const float* xf = nullptr;
        if (h->cfg.math_mode != DQN_MATH_FP32 && !tower_fused_ready(h)) { launch_u8_to_f32(h, h->in_states, h->x_f32, (size_t)m * S); xf = h->x_f32; }
        tower_forward(h, use_target ? h->target : h->online, h->in_states, xf, m, acts);
        DQN_CUDA_OK(cudaMemcpyAsync(q + (size_t)done * A, acts.q, (size_t)m * A * 4, cudaMemcpyDeviceToHost, h->stream));
        done += m;
    }
    h->profile = saved_profile;
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_debug_read(dqn_learner* h, const char* what, float* out, int64_t count) {
    API_BEGIN(h)
    const std::string w = what;
    const float* src = nullptr;
    if (w == "q_online") src = h->acts_on.q;
    else if (w == "q_target") src = h->acts_tg.q;
    else if (w == "y") src = h->b_y;
    else if (w == "max_q") src = h->b_maxq;
    else if (w == "losses") src = h->b_losses_out;
    else if (w == "dq") src = h->b_dq;
    else if (w == "new_priorities") src = h->b_newprio;
    else if (w == "is_weights") src = h->b_isw;
    else if (w == "d_hid") src = h->d_hid;
    else if (w == "hid_online") src = h->acts_on.hid;
    else if (w == "act1_online") src = h->acts_on.act[0];
    else if (w == "act2_online") src = h->acts_on.act[1];
    else if (w == "act3_online") src = h->acts_on.act[2];
    else if (w == "d_act1") src = h->d_act[0];
    else if (w == "d_act2") src = h->d_act[1];
    else if (w == "d_act3") src = h->d_act[2];
    else throw std::string("unknown debug buffer: ") + w;
    DQN_CUDA_OK(cudaMemcpyAsync(out, src, count * 4, cudaMemcpyDeviceToHost, h->stream));
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    API_END(h)
}

extern "C" int dqn_set_option(dqn_learner* h, const char* name, int32_t value) {
    API_BEGIN(h)
    const std::string n = name;
    if (n == "overlap") h->overlap = value != 0;          // side-stream forks inside a step
    else if (n == "fuse_fc_adam") h->fuse_enabled = value != 0;  // optimizer in the dense wgrad epilogue
    else if (n == "batch_graph") h->batch_graph_enabled = value != 0;  // dqn_train_batch replays a captured graph
    else if (n == "fc_split_streams") h->fc_split_streams = value != 0;
    else if (n == "fc_tma_fwd") h->fc_tma_fwd = value != 0;
    else if (n == "fc_fwd_share") h->fc_fwd_share = value;
    else if (n == "fc_tma_ctas") h->fc_tma_ctas = value;
    else if (n == "fc_rows_ctas") h->fc_rows_ctas = value;
    else if (n == "w_scratch") {  // fused dense kernel writes the new dense weights to a scratch copy (starts before the dense dgrad)
        h->w_scratch_on = value != 0;
        if (h->fa_row_maps) { free(h->fa_row_maps); h->fa_row_maps = nullptr; }  // the store maps point somewhere else now
    }
    else if (n == "external_normaliser") h->external_normaliser = value != 0;  // the caller owns last_max_weight (data-parallel)
    else if (n == "l2_keep") h->l2_keep = value;          // L2 residency hints (learner.h)
    else if (n == "w_scratch") {  // fused dense kernel writes the new dense weights to a scratch copy (starts before the dense dgrad)
        h->w_scratch_on = value != 0;
        if (h->fa_row_maps) { free(h->fa_row_maps); h->fa_row_maps = nullptr; }  // the store maps point somewhere else now
    }
    else if (n == "external_normaliser") h->external_normaliser = value != 0;  // the caller owns last_max_weight (data-parallel)
    else if (n == "l2_keep") h->l2_keep = value;          // L2 residency hints (learner.h)
    else if (n == "l2_persist_mb") {                      // experiment: size of the persisting L2 carve-out
        cudaDeviceProp prop; DQN_CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
        size_t want = (size_t)value << 20;
        if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
        DQN_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        fprintf(stderr, "[dqn_b200] persisting L2: max %d MB, set %zu MB, L2 %d MB\n", prop.persistingL2CacheMaxSize >> 20, want >> 20, prop.l2CacheSize >> 20);
    }
    else if (n == "wg1_fit") h->wg1_fit = value != 0;
    else if (n == "tail_join") h->tail_join = value != 0;
    else if (n == "fc_after_dgrad") h->fc_after_dgrad = value != 0;
    else if (n == "keep_grads") h->keep_grads = value != 0;  // dqn_train_batch: dense-kernel gradients stay readable (slot 3)
    else if (n == "fc_tma_adam") h->fc_tma_adam = value != 0;    // dense wgrad (tcgen05) + optimizer, TMA-staged state
    else if (n == "fc_ffma_adam") h->fc_ffma_adam = value != 0;  // dense wgrad+optimizer as one fp32 pass
    else if (n == "pdl") h->pdl = value != 0;             // programmatic dependent launch between consecutive kernels
    else if (n == "ts_gemm") h->ts_gemm = value != 0;     // A operand through tensor memory (tsgemm.cuh) vs all-smem (tcgemm.cuh)
    else if (n == "repack") { launch_pack_conv(h, 0); launch_pack_conv(h, 1); }  // after writing weights through dqn_device_ptr
    else if (n == "img_conv") h->img_conv = value != 0;   // conv layers through the image-resident kernel (cvgemm.cuh)
    else if (n == "tower_fused") h->tower_fused = value != 0;  // conv1..conv3 of both towers as one kernel on the uint8 batch (cvtower.cuh)
    else if (n == "tower_ext2") h->tower_ext2 = value != 0;
    else if (n == "fc_prefetch") h->fc_prefetch = value != 0;
    else if (n == "fc_rows_pf") { DQN_REQUIRE(value >= 0 && value <= 16, "fc_rows_pf: 0..16 tiles per CTA ahead"); h->fc_rows_pf = (int)value; }
    else if (n == "tower_cluster") { DQN_REQUIRE(value == 1 || value == 2 || value == 4 || value == 8, "tower_cluster: 1, 2, 4 or 8"); h->tower_cluster = (int)value; }
    else if (n == "img_dgrad") h->img_dgrad = value != 0;
    else if (n == "fuse_td") h->fuse_td = value != 0;
    else if (n == "pack_in_tail") h->pack_in_tail = value != 0;
    else if (n == "img_wgrad1") h->img_wgrad1 = value != 0;
    else if (n == "merge_towers") h->merge_towers = value != 0;
    else if (n == "prefetch") h->prefetch = value != 0;   // dqn_train_steps samples the next batch inside the current step
    else if (n == "defer_adam") h->defer_adam = value != 0;  // dense-kernel optimizer pass of step i under the conv forward of step i+1
    else if (n == "priorities") h->use_prio = value != 0; // critical-chain kernels dispatch ahead of the multi-wave side kernels
    else throw std::string("unknown option: ") + n;
    h->graph_valid = false; h->graph_valid_batch = false; h->graph_valid_u = false; h->graph_valid_u_alt = false;
    if (h->prof_graph) { cudaGraphExecDestroy(h->prof_graph); h->prof_graph = nullptr; }
    API_END(h)
}

extern "C" int dqn_launch_count(const dqn_learner* h, int64_t* n) {
    if (!h) return 2;
    if (n) *n = h->launches;
    return 0;
}

extern "C" int dqn_set_profile(dqn_learner* h, int32_t enabled) {
    if (!h) return 2;
    h->profile = enabled;
    h->prof_acc.clear();
    h->prof_steps = 0;
    return 0;
}

extern "C" int dqn_get_profile(dqn_learner* h, int32_t* n, const char** names, float* ms, int32_t cap, int64_t* steps) {
    if (!h) return 2;
    int k = 0;
    for (auto& kv : h->prof_acc) {
        if (k >= cap) break;
        names[k] = kv.first;
        ms[k] = kv.second;
        ++k;
    }
    if (n) *n = k;
    if (steps) *steps = h->prof_steps;
    return 0;
}

// ---- persistence -------------------------------------------------------------------------------------------
struct CkptHeader {
    char magic[8];
    int32_t version, n_tensors;
    int64_t slab, step, game_count;
    float b1p, b2p;
    int32_t H, W, C, A, hidden, dueling, variant, pad;
    uint64_t rng_counter;  // version >= 2: device Philox counter, so that a resumed device_rng run continues its draw sequence
    uint64_t reserved[3];
};
constexpr size_t CKPT_V1_HEADER = 80;  // sizeof(CkptHeader) of version 1 (everything up to and including `pad`)

extern "C" int dqn_save(dqn_learner* h, const char* prefix) {
    API_BEGIN(h)
    read_scalars(h);
    const std::string path = std::string(prefix) + ".dqnb200";
    FILE* f = fopen((path + ".tmp").c_str(), "wb");
    DQN_REQUIRE(f != nullptr, "cannot open checkpoint for writing");
    CkptHeader hd; memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, "DQNB200", 8); hd.version = 2; hd.n_tensors = (int32_t)h->net.tensors.size();
    hd.slab = h->net.slab; hd.step = h->h_sc->step; hd.game_count = h->h_sc->game_count;
    hd.b1p = h->h_sc->b1p; hd.b2p = h->h_sc->b2p;
    hd.H = h->net.H; hd.W = h->net.W; hd.C = h->net.C; hd.A = h->net.A; hd.hidden = h->net.hidden;
    hd.dueling = h->net.dueling; hd.variant = h->cfg.conv_variant;
    hd.rng_counter = h->h_sc->rng_counter;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
    std::vector<float> buf((size_t)h->net.slab);
    for (float* slab : {h->online, h->target, h->slot_m, h->slot_v}) {
        DQN_CUDA_OK(cudaMemcpyAsync(buf.data(), slab, buf.size() * 4, cudaMemcpyDeviceToHost, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        ok = ok && fwrite(buf.data(), 4, buf.size(), f) == buf.size();
    }
    fclose(f);
    DQN_REQUIRE(ok, "short write on checkpoint");
    DQN_REQUIRE(rename((path + ".tmp").c_str(), path.c_str()) == 0, "cannot move checkpoint into place");
    API_END(h)
}

__global__ void set_rng_kernel(dqn_learner::Scalars* sc, unsigned long long v) { sc->rng_counter = v; }

extern "C" int dqn_restore(dqn_learner* h, const char* prefix) {
    if (!h) return 2;
    const std::string path = std::string(prefix) + ".dqnb200";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return 3;  // "model does not exist" (rl/model/model.py:76-78)
    try {
        DQN_CUDA_OK(cudaSetDevice(h->device));
        // steps may still be in flight (dqn_train_steps returns after enqueueing); the look-ahead batch was sampled for the
        // old weights' step counter
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        h->pf_valid = false;
        CkptHeader hd; memset(&hd, 0, sizeof(hd));
        DQN_REQUIRE(fread(&hd, CKPT_V1_HEADER, 1, f) == 1 && memcmp(hd.magic, "DQNB200", 8) == 0, "not a dqn_b200 checkpoint");
        DQN_REQUIRE(hd.version == 1 || hd.version == 2, "unsupported dqn_b200 checkpoint version");
        if (hd.version >= 2)
            DQN_REQUIRE(fread(reinterpret_cast<char*>(&hd) + CKPT_V1_HEADER, sizeof(hd) - CKPT_V1_HEADER, 1, f) == 1, "truncated checkpoint header");
        DQN_REQUIRE(hd.slab == h->net.slab && hd.H == h->net.H && hd.W == h->net.W && hd.C == h->net.C && hd.A == h->net.A &&
                        hd.hidden == h->net.hidden && hd.dueling == h->net.dueling && hd.variant == h->cfg.conv_variant &&
                        hd.n_tensors == (int32_t)h->net.tensors.size(),
                    "checkpoint was written for a different network configuration");
        std::vector<float> buf((size_t)h->net.slab);
        for (float* slab : {h->online, h->target, h->slot_m, h->slot_v}) {
            DQN_REQUIRE(fread(buf.data(), 4, buf.size(), f) == buf.size(), "truncated checkpoint");
            DQN_CUDA_OK(cudaMemcpyAsync(slab, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice, h->stream));
            DQN_CUDA_OK(cudaStreamSynchronize(h->stream));  // buf is reused
        }
        fclose(f); f = nullptr;
        set_ll(h, &h->d_sc->step, hd.step);
        set_ll(h, &h->d_sc->game_count, hd.game_count);
        set_f2_kernel<<<1, 1, 0, h->stream>>>(&h->d_sc->b1p, hd.b1p, &h->d_sc->b2p, hd.b2p);
        if (hd.version >= 2) set_rng_kernel<<<1, 1, 0, h->stream>>>(h->d_sc, hd.rng_counter);
        h->host_step = hd.step;
        h->has_max_weight = false;  // PER normaliser: recomputed by the first sample, as after a reference restart (:505 "is None")
        launch_pack_conv(h, 0); launch_pack_conv(h, 1);
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    } catch (const std::string& e) {
        if (f) fclose(f);
        h->err = e;
        return 1;
    }
    return 0;
}


// ---- replay persistence ----------------------------------------------------------------------------------------------
// The reference writes its RAM replay to <prefix>_replay_memory.hdf5 at exit and reloads it at the next start
// (rl/deep_q_agent.py:410-432, :211-222): rows 0 .. len-1 in INDEX order, all six columns; the sum tree is not stored, it is
// rebuilt from the float16 `losses` column by ReplaySamplerPriority.add_losses (rl/data/replay_sampler_priority.py:57-59).
// Same contents and the same reload semantics here, in this library's own container (<prefix>_replay_memory.dqnb200;
// HDF5 itself cannot be produced or checked in this environment).
struct ReplayHeader {
    char magic[8];                 // "DQNRPLY"
    int32_t version, layout;       // layout 0 = stacked s / s', 1 = frame ring + 8 slots per row
    int32_t H, W, C, pad;
    int64_t max_size, rows, cur_idx, is_full;
    int64_t frame_capacity, frames_written;
};

__global__ void half_to_float_kernel(const __half* __restrict__ src, float* __restrict__ dst, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __half2float(src[i]);
}

static void stream_out(dqn_learner* h, FILE* f, const void* d_src, size_t bytes, std::vector<uint8_t>& buf) {
    for (size_t off = 0; off < bytes; off += buf.size()) {
        const size_t m = std::min(buf.size(), bytes - off);
        DQN_CUDA_OK(cudaMemcpyAsync(buf.data(), static_cast<const uint8_t*>(d_src) + off, m, cudaMemcpyDeviceToHost, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        DQN_REQUIRE(fwrite(buf.data(), 1, m, f) == m, "short write on replay file");
    }
}
static void stream_in(dqn_learner* h, FILE* f, void* d_dst, size_t bytes, std::vector<uint8_t>& buf) {
    for (size_t off = 0; off < bytes; off += buf.size()) {
        const size_t m = std::min(buf.size(), bytes - off);
        DQN_REQUIRE(fread(buf.data(), 1, m, f) == m, "truncated replay file");
        DQN_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + off, buf.data(), m, cudaMemcpyHostToDevice, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    }
}

extern "C" int dqn_save_replay(dqn_learner* h, const char* prefix) {
    API_BEGIN(h)
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    const std::string path = std::string(prefix) + "_replay_memory.dqnb200";
    FILE* f = fopen((path + ".tmp").c_str(), "wb");
    DQN_REQUIRE(f != nullptr, "cannot open replay file for writing");
    try {
        ReplayHeader hd; memset(&hd, 0, sizeof(hd));
        memcpy(hd.magic, "DQNRPLY", 8); hd.version = 1; hd.layout = h->dedup ? 1 : 0;
        hd.H = h->net.H; hd.W = h->net.W; hd.C = h->net.C;
        hd.max_size = h->cap; hd.rows = h->ring_size; hd.cur_idx = h->ring_cur; hd.is_full = h->ring_full ? 1 : 0;
        hd.frame_capacity = h->fcap; hd.frames_written = h->f_written;
        DQN_REQUIRE(fwrite(&hd, sizeof(hd), 1, f) == 1, "short write on replay file");
        std::vector<uint8_t> buf(64u << 20);
        const size_t n = (size_t)h->ring_size;
        stream_out(h, f, h->r_actions, n, buf); stream_out(h, f, h->r_rewards, n, buf); stream_out(h, f, h->r_cont, n, buf);
        stream_out(h, f, h->r_losses, n * sizeof(__half), buf);
        if (h->dedup) {
            stream_out(h, f, h->r_fidx, n * 8 * sizeof(uint32_t), buf);
            stream_out(h, f, h->r_frames, (size_t)std::min<int64_t>(h->f_written, h->fcap) * h->frame_bytes, buf);
        } else {
            stream_out(h, f, h->r_states, n * h->state_bytes, buf);
            stream_out(h, f, h->r_next, n * h->state_bytes, buf);
        }
    } catch (...) { fclose(f); remove((path + ".tmp").c_str()); throw; }
    fclose(f);
    DQN_REQUIRE(rename((path + ".tmp").c_str(), path.c_str()) == 0, "cannot move replay file into place");
    API_END(h)
}

// returns 3 when the file does not exist.  The ring must be empty (a freshly created handle): rows are appended from index 0
// like ReplayMemory.extend does, the sum tree is rebuilt by add(losses[i], i) in index order.
extern "C" int dqn_restore_replay(dqn_learner* h, const char* prefix, int64_t* rows_loaded) {
    if (!h) return 2;
    const std::string path = std::string(prefix) + "_replay_memory.dqnb200";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return 3;
    try {
        DQN_CUDA_OK(cudaSetDevice(h->device));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        h->pf_valid = false;
        DQN_REQUIRE(h->ring_size == 0 && h->tree_size == 0, "replay restore needs an empty ring");
        ReplayHeader hd;
        DQN_REQUIRE(fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, "DQNRPLY", 8) == 0 && hd.version == 1, "not a dqn_b200 replay file");
        DQN_REQUIRE(hd.H == h->net.H && hd.W == h->net.W && hd.C == h->net.C, "replay file holds frames of another geometry");
        DQN_REQUIRE(hd.layout == (h->dedup ? 1 : 0), "replay file layout (stacked / frame ring) differs from this handle's");
        DQN_REQUIRE(hd.rows <= h->cap, "replay file holds more rows than this handle's capacity");
        DQN_REQUIRE(!h->dedup || hd.frame_capacity == h->fcap, "frame ring capacity differs from the file's");
        std::vector<uint8_t> buf(64u << 20);
        const size_t n = (size_t)hd.rows;
        stream_in(h, f, h->r_actions, n, buf); stream_in(h, f, h->r_rewards, n, buf); stream_in(h, f, h->r_cont, n, buf);
        stream_in(h, f, h->r_losses, n * sizeof(__half), buf);
        if (h->dedup) {
            stream_in(h, f, h->r_fidx, n * 8 * sizeof(uint32_t), buf);
            stream_in(h, f, h->r_frames, (size_t)std::min<int64_t>(hd.frames_written, hd.frame_capacity) * h->frame_bytes, buf);
            h->f_written = hd.frames_written;
        } else {
            stream_in(h, f, h->r_states, n * h->state_bytes, buf);
            stream_in(h, f, h->r_next, n * h->state_bytes, buf);
        }
        fclose(f); f = nullptr;
        // ReplayMemory.extend: n appends from index 0 (rl/data/replay_memory.py:70-75)
        h->ring_cur = hd.rows % h->cap; h->ring_full = hd.rows >= h->cap; h->ring_size = hd.rows;
        if (h->cfg.use_per && n > 0) {
            // ReplaySamplerPriority.add_losses: sum_tree.add(losses[i], i) for i in range(len) -- the priorities are the stored
            // float16 values (rl/data/replay_sampler_priority.py:57-59)
            const int64_t chunk = (int64_t)(h->dev_stage_bytes / 4);
            for (int64_t s0 = 0; s0 < hd.rows; s0 += chunk) {
                const int64_t m = std::min(chunk, hd.rows - s0);
                float* d_sc = (float*)h->dev_stage;
                half_to_float_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, h->stream>>>(h->r_losses + s0, d_sc, m);
                launch_tree_add(h, m, d_sc, nullptr, s0);
                DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
            }
            h->tree_max_reached = hd.rows >= h->cap;
            h->tree_size = hd.rows; h->tree_write = hd.rows % h->cap;
        }
        h->has_max_weight = false;
        push_sizes(h);
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        if (rows_loaded) *rows_loaded = hd.rows;
    } catch (const std::string& e) {
        if (f) fclose(f);
        h->err = e;
        return 1;
    }
    return 0;
}


// GEMM engine self-test (tests/test_gpu_tcgemm.py): C[M,N] = A * B with host matrices, through the FFMA kernel
// (mode 0) or the tcgen05 kernel (mode 1: TF32 x3, mode 2: TF32 x1)
extern "C" int dqn_selftest_gemm(dqn_learner* h, int32_t mode, int32_t M, int32_t N, int32_t K, int32_t a_kcontig,
                                 int32_t b_kcontig, const float* hA, const float* hB, float* hC) {
    API_BEGIN(h)
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    dmalloc(dA, (size_t)M * K); dmalloc(dB, (size_t)N * K); dmalloc(dC, (size_t)M * N);
    try {
        DQN_CUDA_OK(cudaMemcpyAsync(dA, hA, (size_t)M * K * 4, cudaMemcpyHostToDevice, h->stream));
        DQN_CUDA_OK(cudaMemcpyAsync(dB, hB, (size_t)N * K * 4, cudaMemcpyHostToDevice, h->stream));
        DQN_CUDA_OK(cudaMemsetAsync(dC, 0xFF, (size_t)M * N * 4, h->stream));
        selftest_gemm(h, mode, M, N, K, a_kcontig, b_kcontig, dA, dB, dC);
        DQN_CUDA_OK(cudaMemcpyAsync(hC, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    } catch (...) {
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
        throw;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    API_END(h)
}


// GEMM engine micro-benchmark: device-resident random-free operands, `iters` launches timed with CUDA events.
extern "C" int dqn_bench_gemm(dqn_learner* h, int32_t mode, int32_t M, int32_t N, int32_t K, int32_t a_kcontig,
                              int32_t b_kcontig, int32_t iters, float* ms_per_launch) {
    API_BEGIN(h)
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    dmalloc(dA, (size_t)M * K); dmalloc(dB, (size_t)N * K); dmalloc(dC, (size_t)M * N);
    cudaEvent_t e0, e1;
    DQN_CUDA_OK(cudaEventCreate(&e0)); DQN_CUDA_OK(cudaEventCreate(&e1));
    try {
        DQN_CUDA_OK(cudaMemsetAsync(dA, 0x3c, (size_t)M * K * 4, h->stream));
        DQN_CUDA_OK(cudaMemsetAsync(dB, 0x3c, (size_t)N * K * 4, h->stream));
        for (int i = 0; i < 3; ++i) selftest_gemm(h, mode, M, N, K, a_kcontig, b_kcontig, dA, dB, dC);
        DQN_CUDA_OK(cudaEventRecord(e0, h->stream));
        for (int i = 0; i < iters; ++i) selftest_gemm(h, mode, M, N, K, a_kcontig, b_kcontig, dA, dB, dC);
        DQN_CUDA_OK(cudaEventRecord(e1, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        DQN_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        *ms_per_launch = ms / iters;
    } catch (...) {
        cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaEventDestroy(e0); cudaEventDestroy(e1);
        throw;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaEventDestroy(e0); cudaEventDestroy(e1);
    API_END(h)
}


// In-kernel timeline of the first CTA of one tcgen05 GEMM launch: out[kb*8 + {0..3}] = producer thread 0 clocks
// (loop top, after empty wait, after stores, after arrive+next loads), {4..7} = MMA thread (before full wait, after,
// after fences, after issue+commit).  Debug aid for the k-loop pipeline.
extern "C" int dqn_debug_timeline(dqn_learner* h, int32_t mode, int32_t M, int32_t N, int32_t K, int32_t a_kcontig,
                                  int32_t b_kcontig, unsigned long long* out512) {
    API_BEGIN(h)
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    unsigned long long* dT = nullptr;
    if (mode >= 16) {  // conv layer (mode - 16) of the online tower over M images, in place on the tower's buffers
        // mode 20: the fused conv tower (cvtower.cuh) of the online tower over M images of the sampled batch; 1024 marks
        // modes 21-23: the image-resident weight gradient of conv layer 1-3 (1024 marks as well)
        DQN_REQUIRE(mode - 16 <= 7 && M >= 1 && M <= 2 * h->max_rows, "debug conv layer / rows");
        const int marks = mode >= 20 ? 1024 : 512;
        dmalloc(dT, 1024);
        DQN_CUDA_OK(cudaMemsetAsync(dT, 0, 1024 * 8, h->stream));
        debug_conv_layer(h, mode - 16, M);
        h->tc_dbg = dT;
        try { debug_conv_layer(h, mode - 16, M); } catch (...) { h->tc_dbg = nullptr; throw; }
        h->tc_dbg = nullptr;
        DQN_CUDA_OK(cudaMemcpyAsync(out512, dT, marks * 8, cudaMemcpyDeviceToHost, h->stream));
        DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
        cudaFree(dT);
        return 0;
    }
    dmalloc(dA, (size_t)M * K); dmalloc(dB, (size_t)N * K); dmalloc(dC, (size_t)M * N); dmalloc(dT, 512);
    DQN_CUDA_OK(cudaMemsetAsync(dA, 0x3c, (size_t)M * K * 4, h->stream));
    DQN_CUDA_OK(cudaMemsetAsync(dB, 0x3c, (size_t)N * K * 4, h->stream));
    DQN_CUDA_OK(cudaMemsetAsync(dT, 0, 512 * 8, h->stream));
    selftest_gemm(h, mode, M, N, K, a_kcontig, b_kcontig, dA, dB, dC);  // warm
    h->tc_dbg = dT;
    try { selftest_gemm(h, mode, M, N, K, a_kcontig, b_kcontig, dA, dB, dC); } catch (...) { h->tc_dbg = nullptr; throw; }
    h->tc_dbg = nullptr;
    DQN_CUDA_OK(cudaMemcpyAsync(out512, dT, 512 * 8, cudaMemcpyDeviceToHost, h->stream));
    DQN_CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dT);
    API_END(h)
}
