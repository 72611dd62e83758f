/* C restatement of the reference sum tree for full-size parity runs.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/rl/data/sum_tree.py: add :23-42, update_score :45-52, _propagate :101-110,
 * _retrieve :113-129 (incl. the fix-up clamp :120-124), get_with_info :76-83, and the stratified probe of
 * rl/data/replay_sampler_priority.py:30-33 under the pinned in-container promotion (every operand float32).
 * Pinned by tests/test_oracle_replay.py against the same tests/golden fixtures as oracle/sumtree_ref.py.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction, no fast-math). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t capacity, size, write;
    int max_reached;
    float* tree;
    uint32_t* data;
} st_t;

st_t* st_new(int64_t capacity) {
    st_t* t = (st_t*)calloc(1, sizeof(st_t));
    t->capacity = capacity;
    t->tree = (float*)calloc((size_t)(2 * capacity - 1), sizeof(float));
    t->data = (uint32_t*)calloc((size_t)capacity, sizeof(uint32_t));
    return t;
}
void st_free(st_t* t) { free(t->tree); free(t->data); free(t); }
float* st_tree(st_t* t) { return t->tree; }
uint32_t* st_data(st_t* t) { return t->data; }
int64_t st_size(st_t* t) { return t->size; }
int64_t st_write(st_t* t) { return t->write; }
void st_set_state(st_t* t, int64_t size, int64_t write, int max_reached) { t->size = size; t->write = write; t->max_reached = max_reached; }

void st_update(st_t* t, int64_t idx, float score) {
    volatile float change = score - t->tree[idx];
    t->tree[idx] = score;
    while (1) {
        idx = (idx - 1) / 2;
        volatile float v = t->tree[idx] + change;
        t->tree[idx] = v;
        if (idx == 0) break;
    }
}

void st_add(st_t* t, float score, uint32_t data) {
    int64_t idx = t->write + t->capacity - 1;
    t->data[t->write] = data;
    st_update(t, idx, score);
    if (!t->max_reached) t->size += 1;
    t->write += 1;
    if (t->write >= t->capacity) { t->write = 0; t->max_reached = 1; }
}

void st_add_many(st_t* t, int64_t n, const float* scores, const uint32_t* data) {
    for (int64_t i = 0; i < n; ++i) st_add(t, scores[i], data ? data[i] : (uint32_t)t->write);
}

void st_update_many(st_t* t, int64_t n, const int64_t* idx, const float* scores) {
    for (int64_t i = 0; i < n; ++i) st_update(t, idx[i], scores[i]);
}

int64_t st_retrieve(const st_t* t, float score) {
    const int64_t n = 2 * t->capacity - 1;
    int64_t idx = 0;
    while (1) {
        int64_t left = 2 * idx + 1;
        if (left >= n) {
            int64_t over = (idx - t->capacity + 1) - t->size + 1;
            if (over > 0) idx -= over;
            return idx;
        }
        float lv = t->tree[left];
        if (score <= lv) idx = left;
        else { volatile float s = score - lv; score = s; idx = left + 1; }
    }
}

/* B stratified probes: s = f32(f32(u)*size) + f32(f32(i)*size), size = total / f32(B) */
void st_sample(const st_t* t, int32_t batch, const double* u, int64_t* tree_idx, int64_t* data_idx, double* prio, int64_t* mem_idx) {
    volatile float size = t->tree[0] / (float)batch;
    for (int32_t i = 0; i < batch; ++i) {
        volatile float a = (float)u[i] * size;
        volatile float b = (float)i * size;
        volatile float s = a + b;
        int64_t ti = st_retrieve(t, s);
        int64_t di = ti - t->capacity + 1;
        tree_idx[i] = ti; data_idx[i] = di; prio[i] = (double)t->tree[ti]; mem_idx[i] = (int64_t)t->data[di];
    }
}
