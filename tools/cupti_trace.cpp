// Tiny CUPTI activity tracer (developer tool, not part of the product): records every kernel's start/end/stream, also
// for the nodes of a replayed CUDA graph, so the overlap structure of one fused step can be read as a timeline.
//   trace_begin(); ...run...; trace_end("file.csv")
#include <cupti.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <mutex>

struct Rec { std::string name; unsigned long long start, end; unsigned stream, grid; unsigned long long graphNode; };
static std::vector<Rec> g_recs;
static std::mutex g_mu;

static void CUPTIAPI buf_req(uint8_t** buffer, size_t* size, size_t* maxNumRecords) {
    *size = 8 << 20;
    *buffer = (uint8_t*)aligned_alloc(8, *size);
    *maxNumRecords = 0;
}
static void CUPTIAPI buf_done(CUcontext, uint32_t, uint8_t* buffer, size_t, size_t validSize) {
    CUpti_Activity* rec = nullptr;
    while (cuptiActivityGetNextRecord(buffer, validSize, &rec) == CUPTI_SUCCESS) {
        if (rec->kind == CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL || rec->kind == CUPTI_ACTIVITY_KIND_KERNEL) {
            auto* k = (CUpti_ActivityKernel9*)rec;
            std::lock_guard<std::mutex> l(g_mu);
            g_recs.push_back({k->name ? k->name : "?", (unsigned long long)k->start, (unsigned long long)k->end, k->streamId,
                              (unsigned)(k->gridX * k->gridY * k->gridZ), (unsigned long long)k->graphNodeId});
        } else if (rec->kind == CUPTI_ACTIVITY_KIND_MEMCPY) {
            auto* m = (CUpti_ActivityMemcpy5*)rec;
            std::lock_guard<std::mutex> l(g_mu);
            g_recs.push_back({"[memcpy]", (unsigned long long)m->start, (unsigned long long)m->end, m->streamId, (unsigned)m->bytes, 0});
        }
    }
    free(buffer);
}
extern "C" int trace_begin() {
    if (cuptiActivityRegisterCallbacks(buf_req, buf_done) != CUPTI_SUCCESS) return 1;
    if (cuptiActivityEnable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL) != CUPTI_SUCCESS) return 2;
    cuptiActivityEnable(CUPTI_ACTIVITY_KIND_MEMCPY);
    return 0;
}
extern "C" int trace_end(const char* path) {
    cuptiActivityFlushAll(1);
    cuptiActivityDisable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL);
    cuptiActivityDisable(CUPTI_ACTIVITY_KIND_MEMCPY);
    FILE* f = fopen(path, "w");
    if (!f) return 1;
    fprintf(f, "start_ns,end_ns,stream,grid,node,name\n");
    for (auto& r : g_recs) fprintf(f, "%llu,%llu,%u,%u,%llu,\"%s\"\n", r.start, r.end, r.stream, r.grid, r.graphNode, r.name.c_str());
    fclose(f);
    g_recs.clear();
    return 0;
}
