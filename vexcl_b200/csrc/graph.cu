// CUDA-graph capture of a span of asynchronous work (include/vexb200.h, "CUDA graphs").
#include "common.cuh"

struct vexb_graph {
    int dev = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

using namespace vexb;

extern "C" int vexb_graph_begin(int dev, void *stream) {
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CHECK(stream != nullptr, "graph capture needs an explicit stream (not the legacy default stream)");
    // Relaxed mode: other threads / libraries (NCCL's proxy) may make CUDA calls meanwhile.
    VEXB_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeRelaxed));
    return VEXB_OK;
}

extern "C" int vexb_graph_end(int dev, void *stream, vexb_graph **graph) {
    VEXB_CHECK(graph, "graph is NULL");
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    cudaGraph_t gr = nullptr;
    VEXB_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &gr));
    cudaGraphExec_t ex = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ex, gr, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(gr); VEXB_FAIL(VEXB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    auto *G = new vexb_graph();
    G->dev = dev; G->graph = gr; G->exec = ex;
    *graph = G;
    return VEXB_OK;
}

extern "C" int vexb_graph_launch(vexb_graph *graph, void *stream) {
    VEXB_CHECK(graph && graph->exec, "graph is NULL");
    DeviceGuard g(graph->dev); VEXB_CHECK(g.ok, "cannot select device %d", graph->dev);
    VEXB_CUDA(cudaGraphLaunch(graph->exec, (cudaStream_t)stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return VEXB_OK;
}

extern "C" int vexb_graph_destroy(vexb_graph *graph) {
    if (!graph) return VEXB_OK;
    VEXB_RELEASE_GUARD();
    DeviceGuard g(graph->dev);
    if (graph->exec) cudaGraphExecDestroy(graph->exec);
    if (graph->graph) cudaGraphDestroy(graph->graph);
    delete graph;
    return VEXB_OK;
}
