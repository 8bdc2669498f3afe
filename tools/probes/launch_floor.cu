// launch_floor.cu -- what a chain of SIX dependent, EMPTY kernels with the fused step's grid shapes costs per step on this GPU:
// the floor that launch latency + the grid-wide dependency (drain, flush, release of the dependents) put under the step,
// whatever the kernels do.  Three chains: plain stream order, programmatic dependent launch with the dependents released at the
// kernel's start, and released at CTA exit (what the library does in its raster kernels).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/launch_floor tools/probes/launch_floor.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_empty(int late, float* sink)
{
    if (!late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) sink[0] = 1.0f;      // one store, so that there is something to flush
}

static void launch(dim3 g, dim3 b, bool pdl, int late, float* sink, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = g; cfg.blockDim = b; cfg.stream = s;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = a; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_empty, late, sink);
}

int main()
{
    // the fused step at cfg-2: vertex_fwd 384 x 512, hard 960 x 256, soft_fwd 1920 x 128, shade 1536 x 128, soft_bwd 1480 x 128,
    // vertex_bwd 192 x 512
    const int grids[6] = {384, 960, 1920, 1536, 1480, 192}, blocks[6] = {512, 256, 128, 128, 128, 512};
    float* sink; cudaMalloc(&sink, 256);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int steps = 2000;
    for (int mode = 0; mode < 3; ++mode) {
        const bool pdl = mode > 0; const int late = mode == 2;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0, s);
            for (int i = 0; i < steps; ++i)
                for (int k = 0; k < 6; ++k) launch(dim3(grids[k]), dim3(blocks[k]), pdl, late, sink, s);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("%-46s %.2f us per step of 6 empty kernels (%.2f us per kernel)\n",
                            mode == 0 ? "stream order" : (mode == 1 ? "PDL, dependents released at kernel start" : "PDL, dependents released at CTA exit"),
                            1e3 * ms / steps, 1e3 * ms / steps / 6);
        }
    }
    // the same six launches captured once and replayed as a graph (how the bench's API path replays its step)
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int k = 0; k < 6; ++k) launch(dim3(grids[k]), dim3(blocks[k]), true, 1, sink, s);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0, s);
        for (int i = 0; i < steps; ++i) cudaGraphLaunch(ge, s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("%-46s %.2f us per step\n", "graph replay of the late-release chain", 1e3 * ms / steps);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
