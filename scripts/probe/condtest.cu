// Probe: cost of an IDLE rebuild chain inside a step-group graph -- 11 early-exit kernels against one gate kernel plus
// a conditional IF node (CUDA graphs, sm_100a). Prints microseconds per graph launch.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void gate(cudaGraphConditionalHandle h, const int* flag) { cudaGraphSetConditional(h, *flag ? 1u : 0u); }
__global__ void body(int* c) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(c, 1); }
__global__ void chaink(const int* flag, int* c) { if (*flag == 0) return; atomicAdd(c, 1); }
static float run(cudaGraphExec_t x, cudaStream_t s, int n) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int k = 0; k < 20; k++) cudaGraphLaunch(x, s);
    cudaEventRecord(a, s); for (int k = 0; k < n; k++) cudaGraphLaunch(x, s); cudaEventRecord(b, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1000.f / n;
}
int main() {
    int *flag, *cnt; cudaMalloc(&flag, 4); cudaMalloc(&cnt, 4); cudaMemset(cnt, 0, 4); cudaMemset(flag, 0, 4);
    cudaStream_t s, s2; cudaStreamCreate(&s); cudaStreamCreate(&s2);
    for (int steps : {1, 16}) {
        cudaGraph_t g; cudaGraphExec_t xa, xb, xc;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < steps; k++) body<<<8, 256, 0, s>>>(cnt);
        cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&xc, g, 0);
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < steps; k++) body<<<8, 256, 0, s>>>(cnt);
        for (int k = 0; k < 11; k++) chaink<<<8, 256, 0, s>>>(flag, cnt);
        cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&xa, g, 0);
        cudaGraphCreate(&g, 0);
        cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
        cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < steps; k++) body<<<8, 256, 0, s>>>(cnt);
        gate<<<1, 1, 0, s>>>(h, flag);
        cudaStreamCaptureStatus st; const cudaGraphNode_t* deps; size_t nd; cudaGraph_t cg;
        cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &cg, &deps, &nd);
        cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeIf; p.conditional.size = 1;
        cudaGraphNode_t cn; cudaGraphAddNode(&cn, g, deps, nd, &p);
        cudaStreamBeginCaptureToGraph(s2, p.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < 11; k++) chaink<<<8, 256, 0, s2>>>(flag, cnt);
        cudaStreamEndCapture(s2, nullptr);
        cudaStreamUpdateCaptureDependencies(s, &cn, 1, cudaStreamSetCaptureDependencies);
        cudaStreamEndCapture(s, nullptr);
        cudaError_t e = cudaGraphInstantiate(&xb, g, 0);
        printf("steps %2d: steps only %.2f us | + 11 idle kernels %.2f us | + gate + idle IF %.2f us   (%s)\n", steps, run(xc, s, 2000), run(xa, s, 2000), run(xb, s, 2000), cudaGetErrorString(e));
        int one = 1, zero = 0; cudaMemcpy(flag, &one, 4, cudaMemcpyHostToDevice);
        printf("          taken: 11 kernels %.2f us | IF %.2f us\n", run(xa, s, 2000), run(xb, s, 2000));
        cudaMemcpy(flag, &zero, 4, cudaMemcpyHostToDevice);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
