// probe: does this driver run a CUDA-graph WHILE node whose condition a kernel sets?
#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int *ctr, cudaGraphConditionalHandle h) {
  int v = atomicAdd(ctr, 1);
  cudaGraphSetConditional(h, v + 1 < 10);
}
int main() {
  int *d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  printf("handle %d\n", (int)cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {}; np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = h; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
  cudaGraphNode_t node;
  printf("add %d\n", (int)cudaGraphAddNode(&node, g, nullptr, 0, &np));
  cudaGraph_t bodyg = np.conditional.phGraph_out[0];
  printf("cap %d\n", (int)cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  body<<<1,1,0,s>>>(d, h);
  printf("end %d\n", (int)cudaStreamEndCapture(s, nullptr));
  cudaGraphExec_t e; printf("inst %d\n", (int)cudaGraphInstantiate(&e, g, 0));
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d, 0, 4);
    printf("launch %d\n", (int)cudaGraphLaunch(e, s)); printf("sync %d\n", (int)cudaStreamSynchronize(s));
    int v; cudaMemcpy(&v, d, 4, cudaMemcpyDeviceToHost); printf("count %d (expect 10)\n", v);
  }
}
