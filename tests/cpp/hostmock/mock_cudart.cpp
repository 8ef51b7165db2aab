// TEST INFRASTRUCTURE ONLY — a stand-in for the handful of CUDA runtime calls zk_b200/csrc/api.cu makes, so that the
// host orchestration of the C ABI (round loop, transcript hops, buffer swaps, staging, virtual-rank exchanges) can be
// executed on a machine without a GPU against the oracle.  "Device" memory is host memory, every "launch" runs to
// completion inside the call (tests/cpp/hostmock/mock_kernels.cpp), so streams and events have nothing to order.
// Built only by tests/test_hostmock_orchestration.py into build/libzk_b200_hostmock.so; nothing under zk_b200/ links,
// loads or knows about it, and the product library keeps failing without a GPU (tests/test_abi_host.py).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

extern "C" {

cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    p->multiProcessorCount = 3;  // small on purpose: grid-stride paths get exercised
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "host mock"; }

static void* mock_alloc(size_t bytes) {
    const size_t rounded = ((bytes ? bytes : 1) + 255) / 256 * 256;
    return std::aligned_alloc(256, rounded);
}
cudaError_t cudaMalloc(void** p, size_t bytes) { *p = mock_alloc(bytes); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { *p = mock_alloc(bytes); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void** dev, void* host, unsigned) { *dev = host; return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
// no CUDA IPC in the mock: sharded contexts agree on the NCCL all-reduce (the fallback path of api.cu's mailbox_setup)
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
    if (n) std::memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; r++) std::memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
    return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }

// lets a harness make sure it is talking to the mock and not to the product library
__attribute__((visibility("default"))) int zk_b200_is_host_mock(void) { return 1; }
}
