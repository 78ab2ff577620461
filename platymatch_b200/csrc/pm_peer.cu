// pm_peer.cu — peer-memory windows for the multi-GPU exchange steps (no reference analogue: the
// reference is single-process).
//
// The only bulk exchange of the path is "cost-matrix rows -> the rank that solves that matrix" (config 4,
// 20k x 20k: 1.44 GB per matrix).  Instead of computing a row shard locally and shipping it with a collective,
// the chi^2 kernel of every rank stores its rows DIRECTLY into the owner's cost matrix over NVLink: the owner's
// buffer is a cudaMalloc allocation exported with CUDA IPC and mapped into every peer process, and
// pm_chi2_cost simply gets the mapped address (+ row offset) as its output pointer.  The transfer overlaps
// the (compute-bound) kernel tile by tile and nothing but the owner ever holds the matrix.
// A stream-ordered barrier after the kernels (the tiny NCCL all-reduce the host side issues) makes the rows
// visible to the owner: peer stores are performed when the storing kernel has completed.
#include <string.h>
#include "pm_common.cuh"

extern "C" int pm_peer_alloc(size_t bytes, void **ptr) {
    PM_REQUIRE(ptr && bytes > 0, "bad arguments");
    PM_CUDA_TRY(cudaMalloc(ptr, bytes));          // plain cudaMalloc: exportable (a pooled / VMM block is not)
    return PM_OK;
}

extern "C" int pm_peer_free(void *ptr) {
    if (ptr) PM_CUDA_TRY(cudaFree(ptr));
    return PM_OK;
}

extern "C" int pm_peer_export(void *ptr, unsigned char *handle64) {
    PM_REQUIRE(ptr && handle64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == PM_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    PM_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, sizeof(h));
    return PM_OK;
}

extern "C" int pm_peer_open(const unsigned char *handle64, void **ptr) {
    PM_REQUIRE(ptr && handle64, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    PM_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PM_OK;
}

extern "C" int pm_peer_close(void *ptr) {
    if (ptr) PM_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return PM_OK;
}
