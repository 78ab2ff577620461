// pm_labels.cu — label image -> nucleus centroids and sizes.
//
// Reference: platymatch/_dock_widget.py:497-521 (inside EstimateTransform._click_run): for every
// non-zero id of np.unique(label_image), z, y, x = np.where(image == id); centroid = (mean z, mean y,
// mean x); size = anisotropy * len(z) — one full pass over the volume PER ID (O(ids x voxels)).
//
// Here: ONE streaming pass over the volume (HBM-bound: 4 B per voxel read once).  Every thread
// loads four consecutive voxels; lanes of a warp that hold the same non-zero id form a group
// (cooperative-groups labeled_partition = match.any), the group's voxel count and coordinate sums
// are reduced in registers and its leader issues four 64-bit atomic adds on the id's accumulator
// {count, sum z, sum y, sum x}.  The sums are exact integers, so mean = sum / count rounds once and
// equals np.mean of the integer coordinates bit for bit (sums stay far below 2^53).
// A second small kernel compacts the non-empty ids in ascending order (np.unique order).
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include "pm_common.cuh"

namespace cg = cooperative_groups;

template <typename T>
__global__ void __launch_bounds__(256) pm_label_max_kernel(const T *__restrict__ labels, size_t n_vox,
                                                           unsigned *__restrict__ max_id) {
    unsigned m = 0u;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vox; v += stride) {
        const long long id = (long long)labels[v];
        m = max(m, id > 0 ? (unsigned)id : 0u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(max_id, m);
}

template <typename T>
__global__ void __launch_bounds__(256) pm_label_accumulate_kernel(const T *__restrict__ labels, int nz, int ny, int nx,
                                                                  unsigned table_size,
                                                                  unsigned long long *__restrict__ acc) {
    const size_t n_vox = (size_t)nz * ny * nx;
    const size_t plane = (size_t)ny * nx;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; base < n_vox; base += stride) {
        long long ids[4];
        if (base + 3 < n_vox && (reinterpret_cast<size_t>(labels) & 15) == 0) {   // one 16- / 8-byte load
            if (sizeof(T) == 4) {
                const int4 q = *reinterpret_cast<const int4 *>(labels + base);
                ids[0] = q.x; ids[1] = q.y; ids[2] = q.z; ids[3] = q.w;
            } else {
                const ushort4 q = *reinterpret_cast<const ushort4 *>(labels + base);
                ids[0] = q.x; ids[1] = q.y; ids[2] = q.z; ids[3] = q.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) ids[e] = (base + e < n_vox) ? (long long)labels[base + e] : 0;
        }
        if ((ids[0] | ids[1] | ids[2] | ids[3]) == 0) continue;        // background (most of the volume)
        const unsigned z0 = (unsigned)(base / plane);
        const size_t rem = base - (size_t)z0 * plane;
        const unsigned y0 = (unsigned)(rem / nx), x0 = (unsigned)(rem - (size_t)y0 * nx);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (ids[e] <= 0 || (unsigned long long)ids[e] >= table_size) continue;
            unsigned x = x0 + e, y = y0, z = z0;
            while (x >= (unsigned)nx) { x -= nx; if (++y >= (unsigned)ny) { y = 0; ++z; } }
            const cg::coalesced_group active = cg::coalesced_threads();
            const cg::coalesced_group grp = cg::labeled_partition(active, (unsigned)ids[e]);
            const unsigned sz = cg::reduce(grp, z, cg::plus<unsigned>());
            const unsigned sy = cg::reduce(grp, y, cg::plus<unsigned>());
            const unsigned sx = cg::reduce(grp, x, cg::plus<unsigned>());
            if (grp.thread_rank() == 0) {
                unsigned long long *a = acc + (size_t)ids[e] * 4;
                atomicAdd(a + 0, (unsigned long long)grp.size());
                atomicAdd(a + 1, (unsigned long long)sz);
                atomicAdd(a + 2, (unsigned long long)sy);
                atomicAdd(a + 3, (unsigned long long)sx);
            }
        }
    }
}

// one CTA: ids with a non-zero count, ascending, -> ids / centroids (z, y, x) / sizes
__global__ void __launch_bounds__(1024) pm_label_finalize_kernel(const unsigned long long *__restrict__ acc,
                                                                 unsigned table_size, double anisotropy, int capacity,
                                                                 int32_t *__restrict__ ids, double *__restrict__ centroids,
                                                                 double *__restrict__ sizes, int32_t *__restrict__ n_out) {
    __shared__ int s_scan[33];
    __shared__ int s_base;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_base = 0;
    __syncthreads();
    for (unsigned b = 1; b < table_size; b += 1024) {
        const unsigned id = b + t;
        const unsigned long long cnt = (id < table_size) ? acc[(size_t)id * 4] : 0ull;
        const bool has = cnt != 0ull;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        if (t == 0) {
            int a = s_base;
            for (int w = 0; w < 32; ++w) { const int c = s_scan[w]; s_scan[w] = a; a += c; }
            s_scan[32] = a;
        }
        __syncthreads();
        if (has) {
            const int k = s_scan[warp] + __popc(bal & ((1u << lane) - 1u));
            if (k < capacity) {
                const double c = (double)cnt;
                ids[k] = (int32_t)id;
                centroids[3 * (size_t)k + 0] = (double)acc[(size_t)id * 4 + 1] / c;     // np.mean(z)
                centroids[3 * (size_t)k + 1] = (double)acc[(size_t)id * 4 + 2] / c;
                centroids[3 * (size_t)k + 2] = (double)acc[(size_t)id * 4 + 3] / c;
                sizes[k] = anisotropy * c;                                              // anisotropy * len(z)  (:508)
            }
        }
        __syncthreads();
        if (t == 0) s_base = s_scan[32];
        __syncthreads();
    }
    if (t == 0) n_out[0] = s_base;
}

template <typename T>
static int pm_label_run(const T *labels, int nz, int ny, int nx, unsigned table_size, double anisotropy,
                        unsigned long long *acc, int capacity, int32_t *ids, double *centroids, double *sizes,
                        int32_t *n_out, cudaStream_t s) {
    const size_t n_vox = (size_t)nz * ny * nx;
    PM_CUDA_TRY(cudaMemsetAsync(acc, 0, (size_t)table_size * 4 * sizeof(unsigned long long), s));
    int dev = 0, sms = 0;
    PM_CUDA_TRY(cudaGetDevice(&dev));
    PM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    size_t want = (n_vox + 1023) / 1024;
    const int blocks = (int)(want < (size_t)sms * 8 ? (want ? want : 1) : (size_t)sms * 8);   // 8 resident CTAs per SM
    pm_label_accumulate_kernel<T><<<blocks, 256, 0, s>>>(labels, nz, ny, nx, table_size, acc);
    PM_LAUNCH_CHECK();
    pm_label_finalize_kernel<<<1, 1024, 0, s>>>(acc, table_size, anisotropy, capacity, ids, centroids, sizes, n_out);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

template <typename T>
static int pm_label_max(const T *labels, size_t n_vox, unsigned *max_id, cudaStream_t s) {
    PM_CUDA_TRY(cudaMemsetAsync(max_id, 0, sizeof(unsigned), s));
    int dev = 0, sms = 0;
    PM_CUDA_TRY(cudaGetDevice(&dev));
    PM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    pm_label_max_kernel<T><<<sms * 8, 256, 0, s>>>(labels, n_vox, max_id);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

extern "C" int pm_label_max_id(const void *labels, int dtype, size_t n_voxels, uint32_t *max_id, void *stream) {
    PM_REQUIRE(labels && max_id, "null pointer");
    PM_REQUIRE(dtype == PM_LABEL_I32 || dtype == PM_LABEL_U16, "dtype must be PM_LABEL_I32 or PM_LABEL_U16");
    if (dtype == PM_LABEL_I32) return pm_label_max((const int32_t *)labels, n_voxels, max_id, pm_stream(stream));
    return pm_label_max((const uint16_t *)labels, n_voxels, max_id, pm_stream(stream));
}

extern "C" size_t pm_label_workspace_bytes(uint32_t table_size) { return (size_t)table_size * 4 * sizeof(unsigned long long); }

extern "C" int pm_label_centroids(const void *labels, int dtype, int nz, int ny, int nx, uint32_t table_size,
                                  double anisotropy, int capacity, int32_t *ids, double *centroids, double *sizes,
                                  int32_t *n_out, void *workspace, size_t workspace_bytes, void *stream) {
    PM_REQUIRE(labels && ids && centroids && sizes && n_out && workspace, "null pointer");
    PM_REQUIRE(dtype == PM_LABEL_I32 || dtype == PM_LABEL_U16, "dtype must be PM_LABEL_I32 or PM_LABEL_U16");
    PM_REQUIRE(nz >= 1 && ny >= 1 && nx >= 1 && nz < (1 << 26) && ny < (1 << 26) && nx < (1 << 26), "bad volume shape");
    PM_REQUIRE(table_size >= 1 && capacity >= 0, "bad table size / capacity");
    if (workspace_bytes < pm_label_workspace_bytes(table_size)) {
        pm_set_error("pm_label_centroids: workspace too small");
        return PM_ERR_WORKSPACE;
    }
    unsigned long long *acc = (unsigned long long *)workspace;
    if (dtype == PM_LABEL_I32)
        return pm_label_run((const int32_t *)labels, nz, ny, nx, table_size, anisotropy, acc, capacity, ids, centroids,
                            sizes, n_out, pm_stream(stream));
    return pm_label_run((const uint16_t *)labels, nz, ny, nx, table_size, anisotropy, acc, capacity, ids, centroids, sizes,
                        n_out, pm_stream(stream));
}
