// compile check only:  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -c tools/prototypes/sector_walk_check.cu -o /dev/null
#include "sector_walk.cuh"

template <int P, int J>
__global__ void __launch_bounds__(P) sector_forest_kernel(rfsi_proto::SectorForest q, int T, int F, const uint32_t* __restrict__ ranks,
                                                          long long npix, double* __restrict__ mean)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* rs = reinterpret_cast<uint32_t*>(smem) + threadIdx.x;
    const long long p = (long long)blockIdx.x * P + threadIdx.x;
    if (p >= npix) return;
    for (int f = 0; f < F; ++f) rs[f * P] = ranks[(long long)f * npix + p];
    mean[p] = rfsi_proto::traverse_s<P, J>(q, 0, T, 0.0, rs, nullptr) / (double)T;
}
template __global__ void sector_forest_kernel<256, 8>(rfsi_proto::SectorForest, int, int, const uint32_t*, long long, double*);
template __global__ void sector_forest_kernel<256, 4>(rfsi_proto::SectorForest, int, int, const uint32_t*, long long, double*);
