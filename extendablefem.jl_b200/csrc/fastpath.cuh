// fastpath.cuh -- specialised kernels for the headline configurations.
#pragma once
#include "common.cuh"

namespace extfem {

template <int DIM>
__global__ void cell_volumes_kernel(long long ncells, const double *__restrict__ coords, const int *__restrict__ cellnodes,
                                    double *__restrict__ vol)
{
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int *cn = cellnodes + c * (DIM + 1);
    double A[DIM][DIM];
    const double *p0 = coords + (size_t)cn[0] * DIM;
    for (int r = 0; r < DIM; ++r) {
        const double *pr = coords + (size_t)cn[r + 1] * DIM;
        for (int d = 0; d < DIM; ++d) A[d][r] = pr[d] - p0[d];
    }
    double det;
    if (DIM == 1) det = A[0][0];
    else if (DIM == 2) det = A[0][0] * A[1 % DIM][1 % DIM] - A[0][1 % DIM] * A[1 % DIM][0];
    else {
        constexpr int I1 = 1 % DIM, I2 = 2 % DIM;
        det = A[0][0] * (A[I1][I1] * A[I2][I2] - A[I1][I2] * A[I2][I1]) + A[0][I1] * (A[I1][I2] * A[I2][0] - A[I1][0] * A[I2][I2]) +
              A[0][I2] * (A[I1][0] * A[I2][I1] - A[I1][I1] * A[I2][0]);
    }
    const double fact = DIM == 1 ? 1.0 : (DIM == 2 ? 2.0 : 6.0);
    vol[c] = fabs(det) / fact;
}

static inline int launch_cell_volumes(cudaStream_t st, int dim, long long ncells, const double *coords, const int *cellnodes, double *vol)
{
    unsigned g = (unsigned)((ncells + 255) / 256);
    if (dim == 1) cell_volumes_kernel<1><<<g, 256, 0, st>>>(ncells, coords, cellnodes, vol);
    else if (dim == 2) cell_volumes_kernel<2><<<g, 256, 0, st>>>(ncells, coords, cellnodes, vol);
    else cell_volumes_kernel<3><<<g, 256, 0, st>>>(ncells, coords, cellnodes, vol);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

} // namespace extfem
