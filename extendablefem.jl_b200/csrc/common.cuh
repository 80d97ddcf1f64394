// common.cuh -- shared declarations of libextfem_cuda.so (device structs, error helpers).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/extfem_cuda.h"

namespace extfem {

constexpr int MAXARGS = EXTFEM_MAXARGS;
constexpr int MAXOP = 16;      // max total operator length (input / result vector)
constexpr int MAXPARAMS = 40;  // qpinfo.params capacity (6x6 Hooke tensor + slack)
constexpr int MAXREGIONS = 16;
constexpr int MAXLOC = 64;     // max local rows / cols of one operator (all blocks)

// One (FESpace, FunctionOperator) pair on the device: the engine's FEEvaluator.
struct ArgDev {
    int ncomp, nscalar, op, nd;   // nd = ncomp*nscalar local dofs
    int oplen, opoff;             // operator length and offset in the input / result vector
    int locoff;                   // offset of this argument's dofs in the operator-local matrix
    int block;                    // pattern block (row block for test, column block otherwise)
    const int *celldofs;          // [ncells][nd], 0-based, block-local
    long long soloff;             // offset of the block in the global solution vector
    const double *refvals;        // [nq][nscalar]
    const double *refgrads;       // [nq][nscalar][dim]
};

struct OpDev {
    int dim, nq;
    int ntest, nansatz, nargs;
    ArgDev test[MAXARGS], ansatz[MAXARGS], args[MAXARGS];
    int NR, NC;                   // operator-local matrix: NC columns x NR rows
    int nin, nout;                // lengths of input (ansatz/args) and result (test) vectors
    int kernel_id, nparams;
    double params[MAXPARAMS];
    double factor, time, offdiag;
    const double *qw;             // [nq]
    const double *qx;             // [nq][dim]
    int nregions;
    int regions[MAXREGIONS];
    unsigned char coupling[MAXARGS * MAXARGS]; // [nansatz][ntest]
    int lump;
    const double *tabulated;
    // mesh
    long long ncells;
    const double *coords;         // [nnodes][dim]
    const int *cellnodes;         // [ncells][dim+1] 0-based
    const int *cellregions;
    const double *cellvolumes;
    const double *sol;            // global solution vector (args)
};

struct Ctx;
void set_error(Ctx *ctx, int code, const std::string &msg);

#define EXTFEM_CUDA_CHECK(ctx, call)                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            set_error(ctx, EXTFEM_ERR_CUDA,                                                       \
                      std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                          std::to_string(__LINE__) + ")");                                        \
            return EXTFEM_ERR_CUDA;                                                               \
        }                                                                                         \
    } while (0)

} // namespace extfem
