/* tests/abi_driver.c -- a plain C caller of libextfem_cuda.so (include/extfem_cuda.h), fed with arrays in the layout the Julia
 * side holds them: Int64, 1-based, column-per-item (cellnodes[dim+1, ncells], celldofs[nd, ncells], bfacenodes, bfacedofs).
 * It performs TWO consecutive assemble_system!-style assemblies (src/solvers.jl:124-195: zero, NonlinearOperator,
 * BilinearOperator ON_BFACES, LinearOperator, penalties) of Example108's problem for two iterates and writes pattern, values
 * and right-hand sides to a file; tests/test_abi_c_driver.py compares them with the CPU oracle.  No Python, no torch, no
 * numpy between the caller and the library.
 *
 *   abi_driver <input.bin> <output.bin>
 *   input : int64 header {dim, ncells, nnodes, nd, ndofs, nbfaces, ndb, nbd(boundary dofs)} then
 *           coords f64[nnodes*dim], cellnodes i64[ncells*(dim+1)], celldofs i64[ncells*nd], bfacenodes i64[nbfaces*dim],
 *           bfaceregions i32[nbfaces], bfacedofs i64[nbfaces*ndb], bdofs i64[nbd], bvals f64[nbd], sol1 f64[ndofs], sol2 f64[ndofs]
 *   output: int64 nnz, colptr i64[ndofs+1], rowval i64[nnz], then for each of the two assemblies nzval f64[nnz], b f64[ndofs],
 *           residual f64[ndofs]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/extfem_cuda.h"

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int rc_ = (call);                                                                            \
        if (rc_ != EXTFEM_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, extfem_last_error(ctx)); return 2; } \
    } while (0)

static void *rd(FILE *f, size_t bytes)
{
    void *p = malloc(bytes ? bytes : 1);
    if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(3); }
    return p;
}

int main(int argc, char **argv)
{
    if (argc != 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    int64_t h[8];
    if (fread(h, 8, 8, f) != 8) return 3;
    const int dim = (int)h[0], nd = (int)h[3], ndb = (int)h[6];
    const int64_t ncells = h[1], nnodes = h[2], ndofs = h[4], nbfaces = h[5], nbd = h[7];
    double *coords = rd(f, nnodes * dim * 8);
    int64_t *cellnodes = rd(f, ncells * (dim + 1) * 8), *celldofs = rd(f, ncells * nd * 8), *bfacenodes = rd(f, nbfaces * dim * 8);
    int32_t *bfaceregions = rd(f, nbfaces * 4);
    int64_t *bfacedofs = rd(f, nbfaces * ndb * 8), *bdofs = rd(f, nbd * 8);
    double *bvals = rd(f, nbd * 8), *sol[2];
    sol[0] = rd(f, ndofs * 8); sol[1] = rd(f, ndofs * 8);
    fclose(f);

    extfem_ctx *ctx = NULL;
    CHECK(extfem_ctx_create(0, &ctx));
    int mesh, space, pattern;
    CHECK(extfem_mesh_set(ctx, dim, ncells, nnodes, coords, cellnodes, 8, NULL, NULL, &mesh));
    CHECK(extfem_mesh_set_bfaces(ctx, mesh, nbfaces, bfacenodes, 8, bfaceregions, NULL));
    CHECK(extfem_space_set(ctx, mesh, EXTFEM_FE_H1P2, 1, celldofs, 8, nd, ndofs, &space));
    CHECK(extfem_space_set_bfacedofs(ctx, space, bfacedofs, 8, ndb));
    CHECK(extfem_pattern_build(ctx, 1, &space, 1, &space, NULL, &pattern));
    int64_t nrows, ncols, nnz;
    CHECK(extfem_pattern_dims(ctx, pattern, &nrows, &ncols, &nnz));
    int64_t *colptr = malloc((ncols + 1) * 8), *rowval = malloc(nnz * 8);
    CHECK(extfem_pattern_get(ctx, pattern, colptr, rowval));

    /* the three operators of Example108 (examples/Example108_RobinBoundaryCondition.jl:59-63) as flat descriptors */
    extfem_opdesc nl, robin, rhs;
    memset(&nl, 0, sizeof nl); memset(&robin, 0, sizeof robin); memset(&rhs, 0, sizeof rhs);
    nl.ntest = 2; nl.test_op[0] = EXTFEM_OP_ID; nl.test_op[1] = EXTFEM_OP_GRAD;
    nl.nargs = 2; nl.args_op[0] = EXTFEM_OP_ID; nl.args_op[1] = EXTFEM_OP_GRAD;
    nl.kernel_id = extfem_kernel_id("rcd"); nl.factor = 1.0; nl.quadorder = -1;
    const double g[1] = {2.0};
    const int32_t reg1[1] = {1};
    robin.ntest = 1; robin.test_op[0] = EXTFEM_OP_ID; robin.nansatz = 1; robin.ansatz_op[0] = EXTFEM_OP_ID;
    robin.kernel_id = extfem_kernel_id("robin108"); robin.params = g; robin.nparams = 1; robin.factor = 1.0; robin.quadorder = -1;
    robin.regions = reg1; robin.nregions = 1; robin.entities = EXTFEM_ON_BFACES;
    rhs.ntest = 1; rhs.test_op[0] = EXTFEM_OP_ID; rhs.kernel_id = extfem_kernel_id("exp2x"); rhs.factor = 1.0; rhs.quadorder = -1;

    FILE *o = fopen(argv[2], "wb");
    fwrite(&nnz, 8, 1, o); fwrite(colptr, 8, ncols + 1, o); fwrite(rowval, 8, nnz, o);
    double *nzval = malloc(nnz * 8), *b = malloc(ndofs * 8), *res = malloc(ndofs * 8);
    for (int it = 0; it < 2; ++it) {
        CHECK(extfem_values_zero(ctx, pattern, 1, 1));                                  /* fill!(nzval, 0); fill!(b, 0)        */
        CHECK(extfem_assemble_nonlinear(ctx, pattern, &nl, sol[it], 1, NULL, NULL));     /* every operator ADDS to the system   */
        CHECK(extfem_assemble_bilinear(ctx, pattern, &robin, NULL, 1, NULL));
        CHECK(extfem_assemble_linear(ctx, pattern, &rhs, NULL, 1, NULL));
        CHECK(extfem_apply_penalties(ctx, pattern, nbd, bdofs, bvals, 1e30));            /* apply_penalties!                    */
        CHECK(extfem_apply_values(ctx, nbd, bdofs, bvals, sol[it], ndofs));              /* ... assemble_sol leg                */
        CHECK(extfem_values_get(ctx, pattern, nzval, b));
        CHECK(extfem_residual(ctx, pattern, sol[it], res));                              /* compute_nonlinear_residual!         */
        fwrite(nzval, 8, nnz, o); fwrite(b, 8, ndofs, o); fwrite(res, 8, ndofs, o);
    }
    fclose(o);
    CHECK(extfem_ctx_destroy(ctx));
    return 0;
}
