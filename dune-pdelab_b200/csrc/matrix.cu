// matrix.cu — placeholder, filled in below
#include "common.cuh"
namespace pdb {
struct MatrixPlan {};
MatrixPlan* matrix_plan_create(const DevParams&, FemPlan*, cudaStream_t) { throw Error("matrix path not built yet"); }
void matrix_plan_destroy(MatrixPlan* p) { delete p; }
void matrix_pattern_size(MatrixPlan*, int, uint64_t*, uint64_t*) {}
int matrix_pattern_write(MatrixPlan*, int, void*, bool, void*, bool, bool, cudaStream_t) { return 0; }
int matrix_assemble(MatrixPlan*, int, double*, bool, bool, int*, cudaStream_t) { return 0; }
int matrix_mv(MatrixPlan*, int, const double*, const double*, double*, cudaStream_t) { return 0; }
}  // namespace pdb
