// kge_train_inst.cu -- the train-path kernels of ONE model (compiled once per model with -DKGE_TU_MODEL=<0..4>, so
// that the five sets of template instantiations build in parallel).
#include "kge_train_launch.cuh"

#ifndef KGE_TU_MODEL
#error "compile with -DKGE_TU_MODEL=<model id>"
#endif

namespace kge {
template int launch_rows_model<KGE_TU_MODEL>(bool, const RowArgs &, bool, int, size_t, void *, size_t, cudaStream_t);
template int launch_entity_model<KGE_TU_MODEL>(bool, const RowArgs &, const SplitWs &, int64_t, int64_t, int,
                                               cudaStream_t, int);
}  // namespace kge
