// matcher_tc.cu -- tensor-core (tcgen05) matcher.  Placeholder until the kernel lands.
#include "matcher_common.cuh"

namespace sba {

bool knn2_tensor_applicable(int, int, int) { return false; }

int knn2_tensor(sba_ctx*, const float*, int, const float*, int, int, Top2*)
{
    set_error("tensor-core matcher not built");
    return SBA_ERR_UNSUPPORTED;
}

}  // namespace sba
