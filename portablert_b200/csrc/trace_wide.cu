// trace_wide.cu -- the traversal kernels over the compressed 4-wide nodes (prt_math.cuh: Node4),
// used for reordered incoherent batches on large scenes (trace.cu: launch_trace).
#include "prt_trace_kernel.cuh"

namespace prt {

KernelFn trace_kernel_wide(uint32_t mask, bool aos) { return kernel_of<true, false, false>(mask, aos); }

} // namespace prt
