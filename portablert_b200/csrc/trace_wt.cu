// trace_wt.cu -- the traversal kernels of the opt-in WATERTIGHT mode (prt_math.cuh:
// woop_watertight; prt_b200_set_triangle_test / env PRT_B200_WATERTIGHT).  A separate translation
// unit so that its instantiations compile in parallel with the other kernel families.
#include "prt_trace_kernel.cuh"

namespace prt {

KernelFn trace_kernel_wt(uint32_t mask, bool aos) { return kernel_of<false, true, false>(mask, aos); }
KernelFn trace_kernel_wt_exact(uint32_t mask, bool aos) { return kernel_of<false, true, true>(mask, aos); }

} // namespace prt
