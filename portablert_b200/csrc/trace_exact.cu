// trace_exact.cu -- the EXACT traversal kernels: the reference's ray_box_intersect arithmetic
// (bvh.hpp:195-222) at every node.  They trace the rays the fast kernels set aside (non-finite or
// overflowing components), and every ray when PRT_B200_FAST_BOXES=0.  Also the instrumented
// (counting) variants used for the roofline's bytes-per-ray figure.
#include "prt_trace_kernel.cuh"

namespace prt {

KernelFn trace_kernel_exact(uint32_t mask, bool aos) { return kernel_of<false, false, true>(mask, aos); }

KernelFn trace_kernel_count(bool wide, bool wt) {
	if (wt)
		return k_trace<PRT_TAG_ALL, false, true, false, true, false>;
	return wide ? (KernelFn)k_trace<PRT_TAG_ALL, false, true, true, false, false>
	            : (KernelFn)k_trace<PRT_TAG_ALL, false, true, false, false, false>;
}

} // namespace prt
