// trace_wt.cu -- the traversal kernels of the opt-in WATERTIGHT mode (prt_math.cuh:
// woop_watertight; prt_b200_set_triangle_test / env PRT_B200_WATERTIGHT).  A separate translation
// unit so that the 62 extra instantiations compile in parallel with trace.cu.
#include "prt_trace_kernel.cuh"

namespace prt {

template <uint32_t M> struct TableWT {
	static void fill(KernelFn (*t)[2]) {
		t[M][0] = k_trace<M, false, false, false, true>;
		t[M][1] = k_trace<M, true, false, false, true>;
		TableWT<M - 1>::fill(t);
	}
};
template <> struct TableWT<0> {
	static void fill(KernelFn (*)[2]) {}
};

KernelFn trace_kernel_wt(uint32_t mask, bool aos, bool count) {
	static KernelFn table[32][2];
	static bool ready = false;
	if (!ready) {
		TableWT<31>::fill(table);
		ready = true;
	}
	if (count)
		return k_trace<PRT_TAG_ALL, false, true, false, true>;
	return table[mask][aos ? 1 : 0];
}

} // namespace prt
