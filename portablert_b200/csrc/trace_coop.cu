// trace_coop.cu -- the cooperative-tail kernels (prt_trace_kernel.cuh: k_coop): every ray a fast
// binary kernel was still tracing long after the end of its batch is finished by a whole warp.
#include "prt_trace_kernel.cuh"

namespace prt {

template <bool WT, uint32_t M> struct CoopTable {
	static void fill(KernelFn (*t)[2]) {
		t[M][0] = k_coop<M, false, WT>;
		t[M][1] = k_coop<M, true, WT>;
		CoopTable<WT, M - 1>::fill(t);
	}
};
template <bool WT> struct CoopTable<WT, 0> {
	static void fill(KernelFn (*)[2]) {}
};
template <bool WT> static KernelFn coop_of(uint32_t mask, bool aos) {
	struct Filled {
		KernelFn t[32][2];
		Filled() { CoopTable<WT, 31>::fill(t); }
	};
	static const Filled table;
	return table.t[mask][aos ? 1 : 0];
}

KernelFn coop_kernel(uint32_t mask, bool aos, bool wt) {
	return wt ? coop_of<true>(mask, aos) : coop_of<false>(mask, aos);
}

} // namespace prt
