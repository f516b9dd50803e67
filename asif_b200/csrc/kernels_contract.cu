// kernels_contract.cu -- kernels of the segway and inverted-pendulum models, compiled WITH FMA contraction (-fmad=true).
//
// The double-integrator kernels (engine.cu) are compiled with -fmad=false so that every multiply and add rounds as in
// the reference's GCC x86-64 build; their trajectories, critical indices and constraint rows are bit-identical to that
// build.  The segway callbacks (examples/segway_implicit_tb.cpp:69-212) call sin, cos and tanh on every Euler step:
// CUDA's and glibc's implementations differ in the last bit, so bit identity with the reference is unattainable for
// this model whatever the contraction setting, and parity is judged by the tolerances of BASELINE.json (u within
// 1e-6 + 1e-5|u|, identical return codes, barrier values within 1e-9; tests/test_gpu_parity.py, test_gpu_rollout.py).
// Contraction perturbs results by the same last-bit amount and removes a quarter of the FP64 instructions:
// C5 filter 13.6 -> 12.1 ms per 1e6 states, fleet rollout 1.72 -> 1.53 s (B200, profiles/).  The same holds for the
// inverted pendulum of the implicit filter (examples/InvertedPendulum_Implicit.cpp:27-46, sin and cos per step); there
// the normalisation of the soft saturation also multiplies by 1/range instead of dividing (SAT_RECIP: the quotient
// only feeds two comparisons and the rare bevel leg, and the FP64 division was 14 % of the kernel's stall samples).
#include "engine_internal.cuh"

namespace asifb {

int launch_tb_segway(asif_engine *e, bool shipped, int64_t n, const double *x, const double *ud, double *ua, double *relax,
                     int32_t *rc, double *diag, cudaStream_t st)
{
	if (e->cfg.npBTSS != 4) // any other count: the run-time-count instantiation (centred and shipped backup sets)
		return shipped ? launch_tb<SegwayTB<false>, TB_NPBTSS_RUNTIME>(e, n, x, ud, ua, relax, rc, diag, st)
		               : launch_tb<SegwayTB<true>, TB_NPBTSS_RUNTIME>(e, n, x, ud, ua, relax, rc, diag, st);
	return shipped ? launch_tb<SegwayTB<false>, 4>(e, n, x, ud, ua, relax, rc, diag, st)
	               : launch_tb<SegwayTB<true>, 4>(e, n, x, ud, ua, relax, rc, diag, st);
}

int launch_tb_rollout_segway(asif_engine *e, bool shipped, int64_t n, int32_t steps, double dt, double *x, const double *ud,
                             double *ua, int32_t *rc, cudaStream_t st)
{
	if (e->cfg.npBTSS != 4)
		return shipped ? launch_tb_rollout<SegwayTB<false>, TB_NPBTSS_RUNTIME>(e, n, steps, dt, x, ud, ua, rc, st)
		               : launch_tb_rollout<SegwayTB<true>, TB_NPBTSS_RUNTIME>(e, n, steps, dt, x, ud, ua, rc, st);
	return shipped ? launch_tb_rollout<SegwayTB<false>, 4>(e, n, steps, dt, x, ud, ua, rc, st)
	               : launch_tb_rollout<SegwayTB<true>, 4>(e, n, steps, dt, x, ud, ua, rc, st);
}

int launch_implicit_ip(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                       double *diag, cudaStream_t st)
{
	using M = InvertedPendulumImplicit;
	if (e->cfg.npBTSS != 10) return launch_implicit_t<M, IMP_NPBTSS_RUNTIME, false, SAT_RECIP>(e, n, x, ud, ua, relax, rc, diag, st);
	return launch_implicit_t<M, 10, false, SAT_RECIP>(e, n, x, ud, ua, relax, rc, diag, st);
}

// ASIFimplicitRB on the same callbacks (zero-order-hold backup input, interval lower bound of the safety rows)
int launch_implicit_rb_ip(asif_engine *e, int64_t n, const double *x, const double *ud, double *ua, double *relax, int32_t *rc,
                          double *diag, cudaStream_t st)
{
	using M = InvertedPendulumImplicit;
	if (e->cfg.npBTSS != 10) return launch_implicit_t<M, IMP_NPBTSS_RUNTIME, true, SAT_RECIP>(e, n, x, ud, ua, relax, rc, diag, st);
	return launch_implicit_t<M, 10, true, SAT_RECIP>(e, n, x, ud, ua, relax, rc, diag, st);
}

} // namespace asifb
