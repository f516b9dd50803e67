/* qp_select_shim.h -- force-included (-include) when the reference's filter classes are compiled for
 * oracle/_ref/libasif_ref_b200.so (TEST INFRASTRUCTURE ONLY).
 *
 * Purpose: run the reference's OWN classes (ASIF, ASIFimplicitTB, ASIFimplicit, ... compiled from the unmodified
 * /root/reference/src where they lie) on the B200 QP backend, to show that ASIF::QPWrapperB200 drops in behind
 * ASIF::QPWrapperAbstract.  The reference selects its backend with
 *     switch (qpSolverType) { case QPSOLVER::OSQP: QPsolver_ = new QPWrapperOsqp(nv_, npTC_, diagonalCost); ... }
 * (include/qpwrappers.h:6-9, src/asif_implicit_tb.cpp:134-146 and the five sibling constructors).  A maintainer adds
 * `B200 = 1` to the enum and a second case.  Editing the sources is not allowed here, so the
 * same effect is obtained at compile time: inside the filter classes' translation units the NAME QPWrapperOsqp stands
 * for QPWrapperSelect, a QPWrapperAbstract that forwards every call either to the real QPWrapperOsqp (OSQP stand-in)
 * or to QPWrapperB200, whichever g_ref_qp_backend says when the filter object is constructed.  Not one line of the
 * reference is changed, and the real QPWrapperOsqp class (src/qpwrapper_osqp.cpp) is compiled without this header. */
#ifndef ASIF_QP_SELECT_SHIM_H
#define ASIF_QP_SELECT_SHIM_H
#include <string>
#include "qpwrapper_abstract.h"
#include "qpwrapper_osqp.h" /* the real class, under its real name (the include guard keeps qpwrappers.h from re-reading it) */
#include "asif_b200.hpp"    /* ASIF::QPWrapperB200; its own QPWrapperAbstract declaration is skipped (same include guard) */

extern "C" int g_ref_qp_backend; /* 0 = OSQP stand-in, 1 = B200 (set through ref_select_backend) */

namespace ASIF
{
	class QPWrapperSelect : public QPWrapperAbstract
	{
	public:
		QPWrapperSelect(const uint32_t nv, const uint32_t nc, const bool diagonalCost)
		    : QPWrapperAbstract(nv, nc, diagonalCost),
		      impl_(g_ref_qp_backend == 1 ? static_cast<QPWrapperAbstract *>(new QPWrapperB200(nv, nc, diagonalCost))
		                                  : static_cast<QPWrapperAbstract *>(new QPWrapperOsqp(nv, nc, diagonalCost)))
		{
		}
		virtual ~QPWrapperSelect(void) { delete impl_; }
		virtual int32_t initialize(const double H[], const double c[], const double A[], const double b[], const double lb[],
		                           const double ub[], const bool be[] = nullptr)
		{
			return impl_->initialize(H, c, A, b, lb, ub, be);
		}
		virtual int32_t updateCost(const double H[], const double c[]) { return impl_->updateCost(H, c); }
		virtual int32_t updateA(const double A[]) { return impl_->updateA(A); }
		virtual int32_t updateb(const double b[]) { return impl_->updateb(b); }
		virtual int32_t updateBounds(const double lb[], const double ub[]) { return impl_->updateBounds(lb, ub); }
		virtual int32_t solve(void) { return impl_->solve(); }
		virtual int32_t getSolution(double sol[]) { return impl_->getSolution(sol); }

	private:
		QPWrapperAbstract *impl_;
	};
}
#define QPWrapperOsqp QPWrapperSelect
#endif
