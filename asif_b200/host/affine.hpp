// affine.hpp -- a small host-side affine-arithmetic evaluator (header only).
//
// Why: ASIFrobust / ASIFrealizable take the user's dynamics as an *interval* callback
// (std::function<void(const interval_t*, interval_t*, interval_t*)>, include/asif_realizable.h:43-47, with
// interval_t = libaffa's AAF).  For ASIFrealizable the callback is only ever evaluated over the facets of the
// polytope kernel, which do not depend on the state - so the batched engine needs it once, at initialize, on the host
// (realizable_host.hpp builds the facet table with it).  The reference vendors libaffa 0.9.6 (LGPL) for this; this file
// is an independent implementation of the handful of operations the dynamics callbacks of the named models use,
// with the same definitions so that the resulting intervals agree with libaffa's:
//   value      x = c + sum_i x_i e_i,  e_i in [-1, 1]                 (lib/libaffa/src/aa_aaf.h:44-58)
//   interval   -> c = (hi+lo)/2, one fresh symbol with (hi-lo)/2        (aa_aafcommon.cpp:81-101)
//   + - scalar ops: exact on centre and coefficients                     (aa_aafarithm.cpp:35-201)
//   x * y      centre c_x c_y, coefficients c_x y_i + c_y x_i, fresh symbol rad(x) rad(y)   (aa_aafapprox.cpp:32-99)
//   sin(x)     least-squares line through 8 samples of the interval, max residual on a fresh symbol;
//              width < 1e-10 -> the point sin(mid); width >= 2 pi -> [-1, 1]    (aa_aaftrigo.cpp:42-135)
//   lo/hi      c -+ sum_i |x_i|                                          (aa_aafcommon.cpp:217-245)
// Unlike libaffa the symbol counter is per Context object, not a global, so evaluations are
// re-entrant.
#ifndef ASIF_B200_AFFINE_HPP
#define ASIF_B200_AFFINE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace ASIF
{
namespace b200
{
	class Affine
	{
	public:
		struct Context {
			uint32_t last = 0;
			uint32_t fresh(void) { return ++last; }
		};
		// the context used by values created without an explicit one (one per thread)
		static Context &defaultContext(void)
		{
			static thread_local Context ctx;
			return ctx;
		}

		Affine(const double v = 0.0) : c_(v) {}
		Affine(const double lo, const double hi, Context &ctx = defaultContext()) : c_((hi + lo) / 2)
		{
			t_.push_back(std::make_pair(ctx.fresh(), (hi - lo) / 2));
		}
		static Affine point(const double v, Context &ctx = defaultContext()) { return Affine(v, v, ctx); } // interval(v): zero-width symbol

		double center(void) const { return c_; }
		double rad(void) const
		{
			double s = 0;
			for (size_t i = 0; i < t_.size(); i++) s += (t_[i].second >= 0.0) ? t_[i].second : -t_[i].second;
			return s;
		}
		double lo(void) const { return c_ - rad(); }
		double hi(void) const { return c_ + rad(); }
		double mid(void) const { return lo() * 0.5 + hi() * 0.5; } // interval::mid, aa_interval.cpp:82-95

		Affine operator+(const Affine &p) const { return merge(p, c_ + p.c_, 1.0); }
		Affine operator-(const Affine &p) const { return merge(p, c_ - p.c_, -1.0); }
		Affine operator-(void) const
		{
			Affine r(*this);
			r.c_ = -r.c_;
			for (size_t i = 0; i < r.t_.size(); i++) r.t_[i].second = -r.t_[i].second;
			return r;
		}
		Affine operator*(const double k) const
		{
			Affine r(*this);
			r.c_ = k * c_;
			for (size_t i = 0; i < r.t_.size(); i++) r.t_[i].second = k * r.t_[i].second;
			return r;
		}
		// non-affine product: one fresh symbol carries rad(x) rad(y)
		Affine mul(const Affine &p, Context &ctx = defaultContext()) const
		{
			Affine r(c_ * p.c_);
			size_t a = 0, b = 0;
			while (a < t_.size() || b < p.t_.size()) {
				if (b == p.t_.size() || (a < t_.size() && t_[a].first < p.t_[b].first)) {
					r.t_.push_back(std::make_pair(t_[a].first, p.c_ * t_[a].second));
					a++;
				} else if (a == t_.size() || p.t_[b].first < t_[a].first) {
					r.t_.push_back(std::make_pair(p.t_[b].first, c_ * p.t_[b].second));
					b++;
				} else {
					r.t_.push_back(std::make_pair(t_[a].first, c_ * p.t_[b].second + p.c_ * t_[a].second));
					a++;
					b++;
				}
			}
			r.t_.push_back(std::make_pair(ctx.fresh(), rad() * p.rad()));
			return r;
		}
		Affine operator*(const Affine &p) const { return mul(p); }

		friend Affine operator*(const double k, const Affine &p) { return p * k; }
		friend Affine operator+(const double k, const Affine &p) { return Affine(k) + p; }
		friend Affine operator-(const double k, const Affine &p) { return Affine(k) - p; }

		friend Affine sin(const Affine &p) { return sinCtx(p, defaultContext()); }
		static Affine sinCtx(const Affine &p, Context &ctx)
		{
			const double a = p.lo(), b = p.hi(), w = b - a;
			const double twoPi = 2 * (4 * std::atan(1.0));
			if (w >= twoPi) return Affine(-1.0, 1.0, ctx);
			if (w < 1e-10) {
				const double tmp = std::sin(a * 0.5 + b * 0.5);
				return Affine(tmp, tmp, ctx);
			}
			const int NPTS = 8;
			double x[NPTS], y[NPTS];
			x[0] = a;
			y[0] = std::sin(a);
			x[NPTS - 1] = b;
			y[NPTS - 1] = std::sin(b);
			const double pas = w / (NPTS - 1);
			for (int i = 1; i < NPTS - 1; i++) {
				x[i] = x[i - 1] + pas;
				y[i] = std::sin(x[i]);
			}
			double xm = 0, ym = 0;
			for (int i = 0; i < NPTS; i++) {
				xm = xm + x[i];
				ym = ym + y[i];
			}
			xm = xm / NPTS;
			ym = ym / NPTS;
			double temp2 = 0, alpha = 0;
			for (int i = 0; i < NPTS; i++) {
				const double temp1 = x[i] - xm;
				alpha += y[i] * temp1;
				temp2 += temp1 * temp1;
			}
			alpha = alpha / temp2;
			const double dzeta = ym - alpha * xm;
			double delta = 0;
			for (int i = 0; i < NPTS; i++) delta = std::max(delta, std::fabs(y[i] - (dzeta + alpha * x[i])));
			// z = alpha x + dzeta, error delta on a fresh symbol (affine constructor, aa_aafarithm.cpp:230-259)
			Affine r(alpha * p.c_ + dzeta);
			for (size_t i = 0; i < p.t_.size(); i++) r.t_.push_back(std::make_pair(p.t_[i].first, alpha * p.t_[i].second));
			r.t_.push_back(std::make_pair(ctx.fresh(), delta));
			return r;
		}
		// cos(x) = sin(x + pi/2) (aa_aaftrigo.cpp:138-146)
		friend Affine cos(const Affine &p) { return sin(p + Affine(2 * std::atan(1.0))); }

	private:
		Affine merge(const Affine &p, const double c, const double sign) const
		{
			Affine r(c);
			size_t a = 0, b = 0;
			while (a < t_.size() || b < p.t_.size()) {
				if (b == p.t_.size() || (a < t_.size() && t_[a].first < p.t_[b].first)) {
					r.t_.push_back(t_[a]);
					a++;
				} else if (a == t_.size() || p.t_[b].first < t_[a].first) {
					r.t_.push_back(std::make_pair(p.t_[b].first, sign > 0 ? p.t_[b].second : -p.t_[b].second));
					b++;
				} else {
					r.t_.push_back(std::make_pair(t_[a].first, sign > 0 ? t_[a].second + p.t_[b].second : t_[a].second - p.t_[b].second));
					a++;
					b++;
				}
			}
			return r;
		}
		double c_;
		std::vector<std::pair<uint32_t, double> > t_; // (symbol, coefficient), symbols ascending
	};
} // namespace b200
} // namespace ASIF
#endif
