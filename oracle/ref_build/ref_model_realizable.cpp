/* ref_model_realizable.cpp -- ASIFrealizable on the InvertedPendulum interval dynamics of
 * examples/InvertedPendulum_RealizableSampled.cpp:46-53 and the in-tree polytope kernel
 * include/RealizableKernelData_100Hz_50pt.h (config 4; TEST INFRASTRUCTURE ONLY).
 * Deviation D5: the example loads its kernel from a CSV that is not in the repository and its
 * realizable loop is dead (for(i=3;i<3;...), :216); the in-tree table is used instead, with
 * uncertaintyBounds {0.032, 0.027} (:23), npSSmax = 2 and Options relaxCost 50 / relaxDes 1 (:232-233). */
#include "ref_std_includes.h"

namespace ex_ip_realizable {
#include "examples/InvertedPendulum_RealizableSampled.cpp"
}
namespace kd_realizable_100hz_50pt {
#include "RealizableKernelData_100Hz_50pt.h"
}

namespace {

struct RealizableAccess : ASIF::ASIFrealizable {
	using ASIF::ASIFrealizable::ASIFrealizable;
	const double *A() const { return A_; }
	const double *b() const { return b_; }
	const kernel_t &kern() const { return kernel_; }
	uint32_t nc() const { return nc_; }
	uint32_t nv() const { return nv_; }
	uint32_t npSS() const { return npSS_; }
	uint32_t npSSmax() const { return npSSmax_; }
	void dyn(const interval_t *x, interval_t *f, interval_t *g) const { dynamics_(x, f, g); }
};

/* diag: [nCrit, critFacet[maxCrit] (-1 absent), barrierFacet[npSSmax],
 *        per slot s < maxCrit*maxAct: LgLo, LgHi, LfLo, LfHi   (read back from A_, src/asif_realizable.cpp:503-521),
 *        per barrier row: Lgh, b                                 (:596-603)] */
struct IpRealizable : RefFilter {
	RealizableAccess *f;
	int maxCrit, maxAct, npSS, npSSmax;
	IpRealizable(const double *opts, int n_opts)
	{
		const ASIF::ASIFrealizable::kernel_t &k = kd_realizable_100hz_50pt::kernel;
		double unc[2] = {0.032, 0.027};
		ASIF::ASIFrealizable::Options o;
		o.relaxCost = 50.0;
		o.relaxDes = 1.;
		npSSmax = 2;
		if (opts && n_opts >= 6) { /* [relaxDes, relaxOffset, relaxCost, unc0, unc1, npSSmax] */
			o.relaxDes = opts[0];
			o.relaxOffset = opts[1];
			o.relaxCost = opts[2];
			unc[0] = opts[3];
			unc[1] = opts[4];
			npSSmax = (int)opts[5];
		}
		f = new RealizableAccess(2, 1, unc, k, ex_ip_realizable::dynamics, (uint32_t)npSSmax);
		f->initialize(ex_ip_realizable::lb, ex_ip_realizable::ub, o);
		maxCrit = (int)k.maxCriticalFacets;
		maxAct = (int)k.maxActiveConstraints;
		npSS = maxCrit * maxAct;
		nx = 2; nu = 1; n_relax = 2;
		nc = 2 * npSS + npSSmax; nv = 2; /* dimensions of the REDUCED problem */
		n_diag = 1 + maxCrit + npSSmax + 4 * npSS + 2 * npSSmax;
	}
	~IpRealizable() { delete f; }
	int32_t filter(const double *x, const double *u_des, double *u_act, double *relax, double *diag) override
	{
		AAF::set_default(0);
		int32_t rc = f->filter(x, u_des, u_act, relax);
		if (diag) {
			const double *A = f->A(), *b = f->b();
			const int ncF = (int)f->nc(), nvF = (int)f->nv();
			int o = 0;
			diag[o++] = (double)f->nCriticalFacets_;
			for (int i = 0; i < maxCrit; i++) diag[o++] = i < (int)f->nCriticalFacets_ ? (double)f->criticalFacets_[i] : -1.0;
			for (int i = 0; i < npSSmax; i++) diag[o++] = (double)f->criticalBarrierFacets_[i];
			int col = 1; /* iColLambda = nu */
			for (int s = 0; s < npSS; s++, col += 4) {
				const int row = 3 * s;
				diag[o++] = A[row + (col + 0) * ncF];
				diag[o++] = -A[row + (col + 2) * ncF];
				diag[o++] = A[row + (col + 1) * ncF];
				diag[o++] = -A[row + (col + 3) * ncF];
			}
			for (int i = 0; i < npSSmax; i++) {
				diag[o++] = A[3 * npSS + i + 0 * ncF];
				diag[o++] = b[3 * npSS + i];
			}
			(void)nvF;
		}
		return rc;
	}
	void plant(const double *x, double *fo, double *go) override { ex_ip_realizable::dynamicsExact(x, fo, go); }
};
} // namespace

RefFilter *make_ip_realizable(const double *opts, int n_opts) { return new IpRealizable(opts, n_opts); }

/* Kernel geometry and the x-independent facet table, evaluated with the reference's own types:
 * for facet i and its j-th active constraint, [LfLo, LfHi, LgLo, LgHi] of
 * Lfh = sum_k f_k(xFaceInt_i) * Dh_k, Lgh = sum_k g_k(xFaceInt_i) * Dh_k, Dh = -normal(active_j)
 * exactly as src/asif_realizable.cpp:470-500 computes them (same operand order, same libaffa calls). */
extern "C" int ref_realizable_export(void *h, int32_t *dims /*[nVertices,nFacets,maxCrit,maxAct]*/, double *vertices,
                                     double *normals, int32_t *facet_vertices, int32_t *facet_active, double *facet_lie)
{
	IpRealizable *r = dynamic_cast<IpRealizable *>((RefFilter *)h);
	if (!r) return -1;
	const ASIF::ASIFrealizable::kernel_t &k = r->f->kern();
	const int nV = (int)k.vertices.size(), nF = (int)k.facets.size();
	dims[0] = nV; dims[1] = nF; dims[2] = r->maxCrit; dims[3] = r->maxAct;
	if (!vertices) return 0;
	for (int i = 0; i < nV; i++)
		for (int j = 0; j < 2; j++) vertices[2 * i + j] = k.vertices[i][j];
	AAF::set_default(0);
	for (int i = 0; i < nF; i++) {
		for (int j = 0; j < 2; j++) {
			normals[2 * i + j] = k.facets[i].normal[j];
			facet_vertices[2 * i + j] = (int32_t)k.facets[i].verticesIdx[j];
		}
		for (int j = 0; j < r->maxAct; j++) {
			const bool have = j < (int)k.facets[i].activeConstraintsSet.size();
			facet_active[r->maxAct * i + j] = have ? (int32_t)k.facets[i].activeConstraintsSet[j] : -1;
			double *out = facet_lie + 4 * (r->maxAct * i + j);
			out[0] = out[1] = out[2] = out[3] = 0.0;
			if (!have) continue;
			const std::vector<double> &normal = k.facets[k.facets[i].activeConstraintsSet[j]].normal;
			interval_t DhInt[2] = {interval(-normal[0]), interval(-normal[1])};
			interval_t fI[2], gI[2];
			r->f->dyn(k.facets[i].xFaceInt.data(), fI, gI);
			interval_t Lfh = 0.;
			for (int kk = 0; kk < 2; kk++) Lfh = Lfh + fI[kk] * DhInt[kk];
			interval_t Lgh = 0.;
			for (int kk = 0; kk < 2; kk++) Lgh = Lgh + gI[kk] * DhInt[kk];
			const interval a = Lfh.convert(), c = Lgh.convert();
			out[0] = a.left(); out[1] = a.right(); out[2] = c.left(); out[3] = c.right();
		}
	}
	return 0;
}
