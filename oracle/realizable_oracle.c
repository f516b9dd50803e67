/*
 * realizable_oracle.c -- CPU restatement of ASIFrealizable::filter for config 4
 * (InvertedPendulum interval dynamics + include/RealizableKernelData_100Hz_50pt.h), REDUCED form.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Reference: src/asif_realizable.cpp:314-352 (filter), :375-610 (updateConstraints).
 *  (i)   hFull_i = 1 - n_i.x for every facet                                             (:391-398)
 *  (ii)  critical facets, in index order, at most maxCriticalFacets: bounding box inflated by the
 *        uncertainty bounds (:407-415), then feasibility of { lambda in [0,1]^nx, sum lambda = 1,
 *        |V lambda - x| <= unc } which the reference poses as a QP for OSQP (:419-440).  For nx = 2
 *        the facet is the segment t v0 + (1-t) v1 and the test is an interval intersection in t.
 *  (iii) per critical facet and active constraint j: the interval Lie derivatives over the facet
 *        [LfLo, LfHi], [LgLo, LgHi] - x independent, so they come from a table computed once by the
 *        reference build itself (oracle_tables.h, tests/golden/make_tables.py) - enter the LP-dual
 *        rows (:503-521); eliminating the zero-cost multipliers exactly as for ASIFrobust
 *        (asif_oracle.c) leaves   LgLo u >= -LfLo  and  LgHi u >= -LfLo   (no relax variable).
 *  (iv)  barrier rows on the npSSmax smallest hFull:  Lgh u + eps >= -Lfh - relaxDes (h - relaxOffset)
 *        with the mid-point dynamics (:524-604); eps in [0, inf], cost relaxCost eps^2.
 *  (v)   rc -2 when no facet is critical and some h < 0 (:606-607, :324-326), else 1 / -1 from the QP.
 * relax[0] of the reference is a multiplier (non-unique): reported as 0 here.
 */
#include "asif_oracle.h"
#include "oracle_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NF ORACLE_RZ_NF
#define MAXCRIT ORACLE_RZ_MAXCRIT
#define MAXACT ORACLE_RZ_MAXACT

static int segment_meets_box(const double *v0, const double *v1, const double *x, const double *unc)
{
	/* p(t) = t v0 + (1-t) v1, t in [0,1];  x_k - unc_k <= p_k(t) <= x_k + unc_k for k = 0,1 */
	double tlo = 0.0, thi = 1.0;
	for (int k = 0; k < 2; k++) {
		const double d = v0[k] - v1[k], lo = x[k] - unc[k] - v1[k], hi = x[k] + unc[k] - v1[k];
		if (d == 0.0) {
			if (lo > 0.0 || hi < 0.0) return 0;
		} else {
			double a = lo / d, b = hi / d;
			if (a > b) {
				double t = a;
				a = b;
				b = t;
			}
			if (a > tlo) tlo = a;
			if (b < thi) thi = b;
		}
	}
	return tlo <= thi;
}

int32_t oracle_realizable_filter(const double *opts /* relaxDes, relaxOffset, relaxCost, unc0, unc1, npSSmax, pMin, pMax */,
                                 const double *x, const double *uDes, double *uAct, double *relax, double *diag)
{
	const double relaxDes = opts[0], relaxOffset = opts[1], relaxCost = opts[2];
	const double unc[2] = {opts[3], opts[4]};
	const int npSSmax = (int)opts[5];
	const double pMin = opts[6], pMax = opts[7];
	const double lbU = -1.5, ubU = 1.5, inf = 1e20;
	const int npSS = MAXCRIT * MAXACT;

	double hFull[NF];
	for (int i = 0; i < NF; i++) {
		hFull[i] = 1.;
		for (int j = 0; j < 2; j++) hFull[i] -= oracle_rz_normals[2 * i + j] * x[j];
	}
	int crit[MAXCRIT], nCrit = 0;
	for (int i = 0; i < NF && nCrit < MAXCRIT; i++) {
		const double *v0 = oracle_rz_vertices + 2 * oracle_rz_facet_vertices[2 * i];
		const double *v1 = oracle_rz_vertices + 2 * oracle_rz_facet_vertices[2 * i + 1];
		int potential = 1;
		for (int j = 0; j < 2; j++) {
			const double blo = v0[j] < v1[j] ? v0[j] : v1[j], bhi = v0[j] > v1[j] ? v0[j] : v1[j];
			if (x[j] < (blo - unc[j]) || x[j] > (bhi + unc[j])) {
				potential = 0;
				break;
			}
		}
		if (potential && segment_meets_box(v0, v1, x, unc)) crit[nCrit++] = i;
	}
	/* rows of the reduced problem in v = (u, eps) */
	const int nc = 2 * npSS + npSSmax, nv = 2;
	double *A = (double *)calloc((size_t)nc * nv, sizeof(double));
	double *b = (double *)calloc(nc, sizeof(double));
	double lie[4 * MAXCRIT * MAXACT];
	memset(lie, 0, sizeof(lie));
	int slot = 0;
	for (int c = 0; c < nCrit; c++)
		for (int j = 0; j < MAXACT; j++) {
			if (oracle_rz_facet_active[MAXACT * crit[c] + j] < 0) continue;
			const double *t = oracle_rz_facet_lie + 4 * (MAXACT * crit[c] + j); /* LfLo, LfHi, LgLo, LgHi */
			lie[4 * slot + 0] = t[2];
			lie[4 * slot + 1] = t[3];
			lie[4 * slot + 2] = t[0];
			lie[4 * slot + 3] = t[1];
			A[2 * slot] = t[2];
			A[2 * slot + 1] = t[3];
			b[2 * slot] = -t[0];
			b[2 * slot + 1] = -t[0];
			slot++;
		}
	/* barrier rows */
	int bar[8];
	double barL[8], barB[8];
	{
		/* npSSmax smallest hFull, ascending, ties keep the lower index (std::sort leaves ties unspecified) */
		int cnt = 0;
		for (int i = 0; i < NF; i++) {
			int pos = cnt;
			while (pos > 0 && hFull[i] < hFull[bar[pos - 1]]) pos--;
			if (pos < npSSmax) {
				int last = cnt < npSSmax ? cnt : npSSmax - 1;
				for (int j = last; j > pos; j--) bar[j] = bar[j - 1];
				bar[pos] = i;
				if (cnt < npSSmax) cnt++;
			}
		}
		/* mid-point dynamics: f = [x1, sin x0] (a point), g = [0, mid([pMin, pMax])] (:545-551) */
		const double f[2] = {x[1], sin(x[0])};
		const double gc = (pMax + pMin) / 2, gr = (pMax - pMin) / 2;
		const double g[2] = {0., ((gc - gr) + (gc + gr)) / 2};
		for (int i = 0; i < npSSmax; i++) {
			const double Dh0 = -oracle_rz_normals[2 * bar[i]], Dh1 = -oracle_rz_normals[2 * bar[i] + 1];
			double Lfh = 0.0, Lgh = 0.0;
			Lfh = Lfh + Dh0 * f[0];
			Lfh = Lfh + Dh1 * f[1];
			Lgh = Lgh + Dh0 * g[0];
			Lgh = Lgh + Dh1 * g[1];
			const int row = 2 * npSS + i;
			A[row] = Lgh;
			A[row + nc] = 1.0;
			b[row] = -Lfh - relaxDes * (hFull[bar[i]] - relaxOffset);
			barL[i] = Lgh;
			barB[i] = b[row];
		}
	}
	if (diag) {
		int o = 0;
		diag[o++] = (double)nCrit;
		for (int i = 0; i < MAXCRIT; i++) diag[o++] = i < nCrit ? (double)crit[i] : -1.0;
		for (int i = 0; i < npSSmax; i++) diag[o++] = (double)bar[i];
		for (int s = 0; s < npSS; s++) {
			diag[o++] = lie[4 * s + 0]; /* LgLo */
			diag[o++] = lie[4 * s + 1]; /* LgHi */
			diag[o++] = lie[4 * s + 2]; /* LfLo */
			diag[o++] = lie[4 * s + 3]; /* LfHi */
		}
		for (int i = 0; i < npSSmax; i++) {
			diag[o++] = barL[i];
			diag[o++] = barB[i];
		}
	}
	int anyNeg = 0;
	for (int i = 0; i < NF; i++) anyNeg |= (hFull[i] < 0.);
	int32_t rc;
	if (nCrit == 0 && anyNeg) {
		rc = -2; /* uAct untouched */
	} else {
		double H[4] = {1.0, 0.0, 0.0, npSSmax > 0 ? relaxCost : 1.0}, c[2] = {-2.0 * uDes[0], 0.0};
		double lb[2] = {lbU, 0.0}, ub[2] = {ubU, inf}, sol[2];
		int st = oracle_qp_solve(nv, nc, 1, H, c, A, b, lb, ub, 0, sol);
		if (st == 1) {
			uAct[0] = sol[0] > ubU ? ubU : (sol[0] < lbU ? lbU : sol[0]);
			relax[0] = 0.0;
			relax[1] = sol[1];
			rc = 1;
		} else
			rc = -1;
	}
	free(A);
	free(b);
	return rc;
}
