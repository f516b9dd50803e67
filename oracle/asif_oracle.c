/*
 * asif_oracle.c -- CPU restatement of the reference filter classes (TEST INFRASTRUCTURE ONLY).
 *
 * Follows the reference's structure on purpose (store the whole backup trajectory, sort the
 * time indices, then assemble rows) - the CUDA kernels stream instead, so the two
 * implementations share neither code nor control flow.  See asif_oracle.h for the rules on
 * who may call this.
 */
#include "asif_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NX ORACLE_NXMAX
#define NU ORACLE_NUMAX

/* include/asif_utils.h:22-43 (column-major, k ascending, accumulator starts at 0.0) */
static void mat_mul(const double *A, int nlA, int ncA, const double *B, int ncB, double *AB)
{
	for (int i = 0; i < nlA; i++)
		for (int j = 0; j < ncB; j++) {
			int idx = i + j * nlA;
			AB[idx] = 0.0;
			for (int k = 0; k < ncA; k++) AB[idx] = AB[idx] + A[i + k * nlA] * B[k + j * ncA];
		}
}

/* include/asif_utils.h:45-62 */
static void mat_vec(const double *A, int nlA, int ncA, const double *b, double *Ab)
{
	for (int i = 0; i < nlA; i++) {
		Ab[i] = 0.0;
		for (int k = 0; k < ncA; k++) Ab[i] = Ab[i] + A[i + k * nlA] * b[k];
	}
}

/* include/asif_utils.h:64-73 */
static double vec_norm(const double *v, int len)
{
	double tmp = 0;
	for (int i = 0; i < len; i++) tmp += v[i] * v[i];
	return sqrt(tmp);
}

static double min_elem(const double *v, int n)
{
	double m = v[0];
	for (int i = 1; i < n; i++)
		if (v[i] < m) m = v[i];
	return m;
}

/* The filter(x, H, c, ...) overloads (src/asif_implicit_tb.cpp:252-363, src/asif.cpp:153-210, src/asif_implicit.cpp:296-356):
 * updateH copies the caller's nu x nu block into H_ (:748-762) and the caller's c replaces c_ entirely.  The batch
 * entry below sets this per state; the filters apply it right before the solve. */
static struct {
	const double *H; /* nu x nu column-major or NULL */
	const double *c; /* nv entries or NULL */
	/* ASIF::filter(x, uDes, uAct, Lfh, Lgh[, relax]) (src/asif.cpp:125-141,287-292): caller-supplied Lie derivatives,
	 * npSSmax entries / npSSmax x nu column-major, replacing the computed ones row for row */
	const double *Lfh, *Lgh;
} g_cost;

static void apply_cost_override(int nv, int nu, double *H, double *c)
{
	if (g_cost.H)
		for (int j = 0; j < nu; j++)
			for (int i = 0; i < nu; i++) H[i + j * nv] = g_cost.H[i + j * nu];
	if (g_cost.c)
		for (int i = 0; i < nv; i++) c[i] = g_cost.c[i];
}

/* src/asif_implicit_tb.cpp:821-830 (same in every class) */
static void input_saturate(const oracle_model *md, double *u)
{
	for (int i = 0; i < md->nu; i++) {
		if (u[i] > md->ub[i]) u[i] = md->ub[i];
		else if (u[i] < md->lb[i]) u[i] = md->lb[i];
	}
}

/* src/asif_implicit_tb.cpp:764-819 */
static void input_saturate_soft(const oracle_model *md, double r, const double *u, double *uSat, double *DuSat)
{
	const double alpha = M_PI / 8;
	const double beta = M_PI / 4;
	for (int i = 0; i < md->nu; i++) {
		double mi = md->lb[i], ma = md->ub[i];
		double range = ma - mi;
		double middle = (ma + mi) / 2;
		double uc = 2 * (u[i] - middle) / range;
		double bevelL = r * tan(alpha);
		double bevelStart = 1 - cos(beta) * bevelL;
		double bevelStop = 1 + bevelL;
		double bevelXc = bevelStop;
		double bevelYc = 1 - r;
		if (uc >= bevelStop) {
			uSat[i] = ma;
			DuSat[i] = 0;
		} else if (uc <= -bevelStop) {
			uSat[i] = mi;
			DuSat[i] = 0;
		} else if (uc <= bevelStart && uc >= -bevelStart) {
			uSat[i] = u[i];
			DuSat[i] = 1;
		} else if (uc > bevelStart) {
			uSat[i] = sqrt(r * r - (uc - bevelXc) * (uc - bevelXc)) + bevelYc;
			DuSat[i] = (bevelXc - uc) / sqrt(r * r - (uc - bevelXc) * (uc - bevelXc));
			uSat[i] = 0.5 * uSat[i] * range + middle;
		} else if (uc < -bevelStart) {
			uSat[i] = -sqrt(r * r - (uc + bevelXc) * (uc + bevelXc)) - bevelYc;
			DuSat[i] = (bevelXc + uc) / sqrt(r * r - (uc + bevelXc) * (uc + bevelXc));
			uSat[i] = 0.5 * uSat[i] * range + middle;
		} else {
			DuSat[i] = 1;
			uSat[i] = u[i];
		}
	}
}

/* src/asif_implicit_tb.cpp:833-897 (identical in asif_implicit.cpp:751-815) */
static void backup_cl_dynamics(const oracle_model *md, double sat_sharpness, const double *x, double *fCL, double *DfCL)
{
	const int nx = md->nx, nu = md->nu;
	double f[NX], g[NX * NU], u[NU], Du[NU * NX], uSat[NU], DuSat[NU];
	md->backup_controller(x, u, Du);
	input_saturate_soft(md, sat_sharpness, u, uSat, DuSat);
	if (md->dynamics_with_gradient) {
		double d[NX * NX];
		md->dynamics_with_gradient(x, uSat, f, g, d);
		for (int i = 0; i < nx; i++)
			for (int j = 0; j < nx; j++) {
				int idx = i + j * nx;
				DfCL[idx] = d[idx];
				for (int k = 0; k < nu; k++) DfCL[idx] += g[i + k * nx] * DuSat[k] * Du[k + j * nu];
			}
	} else {
		double Df[NX * NX], Dg[NX * NU * NX];
		md->dynamics(x, f, g);
		md->dynamics_gradients(x, Df, Dg);
		for (int i = 0; i < nx; i++)
			for (int j = 0; j < nx; j++) {
				int idx = i + j * nx;
				DfCL[idx] = Df[idx];
				for (int k = 0; k < nu; k++)
					DfCL[idx] += Dg[i + k * nx + j * nx * nu] * uSat[k] + g[i + k * nx] * DuSat[k] * Du[k + j * nu];
			}
	}
	mat_vec(g, nx, nu, uSat, fCL);
	for (int i = 0; i < nx; i++) fCL[i] += f[i];
}

/* src/asif_implicit_tb.cpp:899-909 */
static void ode_rhs(const oracle_model *md, double sat_sharpness, const double *X, double *Xdot)
{
	double DfCL[NX * NX];
	backup_cl_dynamics(md, sat_sharpness, X, Xdot, DfCL);
	mat_mul(DfCL, md->nx, md->nx, X + md->nx, md->nx, Xdot + md->nx);
}

/* ASIFimplicitRB: zero-order-hold backup controller, src/asif_implicit_robust.cpp:878-952.  The hold is refreshed
 * when t >= t_last + backContDt - 0.0001 (and t_last is reset to -1 while t <= backTrajDt, i.e. on the first
 * step of every trajectory); the saturation acts on the HELD input.  Quirk kept: the fused-gradient branch uses
 * the held Du (:921), the split branch the CURRENT Du (:939). */
typedef struct {
	double backContDt, backTrajDt, t_last;
	double u[NU], Du[NU * NX];
} zoh_state;

static void backup_cl_dynamics_zoh(const oracle_model *md, double sat_sharpness, const double *x, double t, zoh_state *z,
                                   double *fCL, double *DfCL)
{
	const int nx = md->nx, nu = md->nu;
	double f[NX], g[NX * NU], u[NU], Du[NU * NX], uSat[NU], DuSat[NU];
	md->backup_controller(x, u, Du);
	if (t <= z->backTrajDt) z->t_last = -1.;
	if (t >= (z->t_last + z->backContDt - 0.0001)) {
		for (int i = 0; i < nu; i++) z->u[i] = u[i];
		for (int i = 0; i < nu * nx; i++) z->Du[i] = Du[i];
		z->t_last = t;
	}
	input_saturate_soft(md, sat_sharpness, z->u, uSat, DuSat);
	if (md->dynamics_with_gradient) {
		double d[NX * NX];
		md->dynamics_with_gradient(x, uSat, f, g, d);
		for (int i = 0; i < nx; i++)
			for (int j = 0; j < nx; j++) {
				int idx = i + j * nx;
				DfCL[idx] = d[idx];
				for (int k = 0; k < nu; k++) DfCL[idx] += g[i + k * nx] * DuSat[k] * z->Du[k + j * nu];
			}
	} else {
		double Df[NX * NX], Dg[NX * NU * NX];
		md->dynamics(x, f, g);
		md->dynamics_gradients(x, Df, Dg);
		for (int i = 0; i < nx; i++)
			for (int j = 0; j < nx; j++) {
				int idx = i + j * nx;
				DfCL[idx] = Df[idx];
				for (int k = 0; k < nu; k++)
					DfCL[idx] += Dg[i + k * nx + j * nx * nu] * uSat[k] + g[i + k * nx] * DuSat[k] * Du[k + j * nu];
			}
	}
	mat_vec(g, nx, nu, uSat, fCL);
	for (int i = 0; i < nx; i++) fCL[i] += f[i];
}

/* open-loop f,g at the current state as the filter sees them: for the fused-gradient
 * constructor the library wraps dynamicsWithGradient(x, u = 0) (src/asif_implicit_tb.cpp:79-86) */
static void open_loop_dynamics(const oracle_model *md, const double *x, double *f, double *g)
{
	if (md->dynamics_with_gradient) {
		double u0[NU] = {0.0}, d[NX * NX];
		md->dynamics_with_gradient(x, u0, f, g, d);
	} else
		md->dynamics(x, f, g);
}

/* Euler trajectory with sensitivities, src/asif_implicit_tb.cpp:421-429,464-487.
 * traj: N x (nx+nx*nx), hFull: N x npSS, DhFull: N x npSS*nx, hFullMin: N, t: N */
static void integrate_backup_trajectory_z(const oracle_model *md, double sat_sharpness, double dt, int N, const double *x,
                                          double *traj, double *t, double *hFull, double *DhFull, double *hFullMin,
                                          zoh_state *z);
static void integrate_backup_trajectory(const oracle_model *md, double sat_sharpness, double dt, int N, const double *x,
                                        double *traj, double *t, double *hFull, double *DhFull, double *hFullMin)
{
	integrate_backup_trajectory_z(md, sat_sharpness, dt, N, x, traj, t, hFull, DhFull, hFullMin, 0);
}

/* z != NULL: ASIFimplicitRB's loop (src/asif_implicit_robust.cpp:547-570), whose rhs receives t = i*backTrajDt (:550) */
static void integrate_backup_trajectory_z(const oracle_model *md, double sat_sharpness, double dt, int N, const double *x,
                                          double *traj, double *t, double *hFull, double *DhFull, double *hFullMin,
                                          zoh_state *z)
{
	const int nx = md->nx, ns = nx + nx * nx, npSS = md->npSS;
	double *X0 = traj;
	for (int i = 0; i < ns; i++) X0[i] = 0.0;
	for (int i = 0; i < nx; i++) X0[i] = x[i];
	for (int i = nx; i < ns; i += nx + 1) X0[i] = 1.0;
	t[0] = 0.0;
	md->safety_set(X0, hFull, DhFull);
	hFullMin[0] = min_elem(hFull, npSS);
	for (int i = 1; i < N; i++) {
		double *Xp = traj + (size_t)(i - 1) * ns, *Xi = traj + (size_t)i * ns;
		t[i] = t[i - 1] + dt;
		if (z) {
			double DfCL[NX * NX];
			backup_cl_dynamics_zoh(md, sat_sharpness, Xp, (double)(uint32_t)i * dt, z, Xi, DfCL);
			mat_mul(DfCL, md->nx, md->nx, Xp + md->nx, md->nx, Xi + md->nx);
		} else
			ode_rhs(md, sat_sharpness, Xp, Xi);
		for (int k = 0; k < ns; k++) Xi[k] = Xi[k] * dt;
		for (int k = 0; k < ns; k++) Xi[k] = Xi[k] + Xp[k];
		md->safety_set(Xi, hFull + (size_t)i * npSS, DhFull + (size_t)i * npSS * nx);
		hFullMin[i] = min_elem(hFull + (size_t)i * npSS, npSS);
	}
}

/* indices of the k smallest keys among [0, n) in ascending key order; ties -> lower index first
 * (the reference uses std::sort, whose tie order is unspecified: src/asif_implicit_tb.cpp:539) */
static void k_smallest(const double *key, int n, int k, int *out)
{
	int cnt = 0;
	for (int i = 0; i < n; i++) {
		int pos = cnt;
		while (pos > 0 && key[i] < key[out[pos - 1]]) pos--;
		if (pos < k) {
			int last = cnt < k ? cnt : k - 1;
			for (int j = last; j > pos; j--) out[j] = out[j - 1];
			out[pos] = i;
			if (cnt < k) cnt++;
		}
	}
}

/* ------------------------------------------------------------------------------------------
 * ASIFimplicitTB
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	double relaxCost, relaxSafeLb, relaxTTS, relaxMinOrtho, backTrajHorizon, backTrajExtend, backTrajDt,
	    backTrajMinOrtho, satSharpness, inf;
	int npBTSS;
} tb_options;

static void tb_default_options(int cfg, tb_options *o)
{
	/* include/asif_implicit_tb.h:19-33 */
	o->relaxCost = 50.0; o->relaxSafeLb = 5.0; o->relaxTTS = 5.0; o->relaxMinOrtho = 5.0;
	o->backTrajHorizon = 1.0; o->backTrajExtend = 0.05; o->backTrajDt = 0.01; o->backTrajMinOrtho = 0.01;
	o->satSharpness = 0.1; o->inf = 1e20; o->npBTSS = 4;
	if (cfg == ORACLE_CFG_DI_IMPLICIT_TB) { /* examples/DoubleIntegrator_implicit_tb.cpp:90-95 */
		o->backTrajHorizon = 2.0; o->backTrajDt = 0.001; o->relaxSafeLb = 10.0; o->relaxTTS = 5.0; o->relaxMinOrtho = 5.0;
	} else if (cfg == ORACLE_CFG_SEGWAY_TB) { /* examples/segway_implicit_tb.cpp:223-230 */
		o->backTrajHorizon = 3.0; o->backTrajDt = 0.01; o->relaxCost = 10; o->relaxSafeLb = 2.0;
		o->relaxTTS = 30.0; o->relaxMinOrtho = 60.0; o->backTrajMinOrtho = 0.001;
	}
}

static void tb_parse_options(int cfg, const double *v, int n, tb_options *o)
{
	tb_default_options(cfg, o);
	if (!v || n < 9) return;
	o->relaxCost = v[0]; o->relaxSafeLb = v[1]; o->relaxTTS = v[2]; o->relaxMinOrtho = v[3];
	o->backTrajHorizon = v[4]; o->backTrajExtend = v[5]; o->backTrajDt = v[6]; o->backTrajMinOrtho = v[7];
	o->satSharpness = v[8];
	if (n >= 13 && v[12] >= 1.0 && v[12] <= 16.0) o->npBTSS = (int)v[12]; /* constructor argument npBTSS */
}

/* src/asif_implicit_tb.cpp:177-182 */
static int tb_npbt(tb_options *o)
{
	int npBT = (int)round(o->backTrajHorizon * (1.0 + o->backTrajExtend) / o->backTrajDt) + 1;
	if (npBT < o->npBTSS) {
		npBT = o->npBTSS;
		o->backTrajDt = o->backTrajHorizon * (1.0 + o->backTrajExtend) / (double)(npBT - 1);
	}
	return npBT;
}

typedef struct {
	double TTS, BTorthoBS, hSafetyNow, hBackupEnd;
	int critIdx[16];
	int nCrit;
} tb_diag;

/* src/asif_implicit_tb.cpp:407-714.  Returns 1, or -1 when the backup set is never reached. */
static int tb_update_constraints(const oracle_model *md, const tb_options *o, int N, const double *x, double *A,
                                 double *b, tb_diag *dg, double *x_end)
{
	const int nx = md->nx, nu = md->nu, npSS = md->npSS, npBTSS = o->npBTSS;
	const int ns = nx + nx * nx, npTC = npBTSS * npSS + 2, nv = nu + 1;
	double f[NX], g[NX * NU];
	open_loop_dynamics(md, x, f, g); /* :416-418 */

	double *traj = (double *)malloc(sizeof(double) * (size_t)N * ns);
	double *t = (double *)malloc(sizeof(double) * N);
	double *hFull = (double *)malloc(sizeof(double) * (size_t)N * npSS);
	double *DhFull = (double *)malloc(sizeof(double) * (size_t)N * npSS * nx);
	double *hFullMin = (double *)malloc(sizeof(double) * N);
	integrate_backup_trajectory(md, o->satSharpness, o->backTrajDt, N, x, traj, t, hFull, DhFull, hFullMin);
	if (x_end) memcpy(x_end, traj + (size_t)(N - 1) * ns, sizeof(double) * nx);

	/* time to safety, :490-536 */
	double hBSnm1 = -1.0, hBS[1] = {0.0}, DhBS[NX], DDhBS[NX * NX];
	int BSHit = 0, idxHit = 0;
	double cosTilde[1] = {0.0}, fClBS[NX], DfClBS[NX * NX], den1 = 0, den2 = 0, den = 0;
	const double *btX = 0;
	for (int i = 1; i < N; i++) {
		if (hBSnm1 < 0.0) {
			md->backup_set(traj + (size_t)i * ns, hBS, DhBS, DDhBS);
			if (hBS[0] >= 0.0) {
				idxHit = i;
				BSHit = 1;
				btX = traj + (size_t)i * ns;
				backup_cl_dynamics(md, o->satSharpness, btX, fClBS, DfClBS);
				mat_mul(DhBS, 1, nx, fClBS, 1, cosTilde);
				den1 = vec_norm(DhBS, nx);
				den2 = vec_norm(fClBS, nx);
				den = den1 * den2;
				dg->BTorthoBS = cosTilde[0] / den;
				if (dg->BTorthoBS > 2.0 * o->backTrajMinOrtho) break;
			}
		}
		hBSnm1 = hBS[0];
	}
	if (!BSHit) {
		dg->BTorthoBS = 0;
		free(traj); free(t); free(hFull); free(DhFull); free(hFullMin);
		return -1;
	}

	/* critical points, :538-552 */
	int crit[16];
	int nsel = npBTSS > idxHit ? idxHit + 1 : npBTSS;
	k_smallest(hFullMin, idxHit + 1, nsel, crit);
	dg->nCrit = nsel;
	for (int i = 0; i < nsel; i++) dg->critIdx[i] = crit[i];

	/* contingent cone rows, :554-641 */
	double *h = (double *)calloc(npTC, sizeof(double));
	double *Dh = (double *)calloc((size_t)npTC * nx, sizeof(double));
	double DhSSDx[64];
	for (int idx = 0; idx < npBTSS; idx++) {
		if (idx > idxHit) {
			for (int i = 0; i < npSS; i++) {
				h[idx * npSS + i] = 1.0;
				for (int j = 0; j < nx; j++) Dh[(idx * npSS + i) + j * npTC] = 0.0;
			}
		} else {
			int c = crit[idx];
			memcpy(h + idx * npSS, hFull + (size_t)c * npSS, sizeof(double) * npSS);
			mat_mul(DhFull + (size_t)c * npSS * nx, npSS, nx, traj + (size_t)c * ns + nx, nx, DhSSDx);
			for (int i = 0; i < npSS; i++)
				for (int j = 0; j < nx; j++) Dh[(idx * npSS + i) + j * npTC] = DhSSDx[i + j * npSS];
		}
	}
	dg->TTS = t[idxHit];
	double hReach = o->backTrajHorizon - t[idxHit];
	const double *btDX = btX + nx;
	double DhBSDx[NX];
	mat_mul(DhBS, 1, nx, btDX, nx, DhBSDx);
	h[npBTSS * npSS] = hReach;
	for (int i = 0; i < nx; i++) Dh[(npBTSS * npSS) + i * npTC] = DhBSDx[i] / cosTilde[0];

	double denSquared = den * den;
	h[npBTSS * npSS + 1] = dg->BTorthoBS - o->backTrajMinOrtho;
	double DxHit[NX * NX];
	mat_mul(fClBS, nx, 1, DhBSDx, nx, DxHit);
	for (int i = 0; i < nx * nx; i++) DxHit[i] = btDX[i] - DxHit[i];
	double Dnum[NX] = {0.0}, Dden1[NX] = {0.0}, Dden2[NX] = {0.0}, Dden[NX];
	for (int i = 0; i < nx; i++)
		for (int k = 0; k < nx; k++) {
			double temp1 = 0.0, temp2 = 0.0;
			for (int l = 0; l < nx; l++) {
				temp1 += DDhBS[k + l * nx] * DxHit[l + i * nx];
				temp2 += DfClBS[k + l * nx] * DxHit[l + i * nx];
			}
			double temp3 = DhBS[k] * temp2;
			double temp4 = temp1 * fClBS[k];
			Dden1[i] += temp3;
			Dden2[i] += temp4;
			Dnum[i] += temp3 + temp4;
		}
	for (int i = 0; i < nx; i++) Dden[i] = den2 * Dden1[i] / den1 + den1 * Dden2[i] / den2;
	for (int i = 0; i < nx; i++) {
		DhBSDx[i] = (Dnum[i] * den - cosTilde[0] * Dden[i]) / denSquared;
		Dh[(npBTSS * npSS + 1) + i * npTC] = DhBSDx[i];
	}

	/* Lfh, Lgh, A, b  :643-674 */
	double *Lfh = (double *)malloc(sizeof(double) * npTC);
	double *Lgh = (double *)malloc(sizeof(double) * (size_t)npTC * nu);
	mat_vec(Dh, npTC, nx, f, Lfh);
	mat_mul(Dh, npTC, nx, g, nu, Lgh);
	for (int i = 0; i < npTC * nv; i++) A[i] = 0.0; /* last two rows of the relax column stay 0 */
	for (int i = 0; i < npTC; i++)
		for (int j = 0; j < nu; j++) A[i + j * npTC] = Lgh[i + j * npTC];
	for (int i = 0; i < npBTSS * npSS; i++) A[i + nu * npTC] = h[i];
	for (int i = 0; i < npTC; i++) b[i] = -Lfh[i];
	b[npBTSS * npSS] -= o->relaxTTS * h[npBTSS * npSS];
	b[npBTSS * npSS + 1] -= o->relaxMinOrtho * h[npBTSS * npSS + 1];

	free(traj); free(t); free(hFull); free(DhFull); free(hFullMin);
	free(h); free(Dh); free(Lfh); free(Lgh);
	return 1;
}

/* src/asif_implicit_tb.cpp:261-363 with the cost of :198-212,735-746 */
static int32_t tb_filter(const oracle_model *md, const tb_options *o, int N, const double *x, const double *uDes,
                         double *uAct, double *relax, double *diag)
{
	const int nx = md->nx, nu = md->nu, npSS = md->npSS, npBTSS = o->npBTSS;
	const int npTC = npBTSS * npSS + 2, nv = nu + 1;
	double H[ORACLE_QP_NVMAX * ORACLE_QP_NVMAX] = {0.0}, c[ORACLE_QP_NVMAX], lb[ORACLE_QP_NVMAX], ub[ORACLE_QP_NVMAX];
	for (int i = 0; i < nu; i++) {
		H[i + i * nv] = 1.0;
		c[i] = -2.0 * uDes[i];
		lb[i] = md->lb[i];
		ub[i] = md->ub[i];
	}
	H[(nv - 1) + (nv - 1) * nv] = o->relaxCost;
	c[nv - 1] = -2.0 * o->relaxCost * o->relaxSafeLb;
	lb[nv - 1] = o->relaxSafeLb;
	ub[nv - 1] = o->inf;

	double *A = (double *)calloc((size_t)npTC * nv, sizeof(double));
	double *b = (double *)calloc(npTC, sizeof(double));
	tb_diag dg;
	memset(&dg, 0, sizeof(dg));
	double hb[1], Dhb[NX], DDhb[NX * NX], hs[16], Dhs[16 * NX], sol[ORACLE_QP_NVMAX], Du[NU * NX];
	md->backup_set(x, hb, Dhb, DDhb);
	md->safety_set(x, hs, Dhs);
	dg.hSafetyNow = min_elem(hs, npSS);
	int32_t rc;
	if (hb[0] >= 0) {
		/* updateConstraintsTrivial :716-733 */
		for (int i = 0; i < npTC; i++) b[i] = -o->inf;
		dg.TTS = 0.0;
		dg.BTorthoBS = 1.0;
		apply_cost_override(nv, nu, H, c);
		int st = oracle_qp_solve(nv, npTC, 1, H, c, A, b, lb, ub, 0, sol);
		if (st == 1) {
			memcpy(uAct, sol, sizeof(double) * nu);
			input_saturate(md, uAct);
			*relax = sol[nu];
			rc = 2;
		} else {
			md->backup_controller(x, uAct, Du);
			input_saturate(md, uAct);
			rc = -1;
		}
	} else {
		double xe[NX];
		if (tb_update_constraints(md, o, N, x, A, b, &dg, xe) == 1) {
			md->backup_set(xe, hb, Dhb, DDhb);
			dg.hBackupEnd = hb[0];
			apply_cost_override(nv, nu, H, c);
			int st = oracle_qp_solve(nv, npTC, 1, H, c, A, b, lb, ub, 0, sol);
			if (st == 1) {
				memcpy(uAct, sol, sizeof(double) * nu);
				input_saturate(md, uAct);
				*relax = sol[nu];
				rc = 1;
			} else {
				md->backup_controller(x, uAct, Du);
				input_saturate(md, uAct);
				rc = st;
			}
		} else {
			md->backup_controller(x, uAct, Du);
			input_saturate(md, uAct);
			rc = -3;
		}
	}
	if (diag) {
		diag[0] = dg.TTS;
		diag[1] = dg.BTorthoBS;
		diag[2] = dg.hSafetyNow;
		diag[3] = dg.hBackupEnd;
		for (int i = 0; i < npBTSS; i++) diag[4 + i] = i < dg.nCrit ? (double)dg.critIdx[i] : -1.0;
		memcpy(diag + 4 + npBTSS, A, sizeof(double) * npTC * nv);
		memcpy(diag + 4 + npBTSS + npTC * nv, b, sizeof(double) * npTC);
	}
	free(A);
	free(b);
	return rc;
}

/* ------------------------------------------------------------------------------------------
 * ASIF (explicit), src/asif.cpp:64-110 (initialize), :176-210 (filter), :233-312 (updateConstraints)
 * npSSmax = npSS (the default of every named config)
 * ---------------------------------------------------------------------------------------- */
static int32_t explicit_filter(const oracle_model *md, double relaxLb, double relaxCost, int npSSmax, const double *x,
                               const double *uDes, double *uAct, double *relax, double *diag)
{
	const int nx = md->nx, nu = md->nu, npSS = md->npSS, nv = nu + 1;
	const int nsel = (npSSmax > 0 && npSSmax < npSS) ? npSSmax : npSS, nc = nsel;
	double hFull[16], DhFull[16 * NX], h[16], Dh[16 * NX], f[NX], g[NX * NU], Lfh[16], Lgh[16 * NU];
	md->safety_set(x, hFull, DhFull);
	md->dynamics(x, f, g);
	if (nsel < npSS) { /* :250-268: the npSSmax smallest h (std::sort; ties: lower index first here) */
		int idx[16];
		k_smallest(hFull, npSS, nsel, idx);
		for (int i = 0; i < nsel; i++) {
			h[i] = hFull[idx[i]];
			for (int j = 0; j < nx; j++) Dh[i + j * nsel] = DhFull[idx[i] + j * npSS];
		}
	} else {
		memcpy(h, hFull, sizeof(double) * npSS);
		memcpy(Dh, DhFull, sizeof(double) * npSS * nx);
	}
	mat_vec(Dh, nsel, nx, f, Lfh);
	mat_mul(Dh, nsel, nx, g, nu, Lgh);
	if (g_cost.Lfh && g_cost.Lgh) { /* use_custom_ineq_, :287-292 */
		for (int i = 0; i < nsel; i++) Lfh[i] = g_cost.Lfh[i];
		for (int i = 0; i < nsel * nu; i++) Lgh[i] = g_cost.Lgh[i];
	}
	double A[16 * ORACLE_QP_NVMAX], b[16];
	for (int i = 0; i < nsel; i++) {
		for (int j = 0; j < nu; j++) A[i + j * nc] = Lgh[i + j * nsel];
		A[i + nu * nc] = h[i];
		b[i] = -Lfh[i];
	}
	double H[ORACLE_QP_NVMAX * ORACLE_QP_NVMAX] = {0.0}, c[ORACLE_QP_NVMAX], lb[ORACLE_QP_NVMAX], ub[ORACLE_QP_NVMAX], sol[ORACLE_QP_NVMAX];
	for (int i = 0; i < nu; i++) {
		H[i + i * nv] = 1.0;
		c[i] = -2.0 * uDes[i];
		lb[i] = md->lb[i];
		ub[i] = md->ub[i];
	}
	H[nu + nu * nv] = relaxCost;
	c[nu] = -2.0 * relaxCost * relaxLb;
	lb[nu] = relaxLb; /* :88-91 both bounds pinned to relaxLb */
	ub[nu] = relaxLb;
	apply_cost_override(nv, nu, H, c);
	int st = oracle_qp_solve(nv, nc, 1, H, c, A, b, lb, ub, 0, sol);
	if (diag) {
		memcpy(diag, A, sizeof(double) * nc * nv);
		memcpy(diag + nc * nv, b, sizeof(double) * nc);
	}
	if (st == 1) {
		memcpy(uAct, sol, sizeof(double) * nu);
		input_saturate(md, uAct);
		*relax = sol[nu];
		return 1;
	}
	return -1;
}

/* ------------------------------------------------------------------------------------------
 * ASIFimplicit, src/asif_implicit.cpp:194-266 (initialize), :305-356 (filter), :403-611 (updateConstraints)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	double relaxCost, relaxReachLb, relaxSafeLb, backTrajHorizon, backTrajDt, satSharpness, inf;
	int npBTSS;
	/* ASIFimplicitRB (include/asif_implicit_robust.h:22-38): rb != 0 selects that class */
	int rb;
	double backContDt, x_unc[NX];
} imp_options;

/* Learned residual, include/asif_learning_utils.h:34-155: two MLPs (two ReLU hidden layers each) on the input
 * [x ; Dh_index_[0..nx-1] ; 0...]; the drift net's first output is added to Lfh[0], the actuation net's first nu
 * outputs to Lgh[0..nu-1] (:148-154).  Dh_index_ holds DhSS(x_c) Q_c (npSS x nx, column-major) of the FIRST critical
 * point (src/asif_implicit.cpp:533-537), so its first nx entries are column 0 of rows 0..nx-1 - a quirk that is kept.
 * Weights are column-major [rows x cols] as matrixVectorMultiply reads them (include/asif_utils.h:46-62). */
static struct {
	int on;
	uint32_t d[8];
	double *blob;
	const double *w1d, *b1d, *w2d, *b2d, *w3d, *b3d, *w1a, *b1a, *w2a, *b2a, *w3a, *b3a;
} g_learn;

int oracle_set_learning(const uint32_t *dims, const double *blob)
{
	free(g_learn.blob);
	memset(&g_learn, 0, sizeof(g_learn));
	if (!dims || !blob) return 0;
	memcpy(g_learn.d, dims, sizeof(g_learn.d));
	const uint32_t *d = dims;
	const size_t nd = (size_t)d[2] * d[0] + d[2] + (size_t)d[4] * d[2] + d[4] + (size_t)d[6] * d[4] + d[6];
	const size_t na = (size_t)d[3] * d[1] + d[3] + (size_t)d[5] * d[3] + d[5] + (size_t)d[7] * d[5] + d[7];
	g_learn.blob = (double *)malloc(sizeof(double) * (nd + na));
	memcpy(g_learn.blob, blob, sizeof(double) * (nd + na));
	const double *p = g_learn.blob;
	g_learn.w1d = p; p += (size_t)d[2] * d[0];
	g_learn.b1d = p; p += d[2];
	g_learn.w2d = p; p += (size_t)d[4] * d[2];
	g_learn.b2d = p; p += d[4];
	g_learn.w3d = p; p += (size_t)d[6] * d[4];
	g_learn.b3d = p; p += d[6];
	g_learn.w1a = p; p += (size_t)d[3] * d[1];
	g_learn.b1a = p; p += d[3];
	g_learn.w2a = p; p += (size_t)d[5] * d[3];
	g_learn.b2a = p; p += d[5];
	g_learn.w3a = p; p += (size_t)d[7] * d[5];
	g_learn.b3a = p; p += d[7];
	g_learn.on = 1;
	return 0;
}

/* driftNN / actNN, include/asif_learning_utils.h:34-119 */
static void mlp3(const double *w1, const double *b1, const double *w2, const double *b2, const double *w3, const double *b3,
                 int din, int dh1, int dh2, int dout, const double *in, double *out)
{
	double *o1 = (double *)malloc(sizeof(double) * (dh1 + 1)), *o2 = (double *)malloc(sizeof(double) * (dh2 + 1));
	double *o3 = (double *)malloc(sizeof(double) * (dout + 1));
	mat_vec(w1, dh1, din, in, o1);
	for (int i = 0; i < dh1; i++) o1[i] = fmax(0., o1[i] + b1[i]);
	mat_vec(w2, dh2, dh1, o1, o2);
	for (int i = 0; i < dh2; i++) o2[i] = fmax(0., o2[i] + b2[i]);
	mat_vec(w3, dout, dh2, o2, o3);
	for (int i = 0; i < dout; i++) out[i] = o3[i] + b3[i];
	free(o1); free(o2); free(o3);
}

/* update_weights, include/asif_learning_utils.h:121-155 */
static void learned_residual(const double *x, int nx, const double *DhIndex, double *Lfh, double *Lgh, int nu)
{
	const uint32_t *d = g_learn.d;
	double *din = (double *)calloc(d[0] + 1, sizeof(double)), *ain = (double *)calloc(d[1] + 1, sizeof(double));
	double *dout = (double *)calloc(d[6] + 1, sizeof(double)), *aout = (double *)calloc(d[7] + 1, sizeof(double));
	for (int i = 0; i < nx; i++) din[i] = ain[i] = x[i];
	for (int i = 0; i < nx; i++) din[i + nx] = ain[i + nx] = DhIndex[i];
	mlp3(g_learn.w1d, g_learn.b1d, g_learn.w2d, g_learn.b2d, g_learn.w3d, g_learn.b3d, d[0], d[2], d[4], d[6], din, dout);
	mlp3(g_learn.w1a, g_learn.b1a, g_learn.w2a, g_learn.b2a, g_learn.w3a, g_learn.b3a, d[1], d[3], d[5], d[7], ain, aout);
	Lfh[0] += dout[0];
	for (int i = 0; i < nu; i++) Lgh[i] += aout[i];
	free(din); free(ain); free(dout); free(aout);
}

static int32_t implicit_filter(const oracle_model *md, const imp_options *o, int N, const double *x, const double *uDes,
                               double *uAct, double *relax, double *diag)
{
	const int nx = md->nx, nu = md->nu, npSS = md->npSS, npBS = md->npBS, npBTSS = o->npBTSS;
	const int ns = nx + nx * nx, npTC = npBTSS * npSS + npBS, nv = nu + 2;
	double f[NX], g[NX * NU];
	md->dynamics(x, f, g); /* :414-416 */
	double *traj = (double *)malloc(sizeof(double) * (size_t)N * ns);
	double *t = (double *)malloc(sizeof(double) * N);
	double *hFull = (double *)malloc(sizeof(double) * (size_t)N * npSS);
	double *DhFull = (double *)malloc(sizeof(double) * (size_t)N * npSS * nx);
	double *hFullMin = (double *)malloc(sizeof(double) * N);
	zoh_state z;
	memset(&z, 0, sizeof(z));
	z.backContDt = o->backContDt;
	z.backTrajDt = o->backTrajDt;
	z.t_last = -1.;
	integrate_backup_trajectory_z(md, o->satSharpness, o->backTrajDt, N, x, traj, t, hFull, DhFull, hFullMin, o->rb ? &z : 0);
	int crit[64];
	k_smallest(hFullMin, N, npBTSS, crit); /* std::sort over all N points, :487 */

	double *h = (double *)calloc(npTC, sizeof(double));
	double *Dh = (double *)calloc((size_t)npTC * nx, sizeof(double));
	double DhSSDx[64], DhIndex[64] = {0.0};
	for (int idx = 0; idx < npBTSS; idx++) { /* :518-540 */
		int c = crit[idx];
		memcpy(h + idx * npSS, hFull + (size_t)c * npSS, sizeof(double) * npSS);
		/* RB: h <- lower bound of safetySet_int over x_c +- x_unc (src/asif_implicit_robust.cpp:636-647); Dh stays nominal */
		if (o->rb) md->safety_set_lower(traj + (size_t)c * ns, o->x_unc, h + idx * npSS);
		mat_mul(DhFull + (size_t)c * npSS * nx, npSS, nx, traj + (size_t)c * ns + nx, nx, DhSSDx);
		for (int i = 0; i < npSS; i++)
			for (int j = 0; j < nx; j++) Dh[(idx * npSS + i) + j * npTC] = DhSSDx[i + j * npSS];
		if (idx == 0) memcpy(DhIndex, DhSSDx, sizeof(double) * npSS * nx); /* Dh_index_, :533-537 (n_debug == -1) */
	}
	/* backup-set rows at the end of the trajectory, :542-554 */
	const double *btX = traj + (size_t)(N - 1) * ns;
	double DhBS[4 * NX], DhBSDx[4 * NX];
	md->backup_set(btX, h + npBTSS * npSS, DhBS, 0);
	double hBackupEnd = min_elem(h + npBTSS * npSS, npBS);
	mat_mul(DhBS, npBS, nx, btX + nx, nx, DhBSDx);
	for (int i = 0; i < npBS; i++)
		for (int j = 0; j < nx; j++) Dh[(npBTSS * npSS + i) + j * npTC] = DhBSDx[i + j * npBS];

	double *Lfh = (double *)malloc(sizeof(double) * npTC);
	double *Lgh = (double *)malloc(sizeof(double) * (size_t)npTC * nu);
	mat_vec(Dh, npTC, nx, f, Lfh);
	mat_mul(Dh, npTC, nx, g, nu, Lgh);
	if (g_learn.on) learned_residual(x, nx, DhIndex, Lfh, Lgh, nu); /* :585-588 */
	double *A = (double *)calloc((size_t)npTC * nv, sizeof(double));
	double *b = (double *)calloc(npTC, sizeof(double));
	for (int i = 0; i < npTC; i++) /* :590-611 */
		for (int j = 0; j < nu; j++) A[i + j * npTC] = Lgh[i + j * npTC];
	for (int i = 0; i < npBTSS * npSS; i++) A[i + nu * npTC] = h[i];
	for (int i = npBTSS * npSS; i < npTC; i++) A[i + (nu + 1) * npTC] = h[i];
	for (int i = 0; i < npTC; i++) b[i] = -Lfh[i];

	double H[ORACLE_QP_NVMAX * ORACLE_QP_NVMAX] = {0.0}, c[ORACLE_QP_NVMAX], lb[ORACLE_QP_NVMAX], ub[ORACLE_QP_NVMAX], sol[ORACLE_QP_NVMAX];
	for (int i = 0; i < nu; i++) {
		H[i + i * nv] = 1.0;
		c[i] = -2.0 * uDes[i];
		lb[i] = md->lb[i];
		ub[i] = md->ub[i];
	}
	H[(nv - 2) + (nv - 2) * nv] = o->relaxCost; /* :238-254 */
	H[(nv - 1) + (nv - 1) * nv] = o->relaxCost;
	c[nv - 2] = -2.0 * o->relaxCost * o->relaxSafeLb;
	c[nv - 1] = -2.0 * o->relaxCost * o->relaxReachLb;
	lb[nv - 2] = o->relaxSafeLb;
	lb[nv - 1] = o->relaxReachLb;
	ub[nv - 2] = o->inf;
	ub[nv - 1] = o->inf;
	apply_cost_override(nv, nu, H, c);
	int st = oracle_qp_solve(nv, npTC, 1, H, c, A, b, lb, ub, 0, sol);
	int32_t rc;
	if (st == 1) {
		memcpy(uAct, sol, sizeof(double) * nu);
		input_saturate(md, uAct);
		relax[0] = sol[nu];
		relax[1] = sol[nu + 1];
		rc = 1;
	} else {
		double Du[NU * NX];
		md->backup_controller(x, uAct, Du);
		input_saturate(md, uAct);
		rc = -1;
	}
	if (diag) {
		double hs[16], Dhs[16 * NX];
		md->safety_set(x, hs, Dhs);
		diag[0] = min_elem(hs, npSS);
		diag[1] = hBackupEnd;
		for (int i = 0; i < npBTSS; i++) diag[2 + i] = (double)crit[i];
		memcpy(diag + 2 + npBTSS, A, sizeof(double) * npTC * nv);
		memcpy(diag + 2 + npBTSS + npTC * nv, b, sizeof(double) * npTC);
	}
	free(traj); free(t); free(hFull); free(DhFull); free(hFullMin);
	free(h); free(Dh); free(Lfh); free(Lgh); free(A); free(b);
	return rc;
}

/* ------------------------------------------------------------------------------------------
 * ASIFrobust on the InvertedPendulum interval dynamics + half-plane table, in the REDUCED form.
 *
 * Reference: src/asif_robust.cpp:275-367 builds, per safety function k, the LP-dual rows
 *     h_k d + [Lg-_k, Lf-_k].lam+_k - [Lg+_k, Lf+_k].lam-_k >= 0,  lam+_k0 - lam-_k0 = u,  lam+_k1 - lam-_k1 = 1,
 * lam >= 0 with zero cost (:103-150,339-358).  For fixed (u, d) the best lam is lam+ = (u+, 1), lam- = (u-, 0),
 * so row k holds iff  h_k d + min(Lg-_k u, Lg+_k u) + Lf-_k >= 0, i.e. the two ordinary rows
 *     Lg-_k u + h_k d >= -Lf-_k   and   Lg+_k u + h_k d >= -Lf-_k          (SURVEY F8 / 8a-R),
 * and (u*, d*) are identical.  The interval values follow libaffa on a point state:
 * f = [x1, sin x0] has radius 0, g = [0, [pMin, pMax]] is centre (pMin+pMax)/2 with one noise term
 * (pMax-pMin)/2 (lib/libaffa/src/aa_aafcommon.cpp:81-101), Lg_k = Dh_k1 * g1 (aa_aafapprox.cpp:34-101).
 * Pinned against the reference build (402 variables, 300 rows, libaffa) by tests/test_oracle_vs_ref.py.
 * ---------------------------------------------------------------------------------------- */
#include "oracle_tables.h"

static int32_t robust_ip_filter(double relaxLb, double relaxCost, double pMin, double pMax, double inf, const double *x,
                                const double *uDes, double *uAct, double *relax, double *diag)
{
	const int npSS = ORACLE_N_HALFPLANES, nc = 2 * npSS, nv = 2;
	const double *tab = oracle_halfplanes_70_135;
	double *A = (double *)calloc((size_t)nc * nv, sizeof(double));
	double *b = (double *)calloc(nc, sizeof(double));
	const double f0 = x[1], f1 = sin(x[0]);
	const double gc = (pMin + pMax) / 2, gr = (pMax - pMin) / 2; /* AAF(interval): centre and radius */
	for (int k = 0; k < npSS; k++) {
		const double a0 = tab[2 * k], a1 = tab[2 * k + 1];
		const double h = 1. - a0 * x[0] - a1 * x[1]; /* examples/InvertedPendulum_Robust.cpp:53-60 */
		const double Dh0 = -a0, Dh1 = -a1;
		double lf = 0.0;
		lf = lf + Dh0 * f0;
		lf = lf + Dh1 * f1;
		const double lgc = Dh1 * gc, lgr = fabs(Dh1 * gr); /* centre and radius of Dh1 * g1 */
		const double lgLo = lgc - lgr, lgHi = lgc + lgr;
		A[2 * k] = lgLo;
		A[2 * k + 1] = lgHi;
		A[2 * k + nc] = h;
		A[2 * k + 1 + nc] = h;
		b[2 * k] = -lf;
		b[2 * k + 1] = -lf;
		if (diag) {
			diag[5 * k + 0] = h;
			diag[5 * k + 1] = lgLo;
			diag[5 * k + 2] = lgHi;
			diag[5 * k + 3] = lf;
			diag[5 * k + 4] = lf;
		}
	}
	double H[4] = {1.0, 0.0, 0.0, relaxCost}, c[2] = {-2.0 * uDes[0], -2.0 * relaxCost * relaxLb};
	double lb[2] = {-1.5, relaxLb}, ub[2] = {1.5, inf}, sol[2];
	int st = oracle_qp_solve(nv, nc, 1, H, c, A, b, lb, ub, 0, sol);
	free(A);
	free(b);
	if (st == 1) {
		uAct[0] = sol[0] > 1.5 ? 1.5 : (sol[0] < -1.5 ? -1.5 : sol[0]);
		*relax = sol[1];
		return 1;
	}
	return -1; /* uAct untouched, src/asif_robust.cpp:249-251 */
}

int32_t oracle_realizable_filter(const double *opts, const double *x, const double *uDes, double *uAct, double *relax,
                                 double *diag); /* realizable_oracle.c */

/* ------------------------------------------------------------------------------------------ */
typedef struct {
	int cfg;
	const oracle_model *md;
	oracle_model md_copy; /* model with overridden input bounds */
	tb_options tb;
	int N;
	double relaxLb, relaxCost, pMin, pMax;
	int npSSmax;
	imp_options imp;
	double rz[8];
	int nx, nu, n_relax, nc, nv, n_diag;
} ctx_t;

static int make_ctx(int cfg, const double *opts, int n_opts, ctx_t *c)
{
	memset(c, 0, sizeof(*c));
	c->cfg = cfg;
	switch (cfg) {
	case ORACLE_CFG_DI_EXPLICIT:
		c->md = oracle_get_model(cfg, 0);
		c->relaxLb = 5.0; /* include/asif.h:11-17 */
		c->relaxCost = 50.0;
		if (opts && n_opts >= 2) {
			c->relaxLb = opts[0];
			c->relaxCost = opts[1];
		}
		c->npSSmax = (opts && n_opts >= 3 && opts[2] > 0 && opts[2] < c->md->npSS) ? (int)opts[2] : c->md->npSS;
		c->n_relax = 1;
		c->nc = c->npSSmax;
		c->nv = c->md->nu + 1;
		c->n_diag = c->nc * c->nv + c->nc;
		break;
	case ORACLE_CFG_DI_IMPLICIT_TB:
	case ORACLE_CFG_SEGWAY_TB: {
		int variant = 0;
		if (cfg == ORACLE_CFG_SEGWAY_TB && opts && n_opts >= 10) variant = (opts[9] == 0.0);
		if (cfg == ORACLE_CFG_DI_IMPLICIT_TB && opts && n_opts >= 10) variant = (opts[9] != 0.0);
		c->md = oracle_get_model(cfg, variant);
		if (c->md && opts && n_opts >= 12) { /* initialize(lb, ub) with other bounds than the example's: opts[10], opts[11] */
			c->md_copy = *c->md;
			c->md_copy.lb[0] = opts[10];
			c->md_copy.ub[0] = opts[11];
			c->md = &c->md_copy;
		}
		tb_parse_options(cfg, opts, n_opts, &c->tb);
		c->N = tb_npbt(&c->tb);
		c->n_relax = 1;
		c->nc = c->tb.npBTSS * c->md->npSS + 2;
		c->nv = c->md->nu + 1;
		c->n_diag = 4 + c->tb.npBTSS + c->nc * c->nv + c->nc;
		break;
	}
	case ORACLE_CFG_IP_IMPLICIT_RB:
	case ORACLE_CFG_DI_IMPLICIT_RB:
	case ORACLE_CFG_IP_IMPLICIT: {
		c->md = oracle_get_model(cfg, 0);
		imp_options *o = &c->imp;
		/* include/asif_implicit.h:20-34 + examples/InvertedPendulum_Implicit.cpp:93-97 */
		o->relaxCost = 50.0; o->relaxReachLb = 5.0; o->relaxSafeLb = 10.0; o->backTrajHorizon = 5.0;
		o->backTrajDt = 0.001; o->satSharpness = 0.1; o->inf = 1e20; o->npBTSS = 10;
		if (opts && n_opts >= 6) {
			o->relaxCost = opts[0]; o->relaxReachLb = opts[1]; o->relaxSafeLb = opts[2];
			o->backTrajHorizon = opts[3]; o->backTrajDt = opts[4]; o->satSharpness = opts[5];
		}
		if (opts && n_opts >= 7 && opts[6] >= 1.0 && opts[6] <= 16.0) o->npBTSS = (int)opts[6]; /* constructor argument */
		if (cfg != ORACLE_CFG_IP_IMPLICIT) { /* include/asif_implicit_robust.h:22-38 */
			o->rb = 1;
			if (!(opts && n_opts >= 6)) { o->relaxSafeLb = 5.0; o->backTrajHorizon = 1.0; o->backTrajDt = 0.01; }
			o->backContDt = 0.01;
			if (opts && n_opts >= 8) o->backContDt = opts[7];
			for (int i = 0; i < c->md->nx; i++) o->x_unc[i] = (opts && n_opts >= 8 + c->md->nx) ? opts[8 + i] : 0.0;
		}
		/* src/asif_implicit.cpp:211-216 (no backTrajExtend in this class) */
		c->N = (int)round(o->backTrajHorizon / o->backTrajDt) + 1;
		if (c->N < o->npBTSS) {
			c->N = o->npBTSS;
			o->backTrajDt = o->backTrajHorizon / (double)(c->N - 1);
		}
		c->n_relax = 2;
		c->nc = o->npBTSS * c->md->npSS + c->md->npBS;
		c->nv = c->md->nu + 2;
		c->n_diag = 2 + o->npBTSS + c->nc * c->nv + c->nc;
		break;
	}
	case ORACLE_CFG_IP_ROBUST:
		c->md = oracle_get_model(ORACLE_CFG_IP_IMPLICIT, 0); /* dims + open-loop plant only */
		c->relaxLb = 5.0; c->relaxCost = 50.0; c->pMin = 0.8; c->pMax = 1.2;
		if (opts && n_opts >= 4) {
			c->relaxLb = opts[0]; c->relaxCost = opts[1]; c->pMin = opts[2]; c->pMax = opts[3];
		}
		c->n_relax = 1;
		c->nc = 2 * ORACLE_N_HALFPLANES;
		c->nv = 2;
		c->n_diag = 5 * ORACLE_N_HALFPLANES;
		break;
	case ORACLE_CFG_IP_REALIZABLE: {
		c->md = oracle_get_model(ORACLE_CFG_IP_IMPLICIT, 0); /* dims only */
		/* [relaxDes, relaxOffset, relaxCost, unc0, unc1, npSSmax, pMin, pMax]: Options of
		 * examples/InvertedPendulum_RealizableSampled.cpp:232-233 (relaxOffset keeps its default 5, include/asif_realizable.h:16-19) */
		const double d[8] = {1.0, 5.0, 50.0, 0.032, 0.027, 2.0, 0.9, 1.1};
		memcpy(c->rz, d, sizeof(d));
		if (opts && n_opts >= 6)
			for (int i = 0; i < 6; i++) c->rz[i] = opts[i];
		if (c->rz[5] < 0 || c->rz[5] > 8) return -1;
		const int npSS = ORACLE_RZ_MAXCRIT * ORACLE_RZ_MAXACT, npSSmax = (int)c->rz[5];
		c->n_relax = 2;
		c->nc = 2 * npSS + npSSmax;
		c->nv = 2;
		c->n_diag = 1 + ORACLE_RZ_MAXCRIT + npSSmax + 4 * npSS + 2 * npSSmax;
		break;
	}
	default:
		return -1;
	}
	if (!c->md) return -1;
	c->nx = c->md->nx;
	c->nu = c->md->nu;
	return 0;
}

static int32_t filter_one(const ctx_t *c, const double *x, const double *ud, double *ua, double *relax, double *diag)
{
	switch (c->cfg) {
	case ORACLE_CFG_DI_EXPLICIT:
		return explicit_filter(c->md, c->relaxLb, c->relaxCost, c->npSSmax, x, ud, ua, relax, diag);
	case ORACLE_CFG_DI_IMPLICIT_TB:
	case ORACLE_CFG_SEGWAY_TB:
		return tb_filter(c->md, &c->tb, c->N, x, ud, ua, relax, diag);
	case ORACLE_CFG_IP_IMPLICIT:
	case ORACLE_CFG_IP_IMPLICIT_RB:
	case ORACLE_CFG_DI_IMPLICIT_RB:
		return implicit_filter(c->md, &c->imp, c->N, x, ud, ua, relax, diag);
	case ORACLE_CFG_IP_REALIZABLE:
		return oracle_realizable_filter(c->rz, x, ud, ua, relax, diag);
	case ORACLE_CFG_IP_ROBUST:
		return robust_ip_filter(c->relaxLb, c->relaxCost, c->pMin, c->pMax, 1e20, x, ud, ua, relax, diag);
	}
	return -100;
}

int oracle_dims(int cfg, const double *opts, int n_opts, int32_t *dims)
{
	ctx_t c;
	if (make_ctx(cfg, opts, n_opts, &c)) return -1;
	dims[0] = c.nx; dims[1] = c.nu; dims[2] = c.n_relax; dims[3] = c.nc; dims[4] = c.nv; dims[5] = c.n_diag;
	return 0;
}

int oracle_filter_batch(int cfg, const double *opts, int n_opts, int64_t n, const double *x, const double *u_des,
                        double *u_act, double *relax, int32_t *rc, double *diag)
{
	ctx_t c;
	if (make_ctx(cfg, opts, n_opts, &c)) return -1;
	for (int64_t k = 0; k < n; k++) {
		double r[2] = {0.0, 0.0};
		for (int j = 0; j < c.nu; j++) u_act[k * c.nu + j] = 0.0;
		rc[k] = filter_one(&c, x + k * c.nx, u_des + k * c.nu, u_act + k * c.nu, r, diag ? diag + k * c.n_diag : 0);
		for (int j = 0; j < c.n_relax; j++) relax[k * c.n_relax + j] = r[j];
	}
	return 0;
}

int oracle_filter_batch_cost(int cfg, const double *opts, int n_opts, int64_t n, const double *x, const double *H,
                             const double *cvec, double *u_act, double *relax, int32_t *rc, double *diag)
{
	ctx_t c;
	if (make_ctx(cfg, opts, n_opts, &c)) return -1;
	if (cfg == ORACLE_CFG_IP_ROBUST || cfg == ORACLE_CFG_IP_REALIZABLE) return -2; /* reduced formulations: no LP-dual cost */
	double ud0[NU] = {0.0};
	for (int64_t k = 0; k < n; k++) {
		double r[2] = {0.0, 0.0};
		for (int j = 0; j < c.nu; j++) u_act[k * c.nu + j] = 0.0;
		g_cost.H = H;
		g_cost.c = cvec + k * c.nv;
		rc[k] = filter_one(&c, x + k * c.nx, ud0, u_act + k * c.nu, r, diag ? diag + k * c.n_diag : 0);
		g_cost.H = 0;
		g_cost.c = 0;
		for (int j = 0; j < c.n_relax; j++) relax[k * c.n_relax + j] = r[j];
	}
	return 0;
}

int oracle_filter_batch_lie(const double *opts, int n_opts, int64_t n, const double *x, const double *u_des, const double *Lfh,
                            const double *Lgh, double *u_act, double *relax, int32_t *rc, double *diag)
{
	ctx_t c;
	if (make_ctx(ORACLE_CFG_DI_EXPLICIT, opts, n_opts, &c)) return -1;
	for (int64_t k = 0; k < n; k++) {
		double r[2] = {0.0, 0.0};
		for (int j = 0; j < c.nu; j++) u_act[k * c.nu + j] = 0.0;
		g_cost.Lfh = Lfh + k * c.nc;
		g_cost.Lgh = Lgh + k * c.nc * c.nu;
		rc[k] = filter_one(&c, x + k * c.nx, u_des + k * c.nu, u_act + k * c.nu, r, diag ? diag + k * c.n_diag : 0);
		g_cost.Lfh = g_cost.Lgh = 0;
		relax[k] = r[0];
	}
	return 0;
}

int oracle_rollout(int cfg, const double *opts, int n_opts, int64_t n, int32_t steps, double dt, double *x,
                   const double *u_des, double *u_act_last, int32_t *rc_last, int64_t *rc_hist)
{
	ctx_t c;
	if (make_ctx(cfg, opts, n_opts, &c)) return -1;
	const int nx = c.nx, nu = c.nu;
	if (rc_hist)
		for (int i = 0; i < 8; i++) rc_hist[i] = 0;
	for (int64_t k = 0; k < n; k++) {
		double *xk = x + k * nx, ua[NU] = {0.0}, f[NX], g[NX * NU], fcl[NX];
		int32_t rc = 0;
		for (int32_t s = 0; s < steps; s++) {
			double r[2];
			rc = filter_one(&c, xk, u_des + k * nu, ua, r, 0);
			if (rc_hist) rc_hist[(rc >= -3 && rc <= 2) ? rc + 3 : 7]++;
			/* examples/segway_implicit_tb.cpp:265-283 */
			c.md->dynamics(xk, f, g);
			for (int i = 0; i < nx; i++) {
				fcl[i] = 0.0;
				fcl[i] += f[i];
				for (int j = 0; j < nu; j++) fcl[i] += g[i + j * nx] * ua[j];
			}
			for (int i = 0; i < nx; i++) xk[i] += dt * fcl[i];
		}
		for (int j = 0; j < nu; j++) u_act_last[k * nu + j] = ua[j];
		rc_last[k] = rc;
	}
	return 0;
}
