// host_check.cpp -- exercises the C++ host layer the way the reference's example mains do
// (examples/DoubleIntegrator_implicit_tb.cpp:87-131): construct, initialize(lb, ub, opts),
// filter() in a closed loop, updateOptions() half-way; plus the QPWrapperB200 backend through the
// abstract interface, and a filterBatch call.  Prints "host_check ok" and returns 0 on success.
// Needs a GPU at run time; build() only checks that it compiles and links.
#include "asif_b200.hpp"
#include "loop_log.hpp"
#include <fstream>
#include <string>

#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

static int fail(const char *what)
{
	std::fprintf(stderr, "host_check FAILED: %s (%s)\n", what, ASIF::b200::lastError().c_str());
	return 1;
}

// host_check --update-options <in.bin> <out.bin>: the sequence initialize(opts1) -> filterBatch -> updateOptions(opts2) ->
// filterBatch of FilterBatchImplicitTB on caller-given states, written out for tests/test_gpu_dropin.py, which runs the same
// sequence through the reference build's ASIFimplicitTB (src/asif_implicit_tb.cpp:169-232, 365-405).
// in.bin:  int64 n, double opts1[9], double opts2[9], double x[n][2], double uDes[n]   (option order of the reference harness)
// out.bin: int32 code of updateOptions, then twice { double u[n], double relax[n], int32 rc[n] }
static int update_options_mode(const char *in_path, const char *out_path)
{
	using namespace ASIF;
	std::ifstream in(in_path, std::ios::binary);
	int64_t n = 0;
	double o1[9], o2[9];
	in.read((char *)&n, sizeof(n));
	in.read((char *)o1, sizeof(o1));
	in.read((char *)o2, sizeof(o2));
	if (!in || n <= 0) return fail("update-options: bad input file");
	std::vector<double> X(2 * n), U(n), UA(n), R(n);
	std::vector<int32_t> rc(n);
	in.read((char *)X.data(), sizeof(double) * 2 * n);
	in.read((char *)U.data(), sizeof(double) * n);
	if (!in) return fail("update-options: short input file");
	auto to_opts = [](const double *o) {
		b200::FilterBatchImplicitTB::Options t;
		t.relaxCost = o[0]; t.relaxSafeLb = o[1]; t.relaxTTS = o[2]; t.relaxMinOrtho = o[3]; t.backTrajHorizon = o[4];
		t.backTrajExtend = o[5]; t.backTrajDt = o[6]; t.backTrajMinOrtho = o[7]; t.satSharpness = o[8];
		return t;
	};
	b200::FilterBatchImplicitTB asif(b200::Model::DoubleIntegratorTB, 4);
	const double lb[1] = {-1.0}, ub[1] = {1.0};
	if (asif.initialize(lb, ub, to_opts(o1)) != 1) return fail("update-options: initialize");
	std::ofstream out(out_path, std::ios::binary);
	if (asif.filterBatch(n, X.data(), U.data(), UA.data(), R.data(), rc.data()) != 0) return fail("update-options: filterBatch 1");
	std::vector<double> UA1 = UA, R1 = R;
	std::vector<int32_t> rc1 = rc;
	const int32_t code = asif.updateOptions(to_opts(o2));
	if (asif.filterBatch(n, X.data(), U.data(), UA.data(), R.data(), rc.data()) != 0) return fail("update-options: filterBatch 2");
	out.write((const char *)&code, sizeof(code));
	out.write((const char *)UA1.data(), sizeof(double) * n);
	out.write((const char *)R1.data(), sizeof(double) * n);
	out.write((const char *)rc1.data(), sizeof(int32_t) * n);
	out.write((const char *)UA.data(), sizeof(double) * n);
	out.write((const char *)R.data(), sizeof(double) * n);
	out.write((const char *)rc.data(), sizeof(int32_t) * n);
	std::printf("update-options ok: code %d\n", code);
	return out ? 0 : fail("update-options: write");
}

// host_check --latency: microseconds per single-state filter() through the C++ host classes (no Python in the path),
// the way the reference's example mains call it (examples/DoubleIntegrator_implicit_tb.cpp:117-131), and per
// QPWrapperB200::solve(); median and 90th percentile of 2000 calls after 200 warm-up calls.
#include <algorithm>
#include <chrono>
static void report_latency(const char *what, std::vector<double> &us)
{
	std::sort(us.begin(), us.end());
	std::printf("latency %-46s median %6.2f us   p90 %6.2f us   min %6.2f us\n", what, us[us.size() / 2], us[us.size() * 9 / 10], us[0]);
}
static int latency_mode(void)
{
	using namespace ASIF;
	typedef std::chrono::steady_clock clk;
	const int warm = 200, reps = 2000;
	std::vector<double> us(reps);
	{
		b200::FilterBatchExplicit asif(b200::Model::DoubleIntegrator);
		const double lb[1] = {-1.0}, ub[1] = {1.0};
		if (asif.initialize(lb, ub) != 1) return fail("latency: FilterBatchExplicit::initialize");
		double x[2] = {0.3, 0.2}, uDes[1] = {0.9}, uAct[1] = {0.0}, relax = 0.0;
		for (int i = 0; i < warm + reps; i++) {
			x[0] = 0.3 + 1e-4 * (i % 100);
			const clk::time_point t0 = clk::now();
			asif.filter(x, uDes, uAct, relax);
			if (i >= warm) us[i - warm] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
		}
		report_latency("C1 FilterBatchExplicit::filter (one state)", us);
		double uLaunch = uAct[0];
		if (asif.setLowLatency(true) != 0) return fail("latency: setLowLatency (explicit)");
		for (int i = 0; i < warm + reps; i++) {
			x[0] = 0.3 + 1e-4 * (i % 100);
			const clk::time_point t0 = clk::now();
			if (asif.filter(x, uDes, uAct, relax) <= -100) return fail("latency: engine error through the server (explicit)");
			if (i >= warm) us[i - warm] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
		}
		report_latency("C1 ... through the latency server", us);
		if (uAct[0] != uLaunch) return fail("latency: server and launch path differ (explicit)");
	}
	{
		b200::FilterBatchImplicitTB asif(b200::Model::DoubleIntegratorTB, 4);
		b200::FilterBatchImplicitTB::Options o;
		o.relaxSafeLb = 10.0; o.relaxTTS = 5.0; o.relaxMinOrtho = 5.0; o.backTrajHorizon = 10.0; o.backTrajExtend = 0.0; o.backTrajDt = 0.1;
		const double lb[1] = {-1.0}, ub[1] = {1.0};
		if (asif.initialize(lb, ub, o) != 1) return fail("latency: FilterBatchImplicitTB::initialize");
		double x[2] = {0.3, 0.2}, uDes[1] = {0.9}, uAct[1] = {0.0}, relax = 0.0;
		for (int i = 0; i < warm + reps; i++) {
			x[0] = 0.3 + 1e-4 * (i % 100);
			const clk::time_point t0 = clk::now();
			asif.filter(x, uDes, uAct, relax);
			if (i >= warm) us[i - warm] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
		}
		report_latency("C2 FilterBatchImplicitTB::filter (npBT 101)", us);
		const double uLaunch = uAct[0], rLaunch = relax;
		if (asif.setLowLatency(true) != 0) return fail("latency: setLowLatency (TB)");
		for (int i = 0; i < warm + reps; i++) {
			x[0] = 0.3 + 1e-4 * (i % 100);
			const clk::time_point t0 = clk::now();
			if (asif.filter(x, uDes, uAct, relax) <= -100) return fail("latency: engine error through the server (TB)");
			if (i >= warm) us[i - warm] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
		}
		report_latency("C2 ... through the latency server", us);
		if (uAct[0] != uLaunch || relax != rLaunch) return fail("latency: server and launch path differ (TB)");
	}
	{
		QPWrapperB200 qp(2, 18, true);
		double H[4] = {1, 0, 0, 50}, c[2] = {-1.0, -1000}, A[36], b[18], lb[2] = {-1, 10}, ub[2] = {1, 1e20};
		for (int i = 0; i < 18; i++) {
			A[i] = (i % 3) - 1.0;
			A[18 + i] = 0.1 * i;
			b[i] = -1.0 - i;
		}
		if (qp.initialize(H, c, A, b, lb, ub) != 0) return fail("latency: QPWrapperB200::initialize");
		for (int i = 0; i < warm + reps; i++) {
			c[0] = -1.0 + 1e-3 * (i % 100);
			const clk::time_point t0 = clk::now();
			qp.updateCost(nullptr, c);
			qp.updateA(A);
			qp.updateb(b);
			qp.solve();
			if (i >= warm) us[i - warm] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
		}
		report_latency("QPWrapperB200 updateCost+updateA+updateb+solve", us);
	}
	return 0;
}

int main(int argc, char **argv)
{
	using namespace ASIF;
	if (argc == 4 && std::string(argv[1]) == "--update-options") return update_options_mode(argv[2], argv[3]);
	if (argc == 2 && std::string(argv[1]) == "--latency") return latency_mode();
	// --- 1. the QP backend behind the abstract interface: min (v0-2)^2 + 50 (v1-10)^2, v0 + v1 >= 13, 0<=v0<=1
	{
		std::unique_ptr<QPWrapperAbstract> qp(new QPWrapperB200(2, 1, true));
		const double H[4] = {1, 0, 0, 50}, c[2] = {-4, -1000}, A[2] = {1, 1}, b[1] = {13}, lb[2] = {0, 10}, ub[2] = {1, 1e20};
		if (qp->initialize(H, c, A, b, lb, ub) != 0) return fail("QPWrapperB200::initialize");
		if (qp->solve() != (int32_t)QPWrapperAbstract::SOLVER_STATUS::FEASIBLE) return fail("QPWrapperB200::solve");
		double v[2];
		qp->getSolution(v);
		if (std::fabs(v[0] - 1.0) > 1e-9 || std::fabs(v[1] - 12.0) > 1e-9) return fail("QPWrapperB200 solution");
		const double b2[1] = {1e9}; // infeasible with v0 <= 1 only if v1 cannot grow: it can, so still feasible
		qp->updateb(b2);
		if (qp->solve() != 1) return fail("QPWrapperB200 after updateb");
		const double ub2[2] = {1, 11};
		qp->updateBounds(nullptr, ub2);
		if (qp->solve() != -3) return fail("QPWrapperB200 infeasible status"); // OSQP_PRIMAL_INFEASIBLE passes through
	}
	// --- 2. the example main loop on the batched TB filter (one state)
	{
		b200::FilterBatchImplicitTB asif(b200::Model::DoubleIntegratorTB, 4);
		b200::FilterBatchImplicitTB::Options opts;
		opts.backTrajHorizon = 2.0;
		opts.backTrajDt = 0.001;
		opts.relaxSafeLb = 10.0;
		opts.relaxTTS = 5.0;
		opts.relaxMinOrtho = 5.0;
		const double lb[1] = {-1.0}, ub[1] = {1.0};
		if (asif.initialize(lb, ub, opts) != 1) return fail("FilterBatchImplicitTB::initialize");
		double x[2] = {0.1, 0.1}, uDes[1] = {0.9}, uAct[1] = {0.0}, relax = 0.0;
		const double dt = 0.001;
		int n_fail = 0;
		for (int s = 0; s < 400; s++) {
			if (s == 200) {
				opts.backTrajHorizon = 7.0;
				if (asif.updateOptions(opts) != 1) return fail("updateOptions");
			}
			const int32_t rc = asif.filter(x, uDes, uAct, relax);
			if (rc < 0) n_fail++;
			x[0] += dt * x[1]; // examples/DoubleIntegrator_implicit_tb.cpp:139-154
			x[1] += dt * uAct[0];
		}
		if (!(std::fabs(x[0]) <= 1.0 && std::fabs(x[1]) <= 1.0)) return fail("closed loop left the safe set");
		std::printf("closed loop: x = (%.6f, %.6f) uAct = %.6f relax = %.3f filter failures = %d\n", x[0], x[1], uAct[0], relax, n_fail);
		b200::FilterBatchImplicitTB::Options bad = opts;
		bad.satSharpness = 5.0;
		if (asif.updateOptions(bad) != 2) return fail("satSharpness clamp code 2");
		bad.satSharpness = 0.001;
		if (asif.updateOptions(bad) != 3) return fail("satSharpness clamp code 3");
	}
	// --- 3. a batch
	{
		b200::FilterBatchExplicit asif(b200::Model::DoubleIntegrator);
		const double lb[1] = {-1.0}, ub[1] = {1.0};
		if (asif.initialize(lb, ub) != 1) return fail("FilterBatchExplicit::initialize");
		const int n = 1000;
		std::vector<double> X(2 * n), U(n), UA(n), R(n);
		std::vector<int32_t> rc(n);
		for (int i = 0; i < n; i++) {
			X[2 * i] = -0.9 + 1.8 * i / n;
			X[2 * i + 1] = 0.3;
			U[i] = 1.0;
		}
		if (asif.filterBatch(n, X.data(), U.data(), UA.data(), R.data(), rc.data()) != 0) return fail("filterBatch");
		for (int i = 0; i < n; i++)
			if (rc[i] == 1 && (UA[i] > 1.0 || UA[i] < -1.0)) return fail("uAct outside bounds");
		// the same batch from pinned buffers (asif_host_alloc): the same bits
		b200::PinnedBuffer<double> Xp(2 * n), Up(n), UAp(n), Rp(n);
		b200::PinnedBuffer<int32_t> rcp(n);
		if (!Xp.data() || !rcp.data()) return fail("PinnedBuffer");
		std::memcpy(Xp.data(), X.data(), sizeof(double) * 2 * n);
		std::memcpy(Up.data(), U.data(), sizeof(double) * n);
		if (asif.filterBatch(n, Xp.data(), Up.data(), UAp.data(), Rp.data(), rcp.data()) != 0) return fail("filterBatch (pinned)");
		if (std::memcmp(UAp.data(), UA.data(), sizeof(double) * n) || std::memcmp(Rp.data(), R.data(), sizeof(double) * n) ||
		    std::memcmp(rcp.data(), rc.data(), sizeof(int32_t) * n))
			return fail("pinned and pageable batches differ");
	}
	// --- 4. implicit and robust classes compile against the same surface and run one batch each
	{
		b200::FilterBatchImplicit asif(b200::Model::InvertedPendulum, 10);
		b200::FilterBatchImplicit::Options o;
		o.backTrajHorizon = 5.0;
		o.backTrajDt = 0.05;
		o.relaxReachLb = 5.0;
		o.relaxSafeLb = 10.0;
		const double lb[1] = {-1.5}, ub[1] = {1.5};
		if (asif.initialize(lb, ub, o) != 1) return fail("FilterBatchImplicit::initialize");
		double x[2] = {0.2, -0.1}, uDes[1] = {0.4}, uAct[1] = {0.0}, relax[2] = {0.0, 0.0};
		const int32_t rc = asif.filter(x, uDes, uAct, relax);
		if (rc != 1 || relax[0] < 10.0 - 1e-9 || relax[1] < 5.0 - 1e-9) return fail("FilterBatchImplicit::filter");
		{ // ASIFimplicitRB: with x_unc = 0 and a hold of one Euler step it must return what ASIFimplicit returned
			b200::FilterBatchImplicitRB rb(b200::Model::InvertedPendulum, 10);
			b200::FilterBatchImplicitRB::Options ro;
			ro.backTrajHorizon = 5.0;
			ro.backTrajDt = 0.05;
			ro.backContDt = 0.05;
			ro.relaxReachLb = 5.0;
			ro.relaxSafeLb = 10.0;
			if (rb.initialize(lb, ub, ro) != 1) return fail("FilterBatchImplicitRB::initialize");
			double uRb[1] = {0.0}, relaxRb[2] = {0.0, 0.0};
			if (rb.filter(x, uDes, uRb, relaxRb) != rc || uRb[0] != uAct[0] || relaxRb[0] != relax[0]) return fail("FilterBatchImplicitRB::filter");
			// filter(x, H, c, ...) with H = 1 and c = [-2 uDes, -2 relaxCost relaxSafeLb, -2 relaxCost relaxReachLb] is filter(x, uDes, ...)
			const double Hone[1] = {1.0}, cvec[3] = {-2.0 * uDes[0], -2.0 * 50.0 * 10.0, -2.0 * 50.0 * 5.0};
			double uC[1] = {0.0}, relaxC[2] = {0.0, 0.0};
			if (rb.filter(x, Hone, cvec, uC, relaxC) != rc || uC[0] != uAct[0] || relaxC[1] != relax[1]) return fail("FilterBatchImplicitRB::filter(x, H, c)");
			double unc[2] = {0.05, 0.08};
			ro.x_unc = unc;
			ro.backContDt = 0.2;
			ro.satSharpness = 5.0;
			if (rb.updateOptions(ro) != 2) return fail("FilterBatchImplicitRB::updateOptions clamp code");
			if (rb.filter(x, uDes, uRb, relaxRb) != 1) return fail("FilterBatchImplicitRB::filter with x_unc");
			// learned residual: networks whose last layer is all zero except the bias: Lfh[0] += 0.25, Lgh[0] += 0 -> rows move, call succeeds
			const double w1[4 * 3] = {0.1, -0.2, 0.3, 0.0, 0.2, 0.1, -0.1, 0.3, 0.0, 0.1, 0.2, -0.3}, b1[3] = {0.1, 0.0, -0.1};
			const double w2[3 * 2] = {0.5, -0.5, 0.25, 0.1, -0.2, 0.3}, b2[2] = {0.0, 0.1};
			const double w3[2 * 1] = {0.0, 0.0}, b3d[1] = {0.25}, b3a[1] = {0.0};
			b200::LearningData &L = rb.learning_data_;
			L.d_drift_in = L.d_act_in = 4; L.d_drift_hidden = L.d_act_hidden = 3; L.d_drift_hidden_2 = L.d_act_hidden_2 = 2;
			L.d_drift_out = L.d_act_out = 1;
			L.w_1_drift = L.w_1_act = w1; L.b_1_drift = L.b_1_act = b1; L.w_2_drift = L.w_2_act = w2; L.b_2_drift = L.b_2_act = b2;
			L.w_3_drift = L.w_3_act = w3; L.b_3_drift = b3d; L.b_3_act = b3a;
			ro.use_learning = true;
			ro.satSharpness = 0.1;
			if (rb.updateOptions(ro) != 1) return fail("FilterBatchImplicitRB::updateOptions with use_learning");
			double uL[1] = {0.0}, relaxL[2] = {0.0, 0.0};
			const int32_t rcL = rb.filter(x, uDes, uL, relaxL);
			if (rcL != 1 && rcL != -1) return fail("FilterBatchImplicitRB::filter with use_learning");
		}
		const double planes[4] = {0.5, 0.0, -0.5, 0.0}; // |x0| <= 2
		b200::FilterBatchRobust rob(2, planes, 0.8, 1.2);
		if (rob.initialize(lb, ub) != 1) return fail("FilterBatchRobust::initialize");
		double r1 = 0.0;
		if (rob.filter(x, uDes, uAct, r1) != 1) return fail("FilterBatchRobust::filter");
		// --- 5. the example main loop for a small fleet, logged in the example's CSV layout
		// (examples/InvertedPendulum_Implicit.cpp:84-150: ten start points, dt 1e-3)
		const int n = 10;
		asif_loop_config loop = b200::FilterBatchBase::loopDefaults();
		loop.steps = 40;
		loop.dt = 0.001;
		loop.log_stride = 1;
		loop.log_agents = n;
		std::vector<double> X(2 * n), U(n, 0.0), UA(n), R(2 * n);
		std::vector<int32_t> rcs(n);
		for (int i = 0; i < n; i++) {
			X[2 * i] = 0.1 + double(i) * 0.29;
			X[2 * i + 1] = 0.0;
		}
		const int64_t W = asif.logRecordWidth(), NR = asif.logRecords(loop);
		std::vector<double> log((size_t)(n * NR * W));
		int64_t hist[8];
		if (asif.closedLoop(n, loop, X.data(), U.data(), UA.data(), R.data(), rcs.data(), hist, log.data()) != 0)
			return fail("closedLoop");
		int64_t calls = 0;
		for (int i = 0; i < 8; i++) calls += hist[i];
		if (calls != (int64_t)n * loop.steps || NR != loop.steps) return fail("closedLoop: return-code histogram / record count");
		const b200::LogLayout L(asif.nx(), asif.nu(), asif.nRelax());
		if ((int64_t)L.width() != W) return fail("LogLayout width differs from the library's");
		if (b200::writeLoopCsv("/tmp/asif_b200_host_check_0.csv", b200::LogSchema::ImplicitPendulum, L, log.data(), NR) != 0)
			return fail("writeLoopCsv");
		std::ifstream chk("/tmp/asif_b200_host_check_0.csv");
		std::string header, first;
		std::getline(chk, header);
		std::getline(chk, first);
		if (header != "tNow,x,v,uDes,uAct,gammaSafe,gammaReach,rc") return fail("CSV header");
		if (first.compare(0, 13, "0.0010000000,") != 0) return fail("CSV first record (tNow after the first step, 10 digits)");
		std::printf("closedLoop + CSV: %lld records per agent, first line %s\n", (long long)NR, first.c_str());
	}
	std::printf("host_check ok\n");
	return 0;
}
