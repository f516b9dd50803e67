// facet_table_check.cpp -- CPU-only: loads a polytope kernel CSV, evaluates the InvertedPendulum interval dynamics of
// examples/InvertedPendulum_RealizableSampled.cpp:46-53 over every facet with the host-side affine evaluator and prints
// the facet table, one line per (facet, active constraint).  tests/test_host_affine.py compares it with the table the
// reference build (libaffa) exported.
#include "realizable_host.hpp"

#include <cstdio>

using namespace ASIF::b200;

static const double pMin = 0.9, pMax = 1.1;

static void dynamics(const interval_t *x, interval_t *f, interval_t *g)
{
	f[0] = x[1];
	f[1] = sin(x[0]);
	g[0] = 0.;
	g[1] = interval_t(pMin, pMax);
}

int main(int argc, char **argv)
{
	if (argc < 2) {
		std::fprintf(stderr, "usage: %s kernel.csv\n", argv[0]);
		return 2;
	}
	try {
		const RealizableKernel k = loadRealizableKernel(argv[1]);
		const std::vector<double> t = computeFacetTable(k, dynamics);
		std::printf("%zu %zu %u %u\n", k.vertices.size(), k.facets.size(), k.maxCriticalFacets, k.maxActiveConstraints);
		for (size_t i = 0; i < t.size(); i += 4) std::printf("%.17g %.17g %.17g %.17g\n", t[i], t[i + 1], t[i + 2], t[i + 3]);
	} catch (const std::exception &e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
