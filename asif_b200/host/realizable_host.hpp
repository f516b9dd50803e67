// realizable_host.hpp -- host-side setup of the batched ASIFrealizable (header only): the polytope kernel type and
// its CSV loader as in the reference's example (examples/InvertedPendulum_RealizableSampled.cpp:64-166), and the
// x-independent facet table that src/asif_realizable.cpp:137-157,470-500 recomputes on every filter() call.
#ifndef ASIF_B200_REALIZABLE_HOST_HPP
#define ASIF_B200_REALIZABLE_HOST_HPP

#include <fstream>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "affine.hpp"
#include "asif_b200.hpp"

namespace ASIF
{
namespace b200
{
	typedef Affine interval_t;

	// same fields as ASIFrealizable::kernel_t / facet_t (include/asif_realizable.h:22-37) minus the derived ones
	struct RealizableKernel {
		struct Facet {
			std::vector<uint32_t> verticesIdx;
			std::vector<double> normal;
			std::vector<uint32_t> activeConstraintsSet;
		};
		std::vector<std::vector<double> > vertices;
		std::vector<Facet> facets;
		uint32_t maxCriticalFacets = 0;
		uint32_t maxActiveConstraints = 0;
	};

	// File format of the example's loadRealizableKernel(): four header lines (nVertices, nFacets, maxCriticalFacets,
	// maxActiveConstraints, first comma-separated field of each), nVertices lines of coordinates, then per facet three
	// lines: vertex ids, normal, active-constraint facet ids.  Throws std::runtime_error like the example.
	inline RealizableKernel loadRealizableKernel(const std::string &filename, const uint32_t nx = 2)
	{
		std::ifstream fin(filename.c_str());
		if (!fin.is_open()) throw std::runtime_error("Could not open file: " + filename);
		std::string line, word;
		auto firstField = [&](const char *what) -> int {
			if (!std::getline(fin, line)) throw std::runtime_error(std::string("EOF reached reading ") + what);
			std::stringstream s(line);
			if (!std::getline(s, word, ',')) throw std::runtime_error(std::string("Could not parse ") + what);
			return std::stoi(word);
		};
		RealizableKernel k;
		const int nV = firstField("number of vertices");
		const int nF = firstField("number of facets");
		k.maxCriticalFacets = (uint32_t)firstField("max number of critical facets");
		k.maxActiveConstraints = (uint32_t)firstField("max number of active constraints");
		if (nV < 0 || nF < 0) throw std::runtime_error("negative sizes");
		k.vertices.assign(nV, std::vector<double>(nx, 0.0));
		k.facets.resize(nF);
		auto fields = [&](void) -> std::vector<std::string> {
			if (!std::getline(fin, line)) throw std::runtime_error("EOF reached");
			std::vector<std::string> out;
			std::stringstream ss(line);
			while (std::getline(ss, word, ','))
				if (!word.empty() && word != "\r") out.push_back(word);
			return out;
		};
		for (int i = 0; i < nV; i++) {
			const std::vector<std::string> f = fields();
			for (size_t j = 0; j < f.size() && j < nx; j++) k.vertices[i][j] = std::stod(f[j]);
		}
		for (int i = 0; i < nF; i++) {
			std::vector<std::string> f = fields();
			for (size_t j = 0; j < f.size() && j < nx; j++) k.facets[i].verticesIdx.push_back((uint32_t)std::stoi(f[j]));
			f = fields();
			for (size_t j = 0; j < f.size() && j < nx; j++) k.facets[i].normal.push_back(std::stod(f[j]));
			f = fields();
			for (size_t j = 0; j < f.size(); j++) k.facets[i].activeConstraintsSet.push_back((uint32_t)std::stoi(f[j]));
			if (k.facets[i].verticesIdx.size() != nx || k.facets[i].normal.size() != nx) throw std::runtime_error("bad facet record");
		}
		return k;
	}

	// interval dynamics callback, the signature ASIFrealizable / ASIFrobust take (include/asif_realizable.h:43-47)
	typedef std::function<void(const interval_t * /*x*/, interval_t * /*f*/, interval_t * /*g*/)> IntervalDynamics;

	// [facet][active][LfLo, LfHi, LgLo, LgHi] for nu = 1:
	//   xFaceInt = lambda v0 + (1 - lambda) v1 with lambda = [0, 1]                       (src/asif_realizable.cpp:137-157)
	//   Lfh = sum_k f_k(xFaceInt) * (-normal_k(active)),  Lgh likewise with g             (:470-500)
	inline std::vector<double> computeFacetTable(const RealizableKernel &k, const IntervalDynamics &dynamics, const uint32_t nx = 2)
	{
		const size_t nF = k.facets.size(), mA = k.maxActiveConstraints;
		std::vector<double> table(4 * mA * nF, 0.0);
		for (size_t i = 0; i < nF; i++) {
			std::vector<interval_t> xFace(nx);
			for (uint32_t j = 0; j < nx; j++) xFace[j] = interval_t(k.vertices[k.facets[i].verticesIdx[0]][j]);
			for (uint32_t j = 1; j <= nx - 1; j++) {
				const interval_t lambda(0., 1.);
				const std::vector<double> &vertex = k.vertices[k.facets[i].verticesIdx[j]];
				for (uint32_t q = 0; q < nx; q++) xFace[q] = lambda * xFace[q] + (1. - lambda) * vertex[q];
			}
			for (size_t j = 0; j < k.facets[i].activeConstraintsSet.size() && j < mA; j++) {
				const std::vector<double> &normal = k.facets[k.facets[i].activeConstraintsSet[j]].normal;
				std::vector<interval_t> f(nx), g(nx);
				dynamics(xFace.data(), f.data(), g.data());
				interval_t Lfh = 0., Lgh = 0.;
				for (uint32_t q = 0; q < nx; q++) Lfh = Lfh + f[q] * interval_t::point(-normal[q]);
				for (uint32_t q = 0; q < nx; q++) Lgh = Lgh + g[q] * interval_t::point(-normal[q]);
				double *out = &table[4 * (mA * i + j)];
				out[0] = Lfh.lo();
				out[1] = Lfh.hi();
				out[2] = Lgh.lo();
				out[3] = Lgh.hi();
			}
		}
		return table;
	}

	// Constructor arguments of ASIFrealizable (nx, nu, uncertaintyBounds, kernel, dynamics, npSSmax;
	// include/asif_realizable.h:39-52) -> the batched filter.  [pMin, pMax] is the input-gain interval of the
	// compiled-in InvertedPendulum device model used for the mid-point barrier rows.
	inline FilterBatchRealizable *makeFilterBatchRealizable(const double uncertaintyBounds[2], const RealizableKernel &k,
	                                                        const IntervalDynamics &dynamics, const double pMin,
	                                                        const double pMax, const uint32_t npSSmax = 0, const int32_t device = 0)
	{
		const size_t nV = k.vertices.size(), nF = k.facets.size(), mA = k.maxActiveConstraints;
		std::vector<double> v(2 * nV), n(2 * nF);
		std::vector<int32_t> fv(2 * nF), fa(mA * nF, -1);
		for (size_t i = 0; i < nV; i++)
			for (int j = 0; j < 2; j++) v[2 * i + j] = k.vertices[i][j];
		for (size_t i = 0; i < nF; i++) {
			for (int j = 0; j < 2; j++) {
				n[2 * i + j] = k.facets[i].normal[j];
				fv[2 * i + j] = (int32_t)k.facets[i].verticesIdx[j];
			}
			for (size_t j = 0; j < k.facets[i].activeConstraintsSet.size() && j < mA; j++)
				fa[mA * i + j] = (int32_t)k.facets[i].activeConstraintsSet[j];
		}
		const std::vector<double> lie = computeFacetTable(k, dynamics);
		return new FilterBatchRealizable(uncertaintyBounds, (uint32_t)nV, v.data(), (uint32_t)nF, n.data(), fv.data(), fa.data(),
		                                 lie.data(), k.maxCriticalFacets, (uint32_t)mA, pMin, pMax, npSSmax, device);
	}
} // namespace b200
} // namespace ASIF
#endif
