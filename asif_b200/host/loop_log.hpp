// loop_log.hpp -- CSV writers for the closed-loop log of asif_engine_closed_loop, one file per agent, in the column
// layouts the reference's example programs write (so that the plotting scripts made for them keep working):
//   ImplicitPendulum   examples/InvertedPendulum_Implicit.cpp:89-91,139-147
//                      tNow,x,v,uDes,uAct,gammaSafe,gammaReach,rc
//   RealizableSampled  examples/InvertedPendulum_RealizableSampled.cpp:230,275-284
//                      tNow,x,v,xEstim,vEstim,uDes,uAct,relax,relaxHard,rc
//   RealizableSmoothed examples/DoubleIntegrator_RealizableSampled.cpp:94,165-177
//                      tNow,x,v,xEstim,vEstim,uDes,uFilter,uAct,relax,relaxHard,rc,smoothLb,smoothUb
//   SegwayTB           examples/segway_implicit_tb.cpp:219 (commented out there) and :318-327
//                      tNow,x[0],x[1],x[2],x[3],uDes,uAct,TTSmax,TTS,orhto,hMinSS,rc   (header spelling as shipped)
// Numbers are written std::fixed with 10 digits as the examples do.
#ifndef ASIF_B200_LOOP_LOG_HPP
#define ASIF_B200_LOOP_LOG_HPP

#include <cstdint>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

namespace ASIF
{
	namespace b200
	{
		enum class LogSchema { ImplicitPendulum, RealizableSampled, RealizableSmoothed, SegwayTB };

		// offsets of the fields inside one record (include/asif_b200.h: asif_engine_loop_log_dims)
		struct LogLayout {
			uint32_t nx, nu, nRelax;
			LogLayout(const uint32_t nx_, const uint32_t nu_, const uint32_t nRelax_) : nx(nx_), nu(nu_), nRelax(nRelax_) {}
			uint32_t t(void) const { return 0; }
			uint32_t x(const uint32_t i) const { return 1 + i; }
			uint32_t xEstim(const uint32_t i) const { return 1 + nx + i; }
			uint32_t uDes(const uint32_t i) const { return 1 + 2 * nx + i; }
			uint32_t uFilter(const uint32_t i) const { return 1 + 2 * nx + nu + i; }
			uint32_t uAct(const uint32_t i) const { return 1 + 2 * nx + 2 * nu + i; }
			uint32_t relax(const uint32_t i) const { return 1 + 2 * nx + 3 * nu + i; }
			uint32_t rc(void) const { return 1 + 2 * nx + 3 * nu + nRelax; }
			uint32_t smoothLb(void) const { return rc() + 1; }
			uint32_t smoothUb(void) const { return rc() + 2; }
			uint32_t TTS(void) const { return rc() + 3; }
			uint32_t ortho(void) const { return rc() + 4; }
			uint32_t critIdx0(void) const { return rc() + 5; }
			uint32_t width(void) const { return rc() + 6; }
		};

		// Writes the records of one agent.  log points at this agent's first record; backTrajHorizon fills the
		// TTSmax column of the segway layout (the example logs opts.backTrajHorizon there).  Returns 0 or -1.
		inline int32_t writeLoopCsv(const std::string &path, const LogSchema schema, const LogLayout &L, const double *log,
		                            const int64_t nRecords, const double backTrajHorizon = 0.0)
		{
			std::ofstream f(path.c_str(), std::ofstream::out | std::ofstream::trunc);
			if (!f.is_open()) return -1;
			switch (schema) {
			case LogSchema::ImplicitPendulum: f << "tNow,x,v,uDes,uAct,gammaSafe,gammaReach,rc" << std::endl; break;
			case LogSchema::RealizableSampled: f << "tNow,x,v,xEstim,vEstim,uDes,uAct,relax,relaxHard,rc" << std::endl; break;
			case LogSchema::RealizableSmoothed:
				f << "tNow,x,v,xEstim,vEstim,uDes,uFilter,uAct,relax,relaxHard,rc,smoothLb,smoothUb" << std::endl;
				break;
			case LogSchema::SegwayTB: f << "tNow,x[0],x[1],x[2],x[3],uDes,uAct,TTSmax,TTS,orhto,hMinSS,rc" << std::endl; break;
			}
			f << std::fixed << std::setprecision(10);
			const uint32_t W = L.width();
			for (int64_t r = 0; r < nRecords; r++) {
				const double *R = log + r * W;
				const int32_t rc = (int32_t)R[L.rc()];
				f << R[L.t()] << ',';
				switch (schema) {
				case LogSchema::ImplicitPendulum:
					f << R[L.x(0)] << ',' << R[L.x(1)] << ',' << R[L.uDes(0)] << ',' << R[L.uAct(0)] << ',' << R[L.relax(0)] << ','
					  << R[L.relax(1)] << ',' << rc;
					break;
				case LogSchema::RealizableSampled:
					f << R[L.x(0)] << ',' << R[L.x(1)] << ',' << R[L.xEstim(0)] << ',' << R[L.xEstim(1)] << ',' << R[L.uDes(0)] << ','
					  << R[L.uAct(0)] << ',' << R[L.relax(0)] << ',' << R[L.relax(1)] << ',' << rc;
					break;
				case LogSchema::RealizableSmoothed:
					f << R[L.x(0)] << ',' << R[L.x(1)] << ',' << R[L.xEstim(0)] << ',' << R[L.xEstim(1)] << ',' << R[L.uDes(0)] << ','
					  << R[L.uFilter(0)] << ',' << R[L.uAct(0)] << ',' << R[L.relax(0)] << ',' << R[L.relax(1)] << ',' << rc << ','
					  << R[L.smoothLb()] << ',' << R[L.smoothUb()];
					break;
				case LogSchema::SegwayTB:
					for (uint32_t i = 0; i < L.nx; i++) f << R[L.x(i)] << ',';
					f << R[L.uDes(0)] << ',' << R[L.uAct(0)] << ',' << backTrajHorizon << ',' << R[L.TTS()] << ',' << R[L.ortho()] << ',';
					if (rc > 0) f << (int64_t)R[L.critIdx0()]; // the example logs backTrajCritIdx_[0] under the hMinSS heading
					else f << 0;
					f << ',' << rc;
					break;
				}
				f << std::endl;
			}
			return f.good() ? 0 : -1;
		}
	} // namespace b200
} // namespace ASIF
#endif
