/* customTimer.h -- stand-in for the header the reference examples include but do not ship
 * (examples/DoubleIntegrator_implicit_tb.cpp:7, examples/segway_implicit_tb.cpp:9); SURVEY F5 / deviation D2.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_CUSTOM_TIMER_SHIM_H
#define ORACLE_CUSTOM_TIMER_SHIM_H
#include <chrono>
struct CustomTimer {
	std::chrono::steady_clock::time_point t0;
	double dt;
};
static inline void tic(CustomTimer *t) { t->t0 = std::chrono::steady_clock::now(); }
static inline void toc(CustomTimer *t) { t->dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t->t0).count(); }
#endif
