/* temporary: configs not wired yet */
#include "ref_common.h"
RefFilter *make_ip_implicit(const double *, int) { return 0; }
RefFilter *make_ip_robust(const double *, int) { return 0; }
RefFilter *make_ip_realizable(const double *, int) { return 0; }
