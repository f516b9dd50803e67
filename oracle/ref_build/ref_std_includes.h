/* ref_std_includes.h -- everything the reference examples include, pulled in at global scope
 * so that the example file itself can afterwards be included inside a namespace
 * (its own #includes are then no-ops thanks to the include guards).  TEST INFRASTRUCTURE ONLY. */
#ifndef ASIF_REF_STD_INCLUDES_H
#define ASIF_REF_STD_INCLUDES_H
#include <iostream>
#include <fstream>
#include <sstream>
#include <iomanip>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include <random>
#include <functional>
#include <utility>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <asif++.h>
#include "CyberTimer.hpp"
#include "customTimer.h"
#define ASIF_REF_WITH_LEARNING_HELPER 1
#include "ref_common.h"
#endif
