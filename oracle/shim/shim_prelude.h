// oracle/shim/shim_prelude.h — TEST INFRASTRUCTURE ONLY.
// Pulled in first by every shim header.  Includes every system header the reference sources ask
// for BEFORE the optional `#define float double`, so that the standard library is never re-typed;
// only the reference's own translation units are (EKF_SHIM_DOUBLE = the reference in fp64, which is
// the parity target BASELINE.json names).  Also routes rand()/srand() of the 1-point RANSAC loop
// (vslamRansac.cpp:970,989) to an injectable sequence so that runs are reproducible.
#ifndef EKF_SHIM_PRELUDE_H_
#define EKF_SHIM_PRELUDE_H_
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

extern "C" int ekf_shim_rand(void);
extern "C" void ekf_shim_srand(unsigned int);
#define rand ekf_shim_rand
#define srand ekf_shim_srand

// the reference prints dT on every frame to stdout (vslamRansac.cpp:229); keep stdout clean for
// the harness' JSON lines
#define cout cerr
#endif
