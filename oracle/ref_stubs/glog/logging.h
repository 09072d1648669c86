// Stand-in for <glog/logging.h> when the reference's CUDA kernel sources are compiled for oracle/_ref (oracle/build_ref.py): glog
// is not installed in this image, and c10 ships glog-compatible LOG / CHECK macros (c10/util/logging_is_not_google_glog.h), which
// is what these translation units use; the comparison forms c10 does not define are spelled out on top of its CHECK.
// TEST INFRASTRUCTURE - see oracle/__init__.py.
#pragma once
#include <c10/util/Logging.h>
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) CHECK((a) == (b))
#endif
#ifndef CHECK_NE
#define CHECK_NE(a, b) CHECK((a) != (b))
#endif
#ifndef CHECK_LT
#define CHECK_LT(a, b) CHECK((a) < (b))
#endif
#ifndef CHECK_LE
#define CHECK_LE(a, b) CHECK((a) <= (b))
#endif
#ifndef CHECK_GT
#define CHECK_GT(a, b) CHECK((a) > (b))
#endif
#ifndef CHECK_GE
#define CHECK_GE(a, b) CHECK((a) >= (b))
#endif
