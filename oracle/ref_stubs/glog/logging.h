// Stand-in for <glog/logging.h> when the reference's CUDA kernel sources are compiled for oracle/_ref (oracle/build_ref.py): glog
// is not installed in this image, and c10 ships glog-compatible LOG / CHECK* macros (c10/util/logging_is_not_google_glog.h), which
// is all these translation units use.  TEST INFRASTRUCTURE - see oracle/__init__.py.
#pragma once
#include <c10/util/Logging.h>
