// Shadows legkilo/src/common/glog_utils.hpp (log-directory set-up over gflags / boost::filesystem — nothing the
// measurement-update path depends on): KILO.cc only needs LOG().
#pragma once
#include <glog/logging.h>
