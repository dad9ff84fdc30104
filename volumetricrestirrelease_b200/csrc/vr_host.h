// vr_host.h — internal host-side declarations shared by the C-ABI translation units (not installed).
#pragma once
#include <string>
#include "../../include/vrestir.h"

namespace vr {
int setError(int code, const std::string& msg);   // records vrestir_last_error() for this thread, returns code
}
