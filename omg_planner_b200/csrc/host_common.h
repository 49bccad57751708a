// Host-side helpers shared by the translation units of libomgb200.so (defined in omgb200.cu).
#pragma once
#include <string>

namespace omgb {
int host_fail(int code, const std::string &msg);   // records the message omgb_last_error() returns; returns `code`
void host_count_launch();                          // omgb_launch_count() bookkeeping
}  // namespace omgb
