// Per-thread error message behind stcsp_last_error() (include/stcsp_b200.h).
#pragma once
#include <string>
namespace stcsp {
void set_error(const std::string &msg);
}
