#include "../../include/b200_turbomind.h"
#include "host_utils.h"

namespace b200 {
std::string& last_error_ref()
{
    static thread_local std::string s;
    return s;
}
}  // namespace b200

extern "C" const char* b200_last_error(void)
{
    return b200::last_error_ref().c_str();
}

extern "C" const char* b200_version(void)
{
    return "b200tm 0.1.0 sm_100a";
}
