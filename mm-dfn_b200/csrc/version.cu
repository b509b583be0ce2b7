#include "common.cuh"
namespace mmdfn { unsigned long long g_launch_count = 0; }
extern "C" int mmdfn_abi_version(void) { return 2; }
extern "C" long long mmdfn_launch_count(void) { return (long long)mmdfn::g_launch_count; }
