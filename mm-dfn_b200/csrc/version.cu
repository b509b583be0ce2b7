#include "../../include/mmdfn_b200.h"
extern "C" int mmdfn_abi_version(void) { return 1; }
