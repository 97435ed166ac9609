// Version / architecture probe of the C ABI.
#include "../../include/nunet_b200.h"
extern "C" int nunet_version(void) { return NUNET_ABI_VERSION; }
