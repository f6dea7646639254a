// single-precision instantiation of the C ABI (pnfftf_*)
#define PNX(name) pnfftf_##name
#define RT float
#include "api.cuh"
