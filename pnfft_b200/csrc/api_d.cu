// double-precision instantiation of the C ABI (pnfft_*)
#define PNX(name) pnfft_##name
#define RT double
#include "api.cuh"
