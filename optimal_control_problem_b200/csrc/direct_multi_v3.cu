// Throughput-plan variant: 192 threads, three CTAs per SM (see direct_multi.cu).
#include "direct_multi_variant.h"

DIRECT_MULTI_DEFINE(multi_192x3, 192, 3)
