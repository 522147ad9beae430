// Throughput-plan variant: 128 threads, four CTAs per SM (see direct_multi.cu).
#include "direct_multi_variant.h"

DIRECT_MULTI_DEFINE(multi_128x4, 128, 4)
