// swo_internal.hpp — library-internal accessors shared between translation units.
#pragma once
#include <stdint.h>

#include "swiftover_cuda.h"

// device pointers (on devices[0]) of the source-contig name table: offsets [ntc+1] and chars
int swo_internal_tname_table(swo_ctx *ctx, const uint32_t **d_off, const char **d_chars);
