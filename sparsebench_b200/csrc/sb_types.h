// The two compile-time types of the reference (util.h:35-53): matrix values / vectors / scalars (CG_FLOAT) and matrix
// indices / sizes (CG_UINT). All sources are written in terms of them; libsparsebench_b200.so (double, unsigned int)
// and the variant libraries _f32 / _u64 / _f32u64 are builds of the same code with -DPRECISION / -DUINT_TYPE.
#pragma once
#include "sparsebench_b200.h"

typedef CG_FLOAT real_t;
typedef CG_UINT idx_t;
