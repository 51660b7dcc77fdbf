/* The rim of the circular blast (OURS: what the case-local bc.cl of examples/2D/shock_point does, written for
 * tests/test_shock_point_oracle.py and tests/test_gpu_presets.py): the particles within 1.5 kernel supports of the
 * rim r = R are frozen (imove = 0) while the time scheme moves the others, and released again afterwards. */
#include "resources/Scripts/types/types.h"

__kernel void set_fixed(__global int* imove,
                        const __global vec* r,
                        const usize N,
                        const float R)
{
    const usize i = get_global_id(0);
    if(i >= N)
        return;
    const float band = 1.5f * SUPPORT * H;
    if(length(r[i]) > R - band)
        imove[i] = 0;
}

__kernel void unset_fixed(__global int* imove,
                          const usize N)
{
    const usize i = get_global_id(0);
    if(i < N)
        imove[i] = 1;
}
