/* A user script in the reference's OpenCL dialect (OURS, for tests/test_gpu_scripts.py): what a case-local
 * script such as the examples' init.cl / Rescale.cl / spring.cl looks like -- vec arithmetic, swizzles, literals,
 * built-ins, a helper function, <Define>d constants, a local-memory array, and a walk over the cells of the
 * link-list. */
#include "resources/Scripts/types/types.h"

#ifndef DEMO_GAIN
    #define DEMO_GAIN 1.f
#endif

vec_xyz project(vec_xyz a, vec_xyz n)
{
    return dot(a, n) * n;
}

/* u <- u - 2 (u.n) n for the moving particles, r <- r + dt u, speed and a clamped density */
__kernel void reflect(const __global int* imove,
                      __global vec* r,
                      __global vec* u,
                      __global float* speed,
                      __global float* rho,
                      usize N,
                      float dt,
                      vec plane_n)
{
    const usize i = get_global_id(0);
    if(i >= N)
        return;
    if(imove[i] <= 0){
        speed[i] = 0.f;
        return;
    }
    const vec_xyz n = plane_n.XYZ;
    u[i].XYZ = u[i].XYZ - 2.f * project(u[i].XYZ, n);
    r[i] += dt * u[i] + DEMO_GAIN * H * VEC_ONE * 0.f;
    speed[i] = length(u[i].XYZ) + fabs(u[i].x) * 0.f;
    rho[i] = max(min(rho[i], 1010.f), 990.f);
    if(i == 0)
        r[i] = r[i] + (vec)(0.f);
}

/* particles in the cell of particle i and in the next one of its x row (ihoc = first sorted index of a cell,
 * N when empty; icell sorted), through a local-memory copy of the cell index */
__kernel void cell_count(__global unsigned int* count,
                         usize N,
                         LINKLIST_LOCAL_PARAMS)
{
    const usize i = get_global_id(0);
    const usize it = get_local_id(0);
    __local usize c_l[LOCAL_MEM_SIZE];
    if(i < N)
        c_l[it] = icell[i];
    barrier(CLK_LOCAL_MEM_FENCE);
    if(i >= N)
        return;
    unsigned int n = 0;
    for(usize c = c_l[it]; c < c_l[it] + 2 && c < n_cells.w; c++){
        usize j = ihoc[c];
        while((j < N) && (icell[j] == c)){
            n++;
            j++;
        }
    }
    count[i] = n;
}
