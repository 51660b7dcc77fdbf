/* A minimal types header for the run-time-script tests (OURS: the GPU box has no reference tree, and a user's
 * installation brings its own resources/Scripts).  The handful of names a user script of the reference's
 * dialect relies on: vec / vec_xyz / XYZ / VEC_ZERO / VEC_ONE per dimension, the neighbour-list parameters. */
#ifdef HAVE_3D
    #define vec float4
    #define vec_xyz float3
    #define XYZ xyz
    #define VEC_ZERO ((float4)(0.f, 0.f, 0.f, 0.f))
    #define VEC_ONE ((float4)(1.f, 1.f, 1.f, 0.f))
#else
    #define vec float2
    #define vec_xyz float2
    #define XYZ xy
    #define VEC_ZERO ((float2)(0.f, 0.f))
    #define VEC_ONE ((float2)(1.f, 1.f))
#endif
#define svec4 usize4
#define LINKLIST_LOCAL_PARAMS const __global usize* icell, const __global usize* ihoc, svec4 n_cells
