// k_collision_mesh.cu -- the collision kernel built with mesh geoms (CCD_MESH): hull-vertex support function with cached start
// vertex and hill climbing on the hull graph, mesh multi-contact, plane-mesh (reference collision_gjk.py:116, collision_convex.py:1190,
// collision_primitive.py:52).  Same source as k_collision.cu; launch_collision dispatches here when the model has mesh geoms, so
// mesh-free models keep the lean build (smaller per-lane stack: the mesh clip buffers alone are 1.5 KB).
#define CCD_MESH 1
#define MJB_COLLISION_MESH_TU
#include "k_collision.cu"
