// rg_scene.cuh -- device-side scene-graph walk: entities (local TRS + parent) -> TLAS instance records.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/rgb200.h"

namespace rg {

constexpr int kMaxEntityDepth = 64;   // nesting depth of the scene graph the device walk supports

// dTmp: n records, dEmit: n words (scratch); dOut receives the compacted instances in DFS order, *dCount their number.
void launchEntityInstances(const rg_entity* dEntities, uint32_t n, rg_instance* dTmp, uint32_t* dEmit, rg_instance* dOut, uint32_t* dCount, cudaStream_t st);

// One step of the rigid-sphere integrator, in place on device arrays (n entities, n bodies).
void launchStepSpheres(rg_entity* dEntities, rg_sphere_body* dBodies, uint32_t n, float dt, float floorY, cudaStream_t st);

}  // namespace rg
