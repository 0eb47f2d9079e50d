// fk_resident.cuh -- entry points of the resident-kernel translation unit (fk_resident.cu)
#pragma once
#include <cuda_runtime.h>

#include "fk_resident.h"

namespace fk {

// returns 0 or a cudaError_t; zeroes the mailboxes, then one cooperative launch of P.G.nsteps Euler steps
int launch_resident(const ResPlan& P, const TileArgs& A, int exact, int batch, cudaStream_t st);
// CTAs of this shape the device holds at once (0: does not fit)
int resident_capacity(int exact, int nc, int mg, int threads, long long smem_bytes, int num_sms);

// cluster transport: clusters of `ntiles` CTAs of this shape the device holds at once (0: cannot be launched)
int cluster_capacity(int exact, int nc, int ntiles, int threads, long long smem_bytes);

// development (FK_RES_TIMING=1): cycle counters of CTA (0, 0) of the last launch; synchronises the device
int resident_timing(unsigned long long* out8);

}  // namespace fk
