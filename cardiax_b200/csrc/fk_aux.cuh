// fk_aux.cuh -- launchers of the kernels in fk_aux.cu (snapshot resize, Dormand-Prince stage kernels, electrogram),
// called by the C ABI in fk_api.cu.  Every launcher enqueues on `st` and returns 0 or a cudaError_t.
#pragma once
#include <cuda_runtime.h>

#include "fk_ode.h"

namespace fk {

// defined in fk_api.cu: error text of the C ABI + the library's launch counter
int api_fail(int code, const char* msg);
int api_cuda_fail(int e, const char* where);
void api_count_launch(int n);

struct OdeScratch {
    double* partial;   // device: 2 * 3 * ODE_BLOCKS block sums + 2 results
};
enum { ODE_BLOCKS = 148 * 4, ODE_THREADS = 256 };
inline size_t ode_scratch_bytes() { return sizeof(double) * (2 * 3 * ODE_BLOCKS + 2); }

int launch_ode_copy(const P3& dst, long long off, const P3& src, long long n, cudaStream_t st);
int launch_ode_axpy(int exact, const P3& y, float h, const P3& f, const P3& out, long long n, cudaStream_t st);
int launch_ode_init_norms(int exact, const P3& y, const P3& f, float rtol, float atol, long long n, const OdeScratch& S,
                          double* sums2_host, cudaStream_t st);
int launch_ode_diff_norm(int exact, const P3& f1, const P3& f0, const P3& y, float rtol, float atol, long long n,
                         const OdeScratch& S, double* sum_host, cudaStream_t st);
int launch_ode_stage(int exact, const Dopri& T, int i, const P3& y, const P3* k, float dt, const P3& ys, long long n,
                     cudaStream_t st);
int launch_ode_finish(int exact, const Dopri& T, const P3& y, const P3* k, float dt, float rtol, float atol, const P3& yn,
                      const P3* c, long long n, const OdeScratch& S, double* sum_host, cudaStream_t st);
int launch_ode_interp(int exact, const P3* c, float r, const P3& out, long long off, long long n, cudaStream_t st);

}  // namespace fk
