// fk_stream_tu.cu -- the streaming kernels of ONE temporal-blocking depth and numerics mode
// (-DFK_TU_T=1..4 -DFK_TU_EXACT=0|1); see fk_stream.cuh.
#include "fk_stream.cuh"

#if !defined(FK_TU_T) || !defined(FK_TU_EXACT)
#error "compile with -DFK_TU_T=<steps per launch> -DFK_TU_EXACT=<0|1>"
#endif
#define FK_CAT4(a, b, c, d) a##b##c##d
#define FK_NAME(a, b, c, d) FK_CAT4(a, b, c, d)

namespace fk {

int FK_NAME(launch_stream_T, FK_TU_T, _E, FK_TU_EXACT)(const StreamPlan& P, const TileArgs& A, int batch, cudaStream_t st) {
    return P.G.uniformD ? launch_stream_t<FK_TU_EXACT != 0, FK_TU_T, true>(P, A, batch, st)
                        : launch_stream_t<FK_TU_EXACT != 0, FK_TU_T, false>(P, A, batch, st);
}

int FK_NAME(stream_occupancy_T, FK_TU_T, _E, FK_TU_EXACT)(int uni, int NT, long long smem) {
    return uni ? stream_occupancy_t<FK_TU_EXACT != 0, FK_TU_T, true>(NT, smem)
               : stream_occupancy_t<FK_TU_EXACT != 0, FK_TU_T, false>(NT, smem);
}

}  // namespace fk

#ifdef FK_STREAM_TIMING
// development: the per-CTA records of this translation unit's last launch (start ns, end ns, SM) -- 3 values per CTA
extern "C" int FK_NAME(fk_stream_timing_T, FK_TU_T, _E, FK_TU_EXACT)(unsigned long long* out, int nctas) {
    if (nctas > fk::FK_STREAM_TIMING_CTAS) nctas = fk::FK_STREAM_TIMING_CTAS;
    return (int)cudaMemcpyFromSymbol(out, fk::fk_stream_timing_buf, sizeof(unsigned long long) * 3 * (size_t)nctas);
}
#endif
