/* TEST INFRASTRUCTURE ONLY -- see fk_oracle_impl.h.  Build: make -C oracle  (gcc, -ffp-contract=off). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL float
#define SUF f32
#define REAL_IS_FLOAT 1
#include "fk_oracle_impl.h"
#undef REAL
#undef SUF
#undef REAL_IS_FLOAT

#define REAL double
#define SUF f64
#include "fk_oracle_impl.h"
#undef REAL
#undef SUF

int fk_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
