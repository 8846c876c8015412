// Include-path shadow of the reference's include/tsdf_localization/util/constant.h (which hard-codes
// OMP_THREADS = 8, constant.h:4). Placed FIRST on the include path for the "all host cores" baseline
// variants only; the reference sources themselves are not edited. TEST/BENCH INFRASTRUCTURE ONLY.
#ifndef CONSTANT
#define CONSTANT
#ifndef TSDF_REF_OMP_THREADS
#error "build with -DTSDF_REF_OMP_THREADS=<n>"
#endif
constexpr unsigned int OMP_THREADS = TSDF_REF_OMP_THREADS;
#endif
