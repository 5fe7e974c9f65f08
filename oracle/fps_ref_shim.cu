// extern "C" entry around the reference's own host wrapper of its furthest-point-sampling kernel
// (/root/reference/third_party/pointnet2/_ext_src/src/sampling_gpu.cu:178-215, declared in
// _ext_src/include/sampling.h).  Compiled together with that file into oracle/_ref/libfps_ref.so by oracle/build_ref.py.
// TEST INFRASTRUCTURE: the checker for us3d_fps, never on the product path.
#include <cuda_runtime.h>

void furthest_point_sampling_kernel_wrapper(int b, int n, int m, const float *dataset, float *temp, int *idxs);

extern "C" int fps_ref(int b, int n, int m, const float *dataset, float *temp, int *idxs) {
    // the reference launches on torch's current stream; `temp` holds 1e10 per point (sampling.cpp:79-81)
    furthest_point_sampling_kernel_wrapper(b, n, m, dataset, temp, idxs);
    return (int)cudaGetLastError();
}
