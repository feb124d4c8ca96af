// rv_whfast_kernels.cu -- sm_100a kernel of the optional WHFast variant (rv_whfast.cuh): one thread per (walker, leg).
// Every walker takes the same number of fixed steps between two epochs, so there is no step-count divergence to manage.
#include <cuda_runtime.h>
#include "rv_launch.h"
#include "rv_whfast.cuh"

namespace rv {

template <int P, int D>
__global__ void __launch_bounds__(128) whfast_kernel(const WhArgs a) {
    const long long n_items = a.times ? a.W : 2 * a.W;
    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
         item += (long long)gridDim.x * blockDim.x)
        whfast_item<P, D>(a, item);
}

template <int P, int D>
static cudaError_t launch_wh_one(const WhArgs& a, int num_sms, cudaStream_t stream) {
    const long long n_items = a.times ? a.W : 2 * a.W;
    long long blocks = (n_items + 127) / 128;
    const long long cap = (long long)num_sms * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    whfast_kernel<P, D><<<(unsigned)blocks, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_whfast(const WhArgs& a, int P, int D, int num_sms, cudaStream_t stream) {
    switch (P * 10 + D) {
        case 12: return launch_wh_one<1, 2>(a, num_sms, stream);
        case 13: return launch_wh_one<1, 3>(a, num_sms, stream);
        case 22: return launch_wh_one<2, 2>(a, num_sms, stream);
        case 23: return launch_wh_one<2, 3>(a, num_sms, stream);
        case 32: return launch_wh_one<3, 2>(a, num_sms, stream);
        case 33: return launch_wh_one<3, 3>(a, num_sms, stream);
        case 42: return launch_wh_one<4, 2>(a, num_sms, stream);
        case 43: return launch_wh_one<4, 3>(a, num_sms, stream);
        case 52: return launch_wh_one<5, 2>(a, num_sms, stream);
        case 53: return launch_wh_one<5, 3>(a, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace rv
