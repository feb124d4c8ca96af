// rv_launch.h -- host-callable launchers of the sm_100a kernels (internal to librvgpu.so).
#pragma once
#include <cuda_runtime.h>
namespace rv {
struct LoglikArgs;
cudaError_t launch_loglik(const LoglikArgs& a, int P, int D, int mapping, int num_sms, cudaStream_t stream);
cudaError_t launch_finalize(const double* part_chi2, const int* part_status, long long W, double npoints,
                            double* logp, int* status, unsigned long long* item_counter, cudaStream_t stream);
cudaError_t launch_curve_finalize(unsigned long long* item_counter, cudaStream_t stream);
cudaError_t launch_fp64_peak(double* d_out, int blocks, int iters, cudaStream_t stream);
}  // namespace rv
