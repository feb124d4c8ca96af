// rv_smala.cu -- propose / accept kernels of the device SMALA sampler (Smala.step, mcmc.py:167-187); the value +
// gradient + Hessian evaluation in between is rv_var_kernels.cu's var_kernel.  One thread per chain: the per-chain
// linear algebra (rv_smala.cuh) is microseconds against the milliseconds of a variational evaluation.
#include <cuda_runtime.h>
#include "rv_launch.h"
#include "rv_smala.cuh"

namespace rv {

__global__ void smala_propose_kernel(const double* __restrict__ theta, const double* __restrict__ grad,
                                     const double* __restrict__ hess, const int* __restrict__ cur_status, int n,
                                     long long W, double eps, double alpha, unsigned long long seed,
                                     unsigned long long first_id, unsigned step, double* __restrict__ prop,
                                     double* __restrict__ q_fwd, int* __restrict__ geo_status, double* __restrict__ scratch) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double q = 0.0;
    geo_status[w] = smala_propose_one(n, theta + w * n, grad + w * n, hess + (size_t)w * n * n, cur_status[w], eps, alpha,
                                      seed, first_id + (unsigned long long)w, step, prop + w * n, q,
                                      scratch + (size_t)w * 5 * n * n);
    q_fwd[w] = q;
}

__global__ void smala_accept_kernel(double* __restrict__ theta, double* __restrict__ logp, double* __restrict__ grad,
                                    double* __restrict__ hess, const double* __restrict__ prop,
                                    const double* __restrict__ p_logp, const double* __restrict__ p_grad,
                                    const double* __restrict__ p_hess, const int* __restrict__ p_status,
                                    const int* __restrict__ geo_status, const double* __restrict__ q_fwd, int n,
                                    long long W, double eps, double alpha, unsigned long long seed,
                                    unsigned long long first_id, unsigned step, unsigned long long* __restrict__ n_accept,
                                    unsigned char* __restrict__ accepted, int* __restrict__ flag,
                                    double* __restrict__ chain_row, double* __restrict__ chain_logp_row, long long chain_w,
                                    double* __restrict__ scratch, int mala) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    // mala: the proposal carries the current state's (stale) derivatives (mcmc.py:205-206)
    const double* pg = mala ? grad + w * n : p_grad + w * n;
    const double* ph = mala ? hess + (size_t)w * n * n : p_hess + (size_t)w * n * n;
    const bool acc = smala_accept_one(n, theta + w * n, logp[w], prop + w * n, p_logp[w], pg, ph, p_status[w],
                                      geo_status[w], q_fwd[w], eps, alpha, seed, first_id + (unsigned long long)w, step,
                                      flag ? flag + w : nullptr, scratch + (size_t)w * 5 * n * n) != 0;
    if (acc) {
        for (int i = 0; i < n; i++) theta[w * n + i] = prop[w * n + i];
        if (!mala) {
            for (int i = 0; i < n; i++) grad[w * n + i] = p_grad[w * n + i];
            for (int i = 0; i < n * n; i++) hess[(size_t)w * n * n + i] = p_hess[(size_t)w * n * n + i];
        }
        logp[w] = p_logp[w];
        if (n_accept) n_accept[w] += 1ull;
    }
    if (accepted) accepted[w] = acc ? 1 : 0;
    if (chain_row && w < chain_w) {
        for (int i = 0; i < n; i++) chain_row[w * n + i] = theta[w * n + i];
        chain_logp_row[w] = logp[w];
    }
}
static inline unsigned nblk(long long n, int nt) { return (unsigned)((n + nt - 1) / nt > 0 ? (n + nt - 1) / nt : 1); }

cudaError_t launch_smala_propose(const double* theta, const double* grad, const double* hess, const int* cur_status, int n,
                                 long long W, double eps, double alpha, unsigned long long seed, unsigned long long first_id,
                                 unsigned step, double* prop, double* q_fwd, int* geo_status, double* scratch, cudaStream_t s) {
    smala_propose_kernel<<<nblk(W, 64), 64, 0, s>>>(theta, grad, hess, cur_status, n, W, eps, alpha, seed, first_id, step,
                                                    prop, q_fwd, geo_status, scratch);
    return cudaGetLastError();
}

cudaError_t launch_smala_accept(double* theta, double* logp, double* grad, double* hess, const double* prop,
                                const double* p_logp, const double* p_grad, const double* p_hess, const int* p_status,
                                const int* geo_status, const double* q_fwd, int n, long long W, double eps, double alpha,
                                unsigned long long seed, unsigned long long first_id, unsigned step,
                                unsigned long long* n_accept, unsigned char* accepted, int* flag, double* chain_row,
                                double* chain_logp_row, long long chain_w, double* scratch, int mala, cudaStream_t s) {
    smala_accept_kernel<<<nblk(W, 64), 64, 0, s>>>(theta, logp, grad, hess, prop, p_logp, p_grad, p_hess, p_status,
                                                   geo_status, q_fwd, n, W, eps, alpha, seed, first_id, step, n_accept,
                                                   accepted, flag, chain_row, chain_logp_row, chain_w, scratch, mala);
    return cudaGetLastError();
}

}  // namespace rv
