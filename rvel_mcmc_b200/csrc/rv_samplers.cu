// rv_samplers.cu -- propose / accept kernels of the device samplers (the likelihood in between is
// rv_kernels.cu's loglik_kernel).  One thread per walker; these kernels are bandwidth-trivial.
//
//   MH       (mcmc.py:89-121)  theta' = theta + step_size*scales*z ; accept iff exp(logp'-logp) > u
//   stretch  (emcee 2.2.1 EnsembleSampler._propose_stretch, driven from mcmc.py:57-65)
//            zz = ((a-1)u+1)^2/a ; q = c_j - zz (c_j - s) ; accept iff (dim-1) ln zz + lnp(q) - lnp(s) > ln u'
//            (proposal arithmetic: fixed fma sequence shared bit-for-bit with the oracle)
#include <cuda_runtime.h>
#include "rv_launch.h"
#include "rv_rng.cuh"

namespace rv {

__global__ void mh_propose_kernel(const double* __restrict__ theta, const double* __restrict__ scales,
                                  double step_size, int nvars, long long W, unsigned long long seed,
                                  unsigned long long first_id, unsigned step, double* __restrict__ prop) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const unsigned long long id = first_id + (unsigned long long)w;
    for (int j = 0; 2 * j < nvars; j++) {
        double z0, z1;
        normal_pair(seed, id, step, (uint32_t)j, z0, z1);
        const int v0 = 2 * j, v1 = 2 * j + 1;
        prop[w * nvars + v0] = theta[w * nvars + v0] + step_size * scales[v0] * z0;
        if (v1 < nvars) prop[w * nvars + v1] = theta[w * nvars + v1] + step_size * scales[v1] * z1;
    }
}

// accept rule of Mh.step (mcmc.py:112-121): prior / Encounter (status != 0) -> reject; else exp(dlogp) > u
__global__ void mh_accept_kernel(double* __restrict__ theta, double* __restrict__ logp,
                                 const double* __restrict__ prop, const double* __restrict__ prop_logp,
                                 const int* __restrict__ prop_status, int nvars, long long W,
                                 unsigned long long seed, unsigned long long first_id, unsigned step,
                                 unsigned long long* __restrict__ n_accept, unsigned char* __restrict__ accepted,
                                 double* __restrict__ chain_row, double* __restrict__ chain_logp_row, long long chain_w) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const unsigned long long id = first_id + (unsigned long long)w;
    bool acc = false;
    if (prop_status[w] == ST_OK) {
        const U4 r = philox4x32_10(seed, id, step, RNG_ACCEPT);
        const double u = u53(r.x, r.y);
        acc = exp(prop_logp[w] - logp[w]) > u;
    }
    if (acc) {
        for (int v = 0; v < nvars; v++) theta[w * nvars + v] = prop[w * nvars + v];
        logp[w] = prop_logp[w];
        if (n_accept) n_accept[w] += 1ull;
    }
    if (accepted) accepted[w] = acc ? 1 : 0;
    if (chain_row && w < chain_w) {       // rows hold the first chain_w chains (all of them unless the context says otherwise)
        for (int v = 0; v < nvars; v++) chain_row[w * nvars + v] = theta[w * nvars + v];
        chain_logp_row[w] = logp[w];
    }
}

// S: the half being updated (nS walkers, global ids id0_S + i); C: the complementary half (nC walkers).
__global__ void stretch_propose_kernel(const double* __restrict__ S, const double* __restrict__ C, int nvars,
                                       long long nS, long long nC, double a, unsigned long long seed,
                                       unsigned long long id0_S, unsigned step, unsigned half,
                                       double* __restrict__ q, double* __restrict__ zz_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nS) return;
    const unsigned long long id = id0_S + (unsigned long long)i;
    const U4 r = philox4x32_10(seed, id, step, RNG_STRETCH_Z + half);
    const U4 rj = philox4x32_10(seed, id, step, RNG_STRETCH_J + half);
    const double u = u53(r.x, r.y);
    // The proposal is written as an explicit sequence of correctly-rounded operations (fma / mul / div / sub), the
    // same one the CPU oracle executes (oracle/rv_samplers.c), so positions are bit-identical functions of the
    // accept/reject history: a likelihood that differs by rounding can then only change a decision when
    // |lnpdiff - ln u| < ~1e-11, instead of seeding an error that the stretch map amplifies (E[ln zz] > 0).
    const double t = fma(a - 1.0, u, 1.0);
    const double zz = __ddiv_rn(__dmul_rn(t, t), a);
    const long long j = (long long)(((unsigned long long)rj.x * (unsigned long long)nC) >> 32);
    for (int v = 0; v < nvars; v++) {
        const double c = C[j * nvars + v];
        q[i * nvars + v] = fma(-zz, __dsub_rn(c, S[i * nvars + v]), c);
    }
    zz_out[i] = zz;
}

__global__ void stretch_accept_kernel(double* __restrict__ S, double* __restrict__ lnp, const double* __restrict__ q,
                                      const double* __restrict__ q_lnp, const int* __restrict__ q_status,
                                      const double* __restrict__ zz, int nvars, long long nS,
                                      unsigned long long seed, unsigned long long id0_S, unsigned step, unsigned half,
                                      unsigned long long* __restrict__ n_accept, unsigned char* __restrict__ accepted) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nS) return;
    const unsigned long long id = id0_S + (unsigned long long)i;
    const U4 r = philox4x32_10(seed, id, step, RNG_STRETCH_Z + half);
    const double u = u53(r.z, r.w);
    const double newlnp = (q_status[i] == ST_OK) ? q_lnp[i] : -INFINITY;     // lnprob(): -inf on any failure (mcmc.py:28-35)
    const double lnpdiff = (double)(nvars - 1) * log(zz[i]) + newlnp - lnp[i];
    const bool acc = lnpdiff > log(u);
    if (acc) {
        for (int v = 0; v < nvars; v++) S[i * nvars + v] = q[i * nvars + v];
        lnp[i] = newlnp;
        if (n_accept) n_accept[i] += 1ull;
    }
    if (accepted) accepted[i] = acc ? 1 : 0;
}

// Multi-GPU stretch move: accept + exchange in one kernel.  Every GPU of the group holds a full copy of the ensemble
// (positions and lnp); the GPU that owns a slice decides its accepts and writes each accepted walker straight into every
// copy -- its own and, through peer-mapped pointers (NVLink P2P stores), the other GPUs' -- so there is no separate
// all-gather step; cross-device ordering is by events on the streams (rv_stretch_run_multi).
__global__ void stretch_accept_peer_kernel(PeerCopies pc, int self, long long row0, const double* __restrict__ q,
                                           const double* __restrict__ q_lnp, const int* __restrict__ q_status,
                                           const double* __restrict__ zz, int nvars, long long nS, unsigned long long seed,
                                           unsigned long long id0_S, unsigned step, unsigned half,
                                           unsigned long long* __restrict__ n_accept) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nS) return;
    const unsigned long long id = id0_S + (unsigned long long)i;
    const U4 r = philox4x32_10(seed, id, step, RNG_STRETCH_Z + half);
    const double u = u53(r.z, r.w);
    const double newlnp = (q_status[i] == ST_OK) ? q_lnp[i] : -INFINITY;
    const long long row = row0 + i;
    const double lnpdiff = (double)(nvars - 1) * log(zz[i]) + newlnp - pc.lnp[self][row];
    if (lnpdiff > log(u)) {
        for (int g = 0; g < pc.n; g++) {
            double* __restrict__ dst = pc.theta[g] + row * nvars;
            for (int v = 0; v < nvars; v++) dst[v] = q[i * nvars + v];
            pc.lnp[g][row] = newlnp;
        }
        if (n_accept) n_accept[i] += 1ull;
    }
}

__global__ void mask_logp_kernel(double* __restrict__ logp, const int* __restrict__ status, long long W) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < W && status[w] != ST_OK) logp[w] = -INFINITY;
}

static inline unsigned nblk(long long n, int nt) { return (unsigned)((n + nt - 1) / nt > 0 ? (n + nt - 1) / nt : 1); }

cudaError_t launch_mh_propose(const double* theta, const double* scales, double step_size, int nvars, long long W,
                              unsigned long long seed, unsigned long long first_id, unsigned step, double* prop,
                              cudaStream_t s) {
    mh_propose_kernel<<<nblk(W, 128), 128, 0, s>>>(theta, scales, step_size, nvars, W, seed, first_id, step, prop);
    return cudaGetLastError();
}
cudaError_t launch_mh_accept(double* theta, double* logp, const double* prop, const double* prop_logp,
                             const int* prop_status, int nvars, long long W, unsigned long long seed,
                             unsigned long long first_id, unsigned step, unsigned long long* n_accept,
                             unsigned char* accepted, double* chain_row, double* chain_logp_row, long long chain_w,
                             cudaStream_t s) {
    mh_accept_kernel<<<nblk(W, 128), 128, 0, s>>>(theta, logp, prop, prop_logp, prop_status, nvars, W, seed, first_id,
                                                  step, n_accept, accepted, chain_row, chain_logp_row, chain_w);
    return cudaGetLastError();
}
cudaError_t launch_stretch_propose(const double* S, const double* C, int nvars, long long nS, long long nC, double a,
                                   unsigned long long seed, unsigned long long id0_S, unsigned step, unsigned half,
                                   double* q, double* zz, cudaStream_t s) {
    stretch_propose_kernel<<<nblk(nS, 128), 128, 0, s>>>(S, C, nvars, nS, nC, a, seed, id0_S, step, half, q, zz);
    return cudaGetLastError();
}
cudaError_t launch_stretch_accept(double* S, double* lnp, const double* q, const double* q_lnp, const int* q_status,
                                  const double* zz, int nvars, long long nS, unsigned long long seed,
                                  unsigned long long id0_S, unsigned step, unsigned half, unsigned long long* n_accept,
                                  unsigned char* accepted, cudaStream_t s) {
    stretch_accept_kernel<<<nblk(nS, 128), 128, 0, s>>>(S, lnp, q, q_lnp, q_status, zz, nvars, nS, seed, id0_S, step,
                                                        half, n_accept, accepted);
    return cudaGetLastError();
}
cudaError_t launch_stretch_accept_peer(const PeerCopies& pc, int self, long long row0, const double* q, const double* q_lnp,
                                       const int* q_status, const double* zz, int nvars, long long nS, unsigned long long seed,
                                       unsigned long long id0_S, unsigned step, unsigned half, unsigned long long* n_accept,
                                       cudaStream_t s) {
    stretch_accept_peer_kernel<<<nblk(nS, 128), 128, 0, s>>>(pc, self, row0, q, q_lnp, q_status, zz, nvars, nS, seed, id0_S, step,
                                                             half, n_accept);
    return cudaGetLastError();
}
cudaError_t launch_mask_logp(double* logp, const int* status, long long W, cudaStream_t s) {
    mask_logp_kernel<<<nblk(W, 256), 256, 0, s>>>(logp, status, W);
    return cudaGetLastError();
}

}  // namespace rv
