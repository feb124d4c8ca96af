// rv_var_kernels.cu -- sm_100a kernel for State.get_logp_d_dd (value + gradient + Hessian, SMALA).
//
// var_kernel: persistent grid; every CTA pulls (walker, leg) items and integrates the real system together
// with its first- and second-order variational sets, one thread per (set, planet) -- see rv_var.cuh.
#include <cuda_runtime.h>
#include "rv_launch.h"
#include "rv_var.cuh"
#include "rv_var2.cuh"

namespace rv {

template <int P, int D>
struct DevVarExec {
    VarThread<P, D>& th;
    double* red;                 // [2][64] ping-pong maxima, then the item broadcast slot
    int parity;
    int nwarps;
    template <class F>
    __device__ __forceinline__ void each(F&& f) { f(th); }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void stage_max(const VarThread<P, D>&, double a, double b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if ((threadIdx.x & 31) == 0) {
            red[parity * 64 + 2 * (threadIdx.x >> 5)] = a;
            red[parity * 64 + 2 * (threadIdx.x >> 5) + 1] = b;
        }
    }
    __device__ __forceinline__ void read_max(double& a, double& b) {
        a = 0.0; b = 0.0;
        for (int w = 0; w < nwarps; w++) {
            a = fmax(a, red[parity * 64 + 2 * w]);
            b = fmax(b, red[parity * 64 + 2 * w + 1]);
        }
        parity ^= 1;
    }
    __device__ __forceinline__ long long fetch(unsigned long long* ctr) {
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(red + 128);
        __syncthreads();
        if (threadIdx.x == 0) *slot = atomicAdd(ctr, 1ull);
        __syncthreads();
        return (long long)*slot;
    }
    __device__ __forceinline__ void add_work(unsigned long long* wc, unsigned long long nf, unsigned long long na) {
        if (wc && threadIdx.x == 0) { atomicAdd(&wc[0], nf); atomicAdd(&wc[1], na); }
    }
};

template <int P, int D, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) var_kernel(const VarArgs a, const VarLayout L) {
    extern __shared__ __align__(16) double sm[];
    VarThread<P, D> th;
    var_assign(th, (int)threadIdx.x, L);
    DevVarExec<P, D> ex{th, sm + L.o_red, 0, NT / 32};
    var_run_items<P, D>(ex, a, L, sm);
}

// ---- "one lane per variational set" layout (rv_var2.cuh): producer warp + named barriers ----------------------------
template <int P, int D>
struct DevVar2Exec {
    Var2Thread<P, D>& th;
    double* red;                 // [2][64] ping-pong maxima, then the item broadcast slot
    int parity, nwarps, nt;
    bool producer;               // this thread's warp is the producer warp (real + first-order sets)
    template <class F>
    __device__ __forceinline__ void each(F&& f) { f(th); }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void producer_sync() { if (producer) __syncwarp(); }
    // named barriers 1..7 (0 is __syncthreads): the producer warp arrives without waiting, the others wait
    __device__ __forceinline__ void signal(int n) {
        if (producer) asm volatile("bar.arrive %0, %1;" ::"r"(n), "r"(nt) : "memory");
    }
    __device__ __forceinline__ void wait(int n) {
        if (!producer) asm volatile("bar.sync %0, %1;" ::"r"(n), "r"(nt) : "memory");
    }
    __device__ __forceinline__ void stage_max(const Var2Thread<P, D>&, double a, double b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if ((threadIdx.x & 31) == 0) {
            red[parity * 64 + 2 * (threadIdx.x >> 5)] = a;
            red[parity * 64 + 2 * (threadIdx.x >> 5) + 1] = b;
        }
    }
    __device__ __forceinline__ void read_max(double& a, double& b) {
        a = 0.0; b = 0.0;
        for (int w = 0; w < nwarps; w++) {
            a = fmax(a, red[parity * 64 + 2 * w]);
            b = fmax(b, red[parity * 64 + 2 * w + 1]);
        }
        parity ^= 1;
    }
    __device__ __forceinline__ long long fetch(unsigned long long* ctr) {
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(red + 128);
        __syncthreads();
        if (threadIdx.x == 0) *slot = atomicAdd(ctr, 1ull);
        __syncthreads();
        return (long long)*slot;
    }
    __device__ __forceinline__ void add_work(unsigned long long* wc, unsigned long long nf, unsigned long long na) {
        if (wc && threadIdx.x == 0) { atomicAdd(&wc[0], nf); atomicAdd(&wc[1], na); }
    }
};

// MAXR: register cap per thread (__maxnreg__; __launch_bounds__ rounds a 96-thread CTA up to 128 threads when it derives
// the cap from a CTAs-per-SM target, which would leave 3 x 96 threads only 168 registers each)
template <int P, int D, int NT, int MAXR>
__global__ void __maxnreg__(MAXR) var2_kernel(const VarArgs a, const Var2Layout L) {
    extern __shared__ __align__(16) double sm[];
    Var2Thread<P, D> th;
    var2_assign(th, (int)threadIdx.x, L);
    DevVar2Exec<P, D> ex{th, sm + L.o_red, 0, NT / 32, NT, (int)(threadIdx.x >> 5) == L.nso_warps};
    var2_run_items<P, D>(ex, a, L, sm);
}

template <int P, int D, int NT, int MAXR>
static cudaError_t launch_var2_one(const VarArgs& a, int nv, int num_sms, cudaStream_t stream) {
    auto kern = var2_kernel<P, D, NT, MAXR>;
    const Var2Layout L = var2_layout(P, D, nv, NT);
    if (var2_min_threads(nv) > NT) return cudaErrorInvalidConfiguration;
    const size_t smem = sizeof(double) * (size_t)L.total;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    long long blocks = (long long)num_sms * occ;
    if (2 * a.W < blocks) blocks = 2 * a.W;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, NT, smem, stream>>>(a, L);
    return cudaGetLastError();
}

// logp = -(chi2b + chi2f), grad = -(db + df), hess symmetric (state.py:285,292-293)
__global__ void var_finalize_kernel(const double* __restrict__ part, const int* __restrict__ pstat, long long W, int nv,
                                    int nsets, double* __restrict__ logp, double* __restrict__ grad,
                                    double* __restrict__ hess, int* __restrict__ status,
                                    unsigned long long* item_counter) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) *item_counter = 0ull;
    if (idx >= W * nsets) return;
    const long long w = idx / nsets;
    const int s = (int)(idx - w * nsets);
    const int sb = pstat[w], sf = pstat[W + w];
    const int st = (sf != ST_OK) ? sf : sb;
    const bool ok = st == ST_OK;
    const double v = ok ? -(part[w * nsets + s] + part[(W + w) * nsets + s]) : 0.0;
    if (s == 0) {
        status[w] = st;
        logp[w] = ok ? v : -INFINITY;
    } else if (s <= nv) {
        grad[w * nv + (s - 1)] = v;
    } else {
        const int k = s - 1 - nv;
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= k) a++;
        const int b = k - a * (a + 1) / 2;
        hess[(w * nv + a) * nv + b] = v;
        hess[(w * nv + b) * nv + a] = v;
    }
}

template <int P, int D, int NT, int MINB>
static cudaError_t launch_var_one(const VarArgs& a, int nv, int num_sms, cudaStream_t stream) {
    auto kern = var_kernel<P, D, NT, MINB>;
    const VarLayout L = var_layout(P, D, nv, NT);
    if (L.need > NT) return cudaErrorInvalidConfiguration;
    const size_t smem = sizeof(double) * (size_t)L.total;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    long long blocks = (long long)num_sms * occ;
    if (2 * a.W < blocks) blocks = 2 * a.W;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, NT, smem, stream>>>(a, L);
    return cudaGetLastError();
}

// threads needed by a model (host-side check shared with the ABI)
int var_threads_needed(int P, int nv) {
    const VarLayout L = var_layout(P, 2, nv, 32);
    return L.need;
}

// layout: 0 = automatic (set-per-lane kernel where it exists: one or two planets), 1 = thread per (set, planet)
cudaError_t launch_var(const VarArgs& a, int P, int D, int nv, int layout, int num_sms, cudaStream_t stream) {
    if (layout == 2 && P == 2 && D == 2 && var2_min_threads(nv) <= 96)       // tuning: 168 registers -> 3 CTAs per SM
        return launch_var2_one<2, 2, 96, 168>(a, nv, num_sms, stream);
    if (layout == 0 && var2_supported(P, nv)) {
        const int nt = var2_min_threads(nv);
        if (P == 1 && D == 2 && nt <= 64) return launch_var2_one<1, 2, 64, 168>(a, nv, num_sms, stream);
        if (P == 1 && D == 3 && nt <= 64) return launch_var2_one<1, 3, 64, 224>(a, nv, num_sms, stream);
        if (P == 2 && D == 2 && nt <= 64) return launch_var2_one<2, 2, 64, 224>(a, nv, num_sms, stream);
        if (P == 2 && D == 2 && nt <= 96) return launch_var2_one<2, 2, 96, 224>(a, nv, num_sms, stream);
        if (P == 2 && D == 2 && nt <= 160) return launch_var2_one<2, 2, 160, 200>(a, nv, num_sms, stream);
    }
    const int need = var_threads_needed(P, nv);
    if (P == 1 && D == 2 && need <= 64) return launch_var_one<1, 2, 64, 4>(a, nv, num_sms, stream);
    if (P == 1 && D == 3 && need <= 64) return launch_var_one<1, 3, 64, 4>(a, nv, num_sms, stream);
    if (P == 2 && D == 2 && need <= 64) return launch_var_one<2, 2, 64, 4>(a, nv, num_sms, stream);
    if (P == 2 && D == 2 && need <= 160) return launch_var_one<2, 2, 160, 3>(a, nv, num_sms, stream);
    if (P == 2 && D == 3 && need <= 256) return launch_var_one<2, 3, 256, 1>(a, nv, num_sms, stream);
    if (P == 3 && D == 2 && need <= 448) return launch_var_one<3, 2, 448, 1>(a, nv, num_sms, stream);
    if (P == 3 && D == 3 && need <= 448) return launch_var_one<3, 3, 448, 1>(a, nv, num_sms, stream);
    return cudaErrorInvalidConfiguration;
}

cudaError_t launch_var_finalize(const double* part, const int* pstat, long long W, int nv, double* logp, double* grad,
                                double* hess, int* status, unsigned long long* item_counter, cudaStream_t stream) {
    const int nsets = var_nsets(nv);
    const int nt = 256;
    const long long n = W * nsets;
    const unsigned nb = (unsigned)((n + nt - 1) / nt);
    var_finalize_kernel<<<nb ? nb : 1, nt, 0, stream>>>(part, pstat, W, nv, nsets, logp, grad, hess, status, item_counter);
    return cudaGetLastError();
}

}  // namespace rv
