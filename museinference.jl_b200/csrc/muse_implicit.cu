// muse_implicit.cu — the implicit-differentiation branch of get_H! (/root/reference/src/muse.jl:335-405) for the registered
// families: per sim k of the H shard
//     ẑ  = MAP of sim k at θ₀ from zero(z) / the user's z₀, ∇z_logLike_atol = 1e-1 (hard-coded there, :346)
//     H1 = ∂θ_sim [∇θ′ logLike(x(θ_sim), ẑ, θ′)]                — 0 for these families: their score does not depend on x
//     H2 = −(∂θ ∇z logLike)ᵀ · A⁻¹ · (∂θ_sim ∇z logLike),   A = ∇²z logLike, one conjugate-gradient solve per column     (:361-382)
//     H  = H1 + H2.
// The reference gets every derivative by nested AD; for the registered families they are closed forms (oracle/families.py states
// and tests them), so what is left on the device is the MAP pass, a handful of dot products per sim and — for the dense
// correlated Gaussian, whose A = −(I + aP) is not a multiple of the identity — a batched conjugate-gradient solve whose one
// product per iteration is the (sims × d)·(d × d) DGEMM of the lock-step solver (muse_corr.cu: muse_corr_implicit_h).
//   F1: ∂θ∇z = a·ẑ,                 ∂θ_sim∇z = ½σξ,        A = −(1 + a)I  ⇒  H = ½ a σ (ẑ·ξ) / (1 + a)
//   F2: ∂θ∇z = [a·1, 2a(ẑ − μ)],    ∂θ_sim∇z = [1, σξ],    A = −(1 + a)I  ⇒  H = [a·d, aσΣξ; 2aΣ(ẑ−μ), 2aσΣ(ẑ−μ)ξ] / (1 + a)
// (CG on a multiple of the identity ends after one iteration with A⁻¹b exactly: that iteration count is what is reported.)
#include <cmath>
#include <vector>

#include "muse_handle.cuh"

using namespace muse;

int muse_corr_implicit_h(muse_handle* h, const double* theta0, int nsims_H, int start, int cg_maxiter, double* Hs_out, int32_t* cg_iters_out,
                         int32_t* status_out);

namespace {

// per sim: Σξ, Σ(ẑ − μ), Σ(ẑ − μ)·ξ   (thread-strided partial sums, warp butterflies, warps in order: a fixed tree)
__global__ void __launch_bounds__(256) implicit_dots_kernel(const double* __restrict__ xi, const double* __restrict__ zA, const double* __restrict__ zB,
                                                            const int* __restrict__ zstate, int d, int ld, double mu, double* __restrict__ out) {
    __shared__ double red[3][8];
    const int k = blockIdx.x;
    const int st = zstate[1 + k];
    const double* z = st == kZA ? zA + (size_t)(1 + k) * ld : (st == kZB ? zB + (size_t)(1 + k) * ld : nullptr);
    const double* x = xi + (size_t)k * ld;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int j = threadIdx.x; j < d; j += 256) {
        const double w = (z ? z[j] : 0.0) - mu, e = x[j];
        s0 += e;
        s1 += w;
        s2 = fma(w, e, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; red[2][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = red[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) t += red[threadIdx.x][w];
        out[(size_t)k * 3 + threadIdx.x] = t;
    }
}

}  // namespace

extern "C" int muse_b200_implicit_h(muse_handle* h, const double* theta0, int32_t nsims_H, int32_t start, int32_t cg_maxiter, double* Hs_out,
                                    int32_t* cg_iters_out, int32_t* status_out) {
    if (!h || !theta0 || !Hs_out || nsims_H < 0) return MUSE_EINVAL;
    if (start != MUSE_START_ZEROS && start != MUSE_START_USER) { h->err = "implicit_h: start must be ZEROS or USER"; return MUSE_EINVAL; }
    if (h->cfg.nsims_h > 0) { h->err = "implicit_h: not provided on a handle with a separate H shard (multi-GPU)"; return MUSE_EUNSUPPORTED; }
    if (nsims_H > h->cfg.nsims) { h->err = "nsims_H outside the handle's shard"; return MUSE_EINVAL; }
    if (!h->have_draws) { h->err = "no draws installed (set_draws / seed_draws)"; return MUSE_ESTATE; }
    if (start == MUSE_START_USER && !h->have_z0) { h->err = "user z0 not set (muse_b200_set_z0)"; return MUSE_ESTATE; }
    if (nsims_H == 0) return MUSE_OK;
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) { h->err = "cudaSetDevice"; return MUSE_ECUDA; }
    if (h->corr) return muse_corr_implicit_h(h, theta0, nsims_H, start, cg_maxiter > 0 ? cg_maxiter : 100, Hs_out, cg_iters_out, status_out);

    const int nt = h->cfg.ntheta, d = h->cfg.d;
    // (1) the MAP of every H sim at θ₀, ∇z_logLike_atol = 1e-1 (:346); ẑ stays resident
    int rc = muse_b200_map_score_async(h, theta0, theta0, 1e-1, 0, start, 0, nsims_H);
    if (rc != MUSE_OK) return rc;
    std::vector<int32_t> status((size_t)nsims_H);
    rc = muse_b200_fetch(h, nsims_H, nullptr, nullptr, nullptr, nullptr, status.data());
    if (rc != MUSE_OK) return rc;
    if (status_out) std::copy(status.begin(), status.end(), status_out);
    // (2) the dot products
    double* dots_d = nullptr;
    if (cudaMalloc(&dots_d, (size_t)nsims_H * 3 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); h->err = "implicit_h: allocation failed"; return MUSE_ENOMEM; }
    const double mu = h->cfg.family == MUSE_FAMILY_HIERGAUSS ? theta0[0] : 0.0;
    implicit_dots_kernel<<<nsims_H, 256, 0, h->stream>>>(h->xi, h->zA, h->zB, h->zstate, d, h->ld, mu, dots_d);
    h->acc.launches += 1;
    std::vector<double> dots((size_t)nsims_H * 3);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(dots.data(), dots_d, dots.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dots_d);
    if (e != cudaSuccess) { h->err = std::string("implicit_h: ") + cudaGetErrorString(e); return MUSE_ECUDA; }
    // (3) H2 = −(∂θ∇z)ᵀ · A⁻¹ · (∂θ_sim∇z) with A⁻¹ = −1/(1 + a)
    for (int k = 0; k < nsims_H; ++k) {
        const double s_xi = dots[(size_t)k * 3], s_w = dots[(size_t)k * 3 + 1], s_wxi = dots[(size_t)k * 3 + 2];
        double* H = Hs_out + (size_t)k * nt * nt;
        if (h->cfg.family == MUSE_FAMILY_FUNNEL) {
            const double a = std::exp(-theta0[0]), sig = std::exp(0.5 * theta0[0]);
            H[0] = a * (0.5 * sig) * s_wxi / (1.0 + a);
            if (cg_iters_out) cg_iters_out[k] = 1;
        } else {
            const double a = std::exp(-2.0 * theta0[1]), sig = std::exp(theta0[1]), c = 1.0 / (1.0 + a);
            H[0] = a * (double)d * c;          H[1] = a * sig * s_xi * c;
            H[2] = 2.0 * a * s_w * c;          H[3] = 2.0 * a * sig * s_wxi * c;
            if (cg_iters_out) { cg_iters_out[2 * k] = 1; cg_iters_out[2 * k + 1] = 1; }
        }
    }
    return MUSE_OK;
}
