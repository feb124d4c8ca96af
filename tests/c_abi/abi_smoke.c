/* Plain-C client of librvgpu.so: proves that include/rvgpu.h is a C header and that the boundary needs nothing but
 * pointers and sizes (no Python, no torch).  Usage: abi_smoke <file.vels> ; prints logp of the published HD155358
 * solution (KAT-2: -2.41616612321) and the gradient's first component.  Built and run by tests/test_c_abi.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rvgpu.h"

#define CHECK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, rv_last_error(ctx)); return 2; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.vels\n", argv[0]); return 1; }
    FILE* f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 1; }
    static double t[4096], rv[4096], er[4096];
    int n = 0;
    while (n < 4096 && fscanf(f, "%lf %lf %lf", &t[n], &rv[n], &er[n]) == 3) n++;
    fclose(f);
    /* observations.py:52-69: unit factors, array_split (first half gets the extra row), shift */
    const int nb = (n + 1) / 2, nf = n - nb;
    for (int i = 0; i < n; i++) { t[i] *= 0.01720; rv[i] *= 3.355e-5; er[i] *= 3.355e-5; }
    const double shift = t[nb - 1];
    for (int i = 0; i < n; i++) t[i] -= shift;

    rv_ctx* ctx = NULL;
    int rc = rv_ctx_create(0, &ctx);
    if (rc != 0) { fprintf(stderr, "rv_ctx_create failed (%d): %s\n", rc, rv_last_error(NULL)); return 3; }
    rv_obs* obs = NULL;
    CHECK(rv_obs_create(ctx, t + nb, rv + nb, er + nb, nf, t, rv, er, nb, 100.0, &obs));
    /* two planets, free parameters a,h,k,m,l per planet (the reference's get_params() order) */
    double fixed[2 * RV_NELEM];
    memset(fixed, 0, sizeof fixed);
    const int32_t fp[10] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1};
    const int32_t fe[10] = {RV_EL_A, RV_EL_H, RV_EL_K, RV_EL_M, RV_EL_L, RV_EL_A, RV_EL_H, RV_EL_K, RV_EL_M, RV_EL_L};
    rv_model* model = NULL;
    CHECK(rv_model_create(ctx, 2, fixed, 10, fp, fe, 2.0, 0, &model));
    const double theta[10] = {6.57730330e-01, -9.72263877e-02, -7.82798396e-02, 8.84031737e-04, 4.42804990e+00,
                              1.04404207e+00, -2.05622789e-02, -1.08797961e-01, 8.30379710e-04, 1.49919861e+00};
    double logp = 0, logp2 = 0, grad[10], hess[100];
    int32_t status = -1, status2 = -1;
    CHECK(rv_loglik(ctx, model, obs, theta, 1, &logp, &status));
    CHECK(rv_loglik_d_dd(ctx, model, obs, theta, 1, &logp2, grad, hess, &status2));
    printf("status %d logp %.14f status_dd %d logp_dd %.14f grad0 %.10g hess00 %.10g\n", status, logp, status2, logp2, grad[0], hess[0]);
    rv_model_destroy(model);
    rv_obs_destroy(obs);
    rv_ctx_destroy(ctx);
    return 0;
}
