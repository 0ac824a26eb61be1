/* The C ABI used directly from C: the reference's demo voxel (signals.py:307-314: OEF 0.4, DBV 0.12) through
 * qbold_forward_backward, printed next to the known answers of SURVEY.md Appendix B.
 *
 *   nvcc -x c examples/c_abi_demo.c -I include -L qbold_vi_b200 -lqbold -Xlinker -rpath=$PWD/qbold_vi_b200 -o /tmp/c_abi_demo
 *   /tmp/c_abi_demo
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "qbold.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != 0) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, qbold_last_error());   \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void) {
    /* config (DEFAULT section of the reference's INI file) */
    QboldPhysics phys = {2.67513e8, 3.0, 2.64e-7, 0.074, 11.5, 3.0, 1.21, 1.58, 0.34};
    float taus[11];
    for (int i = 0; i < 11; ++i) taus[i] = -0.016f + (float)i * 0.008f;     /* tf.range(-0.016, 0.065, 0.008) */
    if (qbold_abi_version() != QBOLD_ABI_VERSION || qbold_params_sizeof() != (int)sizeof(QboldParams)) {
        fprintf(stderr, "header / library mismatch\n");
        return 1;
    }
    QboldParams* P = (QboldParams*)malloc(sizeof(QboldParams));
    CHECK(qbold_params_init(P, &phys, taus, 11, /*full_model=*/1, /*include_blood=*/1));

    const float h_x[2] = {0.4f, 0.12f};
    float *d_x, *d_sig, *d_grad;
    if (cudaMalloc((void**)&d_x, sizeof(h_x)) || cudaMalloc((void**)&d_sig, 11 * sizeof(float)) ||
        cudaMalloc((void**)&d_grad, 2 * sizeof(float))) {
        fprintf(stderr, "cudaMalloc failed\n");
        return 1;
    }
    cudaMemcpy(d_x, h_x, sizeof(h_x), cudaMemcpyHostToDevice);
    /* g_signal = NULL: gradient of sum_tau S, as tape.gradient(signal, x) in the reference's demo */
    CHECK(qbold_forward_backward(P, d_x, NULL, 1, d_sig, d_grad, NULL));
    float sig[11], grad[2];
    cudaMemcpy(sig, d_sig, sizeof(sig), cudaMemcpyDeviceToHost);
    cudaMemcpy(grad, d_grad, sizeof(grad), cudaMemcpyDeviceToHost);

    const double kat[11] = {0.37586412, 0.40909051, 0.42242562, 0.40909051, 0.37586412, 0.33616669,
                            0.29923692, 0.26730955, 0.23914064, 0.21358386, 0.19047615};
    const double kat_grad[2] = {-3.06411134, -7.29813321};
    double worst = 0.0;
    for (int t = 0; t < 11; ++t) {
        const double rel = fabs(sig[t] - kat[t]) / kat[t];
        if (rel > worst) worst = rel;
        printf("tau %+.3f  S = %.8f  (known answer %.8f)\n", taus[t], sig[t], kat[t]);
    }
    printf("dS/dOEF = %.6f (%.6f)   dS/dDBV = %.6f (%.6f)\n", grad[0], kat_grad[0], grad[1], kat_grad[1]);
    const double gerr = fmax(fabs(grad[0] - kat_grad[0]) / fabs(kat_grad[0]), fabs(grad[1] - kat_grad[1]) / fabs(kat_grad[1]));
    printf("max relative deviation: signal %.2e, gradient %.2e, %lld kernel launches\n", worst, gerr,
           (long long)qbold_launch_count());
    cudaFree(d_x);
    cudaFree(d_sig);
    cudaFree(d_grad);
    free(P);
    return (worst < 1e-5 && gerr < 1e-4) ? 0 : 2;
}
