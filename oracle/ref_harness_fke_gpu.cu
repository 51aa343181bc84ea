// oracle/ref_harness_fke_gpu.cu -- TEST INFRASTRUCTURE (checker / reported baseline, never the product path).
//
// extern "C" entry to the reference's OWN CUDA movement kernel: updateMovement, launched through launchAdvDiffKernel_MultiSp
// (Rietkerk_FPE/kernel_advdiff_FKE2D.cu:92-180, :226-236) after copyToDeviceConstantMemory_AdvTerms (:29-88).  The reference .cu is
// compiled UNMODIFIED, in place, for sm_100a by oracle/Makefile (-DSPB=2: CuSpNV = 1 moving species) and linked with this file into
// oracle/_ref/libtm_ref_fke_gpu.so.  Used by tests/test_zzz_fke_reference_gpu_kernel.py (the product's gather-form movement step
// against the reference's scatter kernel on the same device buffers) and by bench.py --workload fke (the reference kernel timed
// beside the product's, VERDICT r1 "missing" 7).  Nothing under trophic_maelstorm_b200/ refers to it.
#include <cuda_runtime.h>
#include <utility>

// the reference's declarations (Rietkerk_FPE/rietkerk_bjork_basic.h:457-467), C++ linkage
cudaError_t copyToDeviceConstantMemory_AdvTerms(const int *sig_D_ScaledBounds, const int2 *sigma_vD_ScBounds, const double *mu_vel_prefactor,
                                                const int *d_size_gauss_D, const int *d_size_gauss_VXY);
void launchAdvDiffKernel_MultiSp(double *Rho_t, double *Rho_tsar, double *gamma, std::pair<double, double> *v_eff,
                                 double *gaussian_stencil_D, double *gaussian_stencil_VXY, int L);

extern "C" {

// the constant-memory set-up of one moving species: diffusion half-width, advection half-widths (.x rows, .y columns: the CUDA kernel's
// convention, kernel_advdiff_FKE2D.cu:126-127), mu_vel_prefactor; both stencils start at offset 0 of their arrays
int ref_fke_gpu_setup(int bound_D, int bvx, int bvy, double prefactor)
{
    const int bD[1] = {bound_D};
    const int2 bV[1] = {make_int2(bvx, bvy)};
    const double pf[1] = {prefactor};
    const int off[1] = {0};
    return (int)copyToDeviceConstantMemory_AdvTerms(bD, bV, pf, off, off);
}

// The reference's per-step choreography on the device (rietkerk_bjork_FKE.cpp:4282-4299 without the host copies): Rho_tsar zeroed,
// one launch (legacy default stream, as the reference launches it).  `reps` repetitions; *ms_avg = CUDA-event time of one
// (memset + launch), *ms_kernel_avg = of the launch alone.  All pointers are caller-owned DEVICE pointers.  Returns the cudaError_t.
int ref_fke_gpu_move(int L, void *d_rho, void *d_tsar, void *d_gamma, void *d_veff, void *d_gD, void *d_gV, int reps, float *ms_avg,
                     float *ms_kernel_avg)
{
    cudaEvent_t e0, e1, e2;
    cudaError_t err;
    if ((err = cudaEventCreate(&e0)) != cudaSuccess) return (int)err;
    cudaEventCreate(&e1); cudaEventCreate(&e2);
    const size_t bytes = (size_t)L * L * sizeof(double);
    float tot = 0.f, totk = 0.f;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0, 0);
        cudaMemsetAsync(d_tsar, 0, bytes, 0);
        cudaEventRecord(e1, 0);
        launchAdvDiffKernel_MultiSp((double *)d_rho, (double *)d_tsar, (double *)d_gamma, (std::pair<double, double> *)d_veff,
                                    (double *)d_gD, (double *)d_gV, L);
        cudaEventRecord(e2, 0);
        if ((err = cudaEventSynchronize(e2)) != cudaSuccess) break;
        if ((err = cudaGetLastError()) != cudaSuccess) break;
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, e0, e2); cudaEventElapsedTime(&b, e1, e2);
        tot += a; totk += b;
    }
    if (ms_avg) *ms_avg = tot / reps;
    if (ms_kernel_avg) *ms_kernel_avg = totk / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return (int)err;
}

} // extern "C"
