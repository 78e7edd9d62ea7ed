// Test infrastructure: host replay of the per-thread routine of k_tangent_hypel_tiled
// (insilico_b200/csrc/isl_tangent_tiled.cuh) against the defining formula of the hyperelastic tangent.
#include <cstddef>
#include <cstring>

#include "../../insilico_b200/csrc/isl_tangent_tiled.cuh"

template <int DIM>
static void run(const double* Gt, const double* Gc, const double* Q, const double* det, const double* w, int nq, int nt,
                int nc, double* K /*[nt*DIM][nc*DIM]*/) {
    constexpr int MC = 6;
    const int nchunk = (nt + MC - 1) / MC;
    for (int N = 0; N < nc; N++)
        for (int ch = 0; ch < nchunk; ch++) {
            double acc[MC][DIM * DIM];
            isl_hypel_tile<DIM, MC>(Gt, Gc, Q, det, w, nq, nt, nc, N, ch * MC, acc);
            for (int m = 0; m < MC; m++) {
                const int M = ch * MC + m;
                if (M >= nt) continue;
                for (int i = 0; i < DIM; i++)
                    for (int k = 0; k < DIM; k++) K[(size_t)(M * DIM + i) * (nc * DIM) + N * DIM + k] += acc[m][i * DIM + k];
            }
        }
}

extern "C" void emu_hypel_tile(int dim, const double* Gt, const double* Gc, const double* Q, const double* det, const double* w,
                               int nq, int nt, int nc, double* K) {
    if (dim == 3) run<3>(Gt, Gc, Q, det, w, nq, nt, nc, K);
    else run<2>(Gt, Gc, Q, det, w, nq, nt, nc, K);
}
