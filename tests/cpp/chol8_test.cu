// tests/cpp/chol8_test.cu -- stand-alone check and timing of csrc/chol8.h against the 3x3-block LDL' of csrc/qp_solver.h.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cpp/chol8_test tests/cpp/chol8_test.cu && tests/cpp/chol8_test
// One 256-thread CTA per SM factorises a random SPD matrix of n = 3*nb variables held in shared memory and solves one
// right-hand side; prints the residual of both paths and clock64 cycles per factorisation + solve.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ long long g_c8p[8];
__device__ long long g_c8last;
#ifdef C8PROF   // phase marks cost ~500 cycles each (global read-modify-write on the critical path): off for the totals
#define QR_C8P(tag) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long n_ = clock64(); g_c8p[tag] += n_ - g_c8last; g_c8last = n_; } } while (0)
#endif
#include "../../quadruped-robot_b200/csrc/qp_solver.h"

constexpr int NT = 256;

__global__ void __launch_bounds__(NT, 1) k_test(const double* Kd, const double* rhs, int nred, int reps, double* xout,
                                                long long* cyc, int mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nb = (nred + 2) / 3, nt = qr_k8_nt(nred);
    const int ntri3 = nb * (nb + 1) / 2, ntri8 = nt * (nt + 1) / 2;
    double* K = reinterpret_cast<double*>(smem);
    const size_t kd = (size_t)(9 * ntri3 > 64 * ntri8 ? 9 * ntri3 : 64 * ntri8);
    double* y = K + kd;
    double* out = y + 8 * nt + 8;
    double* Dinv = out + 8 * nt + 8;
    double* xs = Dinv + 9 * nb;
    unsigned short* tri = reinterpret_cast<unsigned short*>(xs + 16);
    for (int idx = threadIdx.x; idx < ntri3; idx += NT) {
        int I, J;
        qr_tri_decode(idx, I, J);
        tri[idx] = (unsigned short)((I << 8) | J);
    }
    QrQpWork W;
    W.K = K; W.wv = y; W.Dinv = Dinv; W.tri = tri;
    __syncthreads();
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
        // dense lower triangle (row-major n x n in global) -> the layout of the path under test
        if (mode == 0) {
            for (int e = threadIdx.x; e < 9 * ntri3; e += NT) K[e] = 0.0;
            __syncthreads();
            for (int e = threadIdx.x; e < 3 * nb * 3 * nb; e += NT) {
                const int i = e / (3 * nb), j = e % (3 * nb);
                if (j > i) continue;
                const double v = (i < nred && j < nred) ? Kd[i * nred + j] : (i == j ? 1.0 : 0.0);
                const int I = i / 3, J = j / 3;
                double* blk = K + qr_kblk(nb, I, J);
                blk[3 * (i % 3) + j % 3] = v;
                if (I == J) blk[3 * (j % 3) + i % 3] = v;
            }
            for (int i = threadIdx.x; i < 3 * nb; i += NT) y[i] = i < nred ? rhs[i] : 0.0;
        } else {
            for (int e = threadIdx.x; e < 8 * nt * 8 * nt; e += NT) {
                const int i = e / (8 * nt), j = e % (8 * nt);
                if (j > i) continue;
                const double v = (i < nred && j < nred) ? Kd[i * nred + j] : (i == j ? 1.0 : 0.0);
                K[qr_k8_idx(nt, i, j)] = v;
                if ((i >> 3) == (j >> 3)) K[qr_k8_idx(nt, j, i)] = v;
            }
            for (int i = threadIdx.x; i < 8 * nt; i += NT) y[i] = i < nred ? rhs[i] : 0.0;
        }
        __syncthreads();
        if (mode == 1 && rep == 0 && nred == 47) {   // the diagonal factor alone: one warp, then all eight warps on private tiles
            double* tmp = K + kd - 8 * 64;
            for (int w8 = 0; w8 < 2; ++w8) {
                for (int e = threadIdx.x; e < 8 * 64; e += NT) { const int r = (e >> 3) & 7, c = e & 7; tmp[(e & ~63) + qr_k8_swz(r, c)] = (r == c ? 2.0 : 0.1 / (1 + r + c)); }
                __syncthreads();
                const long long a0 = clock64();
                if (w8 || threadIdx.x < 32) qr_chol8_diag(tmp + 64 * (threadIdx.x >> 5));
                const long long a1 = clock64();
                __syncthreads();
                if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && threadIdx.x < 64) printf("diag alone (%s), warp %d: %lld cycles\n", w8 ? "8 warps" : "1 warp", threadIdx.x >> 5, a1 - a0);
            }
        }
        const long long t0 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) g_c8last = t0;
        if (mode == 0) {
            qr_ldl_factor<NT>(W, nb, 1);
            qr_ldl_backward<NT>(W, nb, out);
        } else {
            qr_chol8_factor<NT>(K, y, tri, nt, 1);
            qr_chol8_backward<NT>(K, y, xs, nt, out, nred);
        }
        __syncthreads();
        total += clock64() - t0;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < nred; i += NT) xout[i] = out[i];
        if (threadIdx.x == 0) *cyc = total / reps;
    }
}

int main(int argc, char** argv) {
    const int reps = 20;
    for (int nb : {16, 20, 24, 32, 40, 45, 56, 64, 72}) {
        const int nred = 3 * nb - (nb % 3);   // not always a multiple of 3 or 8
        std::vector<double> A((size_t)nred * nred), K((size_t)nred * nred), b(nred);
        srand(nb);
        for (auto& v : A) v = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < nred; ++i)
            for (int j = 0; j < nred; ++j) {
                double s = i == j ? 0.5 : 0.0;
                for (int k = 0; k < nred; ++k) s += A[(size_t)i * nred + k] * A[(size_t)j * nred + k] / nred;
                K[(size_t)i * nred + j] = s;
            }
        for (auto& v : b) v = rand() / (double)RAND_MAX - 0.5;
        double *dK, *db, *dx;
        long long* dc;
        cudaMalloc(&dK, K.size() * 8); cudaMalloc(&db, nred * 8); cudaMalloc(&dx, nred * 8); cudaMalloc(&dc, 8);
        cudaMemcpy(dK, K.data(), K.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(db, b.data(), nred * 8, cudaMemcpyHostToDevice);
        const int nt = (nred + 7) / 8, nbb = (nred + 2) / 3;
        size_t kd = std::max((size_t)9 * nbb * (nbb + 1) / 2, (size_t)64 * nt * (nt + 1) / 2);
        size_t smem = (kd + 2 * (8 * nt + 8) + 9 * nbb + 16) * 8 + (size_t)nbb * (nbb + 1) / 2 * 2 + 64;
        cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        for (int mode = 0; mode < 2; ++mode) {
            k_test<<<148, NT, smem>>>(dK, db, nred, reps, dx, dc, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("nb %d mode %d: %s (smem %zu)\n", nb, mode, cudaGetErrorString(e), smem); return 1; }
            std::vector<double> x(nred);
            long long c;
            cudaMemcpy(x.data(), dx, nred * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            double res = 0.0, nrm = 0.0;
            for (int i = 0; i < nred; ++i) {
                double s = -b[i];
                for (int j = 0; j < nred; ++j) s += K[(size_t)i * nred + j] * x[j];
                res = std::max(res, std::fabs(s));
                nrm = std::max(nrm, std::fabs(x[i]));
            }
            if (mode) { long long pr[8]; cudaMemcpyFromSymbol(pr, g_c8p, sizeof(pr)); printf("   per call: first diag %lld, panel %lld, trailing (thread 0: tile 0 + diag %lld) %lld, backward %lld\n", pr[0]/reps, pr[1]/reps, pr[4]/reps, (pr[2]+pr[4])/reps, pr[3]/reps); long long z[8] = {0}; cudaMemcpyToSymbol(g_c8p, z, sizeof(z)); }
            printf("nred %3d (nb %2d, nt %2d) %s: residual %.2e (|x| %.2e)  %lld cycles per factor+solve\n", nred, nbb, nt,
                   mode ? "chol8 DMMA" : "ldl 3x3   ", res, nrm, c);
        }
        cudaFree(dK); cudaFree(db); cudaFree(dx); cudaFree(dc);
    }
    return 0;
}
