// Cholesky whitening of a d x d covariance on the device:  C = L L^T,  T = L^-T  (so that T^T C T = I), as f32.
//
// Why this exists (VERDICT r1 "next" #1).  The reference whitens every iterate with the PCA
// transform V diag(1/sqrt(max(lambda, 1e-10))) (pycleora/__init__.py:145-156).  Every whitening matrix of C has the
// form C^-1/2 Q with Q orthogonal, and the loop body is equivariant under right-multiplication of the iterate by an
// orthogonal Q:  A (X Q) = (A X) Q,  row L2 norms are unchanged,  cov -> Q^T C Q,  eigenvectors -> Q^T V, and the
// PCA-whitened output is the SAME matrix.  So iterations 1 .. T-1 may use ANY exact whitening -- here the Cholesky
// factor, a d^3/3-flop f64 kernel that stays on one SM -- and only the LAST iteration (and every iterate that is
// handed to the caller: callbacks, multiscale taps, the rmse early stop) needs the eigendecomposition.  That removes
// the serial `eigh` (2.3-4.7 ms at d = 256) from the critical path of every iteration on every GPU count.
// Guard: PCA-with-clamp is a true whitening only while lambda_min >= 1e-10.  trace(C^-1) = |L^-1|_F^2 >= 1/lambda_min
// is a by-product here; if it exceeds the limit, or a pivot is not positive, status[0] is raised (sticky) and the
// caller re-runs the loop with the eigensolver in every iteration.
//
// Algorithm: blocked right-looking Cholesky on the augmented matrix M = [C | I] (row operations only), 32 x 32 blocks,
// one CTA, f64 throughout:
//   per block step k:  P1  warp 0 factors the diagonal block (row per lane, shuffles) and inverts the factor;
//                      P2  row block k  <-  inv(L_kk) * row block k   (thread per column; this is U_k = L_.k^T in the
//                          C part and Z_k = rows of L^-1 in the identity part), staged in shared memory;
//                      P3  row blocks i > k:  M_i -= U_ki^T * panel   (warp per 32 x 32 block, lane per column,
//                          32 f64 accumulators per lane, operands broadcast from shared memory).
// Only blocks that are read later are updated (upper triangle of the C part, block columns <= k of the identity part),
// so the flop count is 2 * d^3/3.  d is padded to a multiple of 32 with an identity block.
#include "device.cuh"

namespace cleora {
namespace chol {

constexpr int NB = 32;
constexpr int THREADS = 512;       // 16 warps x 128 registers: the 32 f64 accumulators per lane need them
constexpr int LDS_ = NB + 1;       // padded leading dimension of the small shared blocks
constexpr unsigned FULL = 0xffffffffu;

__global__ void __launch_bounds__(THREADS, 1)
chol_whiten_kernel(const double *__restrict__ cov, int d, int dp, double *M /* [dp][2 dp], read and written */,
                   float *__restrict__ T, int *__restrict__ status, double trace_limit) {
    extern __shared__ __align__(16) double sm[];
    double *Ps = sm;                          // [NB][dp]   scaled row block k (the panel)
    double *Lk = Ps + (size_t)NB * dp;        // [NB][LDS_] diagonal factor
    double *Li = Lk + NB * LDS_;              // [NB][LDS_] its inverse
    double *dinv = Li + NB * LDS_;            // [NB]       1 / L_jj
    __shared__ int s_bad;
    __shared__ double s_red[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = dp / NB, ldm = 2 * dp;

    for (int idx = tid; idx < dp * ldm; idx += THREADS) {              // M = [ pad(C) | I ]
        const int r = idx / ldm, c = idx - r * ldm;
        double v;
        if (c < dp) v = (r < d && c < d) ? cov[(size_t)r * d + c] : (r == c ? 1.0 : 0.0);
        else v = (c - dp == r) ? 1.0 : 0.0;
        M[idx] = v;
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();

    for (int k = 0; k < nb; ++k) {
        // ------------------------------------------------------------------------------------------ P1
        if (warp == 0) {
            double a[NB];                                              // row `lane` of the diagonal block
#pragma unroll
            for (int c = 0; c < NB; ++c) a[c] = M[(size_t)(k * NB + lane) * ldm + k * NB + c];
            bool bad = false;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                double ajj = __shfl_sync(FULL, a[j], j);
                if (!(ajj > 0.0)) { bad = true; ajj = 1.0; }           // not positive definite: flag, keep finite
                const double inv = 1.0 / sqrt(ajj);
                const double lj = a[j] * inv;                          // L[lane][j] (meaningful for lane >= j)
                a[j] = lj;
                if (lane == j) dinv[j] = inv;
#pragma unroll
                for (int c = j + 1; c < NB; ++c) a[c] = fma(-lj, __shfl_sync(FULL, lj, c), a[c]);
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) Lk[lane * LDS_ + c] = (c <= lane) ? a[c] : 0.0;
            __syncwarp();
            double x[NB];                                              // column `lane` of inv(L_kk), forward substitution
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double s = (r == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int q = 0; q < r; ++q) s = fma(-Lk[r * LDS_ + q], x[q], s);
                x[r] = s * dinv[r];
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) Li[r * LDS_ + lane] = x[r];
            if (bad && lane == 0) s_bad = 1;
        }
        __syncthreads();
        // ------------------------------------------------------------------------------------------ P2
        // panel column block pc:  pc < nb-k-1  -> C part, block column k+1+pc;   else identity part, block column
        // pc-(nb-k-1) in 0..k
        const int n_c = nb - k - 1;
        for (int col = tid; col < dp; col += THREADS) {
            const int pc = col >> 5, within = col & 31;
            const bool ident = pc >= n_c;
            const int gcol = ident ? dp + (pc - n_c) * NB + within : (k + 1 + pc) * NB + within;
            double v[NB];
#pragma unroll
            for (int s = 0; s < NB; ++s) v[s] = M[(size_t)(k * NB + s) * ldm + gcol];
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s <= r; ++s) acc = fma(Li[r * LDS_ + s], v[s], acc);
                Ps[(size_t)r * dp + col] = acc;
                if (ident) M[(size_t)(k * NB + r) * ldm + gcol] = acc;   // rows of L^-1: part of the result
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------------------------------ P3
        int item = 0;
        for (int i = k + 1; i < nb; ++i) {
            const int ib = i - k - 1;                                  // panel block holding U_ki
            for (int pc = ib; pc < nb; ++pc, ++item) {                 // C blocks j >= i, then all identity blocks
                if ((item % (THREADS / 32)) != warp) continue;
                double acc[NB];
#pragma unroll
                for (int a = 0; a < NB; ++a) acc[a] = 0.0;
#pragma unroll 4
                for (int r = 0; r < NB; ++r) {
                    const double pb = Ps[(size_t)r * dp + pc * NB + lane];
                    const double2 *pa = reinterpret_cast<const double2 *>(Ps + (size_t)r * dp + ib * NB);
#pragma unroll
                    for (int a2 = 0; a2 < NB / 2; ++a2) {
                        const double2 u = pa[a2];                      // broadcast
                        acc[2 * a2] = fma(u.x, pb, acc[2 * a2]);
                        acc[2 * a2 + 1] = fma(u.y, pb, acc[2 * a2 + 1]);
                    }
                }
                const int gcol = (pc >= n_c ? dp + (pc - n_c) * NB : (k + 1 + pc) * NB) + lane;
                double *dst = M + (size_t)(i * NB) * ldm + gcol;
#pragma unroll
                for (int a = 0; a < NB; ++a) dst[(size_t)a * ldm] -= acc[a];
            }
        }
        __syncthreads();
    }

    // T[i][kk] = (L^-1)[kk][i] as f32 (upper triangular); trace(C^-1) = sum of squares of L^-1
    double tr = 0.0;
    for (int idx = tid; idx < d * d; idx += THREADS) {
        const int i = idx / d, kk = idx - i * d;
        double z = 0.0;
        if (i <= kk) z = M[(size_t)kk * ldm + dp + i];
        tr = fma(z, z, tr);
        T[idx] = (float)z;
    }
    for (int off = 16; off > 0; off >>= 1) tr += __shfl_xor_sync(FULL, tr, off);
    if (lane == 0) s_red[warp] = tr;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) t += s_red[w];
        if (s_bad || !(t <= trace_limit)) status[0] = 1;               // sticky: the caller zeroes it once per loop
    }
}

}  // namespace chol

bool chol_whiten_supported(int64_t d) { return d >= 1 && d <= 512; }

// cov: d x d f64 (symmetric, already scaled by 1/(n-1)); T: d x d f32; status: int[1] (device), raised when the
// covariance is not safely positive definite (lambda_min may be below the reference's 1e-10 clamp).
void launch_chol_whiten(const double *cov, int64_t d, float *T, int *status, cudaStream_t st) {
    using namespace chol;
    if (!chol_whiten_supported(d)) throw CudaFail{"Cholesky whitening supports 1 <= d <= 512"};
    const int dp = (int)((d + NB - 1) / NB * NB);
    double *M = (double *)workspace().chol.get((size_t)dp * 2 * dp * sizeof(double));
    const size_t smem = ((size_t)NB * dp + 2 * NB * LDS_ + NB) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(chol_whiten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // trace(C^-1) <= 1e8  =>  lambda_min >= 1e-8, two orders above the reference's clamp (pycleora/__init__.py:155)
    chol_whiten_kernel<<<1, THREADS, smem, st>>>(cov, (int)d, dp, M, T, status, 1e8);
    LAUNCH_CHECK();
}

}  // namespace cleora
