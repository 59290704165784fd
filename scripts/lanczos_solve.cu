// RUN IN ROUND 2 (profiles/r02_experiments_pair_lanczos.txt): correct to 4e-14, but SLOWER than the Jacobi kernel it was meant to
// replace (4096 matrices of 128 x 128: 318 ms against 124 ms; the QL stage is a 16M-clock sequential chain per matrix).  Not
// adopted: the product's clamped solve became tridiagonalisation + Sturm multi-section + inverse iteration instead
// (pycmf_b200/csrc/tridiag_solve.cuh).  Kept as the record of the experiment.
//
// Experiment harness (diagnostics, NOT product code): the eigenvalue-clamped Newton solve
//   x = S(H) g,  S(H) = Q diag(1 / max(|lambda|, p)) Q^T        (reference _safe_invert, cmf_solvers.py:346-356)
// WITHOUT an eigendecomposition of H (NumPy prototype: scripts/lanczos_clamped_solve.py):
//   stage 1  k Lanczos steps on (H, g) with full reorthogonalisation (classical Gram-Schmidt twice):  H Q = Q T,  Q^T g = |g| e_1
//   stage 2  f(T) e_1 for the tridiagonal T by implicit QL that carries only the FIRST ROW of the eigenvector matrix and records
//            its Givens rotations, replayed in reverse on z = f(theta) * s1  (no eigenvectors formed)
//   x = |g| Q f(T) e_1
//
// The numerical core (stage 2 and a sequential version of stage 1) is __host__ __device__ and is checked ON THE HOST by this
// program against a plain Jacobi eigendecomposition -- that part runs without a GPU and passed when this file was written.
// The CUDA kernel (one CTA per matrix for stage 1, thread 0 for stage 2 in this first version) was written when the round's GPU
// budget was spent; with a GPU present the program compares it with the product's Jacobi solve
// (pycmf_safe_solve from pycmf_b200/libpycmf_b200.so, loaded with dlopen) and times both.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o scripts/lanczos_solve.bin scripts/lanczos_solve.cu -ldl
//   scripts/lanczos_solve.bin [k] [batch]
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

struct Rot { double c, s; };      // rotation of columns (i, i + 1); i is kept in a parallel int array

// ---- stage 2: f(T) e_1 -----------------------------------------------------------------------------------------------
// Implicit QL (EISPACK tql2 / tqli) on d[0..n), e[0..n) (e[i] couples i and i + 1, e[n-1] unused).  On exit d = eigenvalues,
// row0 = first row of the eigenvector matrix S = G_1 ... G_N; the rotations are appended to (rot_i, rot).  Returns N, or -1
// when max_rots is too small or an eigenvalue needs more than 60 sweeps.
__host__ __device__ inline int ql_first_row(int n, double* d, double* e, double* row0, int* rot_i, Rot* rot, int max_rots) {
    const double eps = 2.220446049250313e-16;
    int nrot = 0;
    for (int i = 0; i < n; i++) row0[i] = i == 0 ? 1.0 : 0.0;
    e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        for (int sweep = 0;; sweep++) {
            int m = l;
            for (; m < n - 1; m++)
                if (fabs(e[m]) <= eps * (fabs(d[m]) + fabs(d[m + 1]))) break;
            if (m == l) break;
            if (sweep == 60) return -1;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = hypot(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? r : -r));
            double sn = 1.0, cs = 1.0, pp = 0.0;
            bool broke = false;
            for (int i = m - 1; i >= l; i--) {
                const double ff = sn * e[i], b = cs * e[i];
                r = hypot(ff, g);
                e[i + 1] = r;
                if (r == 0.0) { d[i + 1] -= pp; e[m] = 0.0; broke = true; break; }
                sn = ff / r; cs = g / r;
                g = d[i + 1] - pp;
                r = (d[i] - g) * sn + 2.0 * cs * b;
                pp = sn * r;
                d[i + 1] = g + pp;
                g = cs * r - b;
                const double a0 = row0[i], a1 = row0[i + 1];
                row0[i + 1] = sn * a0 + cs * a1;
                row0[i] = cs * a0 - sn * a1;
                if (nrot >= max_rots) return -1;
                rot_i[nrot] = i; rot[nrot].c = cs; rot[nrot].s = sn; nrot++;
            }
            if (!broke) { d[l] -= pp; e[l] = g; e[m] = 0.0; }
        }
    }
    return nrot;
}

// z <- S z = G_1 (G_2 (... (G_N z)))
__host__ __device__ inline void replay_rotations(int nrot, const int* rot_i, const Rot* rot, double* z) {
    for (int t = nrot - 1; t >= 0; t--) {
        const int i = rot_i[t];
        const double a0 = z[i], a1 = z[i + 1];
        z[i] = rot[t].c * a0 + rot[t].s * a1;
        z[i + 1] = -rot[t].s * a0 + rot[t].c * a1;
    }
}

// y = f(T) e_1 with f(t) = 1 / max(|t|, p);  diag / off are destroyed, y has n entries.  Returns the rotation count or -1.
__host__ __device__ inline int clamped_inverse_e1(int n, double* diag, double* off, double p, double* y, int* rot_i, Rot* rot,
                                                   int max_rots) {
    const int nrot = ql_first_row(n, diag, off, y, rot_i, rot, max_rots);
    if (nrot < 0) return nrot;
    for (int i = 0; i < n; i++) y[i] = y[i] / fmax(fabs(diag[i]), p);
    replay_rotations(nrot, rot_i, rot, y);
    return nrot;
}

// ---- sequential version of the whole solve (host check of the algorithm as the kernel implements it) -------------------
static int lanczos_solve_host(int k, const double* H, const double* g, double p, double* x) {
    std::vector<double> Q(size_t(k) * k), alpha(k), beta(k), w(k), h(k), y(k);
    std::vector<int> ri(size_t(2) * k * k + 64);
    std::vector<Rot> rr(size_t(2) * k * k + 64);
    double nrm = 0.0, scale = p;
    for (int i = 0; i < k; i++) nrm += g[i] * g[i];
    nrm = sqrt(nrm);
    for (int i = 0; i < k; i++) x[i] = 0.0;
    if (nrm == 0.0) return 0;
    for (int i = 0; i < k; i++) {
        double s = 0.0;
        for (int j = 0; j < k; j++) s += fabs(H[size_t(i) * k + j]);
        scale = std::max(scale, s);
    }
    for (int i = 0; i < k; i++) Q[i] = g[i] / nrm;
    int m = 0;
    for (int j = 0; j < k; j++) {
        const double* q = &Q[size_t(j) * k];
        m = j + 1;
        for (int i = 0; i < k; i++) {
            double s = 0.0;
            for (int c = 0; c < k; c++) s += H[size_t(c) * k + i] * q[c];
            w[i] = s;
        }
        double a = 0.0;
        for (int i = 0; i < k; i++) a += q[i] * w[i];
        alpha[j] = a;
        for (int i = 0; i < k; i++) w[i] -= a * q[i] + (j > 0 ? beta[j - 1] * Q[size_t(j - 1) * k + i] : 0.0);
        for (int pass = 0; pass < 2; pass++) {
            for (int t = 0; t < m; t++) {
                double s = 0.0;
                for (int i = 0; i < k; i++) s += Q[size_t(t) * k + i] * w[i];
                h[t] = s;
            }
            for (int i = 0; i < k; i++) {
                double s = 0.0;
                for (int t = 0; t < m; t++) s += Q[size_t(t) * k + i] * h[t];
                w[i] -= s;
            }
        }
        double b = 0.0;
        for (int i = 0; i < k; i++) b += w[i] * w[i];
        b = sqrt(b);
        if (b <= 1e-14 * scale || j == k - 1) break;
        beta[j] = b;
        for (int i = 0; i < k; i++) Q[size_t(j + 1) * k + i] = w[i] / b;
    }
    const int nrot = clamped_inverse_e1(m, alpha.data(), beta.data(), p, y.data(), ri.data(), rr.data(), int(ri.size()));
    if (nrot < 0) return nrot;
    for (int i = 0; i < k; i++) {
        double s = 0.0;
        for (int t = 0; t < m; t++) s += Q[size_t(t) * k + i] * y[t];
        x[i] = nrm * s;
    }
    return nrot;
}

// reference on the host: cyclic two-sided Jacobi eigendecomposition, x = V diag(1 / max(|l|, p)) V^T g
static void jacobi_solve_host(int k, const double* H, const double* g, double p, double* x) {
    std::vector<double> A(H, H + size_t(k) * k), V(size_t(k) * k, 0.0);
    for (int i = 0; i < k; i++) V[size_t(i) * k + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double offn = 0.0;
        for (int i = 0; i < k; i++) for (int j = i + 1; j < k; j++) offn += A[size_t(i) * k + j] * A[size_t(i) * k + j];
        if (offn < 1e-300) break;
        bool any = false;
        for (int pI = 0; pI < k - 1; pI++) for (int q = pI + 1; q < k; q++) {
            const double apq = A[size_t(pI) * k + q];
            if (fabs(apq) <= 1e-18 * sqrt(fabs(A[size_t(pI) * k + pI] * A[size_t(q) * k + q])) + 1e-300) continue;
            any = true;
            const double th = (A[size_t(q) * k + q] - A[size_t(pI) * k + pI]) / (2.0 * apq);
            const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int r = 0; r < k; r++) {
                const double arp = A[size_t(r) * k + pI], arq = A[size_t(r) * k + q];
                A[size_t(r) * k + pI] = c * arp - s * arq; A[size_t(r) * k + q] = s * arp + c * arq;
            }
            for (int r = 0; r < k; r++) {
                const double apr = A[size_t(pI) * k + r], aqr = A[size_t(q) * k + r];
                A[size_t(pI) * k + r] = c * apr - s * aqr; A[size_t(q) * k + r] = s * apr + c * aqr;
            }
            for (int r = 0; r < k; r++) {
                const double vrp = V[size_t(r) * k + pI], vrq = V[size_t(r) * k + q];
                V[size_t(r) * k + pI] = c * vrp - s * vrq; V[size_t(r) * k + q] = s * vrp + c * vrq;
            }
        }
        if (!any) break;
    }
    for (int i = 0; i < k; i++) x[i] = 0.0;
    for (int j = 0; j < k; j++) {
        double dot = 0.0;
        for (int i = 0; i < k; i++) dot += V[size_t(i) * k + j] * g[i];
        const double f = dot / std::max(fabs(A[size_t(j) * k + j]), p);
        for (int i = 0; i < k; i++) x[i] += f * V[size_t(i) * k + j];
    }
}

// ---- the kernel: one CTA per matrix ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum, result broadcast to every thread (red: >= 33 doubles of shared memory)
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double r = lane < nw ? red[lane] : 0.0;
        r = warp_sum_d(r);
        if (lane == 0) red[32] = r;
    }
    __syncthreads();
    return red[32];
}

// H: batch x k x k (float, symmetric), g, x: batch x k (double).  Shared memory: H as float (k * k), Q as double (k * k, row j =
// Lanczos vector j), alpha / beta / w / h / y (5 k doubles), red (40 doubles).  rot scratch: per CTA max_rots entries in global memory.
// clocks[2 b], clocks[2 b + 1]: cycles of stage 1 and stage 2 of matrix b.
__global__ void __launch_bounds__(256)
lanczos_clamped_solve_kernel(int batch, int k, const float* __restrict__ H, const double* __restrict__ g, double* __restrict__ x,
                             double p, int* __restrict__ rot_i_all, Rot* __restrict__ rot_all, int max_rots,
                             long long* __restrict__ clocks, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Q = reinterpret_cast<double*>(smem_raw);
    double* alpha = Q + size_t(k) * k;
    double* beta = alpha + k;
    double* w = beta + k;
    double* h = w + k;
    double* y = h + k;
    double* red = y + k;
    float* Hs = reinterpret_cast<float*>(red + 40);
    __shared__ int m_sh, nrot_sh;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    int* rot_i = rot_i_all + size_t(blockIdx.x) * max_rots;
    Rot* rot = rot_all + size_t(blockIdx.x) * max_rots;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const long long t0 = clock64();
        __syncthreads();
        for (int e = tid; e < k * k; e += nt) Hs[e] = H[size_t(b) * k * k + e];
        double part = 0.0, rowsum = 0.0;
        for (int i = tid; i < k; i += nt) { const double gi = g[size_t(b) * k + i]; part += gi * gi; }
        const double nrm = sqrt(block_sum_d(part, red));
        for (int i = tid; i < k; i += nt) {                       // ||H||_inf bound for the breakdown test
            double s = 0.0;
            for (int c = 0; c < k; c++) s += fabs(double(Hs[c * k + i]));
            rowsum = fmax(rowsum, s);
        }
        {   // block-wide max of the row sums (and p)
            double mx = rowsum;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            __syncthreads();
            if (lane == 0) red[warp] = mx;
            __syncthreads();
            if (tid == 0) { double s = p; for (int ww = 0; ww < nw; ww++) s = fmax(s, red[ww]); red[33] = s; }
            __syncthreads();
        }
        const double scale = red[33];
        if (nrm == 0.0) {
            for (int i = tid; i < k; i += nt) x[size_t(b) * k + i] = 0.0;
            continue;
        }
        for (int i = tid; i < k; i += nt) Q[i] = g[size_t(b) * k + i] / nrm;
        __syncthreads();
        int m = 0;
        for (int j = 0; j < k; j++) {
            const double* q = Q + size_t(j) * k;
            m = j + 1;
            // w = H q (H symmetric: column access H[c][i] is conflict-free), alpha = q . w
            double a_part = 0.0;
            for (int i = tid; i < k; i += nt) {
                double s = 0.0;
                for (int c = 0; c < k; c++) s = fma(double(Hs[c * k + i]), q[c], s);
                w[i] = s;
                a_part = fma(q[i], s, a_part);
            }
            const double a = block_sum_d(a_part, red);
            if (tid == 0) alpha[j] = a;
            const double bprev = j > 0 ? beta[j - 1] : 0.0;
            for (int i = tid; i < k; i += nt) w[i] -= a * q[i] + (j > 0 ? bprev * Q[size_t(j - 1) * k + i] : 0.0);
            __syncthreads();
            for (int pass = 0; pass < 2; pass++) {
                // h = Q^T w : one warp per Lanczos vector, lanes over the entries
                for (int t = warp; t < m; t += nw) {
                    double s = 0.0;
                    for (int i = lane; i < k; i += 32) s = fma(Q[size_t(t) * k + i], w[i], s);
                    s = warp_sum_d(s);
                    if (lane == 0) h[t] = s;
                }
                __syncthreads();
                // w -= Q h
                for (int i = tid; i < k; i += nt) {
                    double s = 0.0;
                    for (int t = 0; t < m; t++) s = fma(Q[size_t(t) * k + i], h[t], s);
                    w[i] -= s;
                }
                __syncthreads();
            }
            double b_part = 0.0;
            for (int i = tid; i < k; i += nt) b_part = fma(w[i], w[i], b_part);
            const double bnorm = sqrt(block_sum_d(b_part, red));
            if (bnorm <= 1e-14 * scale || j == k - 1) break;          // uniform: every thread holds the same bnorm
            if (tid == 0) beta[j] = bnorm;
            for (int i = tid; i < k; i += nt) Q[size_t(j + 1) * k + i] = w[i] / bnorm;
            __syncthreads();
        }
        __syncthreads();
        const long long t1 = clock64();
        if (tid == 0) {
            m_sh = m;
            nrot_sh = clamped_inverse_e1(m, alpha, beta, p, y, rot_i, rot, max_rots);
        }
        __syncthreads();
        const long long t2 = clock64();
        const int mm = m_sh;
        if (nrot_sh < 0) {
            if (tid == 0) atomicAdd(status, 1);
            for (int i = tid; i < k; i += nt) x[size_t(b) * k + i] = nan("");
        } else {
            for (int i = tid; i < k; i += nt) {
                double s = 0.0;
                for (int t = 0; t < mm; t++) s = fma(Q[size_t(t) * k + i], y[t], s);
                x[size_t(b) * k + i] = nrm * s;
            }
        }
        if (tid == 0 && clocks != nullptr) { clocks[2 * b] = t1 - t0; clocks[2 * b + 1] = t2 - t1; }
    }
}

// ---- test matrices (the classes of scripts/lanczos_clamped_solve.py) ---------------------------------------------------------
static double urand(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return double(s >> 11) / 9007199254740992.0; }
static double nrand(uint64_t& s) { const double u = std::max(urand(s), 1e-300), v = urand(s); return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }

static void make_matrix(int cls, int k, uint64_t& s, std::vector<double>& H) {
    H.assign(size_t(k) * k, 0.0);
    auto gram = [&](int r, double scale, double shift, bool nonneg, bool weights) {
        std::vector<double> A(size_t(r) * k), wts(r, 1.0);
        for (auto& a : A) { a = nrand(s); if (nonneg) a = 0.3 * fabs(a); }
        if (weights) for (auto& wv : wts) wv = 0.25 * exp(-fabs(3.0 * nrand(s)) * 3.0);
        for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) {
            double sum = 0.0;
            for (int t = 0; t < r; t++) sum += wts[t] * A[size_t(t) * k + i] * A[size_t(t) * k + j];
            H[size_t(i) * k + j] = scale * sum + (i == j ? shift : 0.0);
        }
    };
    switch (cls) {
        case 0: gram(3 * k, 1.0 / (3 * k), 0.5, false, false); break;           // well conditioned
        case 1: gram(std::max(1, k / 3), 1.0, 0.0, false, false); break;        // rank k / 3
        case 2: gram(400, 0.5, 0.0, true, true); break;                         // saturated logit Hessian
        case 3: gram(400, 0.5, 0.1, true, true); break;                         // same + 0.1 I
        case 4:                                                                 // indefinite symmetric
            for (int i = 0; i < k; i++) for (int j = 0; j <= i; j++) { const double v = nrand(s) / 2.0; H[size_t(i) * k + j] = v; H[size_t(j) * k + i] = v; }
            break;
        default: break;                                                         // zero matrix
    }
}

typedef int (*create_fn)(int, void**);
typedef int (*solve_fn)(void*, int64_t, int64_t, const double*, int64_t, const double*, double*, double);
typedef const char* (*err_fn)(void);

int main(int argc, char** argv) {
    const int k = argc > 1 ? atoi(argv[1]) : 128, batch = argc > 2 ? atoi(argv[2]) : 4096;
    const double p = 0.2;
    uint64_t seed = 12345;
    // ---- host check of the algorithm (no GPU needed)
    {
        double worst = 0.0;
        long long rots = 0;
        for (int kk : {7, 32, 64, k}) {
            for (int cls = 0; cls < 6; cls++) {
                std::vector<double> H, g(kk), x(kk), ref(kk);
                make_matrix(cls, kk, seed, H);
                for (auto& v : g) v = nrand(seed);
                const int nrot = lanczos_solve_host(kk, H.data(), g.data(), p, x.data());
                if (nrot < 0) { printf("host: QL failed (k = %d, class %d)\n", kk, cls); return 1; }
                rots = std::max<long long>(rots, nrot);
                jacobi_solve_host(kk, H.data(), g.data(), p, ref.data());
                double num = 0.0, den = 0.0;
                for (int i = 0; i < kk; i++) { num += (x[i] - ref[i]) * (x[i] - ref[i]); den += ref[i] * ref[i]; }
                worst = std::max(worst, sqrt(num / std::max(den, 1e-300)));
            }
        }
        printf("host check: Lanczos + QL-first-row vs Jacobi eigendecomposition, k in {7, 32, 64, %d}, 6 matrix classes: max rel err %.2e, "
               "max rotations %lld\n", k, worst, rots);
        if (!(worst < 1e-10)) { printf("HOST CHECK FAILED\n"); return 1; }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no GPU: kernel not run\n"); return 0; }
    // ---- GPU: kernel vs the product's Jacobi solve
    std::vector<float> Hf(size_t(batch) * k * k);
    std::vector<double> Hd(size_t(batch) * k * k), g(size_t(batch) * k);
    for (int b = 0; b < batch; b++) {
        std::vector<double> H;
        make_matrix(b % 6, k, seed, H);
        for (size_t e = 0; e < H.size(); e++) { Hf[size_t(b) * k * k + e] = float(H[e]); Hd[size_t(b) * k * k + e] = double(float(H[e])); }
        for (int i = 0; i < k; i++) g[size_t(b) * k + i] = nrand(seed);
    }
    float* dH; double *dHd, *dg, *dx, *dref; int *drot_i, *dstatus; Rot* drot; long long* dclk;
    const int grid = 148, max_rots = 2 * k * k + 64;
    cudaMalloc(&dH, Hf.size() * 4); cudaMalloc(&dHd, Hd.size() * 8); cudaMalloc(&dg, g.size() * 8);
    cudaMalloc(&dx, g.size() * 8); cudaMalloc(&dref, g.size() * 8);
    cudaMalloc(&drot_i, size_t(grid) * max_rots * 4); cudaMalloc(&drot, size_t(grid) * max_rots * sizeof(Rot));
    cudaMalloc(&dclk, size_t(batch) * 16); cudaMalloc(&dstatus, 4); cudaMemset(dstatus, 0, 4);
    cudaMemcpy(dH, Hf.data(), Hf.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dHd, Hd.data(), Hd.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dg, g.data(), g.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(double) * (size_t(k) * k + 5 * k + 40) + sizeof(float) * size_t(k) * k;
    cudaFuncSetAttribute(lanczos_clamped_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_l = 0.f, ms_j = 0.f;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        lanczos_clamped_solve_kernel<<<grid, 256, smem>>>(batch, k, dH, dg, dx, p, drot_i, drot, max_rots, dclk, dstatus);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaEventElapsedTime(&ms_l, e0, e1);
    }
    void* lib = dlopen("pycmf_b200/libpycmf_b200.so", RTLD_NOW);
    if (lib == nullptr) { printf("cannot load pycmf_b200/libpycmf_b200.so (%s): no reference on the device\n", dlerror()); return 1; }
    create_fn create = (create_fn)dlsym(lib, "pycmf_create");
    solve_fn solve = (solve_fn)dlsym(lib, "pycmf_safe_solve");
    err_fn lasterr = (err_fn)dlsym(lib, "pycmf_last_error");
    void* ctx = nullptr;
    if (create(0, &ctx) != 0) { printf("pycmf_create: %s\n", lasterr()); return 1; }
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (solve(ctx, batch, k, dHd, int64_t(k) * k, dg, dref, p) != 0) { printf("pycmf_safe_solve: %s\n", lasterr()); return 1; }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_j, e0, e1);
    }
    std::vector<double> x(g.size()), ref(g.size());
    std::vector<long long> clk(size_t(batch) * 2);
    int status = 0;
    cudaMemcpy(x.data(), dx, x.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref.data(), dref, ref.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(clk.data(), dclk, clk.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&status, dstatus, 4, cudaMemcpyDeviceToHost);
    double worst = 0.0, c1 = 0.0, c2 = 0.0;
    for (int b = 0; b < batch; b++) {
        double num = 0.0, den = 0.0;
        for (int i = 0; i < k; i++) { const double e = x[size_t(b) * k + i] - ref[size_t(b) * k + i]; num += e * e; den += ref[size_t(b) * k + i] * ref[size_t(b) * k + i]; }
        worst = std::max(worst, sqrt(num / std::max(den, 1e-300)));
        c1 += double(clk[2 * b]); c2 += double(clk[2 * b + 1]);
    }
    printf("GPU: k = %d, batch = %d: Lanczos kernel %.2f ms (stage 1 %.0f clk, stage 2 %.0f clk per matrix; QL failures %d), "
           "product Jacobi solve %.2f ms, max rel difference %.2e\n", k, batch, ms_l, c1 / batch, c2 / batch, status, ms_j, worst);
    return worst < 1e-8 ? 0 : 1;
}
