// gla_oracle.cpp -- CPU ORACLE for the Householder-QR / Cholesky-update hot path.
//
// *** TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library, and only as the checker / reported CPU baseline.
//
// PARITY STATUS: "parity unpinned" for the two Julia-stdlib routines reflector! and the
// left reflectorApply! -- they are NOT in /root/reference (they live in Julia's
// un-vendored, un-pinned stdlib LinearAlgebra: Project.toml:13 `julia = "1.6"`, no
// Manifest), and no Julia runtime exists in this image, so the reference itself cannot
// be executed to generate vectors.  They are restated from their published semantics
// (stdlib/LinearAlgebra/src/generic.jl, Julia 1.9-1.12 form)
// and pinned by the hand-derived known-answer vectors in tests/golden/kat.json and by
// LAPACK cross-checks (row-phase normalised).  Everything that IS in the reference tree
// is restated function by function below with its file:line.
//
// The algorithms are restated from the reference's behaviour; no reference source text
// is copied.  Column-major storage, 0-based indices here vs 1-based in the reference.
//
// Build: g++ -O3 -march=x86-64-v3 -fopenmp -shared -fPIC gla_oracle.cpp -o libgla_oracle.so

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <omp.h>

namespace {

using i64 = int64_t;
using zd = std::complex<double>;

template <class T> struct Num { using real = T; static constexpr bool is_complex = false; };
template <> struct Num<zd> { using real = double; static constexpr bool is_complex = true; };

inline float cj(float x) { return x; }
inline double cj(double x) { return x; }
inline zd cj(zd x) { return std::conj(x); }
inline float re(float x) { return x; }
inline double re(double x) { return x; }
inline double re(zd x) { return x.real(); }
inline float abs2(float x) { return x * x; }
inline double abs2(double x) { return x * x; }
inline double abs2(zd x) { return x.real() * x.real() + x.imag() * x.imag(); }
inline float absmax_part(float x) { return std::fabs(x); }
inline double absmax_part(double x) { return std::fabs(x); }
inline double absmax_part(zd x) { return std::max(std::fabs(x.real()), std::fabs(x.imag())); }

// ---------------------------------------------------------------------------------
// Julia stdlib LinearAlgebra.reflector!(x)          call site: src/qr.jl:96
//   xi = x[1]; nrm = norm(x); nrm == 0 -> tau = 0 (x untouched)
//   nu = copysign(nrm, real(xi)); xi += nu; x[1] = -nu; x[2:] /= xi; tau = xi/nu
// norm(x) is restated as sqrt(sum abs2) (Julia's generic_norm2 fast path / <=1.8 form);
// the scaled variants differ only for over/underflowing data.
// ---------------------------------------------------------------------------------
template <class T>
T reflector(T* x, i64 n) {
  using R = typename Num<T>::real;
  if (n == 0) return T(0);
  T xi = x[0];
  R ss = 0;
  for (i64 i = 0; i < n; ++i) ss += abs2(x[i]);
  R nrm = std::sqrt(ss);
  // Julia >= 1.9 calls norm(x), which rescales when the plain sum of squares leaves the safe range; restated as:
  // outside [tiny, huge] redo the sum on x / 2^e with 2^e ~ max|x_i| (an exact power of two, so in-range data are
  // bitwise unaffected and scaled data give exactly the scaled result).
  const R tiny = sizeof(R) == 8 ? R(1e-280) : R(1e-30), huge = sizeof(R) == 8 ? R(1e280) : R(1e30);
  if (!(ss >= tiny && ss <= huge)) {
    R amax = 0;
    for (i64 i = 0; i < n; ++i) amax = std::max(amax, absmax_part(x[i]));
    if (amax > 0 && std::isfinite(amax)) {
      int e;
      (void)std::frexp(amax, &e);
      const R sc = std::ldexp(R(1), -e);
      R s2 = 0;
      for (i64 i = 0; i < n; ++i) s2 += abs2(x[i] * sc);
      nrm = std::ldexp(std::sqrt(s2), e);
    }
  }
  if (nrm == R(0)) return T(0);
  R nu = std::copysign(nrm, re(xi));
  xi += nu;
  x[0] = T(-nu);
  for (i64 i = 1; i < n; ++i) x[i] /= xi;
  return xi / nu;
}

// ---------------------------------------------------------------------------------
// Julia stdlib left LinearAlgebra.reflectorApply!(x, tau, A)   call site: src/qr.jl:102
//   A <- (I - conj(tau) v v^H) A,  v = [1; x[2:]]
// ---------------------------------------------------------------------------------
template <class T>
void reflector_apply_left(const T* x, T tau, T* A, i64 m, i64 n, i64 lda) {
  for (i64 j = 0; j < n; ++j) {
    T* a = A + j * lda;
    T s = a[0];
    for (i64 i = 1; i < m; ++i) s += cj(x[i]) * a[i];
    s = cj(tau) * s;
    a[0] -= s;
    for (i64 i = 1; i < m; ++i) a[i] -= x[i] * s;
  }
}

// ---------------------------------------------------------------------------------
// right reflectorApply!(A, x, tau)                  src/qr.jl:19-42
//   row by row: s = tau*(A[i,1] + sum_j>=2 A[i,j] x[j]); A[i,1] -= s; A[i,j] -= s*conj(x[j])
// returns -2 (second dimension mismatch) like the DimensionMismatch at :21-27
// ---------------------------------------------------------------------------------
template <class T>
int reflector_apply_right(T* A, i64 m, i64 n, i64 lda, const T* x, i64 lenx, T tau) {
  if (lenx != n) return -2;
  for (i64 i = 0; i < m; ++i) {
    T s = A[i];
    for (i64 j = 1; j < n; ++j) s += A[i + j * lda] * x[j];
    s = s * tau;
    A[i] -= s;
    for (i64 j = 1; j < n; ++j) A[i + j * lda] -= s * cj(x[j]);
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// qrUnblocked!(A, tau)                              src/qr.jl:86-111
// (the reference's recursion on view(A,2:m,2:n) is a column loop)
// ---------------------------------------------------------------------------------
template <class T>
void qr_unblocked(T* A, i64 m, i64 n, i64 lda, T* tau) {
  i64 k = 0;
  while (true) {
    T* a = A + k + k * lda;
    i64 mk = m - k, nk = n - k;
    T t1 = reflector(a, mk);                                  // :94-98
    tau[k] = t1;
    if (nk > 1) reflector_apply_left(a, t1, a + lda, mk, nk - 1, lda);   // :101-103
    if (mk > 1 && nk > 1) { ++k; continue; }                  // :106-108
    break;
  }
}

// ---------------------------------------------------------------------------------
// getindex(::QR2, Tuple{:QBlocked})  -> compact-WY T   src/qr.jl:64-83
//   S[i,j] = tau_i * (conj?(F[j,i]) + dot(F[j+1:m,i], F[j+1:m,j])),  i<j
//   inv!(UnitUpperTriangular(S)); T[j,j] = tau_j; T[i,j] *= tau_j
// literal != 0 reproduces the reference text exactly (F[j,i] without conj, :72), which
// is only correct for real element types (SURVEY.md finding 3).
// ---------------------------------------------------------------------------------
template <class T>
void build_T(const T* F, i64 m, i64 n, i64 ldf, const T* tau, T* Tm, i64 ldt, int literal) {
  i64 k = std::min(m, n);
  std::vector<T> U((size_t)k * k, T(0));
  for (i64 j = 0; j < k; ++j)
    for (i64 i = 0; i < j; ++i) {
      T d = T(0);
      for (i64 l = j + 1; l < m; ++l) d += cj(F[l + i * ldf]) * F[l + j * ldf];
      T fji = literal ? F[j + i * ldf] : cj(F[j + i * ldf]);
      U[i + j * k] = tau[i] * (fji + d);                       // :70-74
    }
  // X = inv(I + U), U strictly upper.  Column j: X[:,j] = e_j - X[:,0:j] * U[0:j,j]   (:75)
  std::vector<T> X((size_t)k * k, T(0));
  for (i64 j = 0; j < k; ++j) {
    X[j + j * k] = T(1);
    for (i64 l = 0; l < j; ++l) {
      T u = U[l + j * k];
      if (u == T(0)) continue;
      for (i64 i = 0; i <= l; ++i) X[i + j * k] -= X[i + l * k] * u;
    }
  }
  for (i64 j = 0; j < k; ++j) {                                // :76-81
    for (i64 i = 0; i < k; ++i) Tm[i + j * ldt] = T(0);
    for (i64 i = 0; i < j; ++i) Tm[i + j * ldt] = X[i + j * k] * tau[j];
    Tm[j + j * ldt] = tau[j];
  }
}

// ---------------------------------------------------------------------------------
// lmul!(H, A, M) / lmul!(H', A, M) for HouseholderBlock   src/householder.jl:82-115,119-157
//   M = V1^H A1 + V2^H A2;  M = T M (or T^H M for the adjoint);  A2 -= V2 M;  A1 -= V1 M
// with V1 = unit-lower top b x b of V, V2 the rows below.  The unit-lower product in the
// last step is the reference's own generic lmul!(UnitLowerTriangular, B, alpha)
// (src/juliaBLAS.jl:154-166), rows bottom-up.  Columns of A are independent -> OpenMP,
// standing in for the multithreaded BLAS the reference's trmm/gemm calls would use.
// returns -1 on row mismatch (DimensionMismatch at :87 / :129)
// ---------------------------------------------------------------------------------
template <class T>
int block_apply(const T* V, i64 mV, i64 nV, i64 ldv, const T* Tm, i64 ldt, T* A, i64 mA, i64 nA,
                i64 lda, int adjoint) {
  if (mV != mA) return -1;
  i64 b = std::min(mV, nV);
#pragma omp parallel
  {
    std::vector<T> Mc(b), M2(b);
#pragma omp for schedule(static)
    for (i64 j = 0; j < nA; ++j) {
      T* a = A + j * lda;
      // M = V1^H * A1 (unit lower, conj-transposed => upper unit):  M[i] = a[i] + sum_{l>i} conj(V[l,i]) a[l]
      for (i64 i = 0; i < b; ++i) {
        T s = a[i];
        const T* v = V + i * ldv;
        for (i64 l = i + 1; l < b; ++l) s += cj(v[l]) * a[l];
        // M += V2^H * A2
        for (i64 l = b; l < mA; ++l) s += cj(v[l]) * a[l];
        Mc[i] = s;
      }
      // M = T M  (upper)  or  T^H M (lower)
      if (!adjoint) {
        for (i64 i = 0; i < b; ++i) {
          T s = T(0);
          for (i64 l = i; l < b; ++l) s += Tm[i + l * ldt] * Mc[l];
          M2[i] = s;
        }
      } else {
        for (i64 i = 0; i < b; ++i) {
          T s = T(0);
          for (i64 l = 0; l <= i; ++l) s += cj(Tm[l + i * ldt]) * Mc[l];
          M2[i] = s;
        }
      }
      // A2 -= V2 M
      for (i64 l = 0; l < b; ++l) {
        const T* v = V + l * ldv;
        T ml = M2[l];
        for (i64 i = b; i < mA; ++i) a[i] -= v[i] * ml;
      }
      // M = -V1 M (unit lower, rows bottom-up), A1 += M
      for (i64 i = b - 1; i >= 0; --i) {
        T s = -M2[i];
        for (i64 l = 0; l < i; ++l) s += -V[i + l * ldv] * M2[l];
        a[i] += s;
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// qrBlocked!(A, blocksize, tau, work)               src/qr.jl:113-146
// ---------------------------------------------------------------------------------
template <class T>
void qr_blocked(T* A, i64 m, i64 n, i64 lda, T* tau, i64 bs, int literal) {
  i64 k = 0;
  std::vector<T> Tm((size_t)bs * bs);
  while (true) {
    T* Ak = A + k + k * lda;
    i64 mk = m - k, nk = n - k;
    i64 nb = std::min(nk, bs);
    qr_unblocked(Ak, mk, nb, lda, tau + k);                   // :123-125
    if (nk > bs) {                                            // :128-133
      build_T(Ak, mk, nb, lda, tau + k, Tm.data(), bs, literal);
      block_apply(Ak, mk, nb, lda, Tm.data(), bs, Ak + bs * lda, mk, nk - bs, lda, 1);
    }
    if (mk > bs && nk > bs) { k += bs; continue; }            // :136-143
    break;
  }
}

// ---------------------------------------------------------------------------------
// rankUpdate!(C::Hermitian, A, alpha::Real) generic rank-k  src/juliaBLAS.jl:89-112, 'L' branch
//   for k, for j, for i>=j: C[i,j] += A[i,k]*alpha*conj(A[j,k])
// (this is the method cholRecursive! actually hits, SURVEY.md finding 4)
// ---------------------------------------------------------------------------------
template <class T>
void rank_update_lower(T* C, i64 n, i64 ldc, const T* A, i64 kk, i64 lda, typename Num<T>::real alpha) {
  for (i64 k = 0; k < kk; ++k)
    for (i64 j = 0; j < n; ++j) {
      T ajc = cj(A[j + k * lda]);
      for (i64 i = j; i < n; ++i) C[i + j * ldc] += A[i + k * lda] * alpha * ajc;
    }
}
// Same sums, j-outermost so columns can go to different threads (bitwise identical result:
// every C[i,j] still accumulates its k terms in increasing k).
template <class T>
void rank_update_lower_mt(T* C, i64 n, i64 ldc, const T* A, i64 kk, i64 lda, typename Num<T>::real alpha) {
#pragma omp parallel for schedule(dynamic, 8)
  for (i64 j = 0; j < n; ++j)
    for (i64 k = 0; k < kk; ++k) {
      T ajc = cj(A[j + k * lda]);
      for (i64 i = j; i < n; ++i) C[i + j * ldc] += A[i + k * lda] * alpha * ajc;
    }
}

// rank-1 Hermitian generic  src/juliaBLAS.jl:53-63 ('L')
template <class T>
void rank1_update_lower(T* C, i64 n, i64 ldc, const T* a, typename Num<T>::real alpha) {
  for (i64 j = 0; j < n; ++j) {
    T ajc = cj(a[j]);
    for (i64 i = j; i < n; ++i) C[i + j * ldc] += a[i] * alpha * ajc;
  }
}

// rdiv!(A21, LowerTriangular(A11)')  (stdlib -> trsm Right/Lower/ConjTrans/NonUnit)  call site src/cholesky.jl:48
//   X <- X * L^{-H}:  column j of X: x_j = (x_j - sum_{l<j} x_l conj(L[j,l])) / conj(L[j,j])
template <class T>
void rdiv_lower_adjoint(T* X, i64 m, i64 n, i64 ldx, const T* L, i64 ldl) {
#pragma omp parallel for schedule(static) if (m * n > 4096)
  for (i64 r0 = 0; r0 < m; r0 += 64) {
    i64 r1 = std::min(m, r0 + 64);
    for (i64 j = 0; j < n; ++j) {
      T* xj = X + j * ldx;
      for (i64 l = 0; l < j; ++l) {
        T c = cj(L[j + l * ldl]);
        const T* xl = X + l * ldx;
        for (i64 i = r0; i < r1; ++i) xj[i] -= xl[i] * c;
      }
      T d = cj(L[j + j * ldl]);
      for (i64 i = r0; i < r1; ++i) xj[i] /= d;
    }
  }
}

// cholUnblocked!(A, Val{:L})                        src/cholesky.jl:3-15
template <class T>
int chol_unblocked(T* A, i64 n, i64 lda, i64 off) {
  for (i64 k = 0; k < n; ++k) {
    T* a = A + k + k * lda;
    if (!(re(a[0]) > 0)) return (int)(off + k + 1);   // sqrt of a negative real throws DomainError
    a[0] = std::sqrt(a[0]);
    if (k + 1 < n) {
      auto inv = typename Num<T>::real(1) / re(a[0]);
      for (i64 i = 1; i < n - k; ++i) a[i] *= inv;
      rank1_update_lower(a + 1 + lda, n - k - 1, lda, a + 1, typename Num<T>::real(-1));
    }
  }
  return 0;
}

// cholBlocked!(A, Val{:L}, blocksize)               src/cholesky.jl:17-35
template <class T>
int chol_blocked(T* A, i64 n, i64 lda, i64 bs) {
  for (i64 k = 0; k < n; k += bs) {
    i64 nb = std::min(n - k, bs);
    T* A11 = A + k + k * lda;
    int info = chol_unblocked(A11, nb, lda, k);
    if (info) return info;
    if (n - k > bs) {
      T* A21 = A11 + bs;
      rdiv_lower_adjoint(A21, n - k - bs, bs, lda, A11, lda);
      rank_update_lower_mt(A21 + bs * lda, n - k - bs, lda, A21, bs, lda, typename Num<T>::real(-1));
    }
  }
  return 0;
}

// cholRecursive!(A, Val{:L}, cutoff)                src/cholesky.jl:37-55
// NB the A11 branch drops `cutoff` (defaults to 1), exactly as :46 does.
template <class T>
int chol_recursive(T* A, i64 n, i64 lda, i64 cutoff, i64 off, int mt) {
  if (n == 1) {
    if (!(re(A[0]) > 0)) return (int)(off + 1);
    A[0] = std::sqrt(A[0]);
    return 0;
  } else if (n < cutoff) {
    return chol_unblocked(A, n, lda, off);
  }
  i64 n2 = n / 2;
  int info = chol_recursive(A, n2, lda, 1, off, mt);
  if (info) return info;
  T* A21 = A + n2;
  rdiv_lower_adjoint(A21, n - n2, n2, lda, A, lda);
  T* A22 = A + n2 + n2 * lda;
  if (mt) rank_update_lower_mt(A22, n - n2, lda, A21, n2, lda, typename Num<T>::real(-1));
  else    rank_update_lower(A22, n - n2, lda, A21, n2, lda, typename Num<T>::real(-1));
  return chol_recursive(A22, n - n2, lda, cutoff, off + n2, mt);
}

// ---------------------------------------------------------------------------------
// LDL^H without pivoting                             src/ldlt.jl
//   _ldlt_lower!  :17-35   delta = A[1,1]; l = a/delta; A_-(lower) -= l*delta*l';  recurse
//   _ldlt_upper!  :50-64   delta = A[1,1]; u^T = a^T/delta; A_-(upper) -= conj(u)*delta*u^T; recurse
//   _ldlt_lower_blocked! :80-103, _ldlt_upper_blocked! :122-146  (rdiv!/ldiv! with the unit triangle, then with D,
//   then the k-innermost rank update with workspace = d .* conj(row))
// Only the named triangle is read or written; the diagonal holds D, the strict triangle the unit factor.
// ---------------------------------------------------------------------------------
template <class T>
void ldlt_lower_unblocked(T* A, i64 n, i64 lda) {
  for (i64 k = 0; k + 1 <= n; ++k) {
    T* a = A + k + k * lda;
    const i64 m = n - k;
    const T delta = a[0];
    for (i64 i = 1; i < m; ++i) a[i] = a[i] / delta;
    for (i64 j = 1; j < m; ++j) {
      const T ljd = a[j] * delta;
      for (i64 i = j; i < m; ++i) a[i + j * lda] -= a[i] * cj(ljd);
    }
    if (m <= 2 && k + 2 >= n) { /* the reference recurses only while n > 2: a trailing 1x1 block is left as is */ }
  }
}
template <class T>
void ldlt_upper_unblocked(T* A, i64 n, i64 lda) {
  for (i64 k = 0; k + 1 <= n; ++k) {
    T* a = A + k + k * lda;
    const i64 m = n - k;
    const T delta = a[0];
    for (i64 j = 1; j < m; ++j) {
      const T dl = a[j * lda];
      a[j * lda] = dl / delta;
      for (i64 i = 1; i <= j; ++i) a[i + j * lda] -= cj(a[i * lda]) * dl;
    }
  }
}
template <class T>
void ldlt_lower_blocked(T* A, i64 n, i64 lda, i64 bs) {
  std::vector<T> ws((size_t)bs);
  for (i64 k = 0; k < n; k += bs) {
    T* A11 = A + k + k * lda;
    const i64 m = n - k;
    if (bs >= m) {
      ldlt_lower_unblocked(A11, m, lda);
      return;
    }
    ldlt_lower_unblocked(A11, bs, lda);
    T* A21 = A11 + bs;
    const i64 r = m - bs;
    // rdiv!(A21, UnitLowerTriangular(A11)'):  x_j = x_j - sum_{l<j} x_l conj(L[j,l])
    for (i64 j = 0; j < bs; ++j)
      for (i64 l = 0; l < j; ++l) {
        const T c = cj(A11[j + l * lda]);
        for (i64 i = 0; i < r; ++i) A21[i + j * lda] -= A21[i + l * lda] * c;
      }
    for (i64 j = 0; j < bs; ++j) {
      const T d = A11[j + j * lda];
      for (i64 i = 0; i < r; ++i) A21[i + j * lda] = A21[i + j * lda] / d;
    }
    T* A22 = A11 + bs + bs * lda;
    for (i64 j = 0; j < r; ++j) {
      for (i64 q = 0; q < bs; ++q) ws[q] = A11[q + q * lda] * cj(A21[j + q * lda]);
      for (i64 i = j; i < r; ++i)
        for (i64 q = 0; q < bs; ++q) A22[i + j * lda] -= A21[i + q * lda] * ws[q];
    }
  }
}
template <class T>
void ldlt_upper_blocked(T* A, i64 n, i64 lda, i64 bs) {
  std::vector<T> ws((size_t)bs);
  for (i64 k = 0; k < n; k += bs) {
    T* A11 = A + k + k * lda;
    const i64 m = n - k;
    if (bs >= m) {
      ldlt_upper_unblocked(A11, m, lda);
      return;
    }
    ldlt_upper_unblocked(A11, bs, lda);
    T* U12 = A11 + bs * lda;
    const i64 r = m - bs;
    // ldiv!(UnitUpperTriangular(A11)', U12):  row i: x_i = x_i - sum_{l<i} conj(U[l,i]) x_l
    for (i64 c = 0; c < r; ++c)
      for (i64 i = 0; i < bs; ++i)
        for (i64 l = 0; l < i; ++l) U12[i + c * lda] -= cj(A11[l + i * lda]) * U12[l + c * lda];
    for (i64 c = 0; c < r; ++c)
      for (i64 i = 0; i < bs; ++i) U12[i + c * lda] = U12[i + c * lda] / A11[i + i * lda];
    T* A22 = A11 + bs + bs * lda;
    for (i64 j = 0; j < r; ++j) {
      for (i64 q = 0; q < bs; ++q) ws[q] = A11[q + q * lda] * U12[q + j * lda];
      for (i64 i = 0; i <= j; ++i)
        for (i64 q = 0; q < bs; ++q) A22[i + j * lda] -= cj(U12[q + i * lda]) * ws[q];
    }
  }
}

}  // namespace

// =================================================================================
// Two-sided Householder reductions.  The apply loops carry an OpenMP `for` over independent columns /
// rows (bitwise the serial result) so that the CPU baseline timed beside the GPU uses the host's cores.
// =================================================================================
template <class T>
void apply_left_mt(const T* x, T tau, T* A, i64 m, i64 n, i64 lda) {   // stdlib reflectorApply!(x, tau, A)
#pragma omp parallel for schedule(static) if (m * n > 16384)
  for (i64 j = 0; j < n; ++j) {
    T* a = A + j * lda;
    T s = a[0];
    for (i64 i = 1; i < m; ++i) s += cj(x[i]) * a[i];
    s = cj(tau) * s;
    a[0] -= s;
    for (i64 i = 1; i < m; ++i) a[i] -= x[i] * s;
  }
}
template <class T>
void apply_right_mt(T* A, i64 m, i64 n, i64 lda, const T* x, T tau) {  // reflectorApply!(A, x, tau), src/qr.jl:19-42
#pragma omp parallel for schedule(static) if (m * n > 16384)
  for (i64 i = 0; i < m; ++i) {
    T s = A[i];
    for (i64 j = 1; j < n; ++j) s += A[i + j * lda] * x[j];
    s = s * tau;
    A[i] -= s;
    for (i64 j = 1; j < n; ++j) A[i + j * lda] -= s * cj(x[j]);
  }
}

// reflector! on a strided view (a row of a column-major matrix), optionally conj!(x) first (src/svd.jl:341-343)
template <class T>
T reflector_strided(T* x, i64 n, i64 inc, bool conj_first, std::vector<T>& tmp) {
  tmp.resize((size_t)n);
  for (i64 i = 0; i < n; ++i) tmp[i] = conj_first ? cj(x[i * inc]) : x[i * inc];
  T tau = reflector(tmp.data(), n);
  for (i64 i = 0; i < n; ++i) x[i * inc] = tmp[i];
  return tau;
}

// ---------------------------------------------------------------------------------
// bidiagonalize!(A)                                  src/svd.jl:328-381
// m >= n: upper bidiagonal, taul has n entries, taur n-1;  m < n: lower bidiagonal, taur has m entries, taul m-1
// ---------------------------------------------------------------------------------
template <class T>
void bidiagonalize(T* A, i64 m, i64 n, i64 lda, T* taul, T* taur) {
  std::vector<T> row;
  if (m >= n) {
    for (i64 i = 0; i < n; ++i) {                                         // :334-346
      T* x = A + i + i * lda;
      T t = reflector(x, m - i);
      taul[i] = t;
      apply_left_mt(x, t, x + lda, m - i, n - i - 1, lda);
      if (i < n - 1) {
        T* r = A + i + (i + 1) * lda;
        T tr = reflector_strided(r, n - i - 1, lda, true, row);
        taur[i] = tr;
        apply_right_mt(A + (i + 1) + (i + 1) * lda, m - i - 1, n - i - 1, lda, row.data(), tr);
      }
    }
  } else {
    for (i64 i = 0; i < m; ++i) {                                         // :358-371
      T* r = A + i + i * lda;
      T tr = reflector_strided(r, n - i, lda, true, row);
      taur[i] = tr;
      apply_right_mt(A + (i + 1) + i * lda, m - i - 1, n - i, lda, row.data(), tr);
      if (i < m - 1) {
        T* x = A + (i + 1) + i * lda;
        T t = reflector(x, m - i - 1);
        taul[i] = t;
        apply_left_mt(x, t, x + lda, m - i - 1, n - i - 1, lda);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// _hessenberg!(A)                                    src/eigenGeneral.jl:18-31
//   lmul!(H', A[i+1:n, i+1:n])  (src/householder.jl:61-79),  rmul!(A[:, i+1:n], H)  (src/householder.jl:43-59)
// ---------------------------------------------------------------------------------
template <class T>
void hessenberg(T* A, i64 n, i64 lda, T* tau) {
  std::vector<T> xw((size_t)(n > 0 ? n : 1));
  for (i64 i = 0; i + 1 < n; ++i) {
    T* xi = A + (i + 1) + i * lda;
    const i64 len = n - i - 1;
    T t = reflector(xi, len);
    tau[i] = t;
    apply_left_mt(xi, t, A + (i + 1) + (i + 1) * lda, len, len, lda);     // va = tau'(A[1,j] + v.Aj); column -= va [1; v]
    // rmul!: x = A1 v + a1;  a1 -= tau x;  A1 += x (-tau) v_j'
    T* a1 = A + (i + 1) * lda;
    T* A1 = a1 + lda;
#pragma omp parallel for schedule(static) if (n * len > 16384)
    for (i64 r = 0; r < n; ++r) {
      T s = T(0);
      for (i64 j = 0; j + 1 < len; ++j) s += A1[r + j * lda] * xi[j + 1];
      s += a1[r];
      xw[r] = s;
      a1[r] -= t * s;
      for (i64 j = 0; j + 1 < len; ++j) A1[r + j * lda] += s * (-t) * cj(xi[j + 1]);
    }
  }
}

// ---------------------------------------------------------------------------------
// symtriLower!(AS, tau, u) / symtriUpper!             src/eigenSelfAdjoint.jl:450-503 / :505-564
// ---------------------------------------------------------------------------------
template <class T>
void symtri_lower(T* A, i64 n, i64 lda, T* tau) {
  using R = typename Num<T>::real;
  for (i64 i = 0; i < n; ++i) A[i + i * lda] = T(re(A[i + i * lda]));     // :458-460
  const i64 steps = n - 2 + (Num<T>::is_complex ? 1 : 0);
  std::vector<T> u((size_t)(n > 0 ? n : 1));
  for (i64 k = 0; k < steps; ++k) {
    T* x = A + (k + 1) + k * lda;
    const i64 L = n - k - 1;
    T t = reflector(x, L);
    tau[k] = t;
    const T tmp = x[0];
    x[0] = T(1);
    T* At = A + (k + 1) + (k + 1) * lda;
    // u = tau * Hermitian(At, :L) * x
#pragma omp parallel for schedule(static) if (L * L > 16384)
    for (i64 i = 0; i < L; ++i) {
      T s = T(0);
      for (i64 j = 0; j < i; ++j) s += At[i + j * lda] * x[j];
      s += T(re(At[i + i * lda])) * x[i];
      for (i64 j = i + 1; j < L; ++j) s += cj(At[j + i * lda]) * x[j];
      u[i] = t * s;
    }
    T dot = T(0);
    for (i64 i = 0; i < L; ++i) dot += cj(x[i]) * u[i];
    const R xi = re(cj(t) * dot);                                          // :479
#pragma omp parallel for schedule(static) if (L * L > 16384)
    for (i64 j = 0; j < L; ++j) {                                          // :483-490
      const T xj = x[j], uj = u[j];
      const T xixj = T(xi) * xj;
      for (i64 i = j; i < L; ++i) At[i + j * lda] += x[i] * cj(xixj) - x[i] * cj(uj) - u[i] * cj(xj);
    }
    x[0] = tmp;
  }
}

template <class T>
void symtri_upper(T* A, i64 n, i64 lda, T* tau) {
  using R = typename Num<T>::real;
  for (i64 i = 0; i < n; ++i) A[i + i * lda] = T(re(A[i + i * lda]));
  const i64 steps = n - 2 + (Num<T>::is_complex ? 1 : 0);
  std::vector<T> u((size_t)(n > 0 ? n : 1)), rev((size_t)(n > 0 ? n : 1));
  for (i64 k = 0; k < steps; ++k) {
    const i64 L = n - k - 1;              // x = A[0:L, L]
    T* x = A + L * lda;
    for (i64 i = 0; i < L; ++i) rev[i] = x[L - 1 - i];                    // the reversed view :524-526
    T t = reflector(rev.data(), L);
    for (i64 i = 0; i < L; ++i) x[L - 1 - i] = rev[i];
    tau[k] = t;
    const T tmp = x[L - 1];
    x[L - 1] = T(1);
    // u = tau * Hermitian(A[0:L, 0:L], :U) * x
#pragma omp parallel for schedule(static) if (L * L > 16384)
    for (i64 i = 0; i < L; ++i) {
      T s = T(0);
      for (i64 j = 0; j < i; ++j) s += cj(A[j + i * lda]) * x[j];
      s += T(re(A[i + i * lda])) * x[i];
      for (i64 j = i + 1; j < L; ++j) s += A[i + j * lda] * x[j];
      u[i] = t * s;
    }
    T dot = T(0);
    for (i64 i = 0; i < L; ++i) dot += cj(x[i]) * u[i];
    const R xi = re(cj(t) * dot);
#pragma omp parallel for schedule(static) if (L * L > 16384)
    for (i64 j = 0; j < L; ++j) {                                          // :545-552
      const T xj = x[j], uj = u[j];
      const T xixj = T(xi) * xj;
      for (i64 i = 0; i <= j; ++i) A[i + j * lda] += x[i] * cj(xixj) - x[i] * cj(uj) - u[i] * cj(xj);
    }
    x[L - 1] = tmp;
  }
}

#define ORACLE_API extern "C" __attribute__((visibility("default")))

#define DEFINE_TYPE(P, T, R)                                                                       \
  ORACLE_API void oracle_##P##reflector(T* x, i64 n, T* tau) { *tau = reflector<T>(x, n); }        \
  ORACLE_API void oracle_##P##reflector_apply_left(const T* x, const T* tau, T* A, i64 m, i64 n,  \
                                                   i64 lda) {                                      \
    reflector_apply_left<T>(x, *tau, A, m, n, lda);                                                \
  }                                                                                                \
  ORACLE_API int oracle_##P##reflector_apply_right(T* A, i64 m, i64 n, i64 lda, const T* x,        \
                                                   i64 lenx, const T* tau) {                       \
    return reflector_apply_right<T>(A, m, n, lda, x, lenx, *tau);                                  \
  }                                                                                                \
  ORACLE_API void oracle_##P##qr_unblocked(T* A, i64 m, i64 n, i64 lda, T* tau) {                  \
    qr_unblocked<T>(A, m, n, lda, tau);                                                            \
  }                                                                                                \
  ORACLE_API void oracle_##P##qr_blocked(T* A, i64 m, i64 n, i64 lda, T* tau, i64 bs,              \
                                         int literal) {                                            \
    qr_blocked<T>(A, m, n, lda, tau, bs, literal);                                                 \
  }                                                                                                \
  ORACLE_API void oracle_##P##build_T(const T* F, i64 m, i64 n, i64 ldf, const T* tau, T* Tm,      \
                                      i64 ldt, int literal) {                                      \
    build_T<T>(F, m, n, ldf, tau, Tm, ldt, literal);                                          \
  }                                                                                                \
  ORACLE_API int oracle_##P##block_apply(const T* V, i64 mV, i64 nV, i64 ldv, const T* Tm,         \
                                         i64 ldt, T* A, i64 mA, i64 nA, i64 lda, int adjoint) {    \
    return block_apply<T>(V, mV, nV, ldv, Tm, ldt, A, mA, nA, lda, adjoint);                       \
  }                                                                                                \
  ORACLE_API void oracle_##P##qr_batched(T* A, i64 m, i64 n, i64 batch, T* tau, i64 bs) {          \
    i64 k = std::min(m, n);                                                                        \
    _Pragma("omp parallel for schedule(static)") for (i64 b = 0; b < batch; ++b)                   \
        qr_blocked<T>(A + b * m * n, m, n, m, tau + b * k, bs, 0);                                 \
  }                                                                                                \
  ORACLE_API void oracle_##P##rank_update_lower(T* C, i64 n, i64 ldc, const T* A, i64 k, i64 lda,  \
                                                R alpha, int mt) {                                 \
    if (mt) rank_update_lower_mt<T>(C, n, ldc, A, k, lda, alpha);                                  \
    else rank_update_lower<T>(C, n, ldc, A, k, lda, alpha);                                        \
  }                                                                                                \
  ORACLE_API void oracle_##P##rdiv_lower_adjoint(T* X, i64 m, i64 n, i64 ldx, const T* L,          \
                                                 i64 ldl) {                                        \
    rdiv_lower_adjoint<T>(X, m, n, ldx, L, ldl);                                                   \
  }                                                                                                \
  ORACLE_API int oracle_##P##chol_unblocked(T* A, i64 n, i64 lda) {                                \
    return chol_unblocked<T>(A, n, lda, 0);                                                        \
  }                                                                                                \
  ORACLE_API int oracle_##P##chol_blocked(T* A, i64 n, i64 lda, i64 bs) {                          \
    return chol_blocked<T>(A, n, lda, bs);                                                         \
  }                                                                                                \
  ORACLE_API int oracle_##P##chol_recursive(T* A, i64 n, i64 lda, i64 cutoff, int mt) {            \
    return chol_recursive<T>(A, n, lda, cutoff, 0, mt);                                            \
  }                                                                                                \
  ORACLE_API void oracle_##P##bidiagonalize(T* A, i64 m, i64 n, i64 lda, T* taul, T* taur) {      \
    bidiagonalize<T>(A, m, n, lda, taul, taur);                                                    \
  }                                                                                                \
  ORACLE_API void oracle_##P##hessenberg(T* A, i64 n, i64 lda, T* tau) { hessenberg<T>(A, n, lda, tau); } \
  ORACLE_API void oracle_##P##symtri(T* A, i64 n, i64 lda, T* tau, int upper) {                    \
    if (upper) symtri_upper<T>(A, n, lda, tau);                                                    \
    else symtri_lower<T>(A, n, lda, tau);                                                          \
  }                                                                                                \
  ORACLE_API void oracle_##P##ldlt(T* A, i64 n, i64 lda, i64 bs, int upper) {                      \
    if (upper) ldlt_upper_blocked<T>(A, n, lda, bs);                                               \
    else ldlt_lower_blocked<T>(A, n, lda, bs);                                                     \
  }

DEFINE_TYPE(s, float, float)
DEFINE_TYPE(d, double, double)
DEFINE_TYPE(z, zd, double)

ORACLE_API int oracle_version() { return 3; }
// OpenMP team size actually in effect (bench.py reports it; torchrun exports OMP_NUM_THREADS=1, which bench.py overrides)
ORACLE_API int oracle_max_threads() { return omp_get_max_threads(); }
ORACLE_API void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
