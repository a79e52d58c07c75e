/*
 * ORACLE - TEST INFRASTRUCTURE ONLY.  Included twice by fus_oracle.c with
 *   T   = float / double
 *   SFX = f32   / f64
 *
 * A plain-C restatement, loop for loop, of the reference's CPU algorithm for
 * the hot path.  Every function names the reference lines it follows
 * (paths relative to /root/reference).  Nothing in the product
 * (fenicsx_fus_gpu_b200/) may include, link or call this file.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* numba-cpu/sum_factorisation.py:17-48 (cpp/common/sum_factorisation.hpp:43-49) */
static void FN(transpose3)(int Na, int Nb, int Nc, int offa, int offb, int offc,
                           const T* A, T* B) {
  for (int a = 0; a < Na; ++a)
    for (int b = 0; b < Nb; ++b)
      for (int c = 0; c < Nc; ++c)
        B[offa * a + offb * b + offc * c] = A[a * Nb * Nc + b * Nc + c];
}

/* numba-cpu/sum_factorisation.py:51-95 (cpp/common/sum_factorisation.hpp:70-86)
 * tr != 0: C[a,d] += A[a*Nk+k] * B[k,d];  tr == 0: C[a,d] += A[k*Na+a] * B[k,d] */
static void FN(contract3)(int Nk, int Na, int Nb, int Nc, int tr, const T* A,
                          const T* B, T* C) {
  const int Nd = Nb * Nc;
  if (tr) {
    for (int k = 0; k < Nk; ++k)
      for (int a = 0; a < Na; ++a)
        for (int d = 0; d < Nd; ++d) C[a * Nd + d] += A[a * Nk + k] * B[k * Nd + d];
  } else {
    for (int k = 0; k < Nk; ++k)
      for (int a = 0; a < Na; ++a)
        for (int d = 0; d < Nd; ++d) C[a * Nd + d] += A[k * Na + a] * B[k * Nd + d];
  }
}

/* numba-cpu/operators.py:19-68 (cpp/common/spectral_op.hpp:69-86);
 * GPU twin cuda/operators.py:18-70.  y += x[dofmap] * detJ * c, per entity. */
void FN(orc_mass)(const T* x, const T* coeff, T* y, const T* detJ,
                  const int32_t* dofmap, int64_t nent, int ncols) {
  T* x_ = (T*)malloc(sizeof(T) * (size_t)ncols);
  for (int64_t e = 0; e < nent; ++e) {
    const int32_t* dm = dofmap + e * ncols;
    const T* dj = detJ + e * ncols;
    for (int i = 0; i < ncols; ++i) x_[i] = x[dm[i]];
    for (int i = 0; i < ncols; ++i) x_[i] *= dj[i] * coeff[e];
    for (int i = 0; i < ncols; ++i) y[dm[i]] += x_[i];
  }
  free(x_);
}

/* numba-cpu/operators.py:71-227 (cpp/common/spectral_op.hpp:173-243);
 * GPU twin cuda/operators.py:73-192.
 * Cells [c0, c1).  dphi is (n, n) row-major, dphi[q*n + i] = l_i'(x_q). */
void FN(orc_stiffness_range)(const T* x, const T* coeff, T* y, const T* G,
                             const int32_t* dofmap, const T* dphi, int64_t c0,
                             int64_t c1, int P) {
  const int n = P + 1, N = n * n * n;
  T* buf = (T*)calloc((size_t)N * 11, sizeof(T));
  T *x_ = buf, *T1 = buf + N, *T2 = buf + 2 * N, *T3 = buf + 3 * N, *T4 = buf + 4 * N;
  T *fw0 = buf + 5 * N, *fw1 = buf + 6 * N, *fw2 = buf + 7 * N;
  T *y0 = buf + 8 * N, *y1 = buf + 9 * N, *y2 = buf + 10 * N;
  const size_t bytes = sizeof(T) * (size_t)N;

  for (int64_t c = c0; c < c1; ++c) {
    const int32_t* dm = dofmap + c * N;
    memset(T1, 0, bytes); memset(T2, 0, bytes); memset(T3, 0, bytes); memset(T4, 0, bytes);
    memset(fw0, 0, bytes); memset(fw1, 0, bytes); memset(fw2, 0, bytes);

    for (int i = 0; i < N; ++i) x_[i] = x[dm[i]];                 /* :156-157 */

    FN(contract3)(n, n, n, n, 1, dphi, x_, fw0);                  /* x-dir :160-162 */

    FN(transpose3)(n, n, n, n, n * n, 1, x_, T1);                 /* y-dir :165-169 */
    FN(contract3)(n, n, n, n, 1, dphi, T1, T2);
    FN(transpose3)(n, n, n, n, n * n, 1, T2, fw1);

    FN(transpose3)(n, n, n, 1, n, n * n, x_, T3);                 /* z-dir :172-176 */
    FN(contract3)(n, n, n, n, 1, dphi, T3, T4);
    FN(transpose3)(n, n, n, 1, n, n * n, T4, fw2);

    {                                                              /* transform :91-115 */
      const T* Gc = G + c * (int64_t)N * 6;
      const T cc = coeff[c];
      for (int q = 0; q < N; ++q) {
        const T* G_ = Gc + 6 * q;
        const T w0 = fw0[q], w1 = fw1[q], w2 = fw2[q];
        fw0[q] = cc * (G_[0] * w0 + G_[1] * w1 + G_[2] * w2);
        fw1[q] = cc * (G_[1] * w0 + G_[3] * w1 + G_[4] * w2);
        fw2[q] = cc * (G_[2] * w0 + G_[4] * w1 + G_[5] * w2);
      }
    }

    memset(T1, 0, bytes); memset(T2, 0, bytes); memset(T3, 0, bytes); memset(T4, 0, bytes);
    memset(y0, 0, bytes); memset(y1, 0, bytes); memset(y2, 0, bytes);

    FN(contract3)(n, n, n, n, 0, dphi, fw0, y0);                  /* :192-194 */

    FN(transpose3)(n, n, n, n, n * n, 1, fw1, T1);                /* :197-201 */
    FN(contract3)(n, n, n, n, 0, dphi, T1, T2);
    FN(transpose3)(n, n, n, n, n * n, 1, T2, y1);

    FN(transpose3)(n, n, n, 1, n, n * n, fw2, T3);                /* :204-208 */
    FN(contract3)(n, n, n, n, 0, dphi, T3, T4);
    FN(transpose3)(n, n, n, 1, n, n * n, T4, y2);

    for (int i = 0; i < N; ++i) y[dm[i]] += y0[i] + y1[i] + y2[i]; /* :211-212 */
  }
  free(buf);
}

void FN(orc_stiffness)(const T* x, const T* coeff, T* y, const T* G,
                       const int32_t* dofmap, const T* dphi, int64_t ncells, int P) {
  FN(orc_stiffness_range)(x, coeff, y, G, dofmap, dphi, 0, ncells, P);
}

/* "mpirun -n k" emulation for the CPU baseline: the reference kernels are
 * serial per rank (no prange) and scale through MPI ranks only; here k
 * threads each own a contiguous 1/k of the cells and a private output
 * vector y + t*ystride (no halo cost).  Not a reference function. */
void FN(orc_stiffness_ranks)(const T* x, const T* coeff, T* y, int64_t ystride,
                             const T* G, const int32_t* dofmap, const T* dphi,
                             int64_t ncells, int P, int nranks) {
#pragma omp parallel for num_threads(nranks) schedule(static, 1)
  for (int t = 0; t < nranks; ++t) {
    const int64_t a = ncells * t / nranks, b = ncells * (t + 1) / nranks;
    FN(orc_stiffness_range)(x, coeff, y + (int64_t)t * ystride, G, dofmap, dphi, a, b, P);
  }
}

void FN(orc_mass_ranks)(const T* x, const T* coeff, T* y, int64_t ystride,
                        const T* detJ, const int32_t* dofmap, int64_t nent,
                        int ncols, int nranks) {
#pragma omp parallel for num_threads(nranks) schedule(static, 1)
  for (int t = 0; t < nranks; ++t) {
    const int64_t a = nent * t / nranks, b = nent * (t + 1) / nranks;
    FN(orc_mass)(x, coeff + a, y + (int64_t)t * ystride, detJ + a * ncols,
                 dofmap + a * ncols, b - a, ncols);
  }
}

/* numba-cpu/operators.py:230-300, cuda/operators.py:195-274 */
void FN(orc_axpy)(T alpha, const T* x, T* y, int64_t n) {
  for (int64_t i = 0; i < n; ++i) y[i] = alpha * x[i] + y[i];
}
void FN(orc_copy)(const T* a, T* b, int64_t n) {
  for (int64_t i = 0; i < n; ++i) b[i] = a[i];
}
void FN(orc_fill)(T alpha, T* x, int64_t n) {
  for (int64_t i = 0; i < n; ++i) x[i] = alpha;
}
void FN(orc_pointwise_divide)(const T* a, const T* b, T* c, int64_t n) {
  for (int64_t i = 0; i < n; ++i) c[i] = a[i] / b[i];
}
void FN(orc_square)(const T* a, T* b, int64_t n) {
  for (int64_t i = 0; i < n; ++i) b[i] = a[i] * a[i];
}

/* cuda/scatterer.py:18-101 (numba-cpu/scatterer.py:18-75).  N = size_local. */
void FN(orc_pack_fwd)(const T* in, T* out, const int64_t* idx, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = in[idx[i]];
}
void FN(orc_unpack_fwd)(const T* in, T* out, const int64_t* idx, int64_t n, int64_t N) {
  for (int64_t i = 0; i < n; ++i) out[idx[i] + N] = in[i];
}
void FN(orc_pack_rev)(const T* in, T* out, const int64_t* idx, int64_t n, int64_t N) {
  for (int64_t i = 0; i < n; ++i) out[i] = in[idx[i] + N];
}
void FN(orc_unpack_rev)(const T* in, T* out, const int64_t* idx, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[idx[i]] += in[i];
}

/* ---- geometry: cuda/precompute.py (== numba-cpu/precompute.py) ---------- */

/* J_[d][c] = sum_v dphi[d, q, v] * coord[v][c]     (precompute.py:110, 150) */
static void FN(jac)(const T* dphi, int nq, int q, const T coord[8][3], T J[3][3]) {
  for (int d = 0; d < 3; ++d)
    for (int c = 0; c < 3; ++c) {
      T s = 0;
      for (int v = 0; v < 8; ++v) s += dphi[((int64_t)d * nq + q) * 8 + v] * coord[v][c];
      J[d][c] = s;
    }
}

static T FN(det3)(const T J[3][3]) {
  return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
         J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

/* cuda/precompute.py:76-112 */
void FN(orc_detJ)(T* detJ, const int32_t* x_dofs, const T* x_g, int64_t ncells,
                  const T* dphi, const T* w, int nq) {
  for (int64_t cell = 0; cell < ncells; ++cell) {
    T coord[8][3];
    for (int v = 0; v < 8; ++v)
      for (int c = 0; c < 3; ++c) coord[v][c] = x_g[(int64_t)x_dofs[cell * 8 + v] * 3 + c];
    for (int q = 0; q < nq; ++q) {
      T J[3][3];
      FN(jac)(dphi, nq, q, coord, J);
      detJ[cell * nq + q] = (T)fabs((double)FN(det3)(J)) * w[q];
    }
  }
}

/* cuda/precompute.py:115-163.  G_ = inv(J_)^T inv(J_), upper triangle scaled
 * by |det J_| w_q.  inv by the adjugate (the reference calls LAPACK). */
void FN(orc_G)(T* G, const int32_t* x_dofs, const T* x_g, int64_t ncells,
               const T* dphi, const T* w, int nq) {
  for (int64_t cell = 0; cell < ncells; ++cell) {
    T coord[8][3];
    for (int v = 0; v < 8; ++v)
      for (int c = 0; c < 3; ++c) coord[v][c] = x_g[(int64_t)x_dofs[cell * 8 + v] * 3 + c];
    for (int q = 0; q < nq; ++q) {
      T J[3][3], A[3][3];
      FN(jac)(dphi, nq, q, coord, J);
      const T det = FN(det3)(J);
      const T id = (T)1 / det;
      A[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
      A[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
      A[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      A[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
      A[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
      A[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      A[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
      A[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
      A[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
      const T s = (T)fabs((double)det) * w[q];
      T* g = G + (cell * nq + q) * 6;
      int t = 0;
      for (int a = 0; a < 3; ++a)
        for (int b = a; b < 3; ++b) {
          T acc = 0;
          for (int k = 0; k < 3; ++k) acc += A[k][a] * A[k][b]; /* (A^T A)[a][b] */
          g[t++] = s * acc;
        }
    }
  }
}

/* cuda/precompute.py:17-73.  boundary_data (nf, 2) = (cell, local facet);
 * dphi_f (6, 3, nq_f, 8). */
void FN(orc_detJ_facet)(T* detJ_f, const int32_t* x_dofs, const T* x_g,
                        const int32_t* boundary_data, int64_t nf, const T* dphi_f,
                        const T* w, int nq) {
  static const T R[6][3][2] = {
      {{1, 0}, {0, 1}, {0, 0}}, {{1, 0}, {0, 0}, {0, 1}}, {{0, 0}, {1, 0}, {0, 1}},
      {{0, 0}, {1, 0}, {0, 1}}, {{1, 0}, {0, 0}, {0, 1}}, {{1, 0}, {0, 1}, {0, 0}}};
  for (int64_t i = 0; i < nf; ++i) {
    const int64_t cell = boundary_data[2 * i];
    const int f = boundary_data[2 * i + 1];
    T coord[8][3];
    for (int v = 0; v < 8; ++v)
      for (int c = 0; c < 3; ++c) coord[v][c] = x_g[(int64_t)x_dofs[cell * 8 + v] * 3 + c];
    for (int q = 0; q < nq; ++q) {
      T J[3][3], F[3][2];
      FN(jac)(dphi_f + (int64_t)f * 3 * nq * 8, nq, q, coord, J);
      for (int c = 0; c < 3; ++c)
        for (int t = 0; t < 2; ++t) {
          T s = 0;
          for (int d = 0; d < 3; ++d) s += J[d][c] * R[f][d][t]; /* J_cell.T @ R */
          F[c][t] = s;
        }
      const T cx = F[1][0] * F[2][1] - F[2][0] * F[1][1];
      const T cy = F[2][0] * F[0][1] - F[0][0] * F[2][1];
      const T cz = F[0][0] * F[1][1] - F[1][0] * F[0][1];
      detJ_f[i * nq + q] = (T)sqrt((double)(cx * cx + cy * cy + cz * cz)) * w[q];
    }
  }
}

#undef FN
#undef CAT
#undef CAT_
